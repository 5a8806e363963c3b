// attention.cu -- the dropout + ReLU that follows the output projection of MultiHead
// (/root/reference/pytorch/models.py:664: F.relu_(self.dropout(self.fc(output)))).
// The attention core itself (QK^T, softmax, dropout, PV and its backward) is csrc/attention_tc.cu.
#include "common.cuh"
#include "philox.cuh"

namespace sed {
namespace {

// y = relu(dropout_p(x)) elementwise (models.py:664: F.relu_(self.dropout(self.fc(output))))
__global__ void dropout_relu_fwd_kernel(const float* __restrict__ x, long long n, float p, unsigned long long seed,
                                        unsigned long long offset, const unsigned long long* __restrict__ state,
                                        float* __restrict__ y) {
  if (state != nullptr) { seed = state[0]; offset += state[1]; }
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (p > 0.f) v = keep_elem(seed, offset, (unsigned long long)i, p) ? v * keep_scale : 0.f;
    y[i] = fmaxf(v, 0.f);
  }
}
// dx = dy * [y > 0] / (1 - p): y > 0 implies the element was kept
__global__ void dropout_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float p,
                                        float* __restrict__ dx) {
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] * keep_scale : 0.f;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_dropout_relu_fwd(const float* x, long long n, float p_drop, unsigned long long seed, unsigned long long offset,
                         const unsigned long long* philox_state, float* y, sed_stream_t stream) {
  SED_REQUIRE(x && y && n >= 0 && p_drop >= 0.f && p_drop < 1.f, "sed_dropout_relu_fwd: bad arguments");
  if (n == 0) return 0;
  const int grid = (int)min((n + 255) / 256, (long long)sm_count() * 16);
  dropout_relu_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, p_drop, seed, offset, philox_state, y);
  SED_LAUNCH_CHECK("dropout_relu_fwd_kernel");
  return 0;
}

int sed_dropout_relu_bwd(const float* dy, const float* y, long long n, float p_drop, float* dx, sed_stream_t stream) {
  SED_REQUIRE(dy && y && dx && n >= 0 && p_drop >= 0.f && p_drop < 1.f, "sed_dropout_relu_bwd: bad arguments");
  if (n == 0) return 0;
  const int grid = (int)min((n + 255) / 256, (long long)sm_count() * 16);
  dropout_relu_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, y, n, p_drop, dx);
  SED_LAUNCH_CHECK("dropout_relu_bwd_kernel");
  return 0;
}

}  // extern "C"
