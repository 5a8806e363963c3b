// prep.cu -- the stage between the log-mel and the first convolution, fused:
//   bn0 (BatchNorm2d over the mel axis)  ->  SpecAugment stripes  ->  mixup of clip pairs.
//
// Replaces /root/reference/pytorch/models.py:202-211 (transpose / bn0 / transpose,
// `self.spec_augmenter(x)`, `do_mixup(x, mixup_lambda)`), i.e. torchlibrosa's
// SpecAugmentation.forward (4*B2 slice-fill launches) and pytorch_utils.py:80-93 (5 elementwise
// kernels), as ONE read of the log-mel and ONE write of the conv1 input.
// The stripe (begin, width) table is drawn on the host from the torch CPU generator in the
// reference's order (bit-exact indices) and uploaded once per step.
#include "common.cuh"

namespace sed {
namespace {

// per-column (sum, sumsq) partials of a row-major fp32 (rows, C) matrix, C <= 256
__global__ void colstats_kernel(const float* __restrict__ x, long long rows, int C, float* __restrict__ partial) {
  extern __shared__ float s_red[];                         // [lanes][2*C]
  const int lanes = blockDim.x / C;
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float s = 0.f, ss = 0.f;
  if (pl < lanes)
    for (long long r = r0 + pl; r < r1; r += lanes) {
      const float v = x[r * C + c];
      s += v;
      ss += v * v;
    }
  if (pl < lanes) {
    s_red[pl * 2 * C + c] = s;
    s_red[pl * 2 * C + C + c] = ss;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * C + i];
    partial[(long long)blockIdx.x * 2 * C + i] = a;
  }
}

struct Stripes {
  const int* t;   // (B2, nt, 2) (begin, width) along time, or nullptr
  const int* f;   // (B2, nf, 2) along mel, or nullptr
  int nt, nf;
};

__device__ __forceinline__ bool masked(const Stripes& s, int clip, int t, int m) {
  bool hit = false;
  if (s.t)
    for (int i = 0; i < s.nt; ++i) {
      const int b = s.t[(clip * s.nt + i) * 2], w = s.t[(clip * s.nt + i) * 2 + 1];
      hit |= (t >= b) & (t < b + w);
    }
  if (s.f)
    for (int i = 0; i < s.nf; ++i) {
      const int b = s.f[(clip * s.nf + i) * 2], w = s.f[(clip * s.nf + i) * 2 + 1];
      hit |= (m >= b) & (m < b + w);
    }
  return hit;
}

// out[j] = aug(bn0(x[j]))                                   (lam == nullptr, Bout = B2)
// out[i] = aug(bn0(x[2i]))*lam[2i] + aug(bn0(x[2i+1]))*lam[2i+1]   (mixup,  Bout = B2/2)
__global__ void bn0_aug_mix_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                       const float* __restrict__ shift, Stripes st, const float* __restrict__ lam,
                                       int Bout, int T, int M, float* __restrict__ out) {
  const long long total = (long long)Bout * T * M;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i % M);
    const long long r = i / M;
    const int t = (int)(r % T);
    const int b = (int)(r / T);
    const float sc = scale[m], sh = shift[m];
    if (lam) {
      const int c0 = 2 * b, c1 = 2 * b + 1;
      const long long per = (long long)T * M;
      float a0 = masked(st, c0, t, m) ? 0.f : fmaf(x[c0 * per + (long long)t * M + m], sc, sh);
      float a1 = masked(st, c1, t, m) ? 0.f : fmaf(x[c1 * per + (long long)t * M + m], sc, sh);
      out[i] = __fadd_rn(__fmul_rn(a0, lam[c0]), __fmul_rn(a1, lam[c1]));
    } else {
      out[i] = masked(st, b, t, m) ? 0.f : fmaf(x[i], sc, sh);
    }
  }
}

// bn0 parameter gradients: partial sums over (clip, frame) of d and d*xhat per mel bin, where
// d = dOut[clip or clip/2] * lam[clip] * [not masked].  block = M x lanes threads.
__global__ void bn0_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                      const float* __restrict__ mean, const float* __restrict__ invstd, Stripes st,
                                      const float* __restrict__ lam, int B2, int T, int M,
                                      float* __restrict__ partial) {
  extern __shared__ float s_red[];
  const int lanes = blockDim.x / M;
  const int m = threadIdx.x % M, pl = threadIdx.x / M;
  const long long rows = (long long)B2 * T;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float s = 0.f, sx = 0.f;
  if (pl < lanes) {
    const float mu = mean[m], is = invstd[m];
    for (long long r = r0 + pl; r < r1; r += lanes) {
      const int t = (int)(r % T);
      const int clip = (int)(r / T);
      if (masked(st, clip, t, m)) continue;
      const int ob = lam ? clip / 2 : clip;
      float d = dout[((long long)ob * T + t) * M + m];
      if (lam) d *= lam[clip];
      s += d;
      sx += d * (x[r * M + m] - mu) * is;
    }
  }
  if (pl < lanes) {
    s_red[pl * 2 * M + m] = s;
    s_red[pl * 2 * M + M + m] = sx;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * M; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += s_red[l * 2 * M + i];
    partial[(long long)blockIdx.x * 2 * M + i] = a;
  }
}

// in-place SpecAugment on a contiguous (B, C, T, F) fp32 tensor (stand-alone seam A module)
__global__ void spec_augment_kernel(float* __restrict__ x, int B, int C, int T, int F, Stripes st) {
  const long long total = (long long)B * C * T * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    long long r = i / F;
    const int t = (int)(r % T);
    r /= T;
    const int b = (int)(r / C);
    if (masked(st, b, t, f)) x[i] = 0.f;
  }
}

// do_mixup on a small (B2, n) fp32 matrix (the targets, main.py:246): out[i] = x[2i]*lam[2i] + x[2i+1]*lam[2i+1]
__global__ void mix_pairs_kernel(const float* __restrict__ x, const float* __restrict__ lam, int Bout, int n,
                                 float* __restrict__ out) {
  const long long total = (long long)Bout * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / n), c = (int)(i % n);
    // separate roundings (no FMA contraction): bit-identical to torch's mul, mul, add
    out[i] = __fadd_rn(__fmul_rn(x[(long long)(2 * b) * n + c], lam[2 * b]), __fmul_rn(x[(long long)(2 * b + 1) * n + c], lam[2 * b + 1]));
  }
}

// out[i] = sum over P partial rows (fp64 accumulate) -- generic deterministic second-level reduction
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int P, long long n, float* __restrict__ out,
                                       int accumulate, float scale) {
  const long long i = blockIdx.x * 32LL + threadIdx.x;
  double a, unused;
  colsum2_block(partial, nullptr, P, n, i, i < n, a, unused);
  if (threadIdx.y != 0 || i >= n) return;
  const float v = (float)a * scale;
  out[i] = accumulate ? out[i] + v : v;
}

}  // namespace
}  // namespace sed

using namespace sed;

extern "C" {

int sed_stat_partials(void) { return sm_count() * 4; }

int sed_colstats_f32(const float* x, long long rows, int C, float* partial, sed_stream_t stream) {
  SED_REQUIRE(x && partial && rows >= 1, "sed_colstats_f32: bad arguments");
  SED_REQUIRE(C >= 1 && C <= 256 && 256 % C == 0, "sed_colstats_f32: C=%d must divide 256", C);
  const int lanes = 256 / C;
  colstats_kernel<<<sed_stat_partials(), 256, (size_t)lanes * 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      x, rows, C, partial);
  SED_LAUNCH_CHECK("colstats_kernel");
  return 0;
}

int sed_bn0_aug_mix_fwd(const float* logmel, const float* scale, const float* shift, const int* t_stripes, int nt,
                        const int* f_stripes, int nf, const float* lam, int B2, int T, int M, float* out,
                        sed_stream_t stream) {
  SED_REQUIRE(logmel && scale && shift && out, "sed_bn0_aug_mix_fwd: null pointer");
  SED_REQUIRE(!lam || B2 % 2 == 0, "sed_bn0_aug_mix_fwd: mixup needs an even number of clips (got %d)", B2);
  if (B2 == 0) return 0;
  const int Bout = lam ? B2 / 2 : B2;
  Stripes st{t_stripes, f_stripes, nt, nf};
  const long long total = (long long)Bout * T * M;
  const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 16);
  bn0_aug_mix_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logmel, scale, shift, st, lam, Bout, T, M, out);
  SED_LAUNCH_CHECK("bn0_aug_mix_fwd_kernel");
  return 0;
}

int sed_bn0_bwd_reduce(const float* dout, const float* logmel, const float* mean, const float* invstd,
                       const int* t_stripes, int nt, const int* f_stripes, int nf, const float* lam, int B2, int T,
                       int M, float* partial, sed_stream_t stream) {
  SED_REQUIRE(dout && logmel && mean && invstd && partial, "sed_bn0_bwd_reduce: null pointer");
  SED_REQUIRE(M >= 1 && M <= 256 && 256 % M == 0, "sed_bn0_bwd_reduce: M=%d must divide 256", M);
  Stripes st{t_stripes, f_stripes, nt, nf};
  const int lanes = 256 / M;
  bn0_bwd_reduce_kernel<<<sed_stat_partials(), 256, (size_t)lanes * 2 * M * sizeof(float), (cudaStream_t)stream>>>(
      dout, logmel, mean, invstd, st, lam, B2, T, M, partial);
  SED_LAUNCH_CHECK("bn0_bwd_reduce_kernel");
  return 0;
}

int sed_spec_augment_f32(float* x, int B, int C, int T, int F, const int* t_stripes, int nt, const int* f_stripes,
                         int nf, sed_stream_t stream) {
  SED_REQUIRE(x, "sed_spec_augment_f32: null pointer");
  if ((long long)B * C * T * F == 0) return 0;
  Stripes st{t_stripes, f_stripes, nt, nf};
  const long long total = (long long)B * C * T * F;
  const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 16);
  spec_augment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, C, T, F, st);
  SED_LAUNCH_CHECK("spec_augment_kernel");
  return 0;
}

int sed_mix_pairs_f32(const float* x, const float* lam, int B2, int n, float* out, sed_stream_t stream) {
  SED_REQUIRE(x && lam && out, "sed_mix_pairs_f32: null pointer");
  SED_REQUIRE(B2 % 2 == 0 && n >= 1, "sed_mix_pairs_f32: needs an even number of rows (got %d)", B2);
  if (B2 == 0) return 0;
  const long long total = (long long)(B2 / 2) * n;
  const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 8);
  mix_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, lam, B2 / 2, n, out);
  SED_LAUNCH_CHECK("mix_pairs_kernel");
  return 0;
}

int sed_reduce_partials(const float* partial, int P, long long n, float* out, int accumulate, float scale,
                        sed_stream_t stream) {
  SED_REQUIRE(partial && out && P >= 1 && n >= 1, "sed_reduce_partials: bad arguments");
  SED_REQUIRE(n < (1LL << 36), "sed_reduce_partials: n too large");
  reduce_partials_kernel<<<(unsigned)((n + 31) / 32), dim3(32, kColLanes), 0, (cudaStream_t)stream>>>(partial, P, n, out, accumulate, scale);
  SED_LAUNCH_CHECK("reduce_partials_kernel");
  return 0;
}

}  // extern "C"
