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

// ---- vectorised forms for M % 4 == 0 (the model: M = 64): a thread owns 4 consecutive mel bins of a (clip, frame)
// row; the time-stripe test runs once per row and the frequency-stripe test once per clip change, there is no
// 64-bit division, and every access is a 16-byte vector.
__device__ __forceinline__ bool time_masked(const Stripes& s, int clip, int t) {
  bool hit = false;
  if (s.t)
    for (int i = 0; i < s.nt; ++i) {
      const int2 bw = __ldg(reinterpret_cast<const int2*>(s.t) + clip * s.nt + i);
      hit |= (t >= bw.x) & (t < bw.x + bw.y);
    }
  return hit;
}
// bit e set <=> mel bin m0 + e survives the frequency stripes of `clip`
__device__ __forceinline__ unsigned keep_mask4(const Stripes& s, int clip, int m0) {
  unsigned keep = 0xFu;
  if (s.f)
    for (int i = 0; i < s.nf; ++i) {
      const int2 bw = __ldg(reinterpret_cast<const int2*>(s.f) + clip * s.nf + i);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (m0 + e >= bw.x && m0 + e < bw.x + bw.y) keep &= ~(1u << e);
    }
  return keep;
}
__device__ __forceinline__ float4 bn_keep4(float4 x, float4 sc, float4 sh, unsigned keep, bool tmask) {
  float4 r;
  r.x = (!tmask && (keep & 1u)) ? fmaf(x.x, sc.x, sh.x) : 0.f;
  r.y = (!tmask && (keep & 2u)) ? fmaf(x.y, sc.y, sh.y) : 0.f;
  r.z = (!tmask && (keep & 4u)) ? fmaf(x.z, sc.z, sh.z) : 0.f;
  r.w = (!tmask && (keep & 8u)) ? fmaf(x.w, sc.w, sh.w) : 0.f;
  return r;
}

// block = (M/4, 256/(M/4)): threadIdx.x = mel quad, threadIdx.y = row lane; rows = (output clip, frame)
__global__ void bn0_aug_mix_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                           const float* __restrict__ shift, Stripes st, const float* __restrict__ lam,
                                           int Bout, int T, int M, float* __restrict__ out) {
  const int m0 = threadIdx.x * 4;
  const float4 sc = *reinterpret_cast<const float4*>(scale + m0), sh = *reinterpret_cast<const float4*>(shift + m0);
  const int rows = Bout * T;
  const int row_step = gridDim.x * blockDim.y;
  int cached_b = -1;
  unsigned keep0 = 0xF, keep1 = 0xF;
  float l0 = 1.f, l1 = 0.f;
  for (int row = blockIdx.x * blockDim.y + threadIdx.y; row < rows; row += row_step) {
    const int b = row / T, t = row - b * T;
    if (b != cached_b) {
      cached_b = b;
      if (lam) {
        keep0 = keep_mask4(st, 2 * b, m0); keep1 = keep_mask4(st, 2 * b + 1, m0);
        l0 = lam[2 * b]; l1 = lam[2 * b + 1];
      } else {
        keep0 = keep_mask4(st, b, m0);
      }
    }
    float4 o;
    if (lam) {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + ((long long)(2 * b) * T + t) * M + m0));
      const float4 x1 = __ldg(reinterpret_cast<const float4*>(x + ((long long)(2 * b + 1) * T + t) * M + m0));
      const float4 a0 = bn_keep4(x0, sc, sh, keep0, time_masked(st, 2 * b, t));
      const float4 a1 = bn_keep4(x1, sc, sh, keep1, time_masked(st, 2 * b + 1, t));
      // separate roundings (no FMA contraction): bit-identical to torch's mul, mul, add
      o.x = __fadd_rn(__fmul_rn(a0.x, l0), __fmul_rn(a1.x, l1));
      o.y = __fadd_rn(__fmul_rn(a0.y, l0), __fmul_rn(a1.y, l1));
      o.z = __fadd_rn(__fmul_rn(a0.z, l0), __fmul_rn(a1.z, l1));
      o.w = __fadd_rn(__fmul_rn(a0.w, l0), __fmul_rn(a1.w, l1));
    } else {
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(x + (long long)row * M + m0));
      o = bn_keep4(x0, sc, sh, keep0, time_masked(st, b, t));
    }
    *reinterpret_cast<float4*>(out + (long long)row * M + m0) = o;
  }
}

// rows = (input clip, frame); partial[blk][2][M]
__global__ void bn0_bwd_reduce_vec_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                          const float* __restrict__ mean, const float* __restrict__ invstd, Stripes st,
                                          const float* __restrict__ lam, int B2, int T, int M,
                                          float* __restrict__ partial) {
  extern __shared__ float s_red[];                       // [blockDim.y][2*M]
  const int m0 = threadIdx.x * 4;
  const float4 mu = *reinterpret_cast<const float4*>(mean + m0), is = *reinterpret_cast<const float4*>(invstd + m0);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), sx = s;
  const int rows = B2 * T;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  int cached = -1;
  unsigned keep = 0xF;
  float l = 1.f;
  for (int row = r0 + threadIdx.y; row < r1; row += blockDim.y) {
    const int clip = row / T, t = row - clip * T;
    if (clip != cached) {
      cached = clip;
      keep = keep_mask4(st, clip, m0);
      l = lam ? lam[clip] : 1.f;
    }
    if (time_masked(st, clip, t)) continue;
    const int ob = lam ? clip >> 1 : clip;
    const float4 d = __ldg(reinterpret_cast<const float4*>(dout + ((long long)ob * T + t) * M + m0));
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)row * M + m0));
    const float d0 = (keep & 1u) ? d.x * l : 0.f, d1 = (keep & 2u) ? d.y * l : 0.f;
    const float d2 = (keep & 4u) ? d.z * l : 0.f, d3 = (keep & 8u) ? d.w * l : 0.f;
    s.x += d0; s.y += d1; s.z += d2; s.w += d3;
    sx.x += d0 * (xv.x - mu.x) * is.x; sx.y += d1 * (xv.y - mu.y) * is.y;
    sx.z += d2 * (xv.z - mu.z) * is.z; sx.w += d3 * (xv.w - mu.w) * is.w;
  }
  float* mine = s_red + threadIdx.y * 2 * M;
  *reinterpret_cast<float4*>(mine + m0) = s;
  *reinterpret_cast<float4*>(mine + M + m0) = sx;
  __syncthreads();
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 2 * M; i += blockDim.x * blockDim.y) {
    float a = 0.f;
    for (int ln = 0; ln < (int)blockDim.y; ++ln) a += s_red[ln * 2 * M + i];
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
  if (M % 4 == 0 && 256 % (M / 4) == 0 && (long long)B2 * T < (1LL << 31) && aligned(logmel, 16) && aligned(out, 16) &&
      aligned(scale, 16) && aligned(shift, 16) && (!t_stripes || aligned(t_stripes, 8)) && (!f_stripes || aligned(f_stripes, 8))) {
    const dim3 block(M / 4, 256 / (M / 4));
    const int gridv = (int)min(((long long)Bout * T + block.y - 1) / block.y, (long long)sm_count() * 8);
    bn0_aug_mix_fwd_vec_kernel<<<gridv, block, 0, (cudaStream_t)stream>>>(logmel, scale, shift, st, lam, Bout, T, M, out);
    SED_LAUNCH_CHECK("bn0_aug_mix_fwd_vec_kernel");
    return 0;
  }
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
  if (M % 4 == 0 && 256 % (M / 4) == 0 && (long long)B2 * T < (1LL << 31) && aligned(logmel, 16) && aligned(dout, 16) &&
      aligned(mean, 16) && aligned(invstd, 16) && (!t_stripes || aligned(t_stripes, 8)) && (!f_stripes || aligned(f_stripes, 8))) {
    const dim3 block(M / 4, 256 / (M / 4));
    bn0_bwd_reduce_vec_kernel<<<sed_stat_partials(), block, (size_t)block.y * 2 * M * sizeof(float), (cudaStream_t)stream>>>(
        dout, logmel, mean, invstd, st, lam, B2, T, M, partial);
    SED_LAUNCH_CHECK("bn0_bwd_reduce_vec_kernel");
    return 0;
  }
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
