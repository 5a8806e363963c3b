/* sed_b200.h -- C ABI of libsedb200.so: the B200 (sm_100a) kernels behind the reference's
 * per-clip training path.
 *
 * Boundary rules:
 *  - plain pointers + sizes + an explicit cudaStream_t (passed as void*); no torch types.
 *  - every tensor (inputs, outputs, saved-for-backward, workspaces) is allocated by the caller
 *    (PyTorch's caching allocator on the Python side) and only BORROWED for the duration of the
 *    enqueue; the library never allocates or frees device memory, never synchronises, never
 *    changes the current device.
 *  - every entry point returns 0 on success and a non-zero status otherwise; the message is
 *    retrievable (per thread) with sed_last_error_string().  No C++ exceptions cross the ABI.
 *  - entry points are re-entrant and thread-safe (DataParallel calls forward from one Python
 *    thread per GPU; autograd runs backward on per-device worker threads).
 *  - there is NO CPU fallback inside the library.
 *
 * Each entry point cites the reference interface it replaces, relative to /root/reference.
 * Layouts: activations are NHWC ("pixels x channels", H = time frames, W = mel/freq bins),
 * bf16 unless stated; weights for the tensor-core convolutions are repacked bf16 shadows of the
 * reference's OIHW fp32 master copies.
 */
#ifndef SED_B200_H_
#define SED_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef void* sed_stream_t; /* a cudaStream_t */

/* ---- library ------------------------------------------------------------------------------ */
const char* sed_last_error_string(void);
int sed_abi_version(void);
/* Number of kernels this library has launched on any stream since load (thread-safe counter);
 * bench.py reports the delta over the timed region as "gpu_launches". */
unsigned long long sed_launch_count(void);
int sed_device_sm_count(int* out_sms);

/* ---- front-end: torchlibrosa.stft.Spectrogram + LogmelFilterBank ---------------------------
 * replaces pytorch/models.py:199-200 (Spectrogram.forward -> LogmelFilterBank.forward) and the
 * six sibling call sites; ctor contracts pytorch/models.py:166-173.
 * n_fft is fixed at 1024 (utils/config.py:11); window = periodic Hann; center=True, reflect pad.
 *
 * Mel bank in CSR-by-mel form (built on the host from the module's frozen melW parameter):
 *   mel_lo[m]   first FFT bin with a non-zero tap,  mel_off[m]..mel_off[m+1] its taps in mel_w.
 * out = 10*log10(max(mel, amin)) - db_offset,  db_offset = 10*log10(max(amin, ref)).
 * Limits of the fused kernels (the caller checks them; the tables live in device memory): n_mels <= 128 and
 * sum over filters of ceil(taps / 16) <= 128 (the reference bank: 64 filters, 866 taps, 90 pieces).
 */
int sed_logmel_f32(const float* wave, int n_clips, int n_samples, int hop,
                   const float* mel_w, const int* mel_lo, const int* mel_off, int n_mels,
                   float amin, float db_offset, float* out /* (n_clips, T, n_mels) */,
                   sed_stream_t stream);
/* Same with int16 PCM input; x/32767 (utils/utilities.py:66-67) is fused into the frame gather. */
int sed_logmel_i16(const int16_t* pcm, int n_clips, int n_samples, int hop,
                   const float* mel_w, const int* mel_lo, const int* mel_off, int n_mels,
                   float amin, float db_offset, float* out, sed_stream_t stream);
/* Unfused seam A: Spectrogram.forward alone -> (n_clips, T, 513) power spectrogram. */
int sed_stft_power_f32(const float* wave, int n_clips, int n_samples, int hop,
                       float* out_power, sed_stream_t stream);
/* Unfused seam A: LogmelFilterBank.forward alone on a (rows, 513) power spectrogram. */
int sed_mel_db_f32(const float* power, long long rows, int n_bins,
                   const float* mel_w, const int* mel_lo, const int* mel_off, int n_mels,
                   float amin, float db_offset, int is_log, float* out, sed_stream_t stream);


/* ---- 3x3 convolutions on tensor cores (tcgen05 + TMA) --------------------------------------
 * replaces the cuDNN calls behind ConvBlock.conv1 / conv2 (pytorch/models.py:75-83, :102-103)
 * and their autograd backward (pytorch/main.py:257).  Activations NHWC bf16 [B][H][W][C]
 * (H = frames, W = mel bins); 128 % W == 0, W >= 8; Cin % 64 == 0.
 */
/* fp32 OIHW master weights -> bf16 shadows: fwd [Cout][tap][Cin]; dgrad [Cin][8-tap][Cout]. */
int sed_conv_pack_weights(const float* w_oihw, int Cout, int Cin, void* fwd_pack, void* dgrad_pack,
                          sed_stream_t stream);
/* number of rows of the stats_partial workspace (4 per CTA of the persistent grid: one per TMEM lane quarter). */
int sed_conv3x3_tc_grid(int B, int H, int W, int Cin, int Cout);
/* y = conv3x3(x, w) (bf16 out, fp32 accumulate).  stats_partial (optional):
 * [sed_conv3x3_tc_grid()][2][Cout] partial (sum, sum of squares) rows of the fp32 results, for the
 * training-mode BatchNorm that follows.  The data gradient is the same call with
 * x = dY, wpack = the dgrad pack and Cin/Cout swapped. */
int sed_conv3x3_tc_fwd(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H,
                       int W, int Cin, int Cout, sed_stream_t stream);
/* The same convolution on CTA pairs (tcgen05.mma.cta_group::2, M = 256 per instruction; each CTA stages half of the
 * weight tile).  Same arguments and results as sed_conv3x3_tc_fwd; statistics workspace rows
 * from sed_conv3x3_tc2_grid. */
int sed_conv3x3_tc2_grid(int B, int H, int W, int Cin, int Cout);
int sed_conv3x3_tc2_fwd(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H, int W, int Cin,
                        int Cout, sed_stream_t stream);
/* Cout = 64 variant with the three kw taps stacked in N (one N = 192 instruction per kernel row and K16 step; the kw
 * shift is applied between accumulator rows in the epilogue): same contract and statistics workspace as
 * sed_conv3x3_tc2_fwd; sed_conv3x3_tc2kw_supported tells whether a shape is covered (Cout == 64, Cin in {64, 128}). */
int sed_conv3x3_tc2kw_supported(int W, int Cin, int Cout);
int sed_conv3x3_tc2kw_fwd(const void* x, const void* wpack, void* y, float* stats_partial, int B, int H, int W, int Cin,
                          int Cout, sed_stream_t stream);
/* Data gradient fused with the FIRST pass of the BatchNorm+ReLU(+2x2 avg-pool) backward of the layer below
 * (F.relu_(bn(conv(x))) [+ F.avg_pool2d], pytorch/models.py:102-113): dx (B,H,W,Cout) is that layer's dA; y_below is
 * its raw conv output (B, Hy, W*pool, Cout) with Hy/pool == H (an odd Hy has a floor-mode tail row), scale/shift
 * its BatchNorm affine of this step.  partial: [sed_conv3x3_tc_grid()][2][Cout] per-CTA sums of g and g*y,
 * g = unpool(dx)/pool^2 * [y*scale + shift > 0]; feed them to sed_bn_bwd_finalize with mean_for_gy. */
int sed_conv3x3_tc_dgrad_bnr(const void* dy, const void* wpack_dgrad, void* dx, int B, int H, int W, int Cin,
                             int Cout, const void* y_below, int Hy, const float* scale_below,
                             const float* shift_below, int pool, float* partial, sed_stream_t stream);
/* weight gradient: slabs = [sed_conv3x3_tc_wgrad_splits()][9][Cout][Cin] fp32 split-K partials;
 * sed_conv_unpack_wgrad sums them (fixed order) into the OIHW fp32 gradient. */
int sed_conv3x3_tc_wgrad_splits(int B, int H, int W, int Cin, int Cout);
/* development switch: 1 (default) = CTA-pair kernel for Cout >= 256 / Cin >= 128, 0 = single-CTA kernel; returns the
 * previous setting.  Process-wide; meant for tests and A/B timing. */
int sed_conv3x3_tc_wgrad_use_pairs(int on);
int sed_conv3x3_tc_wgrad(const void* dy, const void* x, float* slabs, int B, int H, int W, int Cin,
                         int Cout, sed_stream_t stream);
int sed_conv_unpack_wgrad(const float* g_tap_major, int slabs, long long slab_stride, int Cout, int Cin,
                          float* grad_oihw, int accumulate, sed_stream_t stream);
/* The two above for up to 8 layers in ONE launch each (arrays of `layers` entries on the host; results bit-identical to
 * the per-layer entry points): weights are packed once at the top of the forward, the split-K slabs of all weight
 * gradients are folded once at the end of the backward. */
int sed_conv_pack_weights_multi(int layers, const float* const* w_oihw, const int* Cout, const int* Cin,
                                void* const* fwd_pack, void* const* dgrad_pack, sed_stream_t stream);
int sed_conv_unpack_wgrad_multi(int layers, const float* const* slabs, const int* n_slabs, const int* Cout, const int* Cin,
                                float* const* grad_oihw, sed_stream_t stream);
int sed_f32_to_bf16(const float* x, void* y, long long n, sed_stream_t stream);


/* ---- BatchNorm2d(train) + ReLU + avg-pool on NHWC bf16 ---------------------------------------
 * replaces F.relu_(self.bnX(...)) + F.avg_pool2d (pytorch/models.py:102-113) and, with
 * (ph, pw) = (1, W), block 4's identity pool + torch.mean(x, dim=3) (models.py:303 and siblings).
 * partial = [P][2][C] per-CTA (sum, sum of squares) rows written by the producing conv kernel. */
int sed_bn_finalize(const float* partial, int P, int C, double count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* scale, float* shift, float* save_mean,
                    float* save_invstd, sed_stream_t stream);
/* eval mode: affine from the running statistics; save_mean / save_invstd (optional) receive the frozen statistics in
 * the form the backward entry points take them */
int sed_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                       float eps, int C, float* scale, float* shift, float* save_mean, float* save_invstd,
                       sed_stream_t stream);
int sed_bn_relu_pool_fwd(const void* y, const float* scale, const float* shift, int B, int H, int W, int C, int ph,
                         int pw, void* out, int out_is_f32, sed_stream_t stream);
/* Backward of the same stage in two passes over y (no saved activation other than y itself):
 *   reduce   -> partial[P][2][C] = per-worker sums of g and g*y,  g = unpool(dA)/(ph*pw) * [y*scale + shift > 0],
 *               P = sed_bn_bwd_partials(shape);
 *   finalize -> dgamma, dbeta, coef[3][C] (pass the layer's batch mean as mean_for_gy);
 *   apply    -> dY = gamma*invstd * (g - mean(g) - xhat*mean(g*xhat))   (bf16).
 * The fast paths (pool 1x1, 1xW, 2x2; C/4 a power of two) are dynamically scheduled persistent kernels built to share
 * an SM with a weight-gradient CTA (<= 64 registers, 4 KB shared memory): `sched` points to
 * sed_bn_bwd_sched_words() int32 words (work tickets + worker-group arrival counters) that are zero before the call
 * and zero again after it (one block per stream that runs these kernels concurrently).  The reduce pass first writes
 * one row per worker, then the last worker of every group of 16 folds the group into one of the first P rows. */
int sed_bn_bwd_partials(int B, int H, int W, int C, int ph, int pw);        /* P: rows finalize reads = first rows of `partial` */
int sed_bn_bwd_workspace_rows(int B, int H, int W, int C, int ph, int pw);  /* rows of [2][C] floats `partial` must hold */
int sed_bn_bwd_sched_words(void);                                           /* int32 words behind `sched` */
int sed_bn_relu_pool_bwd_reduce(const void* y, const void* dA, int grad_is_f32, const float* scale,
                                const float* shift, int B, int H, int W, int C, int ph, int pw, float* partial,
                                int* sched, sed_stream_t stream);
/* mean_for_gy: NULL when the second partial column holds sum(g*xhat) (sed_bn0_bwd_reduce); the layer's batch mean
 * when it holds sum(g*y) (sed_bn_relu_pool_bwd_reduce, sed_conv3x3_tc_dgrad_bnr):
 * sum(g*xhat) = invstd*(sum(g*y) - mean*sum(g)). */
/* frozen_stats != 0: the statistics are constants (eval-mode BatchNorm under autograd): the batch-mean terms of the
 * coefficients are zero, dY = gamma*invstd*g. */
int sed_bn_bwd_finalize(const float* partial, int P, int C, double count, const float* gamma, const float* invstd,
                        const float* mean_for_gy, float* dgamma, float* dbeta, int accumulate, int frozen_stats,
                        float* coef, sed_stream_t stream);
int sed_bn_relu_pool_bwd_apply(const void* y, const void* dA, int grad_is_f32, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const float* coef, int B, int H, int W, int C,
                               int ph, int pw, void* dy, int* sched, sed_stream_t stream);

/* ---- bn0 + SpecAugment + mixup (pytorch/models.py:202-211, pytorch_utils.py:80-93) ------------
 * stripes: int32 (B2, n, 2) = (begin, width), drawn on the host in torchlibrosa's RNG order. */
int sed_stat_partials(void);
int sed_colstats_f32(const float* x, long long rows, int C, float* partial, sed_stream_t stream);
int sed_bn0_aug_mix_fwd(const float* logmel, const float* scale, const float* shift, const int* t_stripes, int nt,
                        const int* f_stripes, int nf, const float* lam, int B2, int T, int M, float* out,
                        sed_stream_t stream);
int sed_bn0_bwd_reduce(const float* dout, const float* logmel, const float* mean, const float* invstd,
                       const int* t_stripes, int nt, const int* f_stripes, int nf, const float* lam, int B2, int T,
                       int M, float* partial, sed_stream_t stream);
/* torchlibrosa.augmentation.SpecAugmentation.forward stand-alone (in place, (B,C,T,F) fp32). */
int sed_spec_augment_f32(float* x, int B, int C, int T, int F, const int* t_stripes, int nt, const int* f_stripes,
                         int nf, sed_stream_t stream);
/* pytorch_utils.do_mixup (pytorch_utils.py:80-93) on a small (B2, n) fp32 matrix -- the targets at
 * main.py:246: out[i] = x[2i]*lam[2i] + x[2i+1]*lam[2i+1], out (B2/2, n). */
int sed_mix_pairs_f32(const float* x, const float* lam, int B2, int n, float* out, sed_stream_t stream);
int sed_reduce_partials(const float* partial, int P, long long n, float* out, int accumulate, float scale,
                        sed_stream_t stream);

/* ---- first convolution, Cin = 1 (ConvBlock1.conv1, pytorch/models.py:181, :102) -------------- */
int sed_conv_c1_grid(void);
int sed_conv_c1_fwd(const float* x, const float* w, void* y, float* stats_partial, int B, int H, int W, int Cout,
                    sed_stream_t stream);
int sed_conv_c1_wgrad(const float* x, const void* dy, float* partial, int B, int H, int W, int Cout,
                      sed_stream_t stream);
/* sed_bn_relu_pool_bwd_apply (flat, bf16 dA) of block1.bn1 FUSED into the weight-gradient pass: y and dA stream in, dY is
 * formed on chip (same arithmetic, same bf16 rounding), feeds the weight-gradient product and is written to `dy` for
 * sed_conv_c1_dgrad.  coef = the (3, 64) output of sed_bn_bwd_finalize.  One pass over y and dA replaces the apply
 * pass plus the weight gradient's own pass over dY. */
int sed_bn_apply_conv_c1_wgrad(const float* x, const void* y, const void* dA, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const float* coef, void* dy, float* partial, int B,
                               int H, int W, int Cout, sed_stream_t stream);
int sed_conv_c1_dgrad(const void* dy, const float* w, float* dx, int B, int H, int W, int Cout, sed_stream_t stream);

/* ---- heads + loss (pytorch/models.py:118-149, :221-227, :306-312; pytorch/losses.py:5-12) ------ */
int sed_linear_partials(void);
int sed_linear_small_fwd(const float* x, const float* W, const float* bias, long long R, int C, int K, float* out,
                         sed_stream_t stream);
/* Two such maps of the SAME input in one pass over x (AttBlock's att and cla, models.py:137 / :141): out1 (R, K1),
 * out2 (R, K2); K1 + K2 <= 40, C <= 512. */
int sed_linear_pair_fwd(const float* x, const float* W1, const float* bias1, int K1, const float* W2, const float* bias2,
                        int K2, long long R, int C, float* out1, float* out2, sed_stream_t stream);
int sed_linear_small_bwd(const float* dout, const float* x, const float* W, long long R, int C, int K, float* dx,
                         int dx_accumulate, float* partial_w, float* partial_b, sed_stream_t stream);
int sed_head_pool_fwd(const float* logit, int B, int T, int K, int ratio, int mode, float* prob, float* clip,
                      int* argmax, float* frame, sed_stream_t stream);
int sed_head_pool_bwd(const float* prob, const float* dclip, const int* argmax, int B, int T, int K, int mode,
                      float* dlogit, sed_stream_t stream);
int sed_head_att_fwd(const float* att_logit, const float* cla_logit, int B, int T, int K, int ratio, int sigmoid_act,
                     float temperature, float* norm_att, float* cla, float* clip, float* frame, sed_stream_t stream);
int sed_head_att_bwd(const float* att_logit, const float* norm_att, const float* cla, const float* clip,
                     const float* dclip, int B, int T, int K, int sigmoid_act, float temperature, float* d_att_logit,
                     float* d_cla_logit, sed_stream_t stream);
int sed_bce_fwd_bwd(const float* prob, const float* target, long long n, float grad_scale, float* loss, float* dprob,
                    sed_stream_t stream);

/* ---- optimizer: optim.Adam(amsgrad=True).step() (pytorch/main.py:144-145, :258) ----------------
 * bias_corr_dev: NULL, or two device floats {1 - beta1^t, sqrt(1 - beta2^t)} that override the values computed from
 * `step` -- the launch can then be replayed from a CUDA graph with the step number supplied through memory. */
int sed_adam_amsgrad(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                     long long n, float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                     const float* bias_corr_dev, sed_stream_t stream);


/* ---- plain tensor-core GEMMs on the conv pipelines (GRU / attention projections) ---------------
 * sed_gemm_tc:    out[M][N] (fp32) = A[M][K] * Bw[N][K]^T (+ bias[N]);  A, Bw bf16 K-major.
 * sed_gemm_tn_tc: slabs[s][M][N] (fp32 split-K partials) = A[R][lda>=M]^T * Bm[R][ldb>=N]; bf16.
 * sed_split_bf16x3: fp32 (R,K) -> bf16 (R,3K) = [hi|lo|hi] (which=0) or [hi|hi|lo] (which=1) so
 *   that A'.B'^T = hi.hi + lo.hi + hi.lo, an fp32-accurate product on the bf16 tensor cores. */
int sed_gemm_tc(const void* a, const void* bw, const float* bias, float* out, long long M, int N, int K,
                sed_stream_t stream);
int sed_gemm_tn_tc_splits(long long R, int M, int N);
int sed_gemm_tn_tc(const void* a, int lda, const void* bm, int ldb, float* slabs, long long R, int M, int N,
                   sed_stream_t stream);
int sed_split_bf16x3(const float* x, long long R, int K, int which, void* out, sed_stream_t stream);
int sed_transpose_to_bf16(const float* x, int R, int C, void* out, sed_stream_t stream);
int sed_colsum_f32(const float* x, long long rows, int C, float* partial, sed_stream_t stream);

/* ---- bidirectional GRU recurrence: nn.GRU(512, 256, bidirectional) (pytorch/models.py:437-438,
 * :529-530; calls :475, :566).  gx = x W_ih^T + b_ih for both directions (B,T,2,3H), gate order
 * r,z,n; w_hh (2,3H,H), b_hh (2,3H); out (B,T,2H); gates (B,T,2,4,H) saved for the backward.
 * sync_ws: sed_gru_workspace_bytes(B, H, 0) bytes of 16-byte-aligned scratch: the exchange area of the persistent
 * kernel (H = 256: the group's h rows of each step travel through it as flagged 8-byte words; the entry point zeroes it
 * on the stream); NULL selects the per-step kernels.
 * bwd: carry = sed_gru_workspace_bytes(B, H, 1) bytes of 16-byte-aligned scratch (the (2,2,B,H) carry of the per-step
 * kernels, or the dGh exchange area of the persistent one); writes dgx, dgh (B,T,2,3H) and hprev (B,T,2,H). */
long long sed_gru_workspace_bytes(int B, int H, int backward);
int sed_gru_fwd(const float* gx, const float* w_hh, const float* b_hh, float* out, float* gates, void* sync_ws,
                int B, int T, int H, sed_stream_t stream);
/* dgx_bf16 / dgh_bf16 / hprev_bf16: optional (all three or none, H = 256 only) bf16 copies of the three outputs,
 * written by the same kernel: the operands of the weight-gradient GEMMs that follow (saves three conversion passes). */
/* bias_partial: optional (H = 256 only) fp32 (sed_gru_bwd_bias_rows(B), 2, 2, 3H) partial sums over (batch, time) of dgx
 * ([.][0][direction]) and dgh ([.][1][direction]): summed over the first dimension (sed_reduce_partials) they are the
 * gradients of b_ih and b_hh -- saves the two column-sum passes over dgx / dgh. */
int sed_gru_bwd_bias_rows(int B);
int sed_gru_bwd(const float* dout, const float* out, const float* gates, const float* w_hh, float* carry,
                float* dgx, float* dgh, float* hprev, void* dgx_bf16, void* dgh_bf16, void* hprev_bf16,
                float* bias_partial, int B, int T, int H, sed_stream_t stream);

/* ---- scaled-dot-product attention of MultiHead (pytorch/models.py:596-608 inside :641-665) -------
 * q/k/v: fp32 (B*T, ld) row-major, head h in columns [h*64, h*64+64): the outputs of the w_qs / w_ks /
 * w_vs projections used in place (replaces the four permute().contiguous() copies).  T <= 128, d = 64.
 * ctx (B*T, H*64); probs (B,H,T,T) = softmax before dropout (saved for the backward, may be NULL).
 * Every product (Q K^T, P V and the four of the backward) is a warp-level bf16 tensor-core tile product with fp32
 * accumulators and a hi/lo operand split (fp32-class accuracy); q/k/v/dctx must be 16-byte aligned, ld % 4 == 0.
 * Dropout(p_drop) on the probabilities from Philox4x32-10(seed, offset + element/4), element = (score row) * 128 +
 * key (the caller advances its generator by B*H*T*128 draws); p_drop = 0 in eval.
 * philox_state: NULL, or two device words {seed, offset base} read by the kernel at run time (then `seed` is ignored and
 * `offset` is relative to the base): lets a captured CUDA graph be replayed with a fresh generator state per step.
 * bwd: dq/dk/dv use the addressing of q/k/v. */
int sed_attention_fwd(const float* q, const float* k, const float* v, int ldq, int ldk, int ldv, int B, int T, int H,
                      int d, float temperature, float p_drop, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* philox_state, float* ctx, float* probs, sed_stream_t stream);
int sed_attention_bwd(const float* q, const float* k, const float* v, int ldq, int ldk, int ldv, int B, int T, int H,
                      int d, float temperature, float p_drop, unsigned long long seed, unsigned long long offset,
                      const unsigned long long* philox_state, const float* dctx, const float* probs, float* dq, float* dk,
                      float* dv, sed_stream_t stream);
/* y = relu(dropout_p(x)) (models.py:664) and its backward dx = dy * [y > 0] / (1 - p). */
int sed_dropout_relu_fwd(const float* x, long long n, float p_drop, unsigned long long seed, unsigned long long offset,
                         const unsigned long long* philox_state, float* y, sed_stream_t stream);
int sed_dropout_relu_bwd(const float* dy, const float* y, long long n, float p_drop, float* dx, sed_stream_t stream);

/* ---- frame-wise probabilities -> sound events (onset, offset frame indices), bit-exact ------------
 * replaces utils/utilities.py:70-123 (frame_prediction_to_event_prediction) and utils/vad.py:11-134
 * (activity_detection: double threshold, smoothing, salt removal) -- one thread per (clip, class) series.
 * frame (N, T, K) fp32, clip (N, K) fp32 or NULL (no audio-tagging gate), per-class parameter arrays of
 * length K (lo_thres may be NULL: single threshold).  sed_vad_count writes the number of pairs per series
 * and flags (bit 0: the reference would raise IndexError, vad.py:78); the caller turns counts into
 * exclusive offsets; sed_vad_fill writes the (bgn, fin) pairs in (n, k, pair) order. */
int sed_vad_count(const float* frame, const float* clip, int N, int T, int K, const float* at_thres,
                  const float* hi_thres, const float* lo_thres, const int* n_smooth, const int* n_salt,
                  int* counts, int* flags, sed_stream_t stream);
int sed_vad_fill(const float* frame, const float* clip, int N, int T, int K, const float* at_thres,
                 const float* hi_thres, const float* lo_thres, const int* n_smooth, const int* n_salt,
                 const long long* offsets, int* pairs, sed_stream_t stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* SED_B200_H_ */
