/*
 * semseg_b200 — C-ABI of the B200-native hot path of hszhao/semseg.
 *
 * Plain pointers and sizes only: no ATen / pybind types cross this boundary. Every entry point
 * takes device pointers, a cudaStream_t passed as void*, returns 0 on success or a negative
 * SEMSEG_E_* code, and never throws; semseg_last_error() returns the message of the last failure
 * on the calling thread. All kernels are sm_100a-only; there is no CPU fallback behind this ABI.
 *
 * What each group replaces in the reference (paths relative to the hszhao/semseg tree):
 *   - semseg_psamask_*        : lib/psa/src/gpu/operator.h:3-4 (psamask_forward_cuda / psamask_backward_cuda,
 *                               kernels lib/psa/src/gpu/psamask_cuda.cu:8-128) bound by
 *                               lib/psa/functions/psamask.py:19-35.
 *   - semseg_conv_*           : the cuDNN convolutions behind nn.Conv2d at model/resnet.py:63-69,108-113,133-137
 *                               (after the dilation patch model/pspnet.py:49-58), the heads
 *                               model/pspnet.py:65-69,73-77 and PSA 1x1s model/psanet.py:24-51.
 *   - semseg_bn_*             : nn.BatchNorm2d / nn.SyncBatchNorm (training + eval) and the in-place ReLU and
 *                               residual add around them, model/resnet.py:77-92.
 *   - semseg_pack_* / layout  : NCHW fp32 <-> NHWC bf16 at the module boundary (model/pspnet.py:80-105).
 *   - semseg_ppm_* / pool     : model/pspnet.py:12-26 (AdaptiveAvgPool2d, bilinear upsample, concat),
 *                               nn.MaxPool2d at model/resnet.py:115.
 *   - semseg_upsample_ce_*    : F.interpolate + CrossEntropyLoss + argmax, model/pspnet.py:94-103.
 *
 * Activations are NHWC bf16 in HBM; "pitch" arguments are the distance between consecutive pixels in
 * elements (>= channels; lets a kernel read/write a channel slice of a wider concat buffer).
 */
#ifndef SEMSEG_B200_H
#define SEMSEG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMSEG_OK 0
#define SEMSEG_E_INVALID (-1)  /* bad argument (shape, alignment, null pointer) */
#define SEMSEG_E_CUDA (-2)     /* CUDA runtime / driver error (message in semseg_last_error) */
#define SEMSEG_E_UNSUPPORTED (-3)

#define SEMSEG_MAX_TAPS 9

const char* semseg_last_error(void);
int semseg_abi_version(void);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
long long semseg_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * PSA mask (collect / distribute), fp32 NCHW exactly as the reference.
 *   psa_type 0 = collect, 1 = distribute (lib/psa/functions/psamask.py:9).
 *   fwd: in  [N, mH*mW, H, W] -> out [N, H*W, H, W]; every element of out is written (zeros included),
 *        so the caller does NOT need to pre-zero it (the reference requires a zeroed buffer).
 *   bwd: dout [N, H*W, H, W] -> din [N, mH*mW, H, W]; every element of din is written.
 */
int semseg_psamask_fwd(int psa_type, const float* in, float* out, int N, int H, int W, int mH, int mW,
                       void* stream);
int semseg_psamask_bwd(int psa_type, const float* dout, float* din, int N, int H, int W, int mH, int mW,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused point-wise spatial attention (model/psanet.py:81-91: psa_mask -> softmax(dim=1) -> bmm): the
 * [N, HW, HW] attention map never exists in HBM.
 *   attn  fp32 NHWC [N, H*W, a_pitch] (a_pitch >= mH*mW): the attention logits as the 1x1 conv's F32 epilogue writes them
 *   feat / out / dout  activations NHWC [N, H*W, C] (C = 512), plain bf16 or split (hi, lo)
 *   stats fp32 [N, H*W, 2] = (max, 1/sum) of every target's softmax (written by mode 0, read by the backward calls)
 * semseg_psa_attend  mode 0: out[t,:]   = scale * sum_s P[t,s] * feat[s,:]   (forward; P = softmax over sources s)
 *                    mode 1: out[s,:]   = scale * sum_t P[t,s] * feat[t,:]   (feature gradient: pass dout as feat)
 * semseg_psa_attend_bwd_attn: dattn (same shape as attn, every element written: zero where the mask window gives no
 *   gradient) = P * (scale * dout . feat^T - rowsum(dout * out)) scattered back through the mask index map.
 * psa_type 0 = collect, 1 = distribute (lib/psa/functions/psamask.py:9). */
int semseg_psa_attend(int mode, int psa_type, const float* attn, int a_pitch, const void* feat, const void* feat_lo,
                      int feat_pitch, float* stats, void* out, void* out_lo, int out_pitch, int N, int H, int W, int mH,
                      int mW, int C, float scale, void* stream);
int semseg_psa_attend_bwd_attn(int psa_type, const float* attn, int a_pitch, const float* stats, const void* feat,
                               const void* feat_lo, int feat_pitch, const void* out, const void* out_lo, int out_pitch,
                               const void* dout, const void* dout_lo, int dout_pitch, float* dattn, int N, int H, int W,
                               int mH, int mW, int C, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
 *
 * One descriptor drives fprop and dgrad (dgrad = fprop of dY with the transposed/flipped packed
 * weights). The output pixel grid is [N, H, W]; tap t reads input pixel
 * (n*img_mul + img_add[t], h + dh[t], w + dw[t]) of an input tensor [Nin, Hin, Win, Cin] with zero fill
 * outside it, multiplied by packed-weight slab wtap[t] of a [n_wtaps][Cout_rows][Cin_cols] bf16 tensor.
 */
enum {
  SEMSEG_EPI_RAW = 0,    /* y = bf16(acc); optional per-tile BN partial statistics */
  SEMSEG_EPI_AFFINE = 1, /* y = bf16(act(acc*scale[c] + shift[c] + residual)); scale/shift/residual optional */
  SEMSEG_EPI_F32 = 2     /* out_f32[pixel*out_pitch + c] = acc + shift[c] (bias), c < Cout */
};

typedef struct semseg_conv_desc {
  /* output pixel grid and GEMM sizes */
  int32_t N, H, W;
  int32_t Cin, Cout;
  /* input tensor (bf16 NHWC) */
  const void* x;
  int32_t Nin, Hin, Win, x_pitch;
  /* packed weights (bf16 [n_wtaps][w_rows][w_cols], w_cols contiguous) */
  const void* w;
  int32_t n_wtaps, w_rows, w_cols;
  /* taps */
  int32_t taps;
  int32_t dh[SEMSEG_MAX_TAPS], dw[SEMSEG_MAX_TAPS], wtap[SEMSEG_MAX_TAPS];
  int32_t img_mul, img_add[SEMSEG_MAX_TAPS];
  /* epilogue */
  int32_t epi_mode;
  int32_t relu;
  void* y; /* bf16 NHWC output (RAW / AFFINE) */
  int32_t y_pitch;
  const float* scale;   /* [Cout] or NULL */
  const float* shift;   /* [Cout] or NULL (bias in F32 mode) */
  const void* residual; /* bf16 NHWC [N,H,W,*] or NULL */
  int32_t res_pitch;
  float* out_f32; /* F32 mode output */
  int32_t out_pitch;
  /* RAW mode statistics: stats_partial [rows][3][Cout] = per epilogue-warp (sum, sum of squares, count) of the stored
   * bf16 outputs per channel, rows = semseg_conv_stats_rows(); NULL to skip. The kernel zeroes and fills every row. */
  float* stats_partial;
  /* bf16x3 operand mode (split storage): lo planes of x / y / residual (same shapes and pitches as the hi planes);
   * all NULL for plain bf16. With x_lo set, w must hold the lo slab directly behind the hi slab (w_split != 0,
   * semseg_pack_weights(..., split = 1)) and every K block is accumulated as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo;
   * statistics are those of hi + lo. */
  const void* x_lo;
  void* y_lo;
  const void* residual_lo;
  int32_t w_split;
  /* K slicing (F32 epilogue only; 0 or 1 = off): the K blocks (64-channel block x tap) are cut into k_slices ranges,
   * slice s writes its fp32 partial to out_f32 + s*slice_stride (elements); the bias is added by slice 0 only. Used by
   * the bf16x3 mode to bound the length of one tensor-core accumulation chain (semseg_conv_k_slices,
   * semseg_conv_splitk_finish). */
  int32_t k_slices;
  int64_t slice_stride;
} semseg_conv_desc;

/* Number of K slices such that one slice holds at most max_kblocks K blocks (64-channel block x tap). */
int semseg_conv_k_slices(int Cin, int taps, int max_kblocks);
/* Sum the k_slices fp32 partials [k_slices][M][part_pitch] of a K-sliced conv in fp32 (round-to-nearest) and finish
 * like the conv epilogue would have: RAW (y = sum; optional statistics rows [rows][3][C] = (sum, sum of squares, count)
 * per pixel chunk, rows = semseg_conv_splitk_rows(M)) or AFFINE (y = act(sum*scale + shift + residual)). y (and the
 * residual) may be split (hi, lo) or plain. C % 64 == 0. */
int semseg_conv_splitk_rows(int M);
int semseg_conv_splitk_finish(const float* partial, int k_slices, long long slice_stride, int part_pitch, int M, int C,
                              int epi_mode, int relu, const float* scale, const float* shift, const void* residual,
                              const void* residual_lo, int res_pitch, void* y, void* y_lo, int y_pitch,
                              float* stats_partial, void* stream);
/* Rows of the statistics buffer (= 4 x CTAs launched) for an [N,H,W] x Cout output. */
int semseg_conv_stats_rows(int N, int H, int W, int Cout);
int semseg_conv_fprop(const semseg_conv_desc* d, void* stream);

/* wgrad: dw_partial[split][tap][co][ci] (fp32) = sum over the split's pixels of dy[p, co] * x[p + off(tap), ci].
 * Returns the number of splits through *n_splits (query with dw_partial == NULL first; the buffer must hold
 * n_splits*taps*Cout*Cin floats). */
typedef struct semseg_wgrad_desc {
  int32_t N, H, W;
  int32_t Cin, Cout;
  const void* x; /* bf16 NHWC [Nin,Hin,Win,*] */
  int32_t Nin, Hin, Win, x_pitch;
  const void* dy; /* bf16 NHWC [N,H,W,*] */
  int32_t dy_pitch;
  int32_t taps;
  int32_t dh[SEMSEG_MAX_TAPS], dw[SEMSEG_MAX_TAPS];
  int32_t img_mul, img_add[SEMSEG_MAX_TAPS];
  float* dw_partial;
  int32_t n_splits; /* in: 0 = let the library choose; out (via semseg_conv_wgrad_splits) */
  /* bf16x3 operand mode: lo planes of x and dy (both or neither); dy_hi*x_hi + dy_lo*x_hi + dy_hi*x_lo. */
  const void* x_lo;
  const void* dy_lo;
} semseg_wgrad_desc;

int semseg_conv_wgrad_splits(const semseg_wgrad_desc* d);
int semseg_conv_wgrad(const semseg_wgrad_desc* d, void* stream);
/* dw_oihw[co][ci][tap] (+)= sum_s dw_partial[s][tap][co][ci]; accumulate != 0 adds to the existing values. */
int semseg_wgrad_reduce(const float* dw_partial, int n_splits, int taps, int Cout, int Cin, float* dw_oihw,
                        int accumulate, void* stream);

/* Weight packing: fp32 OIHW [Cout][Cin][taps] ->
 *   wf bf16 [taps][rows_f][cols_f]  (wf[t][co][ci], zero padded)   — fprop B operand
 *   wd bf16 [taps][rows_d][cols_d]  (wd[t][ci][co], zero padded)   — dgrad B operand
 * Either output may be NULL. */
int semseg_pack_weights(const float* w_oihw, int Cout, int Cin, int taps, void* wf, int rows_f, int cols_f,
                        void* wd, int rows_d, int cols_d, int split, void* stream);

/* The same packing for every conv of a model in one launch (torch re-packs after each optimizer step; 2 x 63 small
 * launches per step for PSPNet50 otherwise). `items` is an array in DEVICE memory, sorted by tile0; an item's tiles are
 * its 32 x 32 (co, ci) blocks: tiles_ci = ceil(cols_f / 32) per row of ceil(cols_d / 32) rows. wf is
 * [taps][Cout][cols_f], wd is [taps][Cin][cols_d] (cols_* = Cin / Cout rounded up to 8, zero padded); either may be NULL. */
typedef struct semseg_pack_item {
  const float* w;
  void* wf;
  void* wd;
  int Cout, Cin, taps;
  int cols_f, cols_d;
  int tile0, tiles_ci;
  int split; /* != 0: lo slabs follow the hi slabs (wf + taps*Cout*cols_f, wd + taps*Cin*cols_d) */
} semseg_pack_item;
int semseg_pack_weights_multi(const semseg_pack_item* items_dev, int n_items, int n_tiles, int max_taps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout conversion at the module boundary.
 */
int semseg_nchw_f32_to_nhwc_bf16(const float* in, void* out, void* out_lo, int N, int C, int H, int W, int out_pitch,
                                 void* stream);
int semseg_nhwc_bf16_to_nchw_f32(const void* in, const void* in_lo, float* out, int N, int C, int H, int W,
                                 int in_pitch, void* stream);
int semseg_nhwc_f32_to_nchw_f32(const float* in, float* out, int N, int C, int H, int W, int in_pitch,
                                void* stream);

/* Stem conv (3x3, stride 2, pad 1, Cin <= 3 — model/resnet.py:106-108) as a 1x1 conv over input patches:
 *   out[n, ho, wo, (r*3+s)*Cin + c] = x[n, 2ho-1+r, 2wo-1+s, c], zero outside the image and for columns >= 9*Cin;
 *   out is [N, (H-1)/2+1, (W-1)/2+1, 32] bf16. x is NHWC bf16 with x_pitch >= 4 (the first Cin channels are read). */
int semseg_im2col3x3s2(const void* x, int x_pitch, int N, int H, int W, int Cin, void* out, void* stream);

/* 2x2 phase decomposition used to run stride-2 convolutions (model/resnet.py:108 conv1, layer2.0 conv2 and
 * downsample) on the stride-1 tensor-core kernel:
 *   xp [4][N][Hh][Wh][C], Hh = (H+1)/2:  xp[ph*2+pw][n][i][j] = x[n][2i+ph][2j+pw] (zero outside x). */
int semseg_space_to_phases(const void* x, int x_pitch, int N, int H, int W, int C, void* xp, void* stream);
int semseg_phases_to_space(const void* xp, int N, int H, int W, int C, void* x, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm (training statistics, apply, backward) on NHWC bf16 tensors, fp32 statistics.
 */
/* Merge the conv epilogue's per-CTA partials [rows][3][C] (Chan) into per-channel (mean, M2, count): out_stats [3][C]. */
int semseg_bn_merge_partials(const float* stats_partial, int rows, int C, float* out_stats, void* stream);
/* Per-channel statistics of an arbitrary NHWC bf16 tensor [M pixels][C] (used where the producer is not the
 * conv kernel): out_stats [3][C] = (mean, M2, count). */
int semseg_bn_stats(const void* x, int M, int C, int pitch, float* workspace, long long workspace_floats,
                    float* out_stats, void* stream);
/* Scratch floats the two-stage reductions (bn_stats, bn_bwd_reduce) need for an [M][C] tensor. */
long long semseg_bn_workspace_floats(int M, int C);
/* Merge R rank-stat blocks [R][3][C] (R = 1 without SyncBN) and finalise:
 *   mean_invstd [3][C] = (mean, 1/sqrt(var+eps), total samples per channel over all ranks — every finalize entry point
 *   writes the three rows); scale_shift [2][C] with scale = gamma*invstd, shift = beta - mean*scale;
 *   running_mean/var updated in place (momentum, unbiased var) when non-NULL. */
int semseg_bn_finalize(const float* rank_stats, int R, int C, const float* gamma, const float* beta, float eps,
                       float momentum, float* running_mean, float* running_var, float* mean_invstd,
                       float* scale_shift, void* stream);
/* Single-rank fast path: semseg_bn_merge_partials + semseg_bn_finalize (R = 1) in one launch. */
int semseg_bn_finalize_partials(const float* stats_partial, int rows, int C, const float* gamma, const float* beta, float eps, float momentum,
                                float* running_mean, float* running_var, float* mean_invstd, float* scale_shift,
                                void* stream);
/* SyncBatchNorm exchange over NVLink peer memory instead of NCCL (one kernel per exchange). peer_bufs[world] /
 * peer_flags[world] are device pointers into every rank's symmetric (peer-mapped) allocation: a zero-initialised buffer
 * of n_slots*world*slot_floats 8-byte words (every value travels as one {fp32, sequence number} word that the sender
 * stores into sub-block `rank` of the slot in every peer's buffer; the receiver polls its own memory: flag-in-data, no
 * fences) and a uint32 flag array [n_slots][world] (unused by this protocol, kept for ABI stability); `counter` is a zeroed local uint32;
 * `slot` must be unique per exchange within a step and the sequence number strictly increasing per step (same on every
 * rank): it is `seq`, or — when seq_ptr is non-NULL — the uint32 read from that device address when the kernel runs (a
 * device-resident step counter, so that a captured CUDA graph with baked-in slots can be replayed).
 *   finalize_p2p   : per-CTA conv partials -> exchange (mean, M2, n) -> cross-rank merge in rank order -> finalise.
 *   bwd_reduce_p2p : local [sum dz, sum dz*xhat] -> sums_local; exchanged and added in rank order -> sums_total. */
int semseg_bn_finalize_p2p(const float* stats_partial, int rows, int C, const float* gamma, const float* beta,
                           float eps, float momentum, float* running_mean, float* running_var, float* mean_invstd,
                           float* scale_shift, void* const* peer_bufs, void* const* peer_flags, void* counter,
                           int world, int rank, int slot, int slot_floats, unsigned seq, const void* seq_ptr,
                           void* stream);
int semseg_bn_bwd_reduce_p2p(const void* dy, const void* dy_lo, int dy_pitch, const void* y, const void* y_lo,
                             int y_pitch, const void* x, const void* x_lo, int x_pitch, const float* mean_invstd,
                             const float* scale_shift, int M, int C, int relu, float* workspace,
                             long long workspace_floats, float* sums_local, float* sums_total,
                             void* const* peer_bufs, void* const* peer_flags, void* counter, int world, int rank,
                             int slot, int slot_floats, unsigned seq, const void* seq_ptr, void* stream);
/* Eval-mode folding: scale = gamma/sqrt(var+eps), shift = beta - mean*scale. */
int semseg_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, int C, float* scale_shift, void* stream);
/* y = act(x*scale[c] + shift[c] + residual). */
int semseg_bn_apply(const void* x, const void* x_lo, int x_pitch, const float* scale_shift, const void* residual,
                    const void* residual_lo, int res_pitch, void* y, void* y_lo, int y_pitch, int M, int C, int relu,
                    void* stream);
/* Backward reduce: with dz = dy * (y > 0 if relu) and xhat = (x - mean)*invstd,
 *   sums [2][C] = (sum dz, sum dz*xhat). y may be NULL when relu == 0; when relu != 0 and y == NULL the mask is
 *   recomputed as fma(x, scale, shift) > 0 from scale_shift [2][C] (valid when the forward had no residual). */
int semseg_bn_bwd_reduce(const void* dy, const void* dy_lo, int dy_pitch, const void* y, const void* y_lo, int y_pitch,
                         const void* x, const void* x_lo, int x_pitch, const float* mean_invstd,
                         const float* scale_shift, int M, int C, int relu, float* workspace,
                         long long workspace_floats, float* sums, void* stream);
/* Backward apply: dx = gamma*invstd*(dz - sum_dz/count - xhat*sum_dzxhat/count);
 *   dres (optional) = dz; dgamma = sum_dzxhat, dbeta = sum_dz written to dgamma_dbeta [2][C].
 *   count = total number of samples per channel across all ranks; count <= 0 takes it from mean_invstd row 2 (what
 *   the forward exchange measured: correct also when the ranks hold different numbers of pixels, as torch SyncBN). */
int semseg_bn_bwd_apply(const void* dy, const void* dy_lo, int dy_pitch, const void* y, const void* y_lo, int y_pitch,
                        const void* x, const void* x_lo, int x_pitch, const float* mean_invstd, const float* gamma,
                        const float* scale_shift, const float* sums, float count, int M, int C, int relu, void* dx,
                        void* dx_lo, int dx_pitch, void* dres, void* dres_lo, int dres_pitch, float* dgamma_dbeta,
                        void* stream);
/* out = a + b (merges gradient branches; split-aware, unlike an elementwise add of the two planes). */
int semseg_add_act(const void* a, const void* a_lo, int a_pitch, const void* b, const void* b_lo, int b_pitch,
                   void* out, void* out_lo, int out_pitch, int M, int C, void* stream);
/* out[n, p, c] = x[n, p, c] * scale[n*C + c] for p < HW: nn.Dropout2d's per-(image, channel) factor
 * (model/pspnet.py:68,76) and its backward. */
int semseg_scale_nc(const void* x, const void* x_lo, int x_pitch, const float* scale, void* out, void* out_lo,
                    int out_pitch, int N, int HW, int C, void* stream);
/* fp32 rows [M][in_pitch] (C columns used) -> activation rows [M][out_pitch], columns C..Cp-1 zero (Cp % 8 == 0). */
int semseg_f32_to_act(const float* in, int in_pitch, void* out, void* out_lo, int out_pitch, long long M, int C,
                      int Cp, void* stream);
/* activation rows [M][in_pitch] (C % 8 == 0 columns) -> fp32 rows [M][out_pitch]. */
int semseg_act_to_f32(const void* in, const void* in_lo, int in_pitch, float* out, int out_pitch, long long M, int C,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NHWC bf16 (model/resnet.py:115): y [N,Ho,Wo,C] with
 * Ho = (H-1)/2+1; argcode uint8 [N,Ho,Wo,C] (or NULL) = window position 0..8 of the arg-max (first maximum in window
 * order, as ATen). Backward gathers with those codes: dx [N,H,W,C] dense, deterministic, no atomics.
 */
int semseg_maxpool3x3s2_fwd(const void* x, const void* x_lo, void* y, void* y_lo, void* argcode, int N, int H, int W,
                            int C, void* stream);
int semseg_maxpool3x3s2_bwd(const void* argcode, const void* dy, const void* dy_lo, void* dx, void* dx_lo, int N,
                            int H, int W, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pyramid pooling module data movement (model/pspnet.py:12-26), NHWC bf16, all bins in one launch.
 *   bins[nb] = pooled sizes (1,2,3,6); per-bin tensors are [N][b][b][channels] contiguous bf16.
 *   ppm_pool            : pooled_k = AdaptiveAvgPool2d(b_k)(x), window [floor(i*H/b), ceil((i+1)*H/b)).
 *   ppm_pool_bwd        : dx (dense, every element written) = sum_k adjoint of the pooling applied to dpooled_k.
 *   ppm_pool_bwd        : dx = adjoint of ppm_pool (+ `add` [N,H,W,add_pitch], nullable: the identity branch of the
 *                         concat, so the two gradients of x are summed in this kernel instead of by autograd).
 *   ppm_upsample_concat : out[..., 0:C] = x; out[..., C + k*Cr : C + (k+1)*Cr] = bilinear(align_corners=True) of
 *                         feats_k to H x W (the torch.cat of model/pspnet.py:26 written in place).
 *   ppm_upsample_bwd    : dfeats_k = adjoint of the bilinear upsample applied to dout[..., c_off + k*Cr : ...].
 */
int semseg_ppm_pool(const void* x, const void* x_lo, int x_pitch, int N, int H, int W, int C, const int* bins,
                    void* const* pooled, void* const* pooled_lo, int nb, void* stream);
int semseg_ppm_pool_bwd(void* const* dpooled, void* const* dpooled_lo, const int* bins, int nb, int N, int H, int W,
                        int C, void* dx, void* dx_lo, int dx_pitch, const void* add, const void* add_lo, int add_pitch,
                        void* stream);
int semseg_ppm_upsample_concat(const void* x, const void* x_lo, int x_pitch, void* const* feats, void* const* feats_lo,
                               const int* bins, int nb, int N, int H, int W, int C, int Cr, void* out, void* out_lo,
                               int out_pitch, void* stream);
int semseg_ppm_upsample_bwd(const void* dout, const void* dout_lo, int dout_pitch, int c_off, void* const* dfeats,
                            void* const* dfeats_lo, const int* bins, int nb, int N, int H, int W, int Cr, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bilinear resize, align_corners=True, of an NHWC activation [N,Hi,Wi,C] -> [N,Ho,Wo,C] (F.interpolate at
 * model/psanet.py:61,97) and its adjoint (dy [N,Ho,Wo,C] -> dx [N,Hi,Wi,C]; a deterministic gather, no atomics).
 */
int semseg_resize_bilinear_fwd(const void* x, const void* x_lo, int x_pitch, int N, int Hi, int Wi, int C, int Ho,
                               int Wo, void* y, void* y_lo, int y_pitch, void* stream);
int semseg_resize_bilinear_bwd(const void* dy, const void* dy_lo, int dy_pitch, int N, int Hi, int Wi, int C, int Ho,
                               int Wo, void* dx, void* dx_lo, int dx_pitch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused logit upsample (bilinear, align_corners=True, x8) + cross-entropy (ignore_index, mean over valid
 * pixels) + argmax: F.interpolate + CrossEntropyLoss + max(1) of model/pspnet.py:94-103 without the
 * [N, classes, Ho, Wo] tensor. Requires Ho = 8(h-1)+1, Wo = 8(w-1)+1 (zoom_factor 8), classes <= 256.
 *   logits fp32 NHWC [N,h,w,C] (pitch), target int64 [N,Ho,Wo].
 *   fwd: loss_out[0] = mean CE, loss_out[1] = number of non-ignored pixels; argmax int64 [N,Ho,Wo] (or NULL);
 *        lse fp32 [N,Ho,Wo] (saved for backward); workspace: semseg_upsample_ce_workspace_floats() floats.
 *   bwd: dlogits fp32 [N,h,w,C] (dense, every element written) = grad_out[0] * d(mean CE)/dlogits;
 *        workspace: semseg_upsample_ce_bwd_workspace_floats() floats (row-reduced intermediate [N][Ho][w][C]).
 */
long long semseg_upsample_ce_workspace_floats(int N, int Ho, int Wo);
int semseg_upsample_ce_fwd(const float* logits, int pitch, int N, int h, int w, int C, const int64_t* target,
                           int Ho, int Wo, int ignore_index, float* workspace, float* loss_out, int64_t* argmax,
                           float* lse, void* stream);
long long semseg_upsample_ce_bwd_workspace_floats(int N, int Ho, int w, int C);
int semseg_upsample_ce_bwd(const float* logits, int pitch, int N, int h, int w, int C, const int64_t* target,
                           int Ho, int Wo, int ignore_index, const float* lse, const float* loss_info,
                           const float* grad_out, float* workspace, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Step glue: per-class intersection / union / target areas (util/util.py:55-67, called at tool/train.py:286,375).
 *   counts int32 [3][K] (zeroed by the call): [0] = #(pred == target == k), [1] = #(pred == k) with pred forced to
 *   ignore_index where target == ignore_index, [2] = #(target == k); union = [1] + [2] - [0].
 *   write_back != 0 also stores the masked prediction (the reference masks `output` in place).
 */
int semseg_iou_hist(void* pred_i64, const void* target_i64, long long n, int K, long long ignore_index, int write_back,
                    int* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * torch.optim.SGD (momentum, dampening, weight decay, nesterov; tool/train.py:140,274-276) over every parameter tensor in
 * one launch. items_dev: device array sorted by chunk0 (an item's chunks are consecutive blocks of
 * semseg_sgd_chunk_elems() elements); grad_ptrs_dev: device array of n_items gradient pointers (0 = no gradient this
 * step: the parameter is skipped); hyper: per-group hyper-parameters, passed by value. `first` != 0 initialises the
 * momentum buffer with the (decayed) gradient, as torch does on a parameter's first step.
 */
#define SEMSEG_SGD_MAX_GROUPS 16
typedef struct semseg_sgd_item {
  float* w;
  float* buf;
  long long n;
  int group;
  int chunk0;
  int first;
  int reserved;
} semseg_sgd_item;
typedef struct semseg_sgd_hyper {
  float lr[SEMSEG_SGD_MAX_GROUPS];
  float momentum[SEMSEG_SGD_MAX_GROUPS];
  float weight_decay[SEMSEG_SGD_MAX_GROUPS];
  float dampening[SEMSEG_SGD_MAX_GROUPS];
  int nesterov;
} semseg_sgd_hyper;
int semseg_sgd_chunk_elems(void);
int semseg_sgd_multi(const semseg_sgd_item* items_dev, const void* grad_ptrs_dev, int n_items, int n_chunks,
                     const semseg_sgd_hyper* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEMSEG_B200_H */
