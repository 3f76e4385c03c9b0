// Fused logit upsample (bilinear, align_corners=True, x8) + cross-entropy(ignore_index) + argmax.
//
// Replaces F.interpolate -> CrossEntropyLoss -> max(1) at model/pspnet.py:94-103 (same in model/psanet.py:168-177),
// which materialise an [N, classes, H, W] fp32 tensor (2.15 GB at bs16 / 150 classes / 473x473) and stream it
// ~9 times per head. Here the low-resolution logits (fp32 NHWC, a few MB, L2 resident) are staged in shared
// memory and every output pixel's class vector is interpolated on the fly; the only full-resolution tensors
// are the int64 argmax and an fp32 log-sum-exp map kept for the backward pass.
//
// The kernels require Ho = 8*(h-1)+1 and Wo = 8*(w-1)+1 (zoom_factor 8, every shipped config): then the
// align_corners scale (h-1)/(Ho-1) is exactly 1/8, source index = x >> 3 and the weights are (x & 7)/8 — the same
// fp32 values ATen computes. Interpolation order follows ATen's upsample_bilinear2d:
//   v = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11).
//
// Backward is a deterministic, separable gather (rows kernel + cols kernel, see below). No atomics, every dlogits
// element is written exactly once.
#include "host_common.h"

namespace sb {

constexpr int kMaxClasses = 256;

// ---------------------------------------------------------------------------------------------------- forward
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kFwdCols = 128;                    // output columns per CTA (one per thread)
constexpr int kFwdNodes = kFwdCols / 8 + 1;      // low-res node columns a CTA touches

// One CTA per (128 output columns, low-res interval row i0, image); a thread owns one output column and the 8
// output rows of the interval. Per class the horizontal interpolation of the two node rows (top, bot) is done once
// and shared by the 8 rows (v = l0h*top + l1h*bot with compile-time row weights), so a pixel-class costs ~5
// instructions per pass instead of a full 4-tap interpolation. Two passes over the classes: max/argmax, then
// sum of exp2 — one MUFU per pixel-class, no rescaling branches.
__global__ void __launch_bounds__(kFwdCols)
upsample_ce_fwd_kernel(const float* __restrict__ logits, int pitch, int N, int h, int w, int C, int Cs,
                       const long long* __restrict__ target, int Ho, int Wo, int ignore_index,
                       float* __restrict__ partial, long long* __restrict__ argmax_out, float* __restrict__ lse_out) {
  extern __shared__ float S[];  // [2][kFwdNodes][Cs]; Cs odd -> the 4-5 node columns a warp reads hit distinct banks
  __shared__ float red_loss[kFwdCols / 32];
  __shared__ float red_cnt[kFwdCols / 32];
  const int n = blockIdx.z, i0 = blockIdx.y, x0 = blockIdx.x * kFwdCols;
  const int i1 = min(i0 + 1, h - 1);
  const int j_base = x0 >> 3;
  const int nj = min(kFwdNodes, w - j_base);
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 2 * nj * C; idx += kFwdCols) {
    const int c = idx % C;
    const int node = idx / C;
    const int jj = node % nj, rr = node / nj;
    S[(rr * kFwdNodes + jj) * Cs + c] =
        logits[((static_cast<size_t>(n) * h + (rr ? i1 : i0)) * w + (j_base + jj)) * pitch + c];
  }
  __syncthreads();
  float loss = 0.f, cnt = 0.f;
  const int x = x0 + tid;
  const int rows = min(8, Ho - 8 * i0);  // 8, or 1 for the last node row (Ho = 8(h-1)+1)
  if (x < Wo) {
    const int j0 = x >> 3;
    const int j1 = min(j0 + 1, w - 1);
    const float l1w = static_cast<float>(x & 7) * 0.125f, l0w = 1.f - l1w;
    const float* A = S + (j0 - j_base) * Cs;   // node (i0, j0)
    const float* B = S + (j1 - j_base) * Cs;   // node (i0, j1)
    const float* Cc = A + kFwdNodes * Cs;      // node (i1, j0)
    const float* D = B + kFwdNodes * Cs;       // node (i1, j1)
    float m[8], sum[8];
    int am[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      m[r] = -INFINITY;
      am[r] = 0;
      sum[r] = 0.f;
    }
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
      const float top = l0w * A[c] + l1w * B[c];
      const float bot = l0w * Cc[c] + l1w * D[c];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float v = (1.f - 0.125f * r) * top + (0.125f * r) * bot;
        if (v > m[r]) {
          m[r] = v;
          am[r] = c;
        }
      }
    }
    float m2[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) m2[r] = m[r] * kLog2e;
#pragma unroll 2
    for (int c = 0; c < C; ++c) {
      const float top = l0w * A[c] + l1w * B[c];
      const float bot = l0w * Cc[c] + l1w * D[c];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float v = (1.f - 0.125f * r) * top + (0.125f * r) * bot;
        sum[r] += ex2_approx(fmaf(v, kLog2e, -m2[r]));
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < rows) {
        const size_t pix = (static_cast<size_t>(n) * Ho + (8 * i0 + r)) * Wo + x;
        const long long t = target[pix];
        const float lse = m[r] + __logf(sum[r]);
        if (argmax_out) argmax_out[pix] = am[r];
        lse_out[pix] = lse;
        if (t != ignore_index && t >= 0 && t < C) {
          const int tc = static_cast<int>(t);
          const float top = l0w * A[tc] + l1w * B[tc];
          const float bot = l0w * Cc[tc] + l1w * D[tc];
          const float vt = (1.f - 0.125f * r) * top + (0.125f * r) * bot;
          loss += lse - vt;
          cnt += 1.f;
        }
      }
    }
  }
  // deterministic block reduction -> one partial per CTA
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((tid & 31) == 0) {
    red_loss[tid >> 5] = loss;
    red_cnt[tid >> 5] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f, k = 0.f;
    for (int i = 0; i < kFwdCols / 32; ++i) {
      l += red_loss[i];
      k += red_cnt[i];
    }
    const size_t b = (static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partial[2 * b] = l;
    partial[2 * b + 1] = k;
  }
}

// loss_out[0] = sum / max(count, 1) (mean over non-ignored pixels), loss_out[1] = count. Fixed summation order.
__global__ void upsample_ce_reduce_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ loss_out) {
  __shared__ double sl[256];
  __shared__ double sc[256];
  double l = 0.0, k = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    l += partial[2 * i];
    k += partial[2 * i + 1];
  }
  sl[threadIdx.x] = l;
  sc[threadIdx.x] = k;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sl[threadIdx.x] += sl[threadIdx.x + o];
      sc[threadIdx.x] += sc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss_out[0] = static_cast<float>(sl[0] / (sc[0] > 0.0 ? sc[0] : 1.0));
    loss_out[1] = static_cast<float>(sc[0]);
  }
}

// ---------------------------------------------------------------------------------------------------- backward
// Separable, deterministic, no atomics. With g[y,x,c] = softmax_{y,x}[c] - [c == t_{y,x}] (0 for ignored pixels):
//   dL[i,j,c] = gs * sum_y wy(y,i) * sum_x wx(x,j) * g[y,x,c].
// Phase 1 (rows): one CTA per (image, low-res interval row i0), one thread per class. The thread walks x = 0..Wo-1
// with the four node values of its class in registers; per column the horizontal interpolation (top, bot) is shared
// by the interval's 8 output rows and the rows are folded immediately with their compile-time vertical weights, so a
// pixel-class costs ~10 instructions and one MUFU. (lse*log2e, target) of the 8 rows are staged in shared memory as
// one 8-byte word per pixel and read as warp-uniform broadcasts. Output: T2[n][i0][s][j][c], s = 0: the interval's
// contribution to node row i0, s = 1: to node row i0+1 (fp32 workspace, 68 MB at bs16 / 150 classes).
// Phase 2 (cols): dL[i] = gs * (T2[i][0] + T2[i-1][1]).
struct __align__(8) PixInfo {
  float lse2;  // log-sum-exp * log2(e)
  int t;       // target class, -1 = ignored
};

__global__ void __launch_bounds__(256)
upsample_ce_bwd_rows_kernel(const float* __restrict__ logits, int pitch, int N, int h, int w, int C,
                            const long long* __restrict__ target, int Ho, int Wo, int ignore_index,
                            const float* __restrict__ lse, float* __restrict__ T2) {
  extern __shared__ PixInfo s_pix[];  // [8][Wo]
  const int i0 = blockIdx.x, n = blockIdx.y;
  const int i1 = min(i0 + 1, h - 1);
  const int rows = min(8, Ho - 8 * i0);
  for (int r = 0; r < rows; ++r) {
    const size_t rowbase = (static_cast<size_t>(n) * Ho + (8 * i0 + r)) * Wo;
    for (int x = threadIdx.x; x < Wo; x += blockDim.x) {
      const long long t = target[rowbase + x];
      PixInfo pi;
      pi.t = (t == ignore_index || t < 0 || t >= C) ? -1 : static_cast<int>(t);
      pi.lse2 = lse[rowbase + x] * kLog2e;
      s_pix[r * Wo + x] = pi;
    }
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= C) return;
  const float* L0 = logits + (static_cast<size_t>(n) * h + i0) * w * pitch + c;
  const float* L1 = logits + (static_cast<size_t>(n) * h + i1) * w * pitch + c;
  float* T0 = T2 + ((static_cast<size_t>(n) * h + i0) * 2 + 0) * w * C + c;
  float* T1 = T2 + ((static_cast<size_t>(n) * h + i0) * 2 + 1) * w * C + c;
  float a = L0[0], cc = L1[0];      // left node column of the current interval (rows i0 / i1)
  float nb = L0[static_cast<size_t>(min(1, w - 1)) * pitch], nd = L1[static_cast<size_t>(min(1, w - 1)) * pitch];
  float carry0 = 0.f, carry1 = 0.f;  // right-node contributions of the previous interval (node rows i0 / i1)
  for (int j0 = 0; j0 < w; ++j0) {
    const float b = nb, d = nd;     // right node column (j1 = min(j0+1, w-1))
    const int jn = min(j0 + 2, w - 1);
    nb = L0[static_cast<size_t>(jn) * pitch];          // prefetch the next interval's right column
    nd = L1[static_cast<size_t>(jn) * pitch];
    float accL0 = 0.f, accR0 = 0.f, accL1 = 0.f, accR1 = 0.f;
    const int xb = j0 * 8;
    const int nx = min(8, Wo - xb);   // 8, or 1 for the last node column
    if (rows == 8 && nx == 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float l1w = 0.125f * k, l0w = 1.f - l1w;
        const float top = l0w * a + l1w * b;
        const float bot = l0w * cc + l1w * d;
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const PixInfo pi = s_pix[r * Wo + xb + k];
          if (pi.t < 0) continue;  // warp-uniform
          const float v = (1.f - 0.125f * r) * top + (0.125f * r) * bot;
          const float g = ex2_approx(fmaf(v, kLog2e, -pi.lse2)) - (c == pi.t ? 1.f : 0.f);
          g0 = fmaf(1.f - 0.125f * r, g, g0);
          g1 = fmaf(0.125f * r, g, g1);
        }
        accL0 = fmaf(l0w, g0, accL0);
        accR0 = fmaf(l1w, g0, accR0);
        accL1 = fmaf(l0w, g1, accL1);
        accR1 = fmaf(l1w, g1, accR1);
      }
    } else {
      for (int k = 0; k < nx; ++k) {
        const float l1w = 0.125f * k, l0w = 1.f - l1w;
        const float top = l0w * a + l1w * b;
        const float bot = l0w * cc + l1w * d;
        float g0 = 0.f, g1 = 0.f;
        for (int r = 0; r < rows; ++r) {
          const PixInfo pi = s_pix[r * Wo + xb + k];
          if (pi.t < 0) continue;
          const float l1h = 0.125f * r, l0h = 1.f - l1h;
          const float v = l0h * top + l1h * bot;
          const float g = ex2_approx(fmaf(v, kLog2e, -pi.lse2)) - (c == pi.t ? 1.f : 0.f);
          g0 = fmaf(l0h, g, g0);
          g1 = fmaf(l1h, g, g1);
        }
        accL0 = fmaf(l0w, g0, accL0);
        accR0 = fmaf(l1w, g0, accR0);
        accL1 = fmaf(l0w, g1, accL1);
        accR1 = fmaf(l1w, g1, accR1);
      }
    }
    T0[static_cast<size_t>(j0) * C] = carry0 + accL0;
    T1[static_cast<size_t>(j0) * C] = carry1 + accL1;
    carry0 = accR0;
    carry1 = accR1;
    a = b;
    cc = d;
  }
}

__global__ void __launch_bounds__(256)
upsample_ce_bwd_cols_kernel(const float* __restrict__ T2, int N, int h, int w, int C,
                            const float* __restrict__ loss_info, const float* __restrict__ grad_out,
                            float* __restrict__ dlogits) {
  const int i = blockIdx.x, n = blockIdx.y;
  const float cntv = loss_info[1];
  const float gs = grad_out[0] / (cntv > 0.f ? cntv : 1.f);
  const int wc = w * C;
  const float* own = T2 + ((static_cast<size_t>(n) * h + i) * 2 + 0) * wc;                 // interval i, top slot
  const float* prev = i > 0 ? T2 + ((static_cast<size_t>(n) * h + (i - 1)) * 2 + 1) * wc : nullptr;  // interval i-1, bottom
  for (int idx = threadIdx.x; idx < wc; idx += blockDim.x) {
    float acc = own[idx];
    if (prev) acc += prev[idx];
    dlogits[(static_cast<size_t>(n) * h + i) * wc + idx] = acc * gs;
  }
}

}  // namespace sb

using namespace sb;

static int check_tail(const void* logits, int pitch, int N, int h, int w, int C, const void* target, int Ho, int Wo) {
  SB_CHECK_ARG(logits && target, "upsample_ce: null pointer");
  SB_CHECK_ARG(N > 0 && h > 1 && w > 1 && C > 1 && C <= kMaxClasses && pitch >= C, "upsample_ce: bad sizes (C<=%d)",
               kMaxClasses);
  SB_CHECK_ARG(Ho == 8 * (h - 1) + 1 && Wo == 8 * (w - 1) + 1,
               "upsample_ce: fused kernel needs Ho=8(h-1)+1, Wo=8(w-1)+1 (got %dx%d -> %dx%d)", h, w, Ho, Wo);
  return SEMSEG_OK;
}

extern "C" long long semseg_upsample_ce_workspace_floats(int N, int Ho, int Wo) {
  return 2LL * N * ((Ho - 1) / 8 + 1) * cdiv(Wo, kFwdCols);  // (loss, count) per forward CTA
}

extern "C" int semseg_upsample_ce_fwd(const float* logits, int pitch, int N, int h, int w, int C,
                                      const int64_t* target, int Ho, int Wo, int ignore_index, float* workspace,
                                      float* loss_out, int64_t* argmax, float* lse, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int r = check_tail(logits, pitch, N, h, w, C, target, Ho, Wo);
  if (r) return r;
  SB_CHECK_ARG(workspace && loss_out && lse, "upsample_ce_fwd: null output");
  dim3 grid(cdiv(Wo, kFwdCols), h, N);
  const int Cs = C | 1;
  const size_t smem = static_cast<size_t>(2) * kFwdNodes * Cs * sizeof(float);
  upsample_ce_fwd_kernel<<<grid, kFwdCols, smem, stream>>>(
      logits, pitch, N, h, w, C, Cs, reinterpret_cast<const long long*>(target), Ho, Wo, ignore_index, workspace,
      reinterpret_cast<long long*>(argmax), lse);
  SB_LAUNCHED();
  upsample_ce_reduce_kernel<<<1, 256, 0, stream>>>(workspace, static_cast<int>(grid.x * grid.y * grid.z), loss_out);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" long long semseg_upsample_ce_bwd_workspace_floats(int N, int Ho, int w, int C) {
  return 2LL * N * ((Ho - 1) / 8 + 1) * w * C;  // T2[N][h][2][w][C]
}

extern "C" int semseg_upsample_ce_bwd(const float* logits, int pitch, int N, int h, int w, int C,
                                      const int64_t* target, int Ho, int Wo, int ignore_index, const float* lse,
                                      const float* loss_info, const float* grad_out, float* workspace,
                                      float* dlogits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int r = check_tail(logits, pitch, N, h, w, C, target, Ho, Wo);
  if (r) return r;
  SB_CHECK_ARG(lse && loss_info && grad_out && dlogits && workspace, "upsample_ce_bwd: null pointer");
  const int threads = (C + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(8) * Wo * sizeof(PixInfo);
  SB_CHECK_ARG(smem <= 160 * 1024, "upsample_ce_bwd: output width %d too large for the staged rows", Wo);
  static size_t smem_attr = 48 * 1024;
  if (smem > smem_attr) {
    SB_CUDA(cudaFuncSetAttribute(upsample_ce_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    smem_attr = 160 * 1024;
  }
  upsample_ce_bwd_rows_kernel<<<dim3(h, N), threads, smem, stream>>>(
      logits, pitch, N, h, w, C, reinterpret_cast<const long long*>(target), Ho, Wo, ignore_index, lse, workspace);
  SB_LAUNCHED();
  upsample_ce_bwd_cols_kernel<<<dim3(h, N), 256, 0, stream>>>(workspace, N, h, w, C, loss_info, grad_out, dlogits);
  SB_LAUNCHED();
  return SEMSEG_OK;
}
