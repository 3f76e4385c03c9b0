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
// block (32, 8): 32 x 32 output pixels; nodes staged: up to 6 x 6.
__global__ void __launch_bounds__(256)
upsample_ce_fwd_kernel(const float* __restrict__ logits, int pitch, int N, int h, int w, int C,
                       const long long* __restrict__ target, int Ho, int Wo, int ignore_index,
                       float* __restrict__ partial, long long* __restrict__ argmax_out, float* __restrict__ lse_out) {
  extern __shared__ float S[];  // [6*6][C]
  __shared__ float red_loss[8];
  __shared__ float red_cnt[8];
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
  const int i_base = y0 >> 3, j_base = x0 >> 3;
  const int ni = min(6, h - i_base), nj = min(6, w - j_base);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int idx = tid; idx < ni * nj * C; idx += 256) {
    const int c = idx % C;
    const int node = idx / C;
    const int jj = node % nj, ii = node / nj;
    S[(ii * 6 + jj) * C + c] =
        logits[((static_cast<size_t>(n) * h + (i_base + ii)) * w + (j_base + jj)) * pitch + c];
  }
  __syncthreads();
  float loss = 0.f, cnt = 0.f;
  const int x = x0 + threadIdx.x;
  if (x < Wo) {
    const int j0 = x >> 3;
    const int j1 = min(j0 + 1, w - 1);
    const float l1w = static_cast<float>(x & 7) * 0.125f, l0w = 1.f - l1w;
    const int cj0 = (j0 - j_base) * C, cj1 = (j1 - j_base) * C;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
      const int y = y0 + threadIdx.y + r * 8;
      if (y >= Ho) break;
      const int i0 = y >> 3;
      const int i1 = min(i0 + 1, h - 1);
      const float l1h = static_cast<float>(y & 7) * 0.125f, l0h = 1.f - l1h;
      const float* r0 = S + (i0 - i_base) * 6 * C;
      const float* r1 = S + (i1 - i_base) * 6 * C;
      const size_t pix = (static_cast<size_t>(n) * Ho + y) * Wo + x;
      const long long t = target[pix];
      // online log-sum-exp: one interpolation per class; the running sum is rescaled when the maximum moves
      float m = -INFINITY, vt = 0.f, s = 0.f;
      int am = 0;
      for (int c = 0; c < C; ++c) {
        const float v = l0h * (l0w * r0[cj0 + c] + l1w * r0[cj1 + c]) + l1h * (l0w * r1[cj0 + c] + l1w * r1[cj1 + c]);
        if (v > m) {
          s = s * __expf(m - v) + 1.f;
          m = v;
          am = c;
        } else {
          s += __expf(v - m);
        }
        if (c == t) vt = v;
      }
      const float lse = m + __logf(s);
      if (argmax_out) argmax_out[pix] = am;
      lse_out[pix] = lse;
      if (t != ignore_index && t >= 0 && t < C) {
        loss += lse - vt;
        cnt += 1.f;
      }
    }
  }
  // deterministic block reduction -> one partial per CTA
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (threadIdx.x == 0) {
    red_loss[threadIdx.y] = loss;
    red_cnt[threadIdx.y] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    float l = 0.f, k = 0.f;
    for (int i = 0; i < 8; ++i) {
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
// Separable, deterministic, no atomics:  dL[i,j,c] = gs * sum_y wy(y,i) * T[y,j,c],   T[y,j,c] = sum_x wx(x,j) g[y,x,c],
// g[y,x,c] = softmax_{y,x}[c] - [c == t_{y,x}] (0 for ignored pixels).
//
// Phase 1 (rows): one CTA per output row (n, y), one thread per class. The thread walks x = 0..Wo-1; inside a
// low-res interval the four node values of its class stay in registers, so a pixel costs ~12 FMA-class instructions
// and one exp, no shared-memory traffic and no cross-lane reduction; lse / target of the row are staged in smem and
// read as warp-uniform broadcasts. T is an fp32 workspace [N][Ho][w][C] (272 MB at bs16/150 classes).
// Phase 2 (cols): one CTA per node row (n, i), threads over (j, c); fixed-order sum over the <= 15 rows in support.
__global__ void __launch_bounds__(256)
upsample_ce_bwd_rows_kernel(const float* __restrict__ logits, int pitch, int N, int h, int w, int C,
                            const long long* __restrict__ target, int Ho, int Wo, int ignore_index,
                            const float* __restrict__ lse, float* __restrict__ T) {
  extern __shared__ float sm[];
  float* s_lse = sm;                                   // [Wo]
  int* s_t = reinterpret_cast<int*>(sm + Wo);          // [Wo], -1 = ignored
  const int y = blockIdx.x, n = blockIdx.y;
  const size_t rowbase = (static_cast<size_t>(n) * Ho + y) * Wo;
  for (int x = threadIdx.x; x < Wo; x += blockDim.x) {
    const long long t = target[rowbase + x];
    s_t[x] = (t == ignore_index || t < 0 || t >= C) ? -1 : static_cast<int>(t);
    s_lse[x] = lse[rowbase + x];
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= C) return;
  const int i0 = y >> 3;
  const int i1 = min(i0 + 1, h - 1);
  const float l1h = static_cast<float>(y & 7) * 0.125f, l0h = 1.f - l1h;
  const float* L0 = logits + (static_cast<size_t>(n) * h + i0) * w * pitch + c;
  const float* L1 = logits + (static_cast<size_t>(n) * h + i1) * w * pitch + c;
  float* Trow = T + ((static_cast<size_t>(n) * Ho + y) * w) * C + c;
  float a = L0[0], cc = L1[0];      // left node column of the current interval (rows i0 / i1)
  float nb = L0[static_cast<size_t>(min(1, w - 1)) * pitch], nd = L1[static_cast<size_t>(min(1, w - 1)) * pitch];
  float carry = 0.f;                // right-node contribution of the previous interval
  for (int j0 = 0; j0 < w; ++j0) {
    const float b = nb, d = nd;     // right node column (j1 = min(j0+1, w-1))
    const int jn = min(j0 + 2, w - 1);
    nb = L0[static_cast<size_t>(jn) * pitch];          // prefetch the next interval's right column
    nd = L1[static_cast<size_t>(jn) * pitch];
    float accL = 0.f, accR = 0.f;
    const int xb = j0 * 8;
    const int xe = min(xb + 8, Wo);
#pragma unroll 8
    for (int x = xb; x < xe; ++x) {
      const int t = s_t[x];
      if (t < 0) continue;  // warp-uniform
      const float l1w = static_cast<float>(x & 7) * 0.125f, l0w = 1.f - l1w;
      const float v = l0h * (l0w * a + l1w * b) + l1h * (l0w * cc + l1w * d);
      const float g = __expf(v - s_lse[x]) - (c == t ? 1.f : 0.f);
      accL = fmaf(l0w, g, accL);
      accR = fmaf(l1w, g, accR);
    }
    Trow[static_cast<size_t>(j0) * C] = carry + accL;
    carry = accR;
    a = b;
    cc = d;
  }
}

__global__ void __launch_bounds__(256)
upsample_ce_bwd_cols_kernel(const float* __restrict__ T, int N, int h, int w, int C, int Ho,
                            const float* __restrict__ loss_info, const float* __restrict__ grad_out,
                            float* __restrict__ dlogits) {
  const int i = blockIdx.x, n = blockIdx.y;
  const float cntv = loss_info[1];
  const float gs = grad_out[0] / (cntv > 0.f ? cntv : 1.f);
  const int y_lo = max(8 * i - 7, 0), y_hi = min(8 * i + 7, Ho - 1);
  const int wc = w * C;
  for (int idx = threadIdx.x; idx < wc; idx += blockDim.x) {
    float acc = 0.f;
    for (int y = y_lo; y <= y_hi; ++y) {
      const float wy = 1.f - static_cast<float>(abs(y - 8 * i)) * 0.125f;
      acc = fmaf(wy, T[(static_cast<size_t>(n) * Ho + y) * wc + idx], acc);
    }
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
  return 2LL * N * cdiv(Ho, 32) * cdiv(Wo, 32);
}

extern "C" int semseg_upsample_ce_fwd(const float* logits, int pitch, int N, int h, int w, int C,
                                      const int64_t* target, int Ho, int Wo, int ignore_index, float* workspace,
                                      float* loss_out, int64_t* argmax, float* lse, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int r = check_tail(logits, pitch, N, h, w, C, target, Ho, Wo);
  if (r) return r;
  SB_CHECK_ARG(workspace && loss_out && lse, "upsample_ce_fwd: null output");
  dim3 grid(cdiv(Wo, 32), cdiv(Ho, 32), N);
  const size_t smem = static_cast<size_t>(36) * C * sizeof(float);
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(upsample_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 36 * kMaxClasses * (int)sizeof(float)));
    attr = true;
  }
  upsample_ce_fwd_kernel<<<grid, dim3(32, 8), smem, stream>>>(
      logits, pitch, N, h, w, C, reinterpret_cast<const long long*>(target), Ho, Wo, ignore_index, workspace,
      reinterpret_cast<long long*>(argmax), lse);
  SB_LAUNCHED();
  upsample_ce_reduce_kernel<<<1, 256, 0, stream>>>(workspace, static_cast<int>(grid.x * grid.y * grid.z), loss_out);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" long long semseg_upsample_ce_bwd_workspace_floats(int N, int Ho, int w, int C) {
  return static_cast<long long>(N) * Ho * w * C;
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
  const size_t smem = static_cast<size_t>(Wo) * 8;
  upsample_ce_bwd_rows_kernel<<<dim3(Ho, N), threads, smem, stream>>>(
      logits, pitch, N, h, w, C, reinterpret_cast<const long long*>(target), Ho, Wo, ignore_index, lse, workspace);
  SB_LAUNCHED();
  upsample_ce_bwd_cols_kernel<<<dim3(h, N), 256, 0, stream>>>(workspace, N, h, w, C, Ho, loss_info, grad_out, dlogits);
  SB_LAUNCHED();
  return SEMSEG_OK;
}
