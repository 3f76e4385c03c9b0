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
// Backward is a deterministic gather: a CTA *owns* 3x7 low-res nodes, recomputes the softmax of every output
// pixel in their support (32 x 64 pixels), reduces along x with warp shuffles inside the 8-pixel interval groups,
// then along y in a fixed order. No atomics, every dlogits element is written exactly once.
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
      float m = -INFINITY, vt = 0.f;
      int am = 0;
      for (int c = 0; c < C; ++c) {
        const float v = l0h * (l0w * r0[cj0 + c] + l1w * r0[cj1 + c]) + l1h * (l0w * r1[cj0 + c] + l1w * r1[cj1 + c]);
        if (v > m) {
          m = v;
          am = c;
        }
        if (c == t) vt = v;
      }
      float s = 0.f;
      for (int c = 0; c < C; ++c) {
        const float v = l0h * (l0w * r0[cj0 + c] + l1w * r0[cj1 + c]) + l1h * (l0w * r1[cj0 + c] + l1w * r1[cj1 + c]);
        s += __expf(v - m);
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
constexpr int kOwnI = 3;   // owned node rows per CTA  -> 4 row intervals = 32 output rows
constexpr int kOwnJ = 7;   // owned node cols per CTA  -> 8 col intervals = 64 output cols (2 warps per row)

// block 256 threads = 8 warps. For each of the 4 row intervals (8 output rows): warp -> (row in interval r = warp>>1 ...)
// see body. Shared: nodes [5][9][C], T [8 rows][7][C], acc [3][7][C].
__global__ void __launch_bounds__(256)
upsample_ce_bwd_kernel(const float* __restrict__ logits, int pitch, int N, int h, int w, int C,
                       const long long* __restrict__ target, int Ho, int Wo, int ignore_index,
                       const float* __restrict__ lse, const float* __restrict__ loss_info,
                       const float* __restrict__ grad_out, float* __restrict__ dlogits) {
  extern __shared__ float sm[];
  float* S = sm;                          // [(kOwnI+2) * (kOwnJ+2)][C] node values; rows i0-1 .. i0+kOwnI
  float* T = S + (kOwnI + 2) * (kOwnJ + 2) * C;  // [8][kOwnJ][C]
  float* A = T + 8 * kOwnJ * C;           // [kOwnI][kOwnJ][C] accumulators
  const int n = blockIdx.z;
  const int i0 = blockIdx.y * kOwnI, j0 = blockIdx.x * kOwnJ;  // first owned node
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int SW = kOwnJ + 2;
  for (int idx = tid; idx < (kOwnI + 2) * SW * C; idx += 256) {
    const int c = idx % C;
    const int node = idx / C;
    const int jj = node % SW, ii = node / SW;
    const int gi = i0 - 1 + ii, gj = j0 - 1 + jj;
    float v = 0.f;
    if (gi >= 0 && gi < h && gj >= 0 && gj < w)
      v = logits[((static_cast<size_t>(n) * h + gi) * w + gj) * pitch + c];
    S[idx] = v;
  }
  for (int idx = tid; idx < kOwnI * kOwnJ * C; idx += 256) A[idx] = 0.f;
  __syncthreads();
  const float cntv = loss_info[1];
  const float gscale = grad_out[0] / (cntv > 0.f ? cntv : 1.f);

  // 4 row intervals q = 0..3 -> interval index iv = i0 - 1 + q, output rows y = 8*iv + r, r = 0..7.
  // Within an interval, warps (2 per row) process rows r = warp>>1 (+4 on the second pass); half = warp&1 picks
  // the 32-pixel half of the 64-pixel span: pixel x = 8*(j0-1) + half*32 + lane.
#pragma unroll 1
  for (int q = 0; q < kOwnI + 1; ++q) {
    const int iv = i0 - 1 + q;
    // T[r][jo][c] = sum_x Ww[x, j0+jo] * g[y, x, c]
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int r = (warp >> 1) + pass * 4;
      const int half = warp & 1;
      const int y = 8 * iv + r;
      const int x = 8 * (j0 - 1) + half * 32 + lane;
      const bool pvalid = iv >= 0 && y < Ho && y >= 0 && x >= 0 && x < Wo;
      // interpolation setup (clamped like ATen); node rows relative to S: (i - (i0-1))
      const int yi0 = pvalid ? (y >> 3) : max(i0 - 1, 0);
      const int yi1 = min(yi0 + 1, h - 1);
      const float l1h = static_cast<float>(y & 7) * 0.125f, l0h = 1.f - l1h;
      const int xj0 = pvalid ? (x >> 3) : max(j0 - 1, 0);
      const int xj1 = min(xj0 + 1, w - 1);
      const float l1w = static_cast<float>(x & 7) * 0.125f, l0w = 1.f - l1w;
      const float* r0 = S + ((yi0 - (i0 - 1)) * SW) * C;
      const float* r1 = S + ((yi1 - (i0 - 1)) * SW) * C;
      const int cj0 = (xj0 - (j0 - 1)) * C, cj1 = (xj1 - (j0 - 1)) * C;
      float lsev = 0.f;
      long long t = -1;
      bool contributes = false;
      if (pvalid) {
        const size_t pix = (static_cast<size_t>(n) * Ho + y) * Wo + x;
        t = target[pix];
        lsev = lse[pix];
        contributes = (t != ignore_index && t >= 0 && t < C);
      }
      // lane group = interval (x>>3); the group's left node is xj0, right node xj0+1.
      // group index within this warp: lane>>3 (0..3); global interval column jv = (j0-1) + half*4 + (lane>>3).
      const int grp = lane >> 3;
      // KC classes per iteration: all shared-memory reads first, then the shuffle butterflies interleaved, then
      // the stores (T/A alias S for the compiler, so without this the loop serialises on ~7 dependent shuffles).
      constexpr int KC = 5;
      float* stash = A + kOwnI * kOwnJ * C + r * C;
      for (int cb = 0; cb < C; cb += KC) {
        float a[KC], b[KC];
#pragma unroll
        for (int u = 0; u < KC; ++u) {
          const int c = cb + u;
          float g = 0.f;
          if (contributes && c < C) {
            const float v =
                l0h * (l0w * r0[cj0 + c] + l1w * r0[cj1 + c]) + l1h * (l0w * r1[cj0 + c] + l1w * r1[cj1 + c]);
            g = (__expf(v - lsev) - (c == t ? 1.f : 0.f)) * gscale;
          }
          a[u] = l0w * g;  // contribution to the left node of this pixel's interval
          b[u] = l1w * g;  // ... and to the right node
        }
        // reduce inside the 8-lane interval group (fixed butterfly order -> deterministic)
#pragma unroll
        for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
          for (int u = 0; u < KC; ++u) {
            a[u] += __shfl_xor_sync(0xffffffffu, a[u], o);
            b[u] += __shfl_xor_sync(0xffffffffu, b[u], o);
          }
        }
        float bp[KC];
#pragma unroll
        for (int u = 0; u < KC; ++u) bp[u] = __shfl_up_sync(0xffffffffu, b[u], 8);  // b of the previous group
        // Interval group (half, grp) has left node jl = half*4 + grp - 1 (owned index). Owned node jo receives
        // a(group with jl == jo) + b(group with jl == jo - 1); exactly one lane writes each T entry.
        if ((lane & 7) == 0) {
          const int jl = half * 4 + grp - 1;
#pragma unroll
          for (int u = 0; u < KC; ++u) {
            const int c = cb + u;
            if (c < C) {
              if (grp >= 1) T[(r * kOwnJ + jl) * C + c] = a[u] + bp[u];
              else if (half == 1) T[(r * kOwnJ + jl) * C + c] = a[u];  // + b of (half 0, grp 3): stashed below
              if (half == 0 && grp == 3) stash[c] = b[u];              // cross-warp term for owned node 3
            }
          }
        }
      }
    }
    __syncthreads();
    // node jo = 3 += stashed b of (half 0, group 3); then accumulate rows into A with the y weights, fixed order.
    for (int idx = tid; idx < kOwnJ * C; idx += 256) {
      const int c = idx % C;
      const int jo = idx / C;
      float acc_top = 0.f, acc_bot = 0.f;  // contributions to node row iv (weight l0h) and iv+1 (weight l1h)
      for (int r = 0; r < 8; ++r) {
        float tv = T[(r * kOwnJ + jo) * C + c];
        if (jo == 3) tv += A[kOwnI * kOwnJ * C + r * C + c];
        const float l1h = static_cast<float>(r) * 0.125f, l0h = 1.f - l1h;
        acc_top += l0h * tv;
        acc_bot += l1h * tv;
      }
      // interval iv: top node row = iv, bottom = iv+1 (owned rows are i0 .. i0+kOwnI-1)
      const int top = iv - i0, bot = iv + 1 - i0;
      if (top >= 0 && top < kOwnI) A[(top * kOwnJ + jo) * C + c] += acc_top;
      if (bot >= 0 && bot < kOwnI) A[(bot * kOwnJ + jo) * C + c] += acc_bot;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < kOwnI * kOwnJ * C; idx += 256) {
    const int c = idx % C;
    const int node = idx / C;
    const int jo = node % kOwnJ, io = node / kOwnJ;
    const int gi = i0 + io, gj = j0 + jo;
    if (gi < h && gj < w) dlogits[((static_cast<size_t>(n) * h + gi) * w + gj) * C + c] = A[idx];
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

extern "C" int semseg_upsample_ce_bwd(const float* logits, int pitch, int N, int h, int w, int C,
                                      const int64_t* target, int Ho, int Wo, int ignore_index, const float* lse,
                                      const float* loss_info, const float* grad_out, float* dlogits, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int r = check_tail(logits, pitch, N, h, w, C, target, Ho, Wo);
  if (r) return r;
  SB_CHECK_ARG(lse && loss_info && grad_out && dlogits, "upsample_ce_bwd: null pointer");
  dim3 grid(cdiv(w, kOwnJ), cdiv(h, kOwnI), N);
  const size_t floats = static_cast<size_t>((kOwnI + 2) * (kOwnJ + 2) + 8 * kOwnJ + kOwnI * kOwnJ + 8) * C;
  const size_t smem = floats * sizeof(float);
  static bool attr = false;
  if (!attr) {
    const int max_smem = ((kOwnI + 2) * (kOwnJ + 2) + 8 * kOwnJ + kOwnI * kOwnJ + 8) * kMaxClasses * (int)sizeof(float);
    SB_CUDA(cudaFuncSetAttribute(upsample_ce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr = true;
  }
  upsample_ce_bwd_kernel<<<grid, 256, smem, stream>>>(logits, pitch, N, h, w, C,
                                                     reinterpret_cast<const long long*>(target), Ho, Wo, ignore_index,
                                                     lse, loss_info, grad_out, dlogits);
  SB_LAUNCHED();
  return SEMSEG_OK;
}
