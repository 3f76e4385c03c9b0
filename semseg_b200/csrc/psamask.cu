// PSA mask collect / distribute (forward and backward), fp32 NCHW, bit-exact data movement.
//
// Replaces lib/psa/src/gpu/psamask_cuda.cu:8-128 of the reference (one thread per (n,h,w), every load
// its own 32-byte sector). Here the op is decomposed per (n, h, i) into a *shear* of a 2-D slab:
//
//   collect   : out[n, i*W + j, h, w] = in[n, a*mW + b, h, w],   a = i - h + hh,  b = j - w + hw
//   distribute: out[n, h*W + w, i, j] = the same element (collect transposed over (i,j) <-> (h,w))
//
// For fixed (n, h, i) the input elements are rows b of W contiguous floats (row pitch H*W) and the
// output slab is W rows of W contiguous floats, so both sides move whole rows; the shear (and the
// transpose for distribute) happens in shared memory with a conflict-free pitch. The element needed by
// output column w always comes from input column w, so no cross-lane shuffle is required.
// Every output element is written (zeros included): callers need no memset.
#include "host_common.h"

namespace sb {

// Work decomposition: a block owns (n, h) and G consecutive target rows i (forward) or mask rows a (backward); one WARP
// moves one row at a time with lane = position inside the row (W <= 32: the shipped 30x30 / 59x59 geometry; wider maps take
// the strided loop), so there is no per-element index division, and every lane keeps kRowsInFlight independent row loads
// in flight before the first shared-memory write (the kernel is latency-bound: rows are 120 bytes at a 3600-byte pitch).
constexpr int kPsaThreads = 256;
constexpr int kPsaWarps = kPsaThreads / 32;
constexpr int kRowsInFlight = 6;
constexpr int kPsaG = 2;

// TYPE 0 = collect, 1 = distribute. grid = N * H * ceil(H / G) blocks (n, h, i-group).
template <int TYPE>
__global__ void __launch_bounds__(kPsaThreads) psamask_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                  int N, int H, int W, int mH, int mW) {
  extern __shared__ float S[];   // [G][mW][P]
  const int hh = (mH - 1) / 2, hw = (mW - 1) / 2;
  const int P = TYPE == 0 ? ((W | 1) + 1) : (W | 1);
  const int groups = (H + kPsaG - 1) / kPsaG;
  const int ig = blockIdx.x % groups;
  const int h = (blockIdx.x / groups) % H;
  const int n = blockIdx.x / (groups * H);
  const int i0 = ig * kPsaG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t HW = static_cast<size_t>(H) * W;
  const float* src_n = in + static_cast<size_t>(n) * mH * mW * HW + static_cast<size_t>(h) * W;
  const int rows_in = kPsaG * mW;
  for (int w0 = 0; w0 < W; w0 += 32) {          // one pass for W <= 32
    const int w = w0 + lane;
    for (int r0 = warp * kRowsInFlight; r0 < rows_in; r0 += kPsaWarps * kRowsInFlight) {
      float v[kRowsInFlight];
      bool ok[kRowsInFlight];
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        const int r = r0 + u;
        const int g = r / mW, b = r - g * mW;
        const int i = i0 + g, a = i - h + hh, j = b + w - hw;
        ok[u] = r < rows_in && i < H && a >= 0 && a < mH && w < W && j >= 0 && j < W;
        v[u] = ok[u] ? src_n[(static_cast<size_t>(a) * mW + b) * HW + w] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u)
        if (ok[u]) S[(r0 + u) * P + w] = v[u];
    }
  }
  __syncthreads();
  const int rows_out = kPsaG * W;
  for (int r = warp; r < rows_out; r += kPsaWarps) {
    const int g = r / W, k = r - g * W;          // k = j (collect) or w (distribute)
    const int i = i0 + g, a = i - h + hh;
    if (i >= H) break;
    const bool a_ok = a >= 0 && a < mH;
    for (int c = lane; c < W; c += 32) {         // c = w (collect) or j (distribute)
      const int j = TYPE == 0 ? k : c, w = TYPE == 0 ? c : k;
      const int b = j - w + hw;
      const float v = (a_ok && b >= 0 && b < mW) ? S[(g * mW + b) * P + w] : 0.f;
      if (TYPE == 0)
        out[((static_cast<size_t>(n) * HW + static_cast<size_t>(i) * W + j) * H + h) * W + w] = v;
      else
        out[((static_cast<size_t>(n) * HW + static_cast<size_t>(h) * W + w) * H + i) * W + j] = v;
    }
  }
}

// grid = N * H * ceil(mH / G) blocks (n, h, a-group): writes the whole din slabs [mW][W] (zeros where nothing maps).
template <int TYPE>
__global__ void __launch_bounds__(kPsaThreads) psamask_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din,
                                                                  int N, int H, int W, int mH, int mW) {
  extern __shared__ float S[];   // [G][W][P]
  const int hh = (mH - 1) / 2, hw = (mW - 1) / 2;
  const int P = (W | 1) + 1;
  const int groups = (mH + kPsaG - 1) / kPsaG;
  const int ag = blockIdx.x % groups;
  const int h = (blockIdx.x / groups) % H;
  const int n = blockIdx.x / (groups * H);
  const int a0 = ag * kPsaG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t HW = static_cast<size_t>(H) * W;
  const int rows_in = kPsaG * W;
  for (int c0 = 0; c0 < W; c0 += 32) {
    const int c = c0 + lane;
    for (int r0 = warp * kRowsInFlight; r0 < rows_in; r0 += kPsaWarps * kRowsInFlight) {
      float v[kRowsInFlight];
      bool ok[kRowsInFlight];
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u) {
        const int r = r0 + u;
        const int g = r / W, k = r - g * W;       // k = j (collect) or w (distribute); c = the other one
        const int a = a0 + g, i = a + h - hh;
        ok[u] = r < rows_in && a < mH && i >= 0 && i < H && c < W;
        size_t off = 0;
        if (TYPE == 0) off = ((static_cast<size_t>(n) * HW + static_cast<size_t>(i) * W + k) * H + h) * W + c;
        else off = ((static_cast<size_t>(n) * HW + static_cast<size_t>(h) * W + k) * H + i) * W + c;
        v[u] = ok[u] ? dout[off] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kRowsInFlight; ++u)
        if (ok[u]) S[(r0 + u) * P + c] = v[u];
    }
  }
  __syncthreads();
  const int rows_out = kPsaG * mW;
  for (int r = warp; r < rows_out; r += kPsaWarps) {
    const int g = r / mW, b = r - g * mW;
    const int a = a0 + g, i = a + h - hh;
    if (a >= mH) break;
    const bool i_ok = i >= 0 && i < H;
    float* dst = din + ((static_cast<size_t>(n) * mH * mW + static_cast<size_t>(a) * mW + b) * H + h) * W;
    for (int w = lane; w < W; w += 32) {
      const int j = b + w - hw;
      float v = 0.f;
      if (i_ok && j >= 0 && j < W) v = TYPE == 0 ? S[(g * W + j) * P + w] : S[(g * W + w) * P + j];
      dst[w] = v;
    }
  }
}

static int check_psa(int psa_type, const void* a, const void* b, int N, int H, int W, int mH, int mW) {
  SB_CHECK_ARG(psa_type == 0 || psa_type == 1, "psamask: psa_type must be 0 (collect) or 1 (distribute)");
  SB_CHECK_ARG(a && b, "psamask: null pointer");
  SB_CHECK_ARG(N > 0 && H > 0 && W > 0, "psamask: bad feature size");
  SB_CHECK_ARG(mH > 0 && mW > 0 && (mH & 1) && (mW & 1), "psamask: mask dims must be odd (got %d x %d)", mH, mW);
  return SEMSEG_OK;
}

}  // namespace sb

extern "C" int semseg_psamask_fwd(int psa_type, const float* in, float* out, int N, int H, int W, int mH, int mW,
                                  void* stream_) {
  using namespace sb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int r = check_psa(psa_type, in, out, N, H, W, mH, mW);
  if (r) return r;
  const int P = (W | 1) + 1;
  const size_t smem = static_cast<size_t>(kPsaG) * mW * P * sizeof(float);
  SB_CHECK_ARG(smem <= 200 * 1024, "psamask: mask too large for shared memory");
  const unsigned grid = static_cast<unsigned>(N) * H * ((H + kPsaG - 1) / kPsaG);
  if (psa_type == 0) {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_fwd_kernel<0><<<grid, kPsaThreads, smem, stream>>>(in, out, N, H, W, mH, mW);
  } else {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_fwd_kernel<1><<<grid, kPsaThreads, smem, stream>>>(in, out, N, H, W, mH, mW);
  }
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_psamask_bwd(int psa_type, const float* dout, float* din, int N, int H, int W, int mH, int mW,
                                  void* stream_) {
  using namespace sb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int r = check_psa(psa_type, dout, din, N, H, W, mH, mW);
  if (r) return r;
  const int P = (W | 1) + 1;
  const size_t smem = static_cast<size_t>(kPsaG) * W * P * sizeof(float);
  SB_CHECK_ARG(smem <= 200 * 1024, "psamask: feature map too large for shared memory");
  const unsigned grid = static_cast<unsigned>(N) * H * ((mH + kPsaG - 1) / kPsaG);
  if (psa_type == 0) {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_bwd_kernel<0><<<grid, kPsaThreads, smem, stream>>>(dout, din, N, H, W, mH, mW);
  } else {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_bwd_kernel<1><<<grid, kPsaThreads, smem, stream>>>(dout, din, N, H, W, mH, mW);
  }
  SB_LAUNCHED();
  return SEMSEG_OK;
}
