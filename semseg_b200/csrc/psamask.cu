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

// TYPE 0 = collect, 1 = distribute. grid = N*H*H blocks (n, h, i).
template <int TYPE>
__global__ void __launch_bounds__(128) psamask_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          int N, int H, int W, int mH, int mW, int h_fastest) {
  extern __shared__ float S[];
  const int hh = (mH - 1) / 2, hw = (mW - 1) / 2;
  const int P = TYPE == 0 ? ((W | 1) + 1) : (W | 1);
  // block order: with h fastest, concurrently running blocks write the 30 adjacent rows of the same output planes
  const int i = h_fastest ? (blockIdx.x / H) % H : blockIdx.x % H;
  const int h = h_fastest ? blockIdx.x % H : (blockIdx.x / H) % H;
  const int n = blockIdx.x / (H * H);
  const int a = i - h + hh;
  const bool a_ok = a >= 0 && a < mH;
  const size_t HW = static_cast<size_t>(H) * W;
  if (a_ok) {
    const float* src = in + ((static_cast<size_t>(n) * mH * mW + static_cast<size_t>(a) * mW) * H + h) * W;
    for (int idx = threadIdx.x; idx < mW * W; idx += blockDim.x) {
      const int b = idx / W, w = idx - b * W;
      const int j = b + w - hw;
      if (j >= 0 && j < W) S[b * P + w] = src[static_cast<size_t>(b) * HW + w];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < W * W; idx += blockDim.x) {
    int j, w;
    if (TYPE == 0) {
      j = idx / W;
      w = idx - j * W;
    } else {
      w = idx / W;
      j = idx - w * W;
    }
    const int b = j - w + hw;
    const float v = (a_ok && b >= 0 && b < mW) ? S[b * P + w] : 0.f;
    if (TYPE == 0)
      out[((static_cast<size_t>(n) * HW + static_cast<size_t>(i) * W + j) * H + h) * W + w] = v;
    else
      out[((static_cast<size_t>(n) * HW + static_cast<size_t>(h) * W + w) * H + i) * W + j] = v;
  }
}

// grid = N*H*mH blocks (n, h, a): writes the whole din slab [mW][W] (zeros where nothing maps).
template <int TYPE>
__global__ void __launch_bounds__(128) psamask_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din,
                                                          int N, int H, int W, int mH, int mW, int h_fastest) {
  extern __shared__ float S[];
  const int hh = (mH - 1) / 2, hw = (mW - 1) / 2;
  const int P = (W | 1) + 1;
  const int a = h_fastest ? (blockIdx.x / H) % mH : blockIdx.x % mH;
  const int h = h_fastest ? blockIdx.x % H : (blockIdx.x / mH) % H;
  const int n = blockIdx.x / (mH * H);
  const int i = a + h - hh;
  const bool i_ok = i >= 0 && i < H;
  const size_t HW = static_cast<size_t>(H) * W;
  if (i_ok) {
    for (int idx = threadIdx.x; idx < W * W; idx += blockDim.x) {
      const int r = idx / W, c = idx - r * W;
      if (TYPE == 0)  // r = j, c = w
        S[r * P + c] = dout[((static_cast<size_t>(n) * HW + static_cast<size_t>(i) * W + r) * H + h) * W + c];
      else  // r = w, c = j
        S[r * P + c] = dout[((static_cast<size_t>(n) * HW + static_cast<size_t>(h) * W + r) * H + i) * W + c];
    }
  }
  __syncthreads();
  float* dst = din + ((static_cast<size_t>(n) * mH * mW + static_cast<size_t>(a) * mW) * H + h) * W;
  for (int idx = threadIdx.x; idx < mW * W; idx += blockDim.x) {
    const int b = idx / W, w = idx - b * W;
    const int j = b + w - hw;
    float v = 0.f;
    if (i_ok && j >= 0 && j < W) v = TYPE == 0 ? S[j * P + w] : S[w * P + j];
    dst[static_cast<size_t>(b) * HW + w] = v;
  }
}

static int psa_h_fastest() {
  const char* e = getenv("SEMSEG_B200_PSA_ORDER");     // re-read every call: the bench A/Bs the two orders in one run
  return (e && e[0] == '0') ? 0 : 1;
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
  const size_t smem = static_cast<size_t>(mW) * P * sizeof(float);
  SB_CHECK_ARG(smem <= 200 * 1024, "psamask: mask too large for shared memory");
  const unsigned grid = static_cast<unsigned>(N) * H * H;
  if (psa_type == 0) {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_fwd_kernel<0><<<grid, 128, smem, stream>>>(in, out, N, H, W, mH, mW, psa_h_fastest());
  } else {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_fwd_kernel<1><<<grid, 128, smem, stream>>>(in, out, N, H, W, mH, mW, psa_h_fastest());
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
  const size_t smem = static_cast<size_t>(W) * P * sizeof(float);
  SB_CHECK_ARG(smem <= 200 * 1024, "psamask: feature map too large for shared memory");
  const unsigned grid = static_cast<unsigned>(N) * H * mH;
  if (psa_type == 0) {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_bwd_kernel<0><<<grid, 128, smem, stream>>>(dout, din, N, H, W, mH, mW, psa_h_fastest());
  } else {
    if (smem > 48 * 1024)
      SB_CUDA(cudaFuncSetAttribute(psamask_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psamask_bwd_kernel<1><<<grid, 128, smem, stream>>>(dout, din, N, H, W, mH, mW, psa_h_fastest());
  }
  SB_LAUNCHED();
  return SEMSEG_OK;
}
