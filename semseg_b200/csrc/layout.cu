// Layout conversion at the module boundary and weight packing.
//   - NCHW fp32 <-> NHWC bf16 / fp32 (model/pspnet.py:80-105 takes and returns NCHW fp32)
//   - fp32 OIHW master weights -> bf16 [tap][row][col] operand slabs for the implicit-GEMM kernels
#include "host_common.h"
#include "ptx.cuh"
#include "act.cuh"

namespace sb {

// 32x32 tiled transpose between [C][HW] (NCHW plane) and [HW][pitch] (NHWC) per image.
template <typename TIn, typename TOut>
__global__ void nchw_to_nhwc_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, TOut* __restrict__ out_lo, int C,
                                    int HW, int out_pitch) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TIn* src = in + static_cast<size_t>(n) * C * HW;
  TOut* dst = out + static_cast<size_t>(n) * HW * out_pitch;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? static_cast<float>(src[static_cast<size_t>(c) * HW + p]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      const float v = tile[threadIdx.x][r];
      const TOut hi = static_cast<TOut>(v);
      dst[static_cast<size_t>(p) * out_pitch + c] = hi;
      if (out_lo)   // split storage: lo = v - hi
        out_lo[static_cast<size_t>(n) * HW * out_pitch + static_cast<size_t>(p) * out_pitch + c] =
            static_cast<TOut>(v - static_cast<float>(hi));
    }
  }
}

template <typename TIn, typename TOut>
__global__ void nhwc_to_nchw_kernel(const TIn* __restrict__ in, const TIn* __restrict__ in_lo, TOut* __restrict__ out,
                                    int C, int HW, int in_pitch) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TIn* src = in + static_cast<size_t>(n) * HW * in_pitch;
  TOut* dst = out + static_cast<size_t>(n) * C * HW;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      v = static_cast<float>(src[static_cast<size_t>(p) * in_pitch + c]);
      if (in_lo) v += static_cast<float>(in_lo[static_cast<size_t>(n) * HW * in_pitch + static_cast<size_t>(p) * in_pitch + c]);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[static_cast<size_t>(c) * HW + p] = static_cast<TOut>(tile[threadIdx.x][r]);
  }
}

// w[co][ci][t] fp32 -> wf[t][co][ci] bf16 (rows_f x cols_f, zero padded). One thread per output element.
// split != 0: the lo slab (bf16(v - hi)) is written directly behind the hi slab.
__global__ void pack_wf_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, __nv_bfloat16* __restrict__ wf,
                               int rows, int cols, int split) {
  const size_t total = static_cast<size_t>(taps) * rows * cols;
  for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(idx % cols);
    const size_t r = idx / cols;
    const int co = static_cast<int>(r % rows);
    const int t = static_cast<int>(r / rows);
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[(static_cast<size_t>(co) * Cin + ci) * taps + t];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    wf[idx] = hi;
    if (split) wf[total + idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// w[co][ci][t] fp32 -> wd[t][ci][co] bf16 via a 32x32 smem transpose of the (co, ci) plane per tap.
__global__ void pack_wd_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, __nv_bfloat16* __restrict__ wd,
                               int rows, int cols, int split) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    tile[r][threadIdx.x] = (co < Cout && ci < Cin) ? w[(static_cast<size_t>(co) * Cin + ci) * taps + t] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    if (ci < rows && co < cols) {
      const float v = tile[threadIdx.x][r];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const size_t o = (static_cast<size_t>(t) * rows + ci) * cols + co;
      wd[o] = hi;
      if (split) wd[static_cast<size_t>(gridDim.z) * rows * cols + o] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
  }
}

// All conv weights of a model in ONE launch (the per-layer kernels above are launch-latency bound: 2 x 63 launches of
// 2-3 us per step for PSPNet50). A block owns one 32 (co) x 32 (ci) tile of one layer with all of its taps: the fp32
// OIHW rows are read once (32*taps contiguous floats per output channel), parked in shared memory, and written out as
// both bf16 operand slabs (wf[t][co][ci] and wd[t][ci][co], zero padded to the slab widths). items[] lives in device
// memory and is sorted by tile0; the block finds its layer by binary search.
__global__ void __launch_bounds__(256) pack_multi_kernel(const semseg_pack_item* __restrict__ items, int n_items) {
  extern __shared__ float pk_tile[];  // [32][32 * taps + 1]
  __shared__ int s_item;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;
    const int b = static_cast<int>(blockIdx.x);
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].tile0 <= b) lo = mid; else hi = mid - 1;
    }
    s_item = lo;
  }
  __syncthreads();
  const semseg_pack_item it = items[s_item];
  const int lt = static_cast<int>(blockIdx.x) - it.tile0;
  const int co0 = (lt / it.tiles_ci) * 32, ci0 = (lt % it.tiles_ci) * 32;
  const int taps = it.taps, rowlen = 32 * taps, pitch = rowlen + 1;
  const int nci = max(0, min(32, it.Cin - ci0)), nco = max(0, min(32, it.Cout - co0));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < 32; r += 8) {
    const float* src = it.w + (static_cast<size_t>(co0 + r) * it.Cin + ci0) * taps;
    for (int e = lane; e < rowlen; e += 32) pk_tile[r * pitch + e] = (r < nco && e < nci * taps) ? src[e] : 0.f;
  }
  __syncthreads();
  const int total = taps * 1024;
  if (it.wf) {
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(it.wf);
    for (int idx = threadIdx.x; idx < total; idx += 256) {
      const int ci_l = idx & 31, co_l = (idx >> 5) & 31, t = idx >> 10;
      if (co_l < nco && ci0 + ci_l < it.cols_f) {
        const float v = pk_tile[co_l * pitch + ci_l * taps + t];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const size_t o = (static_cast<size_t>(t) * it.Cout + co0 + co_l) * it.cols_f + ci0 + ci_l;
        wf[o] = hi;
        if (it.split) wf[static_cast<size_t>(taps) * it.Cout * it.cols_f + o] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
    }
  }
  if (it.wd) {
    __nv_bfloat16* wd = static_cast<__nv_bfloat16*>(it.wd);
    for (int idx = threadIdx.x; idx < total; idx += 256) {
      const int co_l = idx & 31, ci_l = (idx >> 5) & 31, t = idx >> 10;
      if (ci_l < nci && co0 + co_l < it.cols_d) {
        const float v = pk_tile[co_l * pitch + ci_l * taps + t];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const size_t o = (static_cast<size_t>(t) * it.Cin + ci0 + ci_l) * it.cols_d + co0 + co_l;
        wd[o] = hi;
        if (it.split) wd[static_cast<size_t>(taps) * it.Cin * it.cols_d + o] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
    }
  }
}

// Stem conv (3x3, stride 2, pad 1, <= 3 input channels; model/resnet.py:106-108): patches of the input, so the conv
// becomes ONE 64-wide K block (27 real values) per output pixel instead of 9 taps x 64-wide K blocks that are 7/8 zero
// fill fetched in 16-byte pieces:  P[n, ho, wo, (r*3+s)*3 + c] = x[n, 2ho-1+r, 2wo-1+s, c]  (zero outside, 27..31 zero).
__global__ void __launch_bounds__(256) im2col3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int pitch, int N, int H,
                                                          int W, int Cin, int Ho, int Wo, __nv_bfloat16* __restrict__ out) {
  const long long total = static_cast<long long>(N) * Ho * Wo;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(idx % Wo);
    const int ho = static_cast<int>((idx / Wo) % Ho);
    const int n = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
    __align__(16) __nv_bfloat16 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float2bfloat16_rn(0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hi = 2 * ho - 1 + r;
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const int wi = 2 * wo - 1 + sx;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
          const uint2 px = *reinterpret_cast<const uint2*>(x + ((static_cast<size_t>(n) * H + hi) * W + wi) * pitch);
          const __nv_bfloat16* pc = reinterpret_cast<const __nv_bfloat16*>(&px);   // 4 channels, Cin <= 3 are real
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (c < Cin) v[(r * 3 + sx) * Cin + c] = pc[c];
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + idx * 32);
    const uint4* src = reinterpret_cast<const uint4*>(v);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = src[q];
  }
}

// Stride-2 convolutions run on the stride-1 tensor-core kernel through a 2x2 phase decomposition:
//   xp[(ph*2+pw)*N + n][i][j][c] = x[n][2i+ph][2j+pw][c]   (zero where 2i+ph >= H or 2j+pw >= W)
// so tap (r, s) of a stride-2 conv reads phase ((r+1)&1, (s+1)&1) at a shift of -1 or 0.
__global__ void space_to_phases_kernel(const __nv_bfloat16* __restrict__ x, int pitch, int N, int H, int W, int C,
                                       __nv_bfloat16* __restrict__ xp, int Hh, int Wh) {
  const int groups = C >> 3;
  const long long total = 4LL * N * Hh * Wh * groups;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long p = idx / groups;
    const int j = static_cast<int>(p % Wh);
    p /= Wh;
    const int i = static_cast<int>(p % Hh);
    p /= Hh;
    const int n = static_cast<int>(p % N);
    const int q = static_cast<int>(p / N);
    const int h = 2 * i + (q >> 1), w = 2 * j + (q & 1);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (h < H && w < W)
      v = *reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + h) * W + w) * pitch + g * 8);
    *reinterpret_cast<uint4*>(xp + (idx / groups) * C + g * 8) = v;
  }
}

__global__ void phases_to_space_kernel(const __nv_bfloat16* __restrict__ xp, int N, int H, int W, int C, int Hh,
                                       int Wh, __nv_bfloat16* __restrict__ x) {
  const int groups = C >> 3;
  const long long total = static_cast<long long>(N) * H * W * groups;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long p = idx / groups;
    const int w = static_cast<int>(p % W);
    p /= W;
    const int h = static_cast<int>(p % H);
    const int n = static_cast<int>(p / H);
    const int q = (h & 1) * 2 + (w & 1);
    const size_t src = (((static_cast<size_t>(q) * N + n) * Hh + (h >> 1)) * Wh + (w >> 1)) * C + g * 8;
    *reinterpret_cast<uint4*>(x + (idx / groups) * C + g * 8) = *reinterpret_cast<const uint4*>(xp + src);
  }
}

}  // namespace sb

using namespace sb;
typedef __nv_bfloat16 bf16;

extern "C" int semseg_space_to_phases(const void* x, int x_pitch, int N, int H, int W, int C, void* xp,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(x && xp && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && x_pitch % 8 == 0,
               "space_to_phases: bad args");
  const int Hh = (H + 1) / 2, Wh = (W + 1) / 2;
  const long long total = 4LL * N * Hh * Wh * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  space_to_phases_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(x), x_pitch, N, H,
                                                                           W, C, static_cast<bf16*>(xp), Hh, Wh);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_phases_to_space(const void* xp, int N, int H, int W, int C, void* x, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(x && xp && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "phases_to_space: bad args");
  const int Hh = (H + 1) / 2, Wh = (W + 1) / 2;
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  phases_to_space_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(xp), N, H, W, C,
                                                                           Hh, Wh, static_cast<bf16*>(x));
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_pack_weights(const float* w_oihw, int Cout, int Cin, int taps, void* wf, int rows_f, int cols_f,
                                   void* wd, int rows_d, int cols_d, int split, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(w_oihw && Cout > 0 && Cin > 0 && taps > 0 && taps <= SEMSEG_MAX_TAPS, "pack_weights: bad args");
  if (wf) {
    SB_CHECK_ARG(rows_f >= Cout && cols_f >= Cin && cols_f % 8 == 0, "pack_weights: bad wf dims %d x %d", rows_f,
                 cols_f);
    const size_t total = static_cast<size_t>(taps) * rows_f * cols_f;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    pack_wf_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(w_oihw, Cout, Cin, taps, static_cast<bf16*>(wf),
                                                                     rows_f, cols_f, split);
    SB_LAUNCHED();
  }
  if (wd) {
    SB_CHECK_ARG(rows_d >= Cin && cols_d >= Cout && cols_d % 8 == 0, "pack_weights: bad wd dims %d x %d", rows_d,
                 cols_d);
    dim3 grid(cdiv(rows_d, 32), cdiv(cols_d, 32), taps);
    pack_wd_kernel<<<grid, dim3(32, 8), 0, stream>>>(w_oihw, Cout, Cin, taps, static_cast<bf16*>(wd), rows_d, cols_d,
                                                     split);
    SB_LAUNCHED();
  }
  return SEMSEG_OK;
}

extern "C" int semseg_im2col3x3s2(const void* x, int x_pitch, int N, int H, int W, int Cin, void* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(x && out && N > 0 && H > 0 && W > 0 && Cin >= 1 && Cin <= 3 && x_pitch >= 4 && x_pitch % 4 == 0,
               "im2col3x3s2: needs 1..3 input channels in a pitch that is a multiple of 4 (got Cin=%d pitch=%d)", Cin,
               x_pitch);
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(N) * Ho * Wo;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sb::im2col3x3s2_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const bf16*>(x), x_pitch, N, H, W,
                                                                           Cin, Ho, Wo, static_cast<bf16*>(out));
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_pack_weights_multi(const semseg_pack_item* items_dev, int n_items, int n_tiles, int max_taps,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(items_dev && n_items > 0 && n_tiles > 0 && max_taps > 0 && max_taps <= SEMSEG_MAX_TAPS,
               "pack_weights_multi: bad args");
  const size_t smem = static_cast<size_t>(32) * (32 * max_taps + 1) * sizeof(float);
  sb::pack_multi_kernel<<<static_cast<unsigned>(n_tiles), 256, smem, stream>>>(items_dev, n_items);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_nchw_f32_to_nhwc_bf16(const float* in, void* out, void* out_lo, int N, int C, int H, int W,
                                            int out_pitch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0 && out_pitch >= C, "nchw_f32_to_nhwc_bf16: bad args");
  dim3 grid(cdiv(H * W, 32), cdiv(C, 32), N);
  nchw_to_nhwc_kernel<float, bf16><<<grid, dim3(32, 8), 0, stream>>>(in, static_cast<bf16*>(out),
                                                                     static_cast<bf16*>(out_lo), C, H * W, out_pitch);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_nhwc_bf16_to_nchw_f32(const void* in, const void* in_lo, float* out, int N, int C, int H, int W,
                                            int in_pitch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0 && in_pitch >= C, "nhwc_bf16_to_nchw_f32: bad args");
  dim3 grid(cdiv(H * W, 32), cdiv(C, 32), N);
  nhwc_to_nchw_kernel<bf16, float><<<grid, dim3(32, 8), 0, stream>>>(
      static_cast<const bf16*>(in), static_cast<const bf16*>(in_lo), out, C, H * W, in_pitch);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

extern "C" int semseg_nhwc_f32_to_nchw_f32(const float* in, float* out, int N, int C, int H, int W, int in_pitch,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0 && in_pitch >= C, "nhwc_f32_to_nchw_f32: bad args");
  dim3 grid(cdiv(H * W, 32), cdiv(C, 32), N);
  nhwc_to_nchw_kernel<float, float><<<grid, dim3(32, 8), 0, stream>>>(in, nullptr, out, C, H * W, in_pitch);
  SB_LAUNCHED();
  return SEMSEG_OK;
}
