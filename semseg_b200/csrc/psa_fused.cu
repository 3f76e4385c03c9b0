// Fused point-wise spatial attention: mask gather -> softmax -> aggregation in ONE kernel, so that
// the [N, HW, HW] attention map of model/psanet.py:81-91 (psa_mask -> F.softmax(dim=1) -> torch.bmm, three fp32 round
// trips of a 3.24 MB/image/branch tensor plus an NHWC->NCHW copy of the 12.5 MB logits) never exists in HBM.
//
// Per image:   out[t, :] = (1/norm) * sum_s P[t, s] * feat[s, :],   P[t, :] = softmax_s( L[t, s] )
//   collect    (psa_type 0): L[t, s] = A[t, idx(s - t)]   (the TARGET pixel's own 59x59 attention vector)
//   distribute (psa_type 1): L[t, s] = A[s, idx(t - s)]   (one entry of every SOURCE pixel's vector)
//   idx(d) = (d.y + hh) * mW + (d.x + hw); positions outside the mask window contribute logit 0 (the reference zero-fills
//   psa_mask's output BEFORE the softmax, lib/psa/functions/psamask.py:17).
// A = attention logits fp32 NHWC [N, HW, a_pitch] straight from the 1x1 conv's F32 epilogue (no NCHW copy).
//
// One kernel template covers the forward aggregation AND the feature gradient of the backward pass, which is the same
// contraction with rows and reduction index swapped:  dfeat[s, :] = (1/norm) * sum_t P[t, s] * dout[t, :].
//   kRowOwner : the attention vector of element (row, k) belongs to the row pixel (else to the k pixel)
//   kStatsRow : the softmax statistics (max, 1/sum) of element (row, k) belong to the row (else to k)
//     forward  collect: (1,1)   forward distribute: (0,1)   dfeat collect: (0,0)   dfeat distribute: (1,0)
//
// CTA = 128 rows (whole grid rows of the HxW map: 4 x 30 = 120 live rows for the shipped 30x30 geometry).
//   warp 0      : TMA producer of the B operand: feat/dout K blocks [64 pixels x 512 channels] bf16 as eight
//                 [64 ch, 64 px] boxes = MN-major SWIZZLE_128B operand (exactly the wgrad kernel's operand form)
//   warp 1      : tcgen05.mma issuer: D[128 x 512 fp32, all 512 TMEM columns] += P[128 x 64] * B[64 x 512]
//   warps 2..9  : (a) softmax statistics of the CTA's rows (forward), (b) per K block: gather 128 x 64 logits, exp,
//                 normalise, bf16 -> K-major SWIZZLE_128B A-operand stage in shared memory, (c) epilogue: TMEM -> scale
//                 -> bf16 (or hi/lo pair) -> global.
// bf16x3 (split feat / out): the K loop runs three times (P_hi*B_hi, P_lo*B_hi, P_hi*B_lo) into the same accumulator.
#include "host_common.h"
#include "ptx.cuh"
#include "act.cuh"

namespace sb {

constexpr int kPfRows = 128;
constexpr int kPfK = 64;
constexpr int kPfC = 512;                         // channels of feat / out (mid_channels of the PSA module)
constexpr int kPfBoxBytes = kPfK * 128;           // one [64 ch, 64 px] box
constexpr int kPfBBytes = (kPfC / 64) * kPfBoxBytes;   // 64 KB
constexpr int kPfABytes = kPfRows * 128;          // 16 KB
constexpr int kPfStageBytes = kPfABytes + kPfBBytes;
constexpr int kPfStages = 2;
constexpr int kPfWorkers = 256;                   // warps 2..9
constexpr int kPfThreads = 64 + kPfWorkers;
constexpr int kPfMisc = 256 + 2 * kPfRows * 4 + 2 * 8 * kPfRows * 4;   // barriers, row stats, per-warp partial stats
constexpr int kPfSmem = kPfStages * kPfStageBytes + kPfMisc + 1024;

struct PsaFusedParams {
  const float* A;       // [N][Q][a_pitch]
  float2* stats;        // [N][Q] (max, 1/sum) per target; written when kStatsRow (forward), read otherwise
  __nv_bfloat16* out;   // [N][Q][out_pitch]
  __nv_bfloat16* out_lo;
  int out_pitch;
  int N, H, W, mH, mW, a_pitch;
  int rows_per_tile;    // grid rows per CTA tile: 128 / W
  int tiles_per_img;
  int nseg;             // 1 (bf16) or 3 (bf16x3)
  float scale;          // 1 / normalization_factor
};

__device__ __forceinline__ float pf_logit(const float* __restrict__ An, int a_pitch, int own, int own_i, int own_j,
                                          int oth_i, int oth_j, int hh, int hw, int mH, int mW) {
  const int a = oth_i - own_i + hh, b = oth_j - own_j + hw;
  return (a >= 0 && a < mH && b >= 0 && b < mW) ? __ldg(An + static_cast<size_t>(own) * a_pitch + a * mW + b) : 0.f;
}

template <bool kRowOwner, bool kStatsRow>
__global__ void __launch_bounds__(kPfThreads, 1)
psa_attend_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB_lo,
                  const PsaFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* misc = smem + kPfStages * kPfStageBytes;
  uint64_t* full_b = reinterpret_cast<uint64_t*>(misc);       // TMA bytes of the B tile landed
  uint64_t* full_a = full_b + kPfStages;                      // all workers wrote the P tile
  uint64_t* empty = full_a + kPfStages;                       // MMAs that read the stage completed
  uint64_t* tmem_full = empty + kPfStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_m = reinterpret_cast<float*>(misc + 256);          // [128] row max
  float* s_inv = s_m + kPfRows;                               // [128] row 1/sum

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.tiles_per_img;
  const int tile = blockIdx.x - n * p.tiles_per_img;
  const int Q = p.H * p.W;
  const int row_i0 = tile * p.rows_per_tile;                              // first grid row of this tile
  const int live_rows = min(p.rows_per_tile, p.H - row_i0) * p.W;         // rows of the tile that exist
  const int hh = (p.mH - 1) / 2, hw = (p.mW - 1) / 2;
  const int num_kb = (Q + kPfK - 1) / kPfK;
  const int total_kb = num_kb * p.nseg;
  const float* An = p.A + static_cast<size_t>(n) * Q * p.a_pitch;
  float2* stats_n = p.stats + static_cast<size_t>(n) * Q;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    if (p.nseg > 1) tma_prefetch_desc(&tmB_lo);
    for (int i = 0; i < kPfStages; ++i) {
      mbar_init(&full_b[i], 1);
      mbar_init(&full_a[i], kPfWorkers);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  // rows of the 128-row A tile that do not exist in this CTA's tile stay zero for the whole kernel
  for (int s = 0; s < kPfStages; ++s) {
    uint4* a4 = reinterpret_cast<uint4*>(smem + s * kPfStageBytes);
    for (int i = threadIdx.x; i < kPfABytes / 16; i += kPfThreads) a4[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer (B operand)
    if (elect_one()) {
      for (int it = 0; it < total_kb; ++it) {
        const int s = it % kPfStages;
        const uint32_t par = (it / kPfStages) & 1;
        mbar_wait(&empty[s], par ^ 1);
        const int seg = it / num_kb, kb = it - seg * num_kb;
        const CUtensorMap* m = (seg == 2) ? &tmB_lo : &tmB;
        uint8_t* dst = smem + s * kPfStageBytes + kPfABytes;
        mbar_expect_tx(&full_b[s], kPfBBytes);
#pragma unroll
        for (int bx = 0; bx < kPfC / 64; ++bx) {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
              "[%2];" ::"r"(smem_u32(dst + bx * kPfBoxBytes)),
              "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(&full_b[s])), "r"(bx * 64), "r"(kb * kPfK), "r"(n)
              : "memory");
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kPfRows, 256, 0, 1);   // A K-major, B MN-major
      for (int it = 0; it < total_kb; ++it) {
        const int s = it % kPfStages;
        const uint32_t par = (it / kPfStages) & 1;
        mbar_wait(&full_b[s], par);
        mbar_wait(&full_a[s], par);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * kPfStageBytes);
        const uint32_t b_addr = a_addr + kPfABytes;
        const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
#pragma unroll
        for (int nh = 0; nh < 2; ++nh) {
          const uint64_t bdesc = make_smem_desc_sw128(b_addr + nh * 4 * kPfBoxBytes, kPfBoxBytes, 1024);
#pragma unroll
          for (int k = 0; k < kPfK / 16; ++k)
            umma_bf16(tmem_base + nh * 256, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 128),
                      idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ===================================================================== workers: statistics, P tiles, epilogue
    const int wt = threadIdx.x - 64;          // 0..255
    const int ww = wt >> 5;                    // worker warp 0..7
    // ---- (a) softmax statistics of the tile's rows (forward kernels)
    if (kStatsRow) {
      if (kRowOwner) {
        // collect: row = target, its own attention vector, contiguous along the source column -> one warp per row
        for (int r = ww; r < kPfRows; r += kPfWorkers / 32) {
          float m = -INFINITY, sum = 0.f;
          if (r < live_rows) {
            const int ri = row_i0 + r / p.W, rj = r % p.W, own = ri * p.W + rj;
            for (int q = lane; q < Q; q += 32) m = fmaxf(m, pf_logit(An, p.a_pitch, own, ri, rj, q / p.W, q % p.W, hh, hw, p.mH, p.mW));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            for (int q = lane; q < Q; q += 32)
              sum += __expf(pf_logit(An, p.a_pitch, own, ri, rj, q / p.W, q % p.W, hh, hw, p.mH, p.mW) - m);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) stats_n[own] = make_float2(m, 1.f / sum);
          }
          if (lane == 0) {
            s_m[r] = m;
            s_inv[r] = (r < live_rows) ? 1.f / sum : 0.f;
          }
        }
      } else {
        // distribute: row = target, one entry of every source vector; consecutive rows read consecutive addresses ->
        // lanes along rows, the sources split over the 8 worker warps (online softmax), partials merged through smem
        float* pm = reinterpret_cast<float*>(misc + 256 + 2 * kPfRows * 4);   // [8][128] partial max
        float* ps = pm + 8 * kPfRows;                                           // [8][128] partial sum
        const int q_per = (Q + 7) / 8, q0 = ww * q_per, q1 = min(Q, q0 + q_per);
        for (int r = lane; r < kPfRows; r += 32) {
          float m = -INFINITY, sum = 0.f;
          if (r < live_rows) {
            const int ri = row_i0 + r / p.W, rj = r % p.W;
            for (int q = q0; q < q1; ++q) {
              const float l = pf_logit(An, p.a_pitch, q, q / p.W, q % p.W, ri, rj, hh, hw, p.mH, p.mW);
              const float mn = fmaxf(m, l);
              sum = sum * __expf(m - mn) + __expf(l - mn);
              m = mn;
            }
          }
          pm[ww * kPfRows + r] = m;
          ps[ww * kPfRows + r] = sum;
        }
        named_bar_sync(1, kPfWorkers);
        if (wt < kPfRows) {
          const int r = wt;
          float m = -INFINITY, sum = 0.f;
          if (r < live_rows) {
#pragma unroll
            for (int k = 0; k < 8; ++k) m = fmaxf(m, pm[k * kPfRows + r]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float mk = pm[k * kPfRows + r];
              if (mk > -INFINITY) sum += ps[k * kPfRows + r] * __expf(mk - m);
            }
            const int ri = row_i0 + r / p.W, rj = r % p.W;
            stats_n[ri * p.W + rj] = make_float2(m, 1.f / sum);
          }
          s_m[r] = m;
          s_inv[r] = (r < live_rows) ? 1.f / sum : 0.f;
        }
      }
      named_bar_sync(1, kPfWorkers);
    }
    // ---- (b) P tiles
    for (int it = 0; it < total_kb; ++it) {
      const int s = it % kPfStages;
      const uint32_t par = (it / kPfStages) & 1;
      const int seg = it / num_kb, kb = it - seg * num_kb;
      mbar_wait(&empty[s], par ^ 1);
      uint8_t* a_st = smem + s * kPfStageBytes;
      const bool want_lo = seg == 1;
      if (kRowOwner) {
        // the row pixel owns the attention vector -> addresses are contiguous along k: one warp per (row, 32-column
        // half), lanes along k; the live (row, half) items are dealt round-robin to the 8 worker warps
        // [forward collect, dfeat distribute]
        for (int item = ww; item < 2 * live_rows; item += kPfWorkers / 32) {
          const int r = item >> 1, half = item & 1;
          const int ri = row_i0 + r / p.W, rj = r % p.W, rpos = ri * p.W + rj;
          const int kl = half * 32 + lane, q = kb * kPfK + kl;
          float pv = 0.f;
          if (q < Q) {
            const int qi = q / p.W, qj = q - qi * p.W;
            const float l = pf_logit(An, p.a_pitch, rpos, ri, rj, qi, qj, hh, hw, p.mH, p.mW);
            if (kStatsRow) {
              pv = __expf(l - s_m[r]) * s_inv[r];
            } else {
              const float2 st = stats_n[q];
              pv = __expf(l - st.x) * st.y;
            }
          }
          __nv_bfloat16 hi = __float2bfloat16_rn(pv);
          if (want_lo) hi = __float2bfloat16_rn(pv - __bfloat162float(hi));
          *reinterpret_cast<__nv_bfloat16*>(a_st + r * 128 + (((kl >> 3) ^ (r & 7)) << 4) + (kl & 7) * 2) = hi;
        }
      } else {
        // the k pixel owns the vector -> addresses are contiguous along the row index: lanes along rows; the
        // (32-row group, k column) items are dealt round-robin to the 8 worker warps  [forward distribute, dfeat collect]
        const int row_groups = (live_rows + 31) >> 5;
        for (int item = ww; item < row_groups * kPfK; item += kPfWorkers / 32) {
          const int rg = item % row_groups, kl = item / row_groups;
          const int r = rg * 32 + lane, q = kb * kPfK + kl;
          if (r >= live_rows) continue;
          float pv = 0.f;
          if (q < Q) {
            const int ri = row_i0 + r / p.W, rj = r % p.W;
            const int qi = q / p.W, qj = q - qi * p.W;
            const float l = pf_logit(An, p.a_pitch, q, qi, qj, ri, rj, hh, hw, p.mH, p.mW);
            if (kStatsRow) {
              pv = __expf(l - s_m[r]) * s_inv[r];
            } else {
              const float2 st = stats_n[q];
              pv = __expf(l - st.x) * st.y;
            }
          }
          __nv_bfloat16 hi = __float2bfloat16_rn(pv);
          if (want_lo) hi = __float2bfloat16_rn(pv - __bfloat162float(hi));
          *reinterpret_cast<__nv_bfloat16*>(a_st + r * 128 + (((kl >> 3) ^ (r & 7)) << 4) + (kl & 7) * 2) = hi;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_a[s]);
    }
    // ---- (c) epilogue: TMEM -> registers -> scale -> bf16 (hi/lo) -> global
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int g = warp & 3;                    // TMEM lane quarter of this warp (hardware rule: warp id % 4)
    const int chalf = (warp - 2) >> 2;         // 0: columns 0..255, 1: 256..511
    const int r = g * 32 + lane;
    const bool r_ok = r < live_rows;
    const long long orow = (static_cast<long long>(n) * Q + static_cast<long long>(row_i0) * p.W + r) * p.out_pitch;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[2][32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) + static_cast<uint32_t>(chalf * 256 + ch * 64);
      tmem_ld_32x32(taddr, v[0]);
      tmem_ld_32x32(taddr + 32, v[1]);
      tmem_ld_wait();
      if (r_ok) {
#pragma unroll
        for (int j8 = 0; j8 < 8; ++j8) {
          float f[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(v[(j8 * 8 + q) >> 5][(j8 * 8 + q) & 31]) * p.scale;
          const long long o = orow + chalf * 256 + ch * 64 + j8 * 8;
          if (p.out_lo) act_st8<true>(p.out, p.out_lo, o, f);
          else act_st8<false>(p.out, nullptr, o, f);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool kRowOwner, bool kStatsRow>
static int launch_attend(const CUtensorMap& tmB, const CUtensorMap& tmB_lo, const PsaFusedParams& p, int grid,
                         cudaStream_t stream) {
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    SB_CUDA(cudaFuncSetAttribute(psa_attend_kernel<kRowOwner, kStatsRow>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kPfSmem));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  psa_attend_kernel<kRowOwner, kStatsRow><<<grid, kPfThreads, kPfSmem, stream>>>(tmB, tmB_lo, p);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

}  // namespace sb

// mode 0: out = P * feat (forward; writes stats)      mode 1: dfeat = P^T * dout (backward; reads stats)
extern "C" int semseg_psa_attend(int mode, int psa_type, const float* attn, int a_pitch, const void* feat,
                                 const void* feat_lo, int feat_pitch, float* stats, void* out, void* out_lo, int out_pitch,
                                 int N, int H, int W, int mH, int mW, int C, float scale, void* stream_) {
  using namespace sb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(attn && feat && stats && out && N > 0 && H > 0 && W > 0, "psa_attend: bad args");
  SB_CHECK_ARG(mode == 0 || mode == 1, "psa_attend: mode must be 0 (forward) or 1 (feature gradient)");
  SB_CHECK_ARG(psa_type == 0 || psa_type == 1, "psa_attend: psa_type must be 0 (collect) or 1 (distribute)");
  SB_CHECK_ARG(mH > 0 && mW > 0 && (mH & 1) && (mW & 1) && a_pitch >= mH * mW, "psa_attend: bad mask geometry");
  SB_CHECK_ARG(C == kPfC, "psa_attend: feature width must be %d (got %d)", kPfC, C);
  SB_CHECK_ARG(W <= kPfRows, "psa_attend: feature maps wider than %d are not supported", kPfRows);
  SB_CHECK_ARG(feat_pitch % 8 == 0 && out_pitch % 8 == 0 && feat_pitch >= C && out_pitch >= C, "psa_attend: bad pitch");
  SB_CHECK_ARG((feat_lo != nullptr) == (out_lo != nullptr), "psa_attend: feat and out must use the same storage form");
  PsaFusedParams p;
  memset(&p, 0, sizeof(p));
  p.A = attn; p.stats = reinterpret_cast<float2*>(stats);
  p.out = static_cast<__nv_bfloat16*>(out); p.out_lo = static_cast<__nv_bfloat16*>(out_lo); p.out_pitch = out_pitch;
  p.N = N; p.H = H; p.W = W; p.mH = mH; p.mW = mW; p.a_pitch = a_pitch;
  // grid rows per CTA: at most 128 / W, fewer when that leaves SMs idle (the tensor-core work is negligible, the
  // per-row gather / exp work of the workers is what takes the time)
  p.rows_per_tile = kPfRows / W;
  {
    const int want = (N * H) / num_sms();
    const int rpt = want < 1 ? 1 : want;
    if (rpt < p.rows_per_tile) p.rows_per_tile = rpt;
  }
  p.tiles_per_img = cdiv(H, p.rows_per_tile);
  p.nseg = feat_lo ? 3 : 1;
  p.scale = scale;
  CUtensorMap tmB, tmB_lo;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)H * W, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)feat_pitch * 2, (uint64_t)feat_pitch * 2 * H * W};
    uint32_t box[3] = {64u, (uint32_t)kPfK, 1u};
    int r = encode_tmap_bf16(&tmB, feat, 3, dims, str, box);
    if (r) return r;
    tmB_lo = tmB;
    if (feat_lo && (r = encode_tmap_bf16(&tmB_lo, feat_lo, 3, dims, str, box))) return r;
  }
  const int grid = N * p.tiles_per_img;
  // (row owner, stats on row): forward collect (1,1), forward distribute (0,1), dfeat collect (0,0), dfeat distribute (1,0)
  if (mode == 0) return psa_type == 0 ? launch_attend<true, true>(tmB, tmB_lo, p, grid, stream)
                                      : launch_attend<false, true>(tmB, tmB_lo, p, grid, stream);
  return psa_type == 0 ? launch_attend<false, false>(tmB, tmB_lo, p, grid, stream)
                       : launch_attend<true, false>(tmB, tmB_lo, p, grid, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention-logit gradient of the fused op (the softmax backward, flash-attention style: nothing [HW x HW] is stored):
//   dP[t, s] = scale * sum_c dout[t, c] * feat[s, c]                       (GEMM on tcgen05: M = targets, N = sources, K = C)
//   D[t]     = sum_c dout[t, c] * out[t, c]            (= sum_s P[t, s] * dP[t, s])
//   dL[t, s] = P[t, s] * (dP[t, s] - D[t])             P recomputed from the logits and the saved (max, 1/sum)
//   dA[owner][idx(other - owner)] = dL[t, s]           owner = t (collect) or s (distribute); dA is zero elsewhere (caller
//                                                      zero-fills it: 74 % of a full 59x59 mask never receives gradient)
// CTA = 128 target rows; 4 source blocks of 256 (two TMEM accumulator stages: the epilogue of block j overlaps the MMAs of
// block j+1); operands K-major from TMA ([64 c, 128 rows] of dout, [64 c, 256 rows] of feat), 4-stage ring.
namespace sb {

constexpr int kPgBlockN = 256;
constexpr int kPgABytes = 128 * 128;             // [128 rows][64 c] bf16
constexpr int kPgBBytes = kPgBlockN * 128;       // [256 rows][64 c]
constexpr int kPgStageBytes = kPgABytes + kPgBBytes;   // 48 KB
constexpr int kPgStages = 4;
constexpr int kPgThreads = 64 + 256;
constexpr int kPgSmem = kPgStages * kPgStageBytes + 1024 + 1024;

struct PsaGradParams {
  const float* A;
  const float2* stats;
  float* dA;
  const __nv_bfloat16* dout;
  const __nv_bfloat16* dout_lo;
  const __nv_bfloat16* out;
  const __nv_bfloat16* out_lo;
  int dout_pitch, out_pitch;
  int N, H, W, mH, mW, a_pitch, C;
  int rows_per_tile, tiles_per_img, nseg;
  float scale;
};

template <bool kCollect>
__global__ void __launch_bounds__(kPgThreads, 1)
psa_attn_grad_kernel(const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDO_lo,
                     const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmF_lo,
                     const PsaGradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* misc = smem + kPgStages * kPgStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(misc);
  uint64_t* empty = full + kPgStages;
  uint64_t* tmem_full = empty + kPgStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.tiles_per_img;
  const int tile = blockIdx.x - n * p.tiles_per_img;
  const int Q = p.H * p.W;
  const int row_i0 = tile * p.rows_per_tile;
  const int q_row0 = row_i0 * p.W;                                       // first target position of the tile
  const int live_rows = min(p.rows_per_tile, p.H - row_i0) * p.W;
  const int hh = (p.mH - 1) / 2, hw = (p.mW - 1) / 2;
  const int k_blocks = p.C / 64;
  // one 256-source block per CTA (blockIdx.y): the per-element epilogue (recompute P, scatter) dominates, so the source
  // blocks of a target tile run on different SMs
  const int nb_first = blockIdx.y, nb_last = blockIdx.y + 1;
  const int per_nb = k_blocks * p.nseg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmF);
    for (int i = 0; i < kPgStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int nb = nb_first; nb < nb_last; ++nb) {
        for (int kk = 0; kk < per_nb; ++kk, ++it) {
          const int s = it % kPgStages;
          const uint32_t par = (it / kPgStages) & 1;
          mbar_wait(&empty[s], par ^ 1);
          const int seg = kk / k_blocks, kb = kk - seg * k_blocks;   // 0: do_hi*f_hi, 1: do_lo*f_hi, 2: do_hi*f_lo
          const CUtensorMap* mA = seg == 1 ? &tmDO_lo : &tmDO;
          const CUtensorMap* mB = seg == 2 ? &tmF_lo : &tmF;
          uint8_t* a_dst = smem + s * kPgStageBytes;
          mbar_expect_tx(&full[s], kPgStageBytes);
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
              "[%2];" ::"r"(smem_u32(a_dst)),
              "l"(reinterpret_cast<uint64_t>(mA)), "r"(smem_u32(&full[s])), "r"(kb * 64), "r"(q_row0), "r"(n)
              : "memory");
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
              "[%2];" ::"r"(smem_u32(a_dst + kPgABytes)),
              "l"(reinterpret_cast<uint64_t>(mB)), "r"(smem_u32(&full[s])), "r"(kb * 64), "r"(nb * kPgBlockN), "r"(n)
              : "memory");
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kPgBlockN, 0, 0);   // both operands K-major
      int it = 0;
      for (int nb = nb_first; nb < nb_last; ++nb) {
        const int as = (nb - nb_first) & 1;
        mbar_wait(&tmem_empty[as], (((nb - nb_first) >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kPgBlockN);
        for (int kk = 0; kk < per_nb; ++kk, ++it) {
          const int s = it % kPgStages;
          const uint32_t par = (it / kPgStages) & 1;
          mbar_wait(&full[s], par);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kPgStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(a_addr + kPgABytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                      (kk > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    const int g = warp & 3;
    const int chalf = (warp - 2) >> 2;          // columns [chalf*128, +128) of every 256-source block
    const int r = g * 32 + lane;
    const bool r_ok = r < live_rows;
    const int ti = row_i0 + r / p.W, tj = r % p.W, tpos = ti * p.W + tj;
    const float* An = p.A + static_cast<size_t>(n) * Q * p.a_pitch;
    float* dAn = p.dA + static_cast<size_t>(n) * Q * p.a_pitch;
    // D[t] = <dout[t, :], out[t, :]> and the row's softmax statistics
    float D = 0.f, rm = 0.f, rinv = 0.f;
    if (r_ok) {
      const long long o1 = (static_cast<long long>(n) * Q + tpos) * p.dout_pitch;
      const long long o2 = (static_cast<long long>(n) * Q + tpos) * p.out_pitch;
      for (int c = 0; c < p.C; c += 8) {
        float a[8], b[8];
        if (p.dout_lo) {
          act_ld8<true>(p.dout, p.dout_lo, o1 + c, a);
          act_ld8<true>(p.out, p.out_lo, o2 + c, b);
        } else {
          act_ld8<false>(p.dout, nullptr, o1 + c, a);
          act_ld8<false>(p.out, nullptr, o2 + c, b);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) D = fmaf(a[q], b[q], D);
      }
      const float2 st = p.stats[static_cast<size_t>(n) * Q + tpos];
      rm = st.x;
      rinv = st.y;
    }
    for (int nb = nb_first; nb < nb_last; ++nb) {
      const int as = (nb - nb_first) & 1;
      mbar_wait(&tmem_full[as], ((nb - nb_first) >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(g * 32) << 16) +
                               static_cast<uint32_t>(as * kPgBlockN + chalf * 128 + ch * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (r_ok) {
          const int s0 = nb * kPgBlockN + chalf * 128 + ch * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int s = s0 + j;
            if (s >= Q) continue;
            const int si = s / p.W, sj = s - si * p.W;
            // owner / other of the attention entry
            const int oi = kCollect ? ti : si, oj = kCollect ? tj : sj, own = kCollect ? tpos : s;
            const int a = (kCollect ? si : ti) - oi + hh, b = (kCollect ? sj : tj) - oj + hw;
            if (a >= 0 && a < p.mH && b >= 0 && b < p.mW) {
              const size_t off = static_cast<size_t>(own) * p.a_pitch + a * p.mW + b;
              const float pv = __expf(__ldg(An + off) - rm) * rinv;
              dAn[off] = pv * (p.scale * __uint_as_float(v[j]) - D);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <bool kCollect>
static int launch_attn_grad(const CUtensorMap& a, const CUtensorMap& al, const CUtensorMap& b, const CUtensorMap& bl,
                            const PsaGradParams& p, int grid, cudaStream_t stream) {
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    SB_CUDA(cudaFuncSetAttribute(psa_attn_grad_kernel<kCollect>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPgSmem));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  const dim3 g(grid, cdiv(p.H * p.W, kPgBlockN));
  psa_attn_grad_kernel<kCollect><<<g, kPgThreads, kPgSmem, stream>>>(a, al, b, bl, p);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

}  // namespace sb

extern "C" int semseg_psa_attend_bwd_attn(int psa_type, const float* attn, int a_pitch, const float* stats,
                                          const void* feat, const void* feat_lo, int feat_pitch, const void* out,
                                          const void* out_lo, int out_pitch, const void* dout, const void* dout_lo,
                                          int dout_pitch, float* dattn, int N, int H, int W, int mH, int mW, int C,
                                          float scale, void* stream_) {
  using namespace sb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(attn && stats && feat && out && dout && dattn && N > 0 && H > 0 && W > 0, "psa_attend_bwd_attn: bad args");
  SB_CHECK_ARG(psa_type == 0 || psa_type == 1, "psa_attend_bwd_attn: psa_type must be 0 or 1");
  SB_CHECK_ARG(mH > 0 && mW > 0 && (mH & 1) && (mW & 1) && a_pitch >= mH * mW, "psa_attend_bwd_attn: bad mask geometry");
  SB_CHECK_ARG(C > 0 && C % 64 == 0 && W <= 128, "psa_attend_bwd_attn: C %% 64 == 0 and W <= 128 required");
  SB_CHECK_ARG(feat_pitch % 8 == 0 && out_pitch % 8 == 0 && dout_pitch % 8 == 0, "psa_attend_bwd_attn: bad pitch");
  const bool split = feat_lo != nullptr;
  SB_CHECK_ARG((out_lo != nullptr) == split && (dout_lo != nullptr) == split,
               "psa_attend_bwd_attn: all activations must use the same storage form");
  PsaGradParams p;
  memset(&p, 0, sizeof(p));
  p.A = attn; p.stats = reinterpret_cast<const float2*>(stats); p.dA = dattn;
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.dout_lo = static_cast<const __nv_bfloat16*>(dout_lo);
  p.out = static_cast<const __nv_bfloat16*>(out); p.out_lo = static_cast<const __nv_bfloat16*>(out_lo);
  p.dout_pitch = dout_pitch; p.out_pitch = out_pitch;
  p.N = N; p.H = H; p.W = W; p.mH = mH; p.mW = mW; p.a_pitch = a_pitch; p.C = C;
  p.rows_per_tile = 128 / W;
  {
    const int want = (N * H * cdiv(H * W, kPgBlockN)) / num_sms();     // keep every SM busy (rows <-> epilogue threads)
    const int rpt = want < 1 ? 1 : want;
    if (rpt < p.rows_per_tile) p.rows_per_tile = rpt;
  }
  p.tiles_per_img = cdiv(H, p.rows_per_tile);
  p.nseg = split ? 3 : 1;
  p.scale = scale;
  // the caller's dattn must be zero where no gradient lands
  SB_CUDA(cudaMemsetAsync(dattn, 0, sizeof(float) * static_cast<size_t>(N) * H * W * a_pitch, stream));
  CUtensorMap tmDO, tmDO_lo, tmF, tmF_lo;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)H * W, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)dout_pitch * 2, (uint64_t)dout_pitch * 2 * H * W};
    uint32_t box[3] = {64u, 128u, 1u};
    int r = encode_tmap_bf16(&tmDO, dout, 3, dims, str, box);
    if (r) return r;
    tmDO_lo = tmDO;
    if (split && (r = encode_tmap_bf16(&tmDO_lo, dout_lo, 3, dims, str, box))) return r;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)H * W, (uint64_t)N};
    uint64_t str[2] = {(uint64_t)feat_pitch * 2, (uint64_t)feat_pitch * 2 * H * W};
    uint32_t box[3] = {64u, (uint32_t)kPgBlockN, 1u};
    int r = encode_tmap_bf16(&tmF, feat, 3, dims, str, box);
    if (r) return r;
    tmF_lo = tmF;
    if (split && (r = encode_tmap_bf16(&tmF_lo, feat_lo, 3, dims, str, box))) return r;
  }
  const int grid = N * p.tiles_per_img;
  return psa_type == 0 ? launch_attn_grad<true>(tmDO, tmDO_lo, tmF, tmF_lo, p, grid, stream)
                       : launch_attn_grad<false>(tmDO, tmDO_lo, tmF, tmF_lo, p, grid, stream);
}
