// Implicit-GEMM convolution (fprop and dgrad) for NHWC bf16 activations on tcgen05 tensor cores.
//
//   out[p, co] = sum_t sum_ci x[p + off(t), ci] * w[t][co][ci]
//
// GEMM view: M = pixels (tiles of a bh x bw pixel rectangle inside one image, <= 128 rows),
// N = Cout (BLOCK_N columns per tile), K = taps * Cin in blocks of 64 channels.
//
// Replaces the cuDNN convolutions behind nn.Conv2d in the reference (model/resnet.py:63-69,
// model/pspnet.py:49-58,65-69,73-77). One persistent CTA per SM, warp-specialised:
//   warp 0   : TMA producer   — A tile = 4-D box [64 ch, bw, bh, 1] at the tap-shifted pixel
//                               (TMA zero-fills the halo), B tile = 3-D box [64, BLOCK_N, 1] of the
//                               packed weights; both land in 128B-swizzled shared memory.
//   warp 1   : MMA issuer     — one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16),
//                               accumulators live in TMEM (2 stages so the epilogue overlaps the next tile).
//   warps 2-5: epilogue       — tcgen05.ld -> registers -> (affine / ReLU / residual) -> bf16 ->
//   (+ 6-9)                     swizzled smem -> TMA store; running BatchNorm statistics (sum, sum of squares,
//                               count per channel) of the stored bf16 values, one row per TMEM lane quarter.
//                               With kEpiGroups = 2 a second warpgroup takes every other 64-column chunk of the
//                               tile (own staging buffer, own named barrier): the 1x1 convs with wide outputs are
//                               epilogue-bound and a single warp per scheduler cannot hide its own latencies.
#include "host_common.h"
#include "ptx.cuh"

namespace sb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 bytes = one swizzle span
constexpr int kEpiGroupThreads = 128;  // one epilogue warpgroup = 4 warps = the 4 TMEM lane quarters
constexpr int conv_threads(int epi_groups) { return 64 + epi_groups * kEpiGroupThreads; }
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kStageOutBytes = kBlockM * 64 * 2;    // 16 KB epilogue staging chunk (64 columns)
constexpr int kMiscBytes = 2048;

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kBTileBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStageOutBytes + kMiscBytes + 1024;
  // CTA-pair mode: each CTA stages its own 128 pixel rows of A and HALF of the weight rows -> smaller stages, more of them
  static constexpr int kPairStageBytes = kATileBytes + kBTileBytes / 2;
  static constexpr int kPairStages = (kStages * kStageBytes) / kPairStageBytes;
};

struct ConvKParams {
  int N, H, W;
  int Cin, Cout;
  int taps;
  int bh, bw, tiles_h, tiles_w;
  int n_tiles, num_m_tiles;
  int k_chunks;  // ceil(Cin / 64)
  int dh[SEMSEG_MAX_TAPS], dw[SEMSEG_MAX_TAPS], wtap[SEMSEG_MAX_TAPS], img_add[SEMSEG_MAX_TAPS];
  int img_mul;
  int epi_mode, relu;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* residual_lo;  // split storage: lo plane of the residual
  int res_pitch;
  int nseg;  // operand segments per K block: 1 = bf16, 3 = bf16x3 (x_hi*w_hi, x_lo*w_hi, x_hi*w_lo)
  // K slicing (F32 epilogue only): work item = (pixel tile, channel tile, K slice); slice s accumulates K blocks
  // [s*kb_per_slice, (s+1)*kb_per_slice) and writes its fp32 partial to out_f32 + s*slice_stride. Bounds the length of
  // one tensor-core accumulation chain: tcgen05 accumulates in fp32 with truncation, a bias of ~2^-24 per MMA step
  // towards zero (tools/probe_accum.py), negligible for bf16 but not at the 1e-5 level the bf16x3 mode works at.
  int k_slices, kb_per_slice;
  long long slice_stride;
  float* out_f32;
  int out_pitch;
  float* stats_partial;  // [gridDim.x * 4][3][Cout]: per epilogue warp (sum, sum of squares, count) per channel
};

// kCluster (CTA pair, tcgen05 cta_group::2): two CTAs of a cluster own two neighbouring pixel tiles of the SAME channel
// block and execute ONE M=256 x N=BLOCK_N MMA per K step, issued by the leader CTA. Each CTA stages only its own 128
// pixel rows of A and HALF of the weight rows (the tensor core reads the other half from the peer's shared memory), so
// the operand bytes that cross L2->SM per FLOP drop by a third (16 KB + 16 KB instead of 16 KB + 32 KB per K block at
// BLOCK_N=256) and the freed smem buys more pipeline stages. TMA completions of both CTAs are credited to the leader's
// full barrier; the leader's tcgen05.commit is multicast to both CTAs' empty / tmem_full barriers; the peer's epilogue
// hands its accumulator stage back by arriving on the leader's tmem_empty barrier.
//
// kSplit (bf16x3 operand mode, activations stored as hi/lo bf16 planes — act.cuh): every K block is issued three times,
// (x_hi, w_hi), (x_lo, w_hi), (x_hi, w_lo), into the same fp32 TMEM accumulator; the producer just picks the hi or lo
// tensor map per segment, the MMA warp is unchanged. The epilogue splits its fp32 result into (hi, lo) again and stores
// both planes (two staging tiles, two TMA stores); BatchNorm statistics are taken from hi + lo.
template <int BLOCK_N, bool kCluster, int kEpiGroups, bool kSplit>
__global__ void __launch_bounds__(conv_threads(kEpiGroups), 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_lo, const __grid_constant__ CUtensorMap tmC_lo,
                  const ConvKParams p) {
  static_assert(!kSplit || kEpiGroups == 1, "split storage uses both staging buffers of the single epilogue group");
  using Cfg = ConvCfg<BLOCK_N>;
  constexpr int kNumThreads = conv_threads(kEpiGroups);
  constexpr int kEpiThreads = kEpiGroups * kEpiGroupThreads;
  static_assert(kEpiGroups == 1 || kEpiGroups == 2, "one or two epilogue warpgroups");
  constexpr int kStages = kCluster ? Cfg::kPairStages : Cfg::kStages;
  constexpr int kStageBytes = kCluster ? Cfg::kPairStageBytes : Cfg::kStageBytes;
  const uint32_t cta_rank = kCluster ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  // Work items: (pixel tile, channel tile) or, clustered, (pair of pixel tiles, channel tile) per cluster.
  const int item_first = kCluster ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = kCluster ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_items = (kCluster ? ((p.num_m_tiles + 1) >> 1) : p.num_m_tiles) * p.n_tiles * p.k_slices;
  auto decode_item = [&](int item, int& m_tile, int& n_tile, int& k_slice) -> bool {
    k_slice = item % p.k_slices;
    item /= p.k_slices;
    n_tile = item % p.n_tiles;
    const int mi = item / p.n_tiles;
    const int m_raw = kCluster ? 2 * mi + static_cast<int>(cta_rank) : mi;
    m_tile = m_raw < p.num_m_tiles ? m_raw : p.num_m_tiles - 1;  // odd tail: the second CTA shadows the last tile
    return m_raw < p.num_m_tiles;
  };

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* out_stage = smem + Cfg::kStages * Cfg::kStageBytes;  // 2 x 16 KB (same offset in both modes)
  uint8_t* misc = out_stage + 2 * kStageOutBytes;
  static_assert(kStages <= 16, "barrier area sized for <= 16 stages");
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc);
  uint64_t* empty_bar = full_bar + 16;
  uint64_t* tmem_full = empty_bar + 16;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nseg = kSplit ? p.nseg : 1;
  const int total_kblk = p.taps * p.k_chunks;   // K blocks = (64-channel block, tap) pairs
  auto slice_range = [&](int k_slice, int& kblk0, int& kblk1) {
    kblk0 = k_slice * p.kb_per_slice;
    kblk1 = min(kblk0 + p.kb_per_slice, total_kblk);
  };
  const uint32_t a_bytes = static_cast<uint32_t>(p.bh * p.bw) * 128u;
  // bytes credited to a full barrier per stage: own A + whole B, or (pair mode, leader's barrier) both A tiles + both B halves
  const uint32_t stage_tx = (kCluster ? 2u * a_bytes : a_bytes) + static_cast<uint32_t>(Cfg::kBTileBytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (kSplit) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
      tma_prefetch_desc(&tmC_lo);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kCluster ? 2 * kEpiThreads : kEpiThreads);  // pair: both CTAs' epilogues
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCluster) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_ptr);
    else tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  }
  if (p.stats_partial != nullptr) {  // four statistics rows per CTA: one per epilogue warp (32 accumulator rows each)
    float* row = p.stats_partial + static_cast<size_t>(blockIdx.x) * 4 * 3 * p.Cout;
    for (int i = threadIdx.x; i < 4 * 3 * p.Cout; i += kNumThreads) row[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int it = 0;
      for (int item = item_first; item < num_items; item += item_step) {
        int m_tile, n_tile, k_slice, kblk0, kblk1;
        decode_item(item, m_tile, n_tile, k_slice);
        slice_range(k_slice, kblk0, kblk1);
        const int tiles_per_img = p.tiles_h * p.tiles_w;
        const int img = m_tile / tiles_per_img;
        const int rem = m_tile - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.bh;
        const int w0 = (rem % p.tiles_w) * p.bw;
        const int n0 = n_tile * BLOCK_N;
        for (int kb = kblk0 * nseg; kb < kblk1 * nseg; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t par = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], par ^ 1);
          const int ks = kb / nseg;          // (channel block, tap)
          const int seg = kb - ks * nseg;    // 0: x_hi*w_hi, 1: x_lo*w_hi, 2: x_hi*w_lo
          const int cb = ks / p.taps;
          const int t = ks - cb * p.taps;
          const CUtensorMap* mA = (kSplit && seg == 1) ? &tmA_lo : &tmA;
          const CUtensorMap* mB = (kSplit && seg == 2) ? &tmB_lo : &tmB;
          uint8_t* a_dst = stage_base + s * kStageBytes;
          uint8_t* b_dst = a_dst + kATileBytes;
          if (kCluster) {
            // both CTAs' loads complete on the LEADER's full barrier; only the leader arms it
            const uint32_t lead_bar = mapa_u32(&full_bar[s], 0);
            if (is_leader) mbar_expect_tx(&full_bar[s], stage_tx);
            tma_load_4d_2sm(a_dst, mA, lead_bar, cb * kBlockK, w0 + p.dw[t], h0 + p.dh[t],
                            img * p.img_mul + p.img_add[t]);
            tma_load_3d_2sm(b_dst, mB, lead_bar, cb * kBlockK, n0 + static_cast<int>(cta_rank) * (BLOCK_N / 2),
                            p.wtap[t]);
          } else {
            mbar_expect_tx(&full_bar[s], stage_tx);
            tma_load_4d(a_dst, mA, &full_bar[s], cb * kBlockK, w0 + p.dw[t], h0 + p.dh[t],
                        img * p.img_mul + p.img_add[t]);
            // 3-D weights [taps][rows][cols]: coordinates (k, row, tap)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                "%5}], [%2];" ::"r"(smem_u32(b_dst)),
                "l"(reinterpret_cast<uint64_t>(mB)), "r"(smem_u32(&full_bar[s])), "r"(cb * kBlockK), "r"(n0),
                "r"(p.wtap[t])
                : "memory");
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (pair mode: leader CTA only)
    if ((!kCluster || is_leader) && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kCluster ? 2 * kBlockM : kBlockM, BLOCK_N, 0, 0);
      int it = 0;
      int tile_iter = 0;
      for (int item = item_first; item < num_items; item += item_step, ++tile_iter) {
        const int as = tile_iter & 1;
        const uint32_t apar = (tile_iter >> 1) & 1;
        mbar_wait(&tmem_empty[as], apar ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BLOCK_N);
        int kblk0, kblk1;
        slice_range((item % p.k_slices), kblk0, kblk1);
        const int num_kb = (kblk1 - kblk0) * nseg;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t par = (it / kStages) & 1;
          mbar_wait(&full_bar[s], par);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + s * kStageBytes);
          const uint32_t b_addr = a_addr + kATileBytes;
          const uint64_t adesc = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle span
            if (kCluster)
              umma_bf16_2sm(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            else
              umma_bf16(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem stage once these MMAs have read it (clustered: in both CTAs, the peer multicasts into it)
          if (kCluster) umma_commit_2sm_mcast(&empty_bar[s], static_cast<uint16_t>(3));
          else umma_commit(&empty_bar[s]);
        }
        // accumulator complete (pair mode: each CTA's epilogue waits on its own tmem_full barrier)
        if (kCluster) umma_commit_2sm_mcast(&tmem_full[as], static_cast<uint16_t>(3));
        else umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5 and, two groups, 6..9)
    const int g = warp & 3;             // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int row = g * 32 + lane;      // accumulator row == pixel within the tile
    const int grp = (warp - 2) >> 2;    // epilogue warpgroup: chunks grp, grp + kEpiGroups, ... of every tile
    const int et = ((warp - 2) & 3) * 32 + lane;  // 0..127 inside the group
    const uint32_t bar_id = 1u + static_cast<uint32_t>(grp);
    int tile_iter = 0;
    int store_buf = 0;
    for (int item = item_first; item < num_items; item += item_step, ++tile_iter) {
      int m_tile, n_tile, k_slice;
      const bool tile_live = decode_item(item, m_tile, n_tile, k_slice);
      if (!tile_live) {  // shadow tile of an odd tail: keep the TMEM handshake, store nothing
        const int as_ = tile_iter & 1;
        mbar_wait(&tmem_full[as_], (tile_iter >> 1) & 1);
        tc_fence_after();
        tc_fence_before();
        if (kCluster) mbar_arrive_cluster(mapa_u32(&tmem_empty[as_], 0));
        else mbar_arrive(&tmem_empty[as_]);
        continue;
      }
      const int tiles_per_img = p.tiles_h * p.tiles_w;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.bh;
      const int w0 = (rem % p.tiles_w) * p.bw;
      const int n0 = n_tile * BLOCK_N;
      const int hi = row / p.bw;
      const int wi = row - hi * p.bw;
      const bool row_valid = (row < p.bh * p.bw) && (h0 + hi < p.H) && (w0 + wi < p.W);
      const long long pix = (static_cast<long long>(img) * p.H + (h0 + hi)) * p.W + (w0 + wi);
      const uint32_t row_msk = __ballot_sync(0xffffffffu, row_valid);  // valid rows of this warp's 32-row group
      const int as = tile_iter & 1;
      const uint32_t apar = (tile_iter >> 1) & 1;
      mbar_wait(&tmem_full[as], apar);
      tc_fence_after();

      constexpr int kChunks = BLOCK_N / 64;
#pragma unroll 1
      for (int ch = grp; ch < kChunks; ch += kEpiGroups) {
        const int c0 = n0 + ch * 64;  // first output channel of this chunk
        if (c0 >= p.Cout) break;      // (uniform) nothing to write for padded columns
        // Prefetch this CTA's running statistics for the chunk's columns now; the read-modify-write below then
        // does not expose the global-memory latency in the epilogue's critical path.
        const bool do_stats = p.stats_partial != nullptr && p.epi_mode == SEMSEG_EPI_RAW;
        float st_old[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float* st_dst = nullptr;
        if (do_stats) {  // every epilogue warp owns statistics row (blockIdx.x*4 + g); lane = column pair
          st_dst = p.stats_partial + (static_cast<size_t>(blockIdx.x) * 4 + g) * 3 * p.Cout + c0 + 2 * lane;
          st_old[0] = st_dst[0];
          st_old[1] = st_dst[1];
          st_old[2] = st_dst[p.Cout];
          st_old[3] = st_dst[p.Cout + 1];
          st_old[4] = st_dst[2 * p.Cout];
          st_old[5] = st_dst[2 * p.Cout + 1];
        }
        // Residual operand (AFFINE mode): issue the row's eight 16-byte loads before the TMEM loads so their latency
        // overlaps with tcgen05.ld instead of sitting in front of the first use.
        uint4 rres[8];
        const bool has_res = p.epi_mode == SEMSEG_EPI_AFFINE && p.residual != nullptr && row_valid;
        if (has_res && !kSplit) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.res_pitch + c0);
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) rres[j8] = rp[j8];
        }
        uint32_t v[2][32];
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(g * 32) << 16) + static_cast<uint32_t>(as * BLOCK_N + ch * 64);
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_32x32(taddr + 32, v[1]);
        tmem_ld_wait();

        if (p.epi_mode == SEMSEG_EPI_F32) {
          if (row_valid) {
            float* orow = p.out_f32 + k_slice * p.slice_stride + pix * p.out_pitch;
            const bool bias = p.shift != nullptr && k_slice == 0;
            if (c0 + 64 <= p.Cout && (p.out_pitch & 3) == 0) {   // whole chunk inside: 16-byte stores
#pragma unroll
              for (int j4 = 0; j4 < 16; ++j4) {
                float4 o;
                o.x = __uint_as_float(v[j4 >> 3][(4 * j4 + 0) & 31]);
                o.y = __uint_as_float(v[j4 >> 3][(4 * j4 + 1) & 31]);
                o.z = __uint_as_float(v[j4 >> 3][(4 * j4 + 2) & 31]);
                o.w = __uint_as_float(v[j4 >> 3][(4 * j4 + 3) & 31]);
                if (bias) {
                  o.x += __ldg(p.shift + c0 + 4 * j4);
                  o.y += __ldg(p.shift + c0 + 4 * j4 + 1);
                  o.z += __ldg(p.shift + c0 + 4 * j4 + 2);
                  o.w += __ldg(p.shift + c0 + 4 * j4 + 3);
                }
                *reinterpret_cast<float4*>(orow + c0 + 4 * j4) = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const int c = c0 + j;
                if (c < p.Cout) {
                  float a = __uint_as_float(v[j >> 5][j & 31]);
                  if (bias) a += __ldg(p.shift + c);
                  orow[c] = a;
                }
              }
            }
          }
          continue;
        }

        if (p.epi_mode == SEMSEG_EPI_AFFINE) {
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) {
            float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (has_res) {
              if constexpr (kSplit) {   // hi + lo planes, loaded here (no prefetch: register budget)
                const long long ro = pix * p.res_pitch + c0 + j8 * 8;
                const uint4 rh = *reinterpret_cast<const uint4*>(p.residual + ro);
                const uint4 rl = *reinterpret_cast<const uint4*>(p.residual_lo + ro);
                const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&rh);
                const __nv_bfloat162* pl = reinterpret_cast<const __nv_bfloat162*>(&rl);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 fh = __bfloat1622float2(ph[q]);
                  const float2 fl = __bfloat1622float2(pl[q]);
                  r[2 * q] = fh.x + fl.x;
                  r[2 * q + 1] = fh.y + fl.y;
                }
              } else {
                const uint4 rv = rres[j8];
                const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = __bfloat1622float2(rp[q]);
                  r[2 * q] = f.x;
                  r[2 * q + 1] = f.y;
                }
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int j = j8 * 8 + q;
              const int c = c0 + j;
              float a = __uint_as_float(v[j >> 5][j & 31]);
              const float sc = p.scale ? __ldg(p.scale + c) : 1.f;
              const float sh = p.shift ? __ldg(p.shift + c) : 0.f;
              a = fmaf(a, sc, sh) + r[q];
              if (p.relu) a = fmaxf(a, 0.f);
              v[j >> 5][j & 31] = __float_as_uint(a);
            }
          }
        }

        // registers -> bf16 -> 128B-swizzled staging tile (row = pixel, 64 channels = 128 bytes)
        // one group: two staging buffers used alternately; two groups: one buffer each
        // split storage: buffer 0 = hi plane tile, buffer 1 = lo plane tile
        uint8_t* obuf = out_stage + (kSplit ? 0 : (kEpiGroups == 1 ? store_buf : grp)) * kStageOutBytes;
        if (et == 0) tma_store_wait_read<((kEpiGroups == 1 && !kSplit) ? 1 : 0)>();  // the store(s) that last read the buffer(s) drained
        named_bar_sync(bar_id, kEpiGroupThreads);
#pragma unroll
        for (int j8 = 0; j8 < 8; ++j8) {
          uint4 o;
          const int b = j8 * 8;
          o.x = pack_bf16x2(__uint_as_float(v[b >> 5][(b + 0) & 31]), __uint_as_float(v[b >> 5][(b + 1) & 31]));
          o.y = pack_bf16x2(__uint_as_float(v[b >> 5][(b + 2) & 31]), __uint_as_float(v[b >> 5][(b + 3) & 31]));
          o.z = pack_bf16x2(__uint_as_float(v[b >> 5][(b + 4) & 31]), __uint_as_float(v[b >> 5][(b + 5) & 31]));
          o.w = pack_bf16x2(__uint_as_float(v[b >> 5][(b + 6) & 31]), __uint_as_float(v[b >> 5][(b + 7) & 31]));
          *reinterpret_cast<uint4*>(obuf + row * 128 + ((j8 ^ (row & 7)) << 4)) = o;
          if constexpr (kSplit) {   // lo = value - float(hi), staged in the second buffer
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&o);
            float l[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 hf = __bfloat1622float2(hp[q]);
              l[2 * q] = __uint_as_float(v[b >> 5][(b + 2 * q) & 31]) - hf.x;
              l[2 * q + 1] = __uint_as_float(v[b >> 5][(b + 2 * q + 1) & 31]) - hf.y;
            }
            uint4 ol;
            ol.x = pack_bf16x2(l[0], l[1]);
            ol.y = pack_bf16x2(l[2], l[3]);
            ol.z = pack_bf16x2(l[4], l[5]);
            ol.w = pack_bf16x2(l[6], l[7]);
            *reinterpret_cast<uint4*>(obuf + kStageOutBytes + row * 128 + ((j8 ^ (row & 7)) << 4)) = ol;
          }
        }
        fence_proxy_async_smem();
        if (do_stats) {
          // Per-column statistics of the bf16 values just staged (exactly what BN-apply will read back). Each warp
          // reduces the 32 rows it wrote itself (only a __syncwarp away), lane = column pair (one 4-byte word,
          // conflict-free), one pass of sum and sum of squares, then a read-modify-write of the warp's own running
          // (sum, sum of squares, count) row in global memory: no cross-warp traffic, fixed order -> deterministic.
          __syncwarp();
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint8_t* base = obuf + (lane & 3) * 4;
          const int chunk16 = lane >> 2;
          auto add_row = [&](int rr, int sw) {  // sw = rr & 7 (the row's swizzle phase)
            const __nv_bfloat162 bv =
                *reinterpret_cast<const __nv_bfloat162*>(base + rr * 128 + ((chunk16 ^ sw) << 4));
            float2 f = __bfloat1622float2(bv);
            if constexpr (kSplit) {
              const float2 fl = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(
                  base + kStageOutBytes + rr * 128 + ((chunk16 ^ sw) << 4)));
              f.x += fl.x;
              f.y += fl.y;
            }
            s0 += f.x;
            s1 += f.y;
            q0 = fmaf(f.x, f.x, q0);
            q1 = fmaf(f.y, f.y, q1);
          };
          if ((row_msk & (row_msk + 1u)) == 0u) {
            // valid rows are a prefix of the warp's 32 (always, unless the box is clipped by the right image edge):
            // whole groups of 8 rows run without per-row predicates, the swizzle phase is a compile-time constant
            const int nr = __popc(row_msk);
            const int nfull = nr >> 3;
            for (int b8 = 0; b8 < nfull; ++b8) {
              const int rr0 = g * 32 + b8 * 8;
#pragma unroll
              for (int k = 0; k < 8; ++k) add_row(rr0 + k, k);
            }
            for (int r = nfull * 8; r < nr; ++r) add_row(g * 32 + r, r & 7);
          } else {
#pragma unroll
            for (int r = 0; r < 32; ++r)
              if ((row_msk >> r) & 1u) add_row(g * 32 + r, r & 7);
          }
          const float nt = static_cast<float>(__popc(row_msk));
          st_dst[0] = st_old[0] + s0;
          st_dst[1] = st_old[1] + s1;
          st_dst[p.Cout] = st_old[2] + q0;
          st_dst[p.Cout + 1] = st_old[3] + q1;
          st_dst[2 * p.Cout] = st_old[4] + nt;
          st_dst[2 * p.Cout + 1] = st_old[5] + nt;
        }
        named_bar_sync(bar_id, kEpiGroupThreads);
        if (et == 0) {
          tma_store_4d(&tmC, obuf, c0, w0, h0, img);
          if (kSplit) tma_store_4d(&tmC_lo, obuf + kStageOutBytes, c0, w0, h0, img);
          tma_store_commit();
        }
        store_buf ^= 1;
      }
      // all TMEM reads of this accumulator stage are done -> hand it back to the (leader's) MMA warp
      tc_fence_before();
      if (kCluster) mbar_arrive_cluster(mapa_u32(&tmem_empty[as], 0));
      else mbar_arrive(&tmem_empty[as]);
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster) cluster_sync_all();  // the peer may still signal my barriers until it is done as well
  if (warp == 1) {
    tc_fence_after();
    if (kCluster) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

static bool cluster_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEMSEG_B200_CLUSTER");
    v = (e && e[0] == '0') ? 0 : 1;  // CTA-pair mode by default (validated on B200); SEMSEG_B200_CLUSTER=0 disables
  }
  return v != 0;
}

// Grid (= rows of the statistics buffer) and whether the clustered variant is used.
static int conv_grid(int num_m_tiles, int n_tiles, bool* clustered) {
  const int sms = num_sms();
  const bool cl = cluster_enabled() && num_m_tiles >= 2;
  *clustered = cl;
  if (!cl) {
    const long long tiles = static_cast<long long>(num_m_tiles) * n_tiles;
    return static_cast<int>(tiles < sms ? tiles : sms);
  }
  const long long items = static_cast<long long>((num_m_tiles + 1) / 2) * n_tiles;  // one per cluster
  const long long max_clusters = sms / 2;
  return static_cast<int>(2 * (items < max_clusters ? items : max_clusters));
}

// Two epilogue warpgroups by default for tiles with >= 2 column chunks (SEMSEG_B200_EPI_GROUPS=1 keeps one).
static int epi_groups_wanted() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEMSEG_B200_EPI_GROUPS");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

struct ConvMaps {
  CUtensorMap a, b, c, a_lo, b_lo, c_lo;
};

template <int BLOCK_N, int kEpiGroups, bool kSplit>
static int launch_conv_g(const ConvMaps& tm, const ConvKParams& kp, bool clustered, int grid, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N>;
  // the opt-in to > 48 KB of dynamic shared memory is per device: set it once for every device this process uses
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
    SB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, false, kEpiGroups, kSplit>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    SB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, true, kEpiGroups, kSplit>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (dev >= 0 && dev < 64) attr_set[dev].store(true, std::memory_order_release);
  }
  constexpr int kThreads = conv_threads(kEpiGroups);
  if (!clustered) {
    conv_igemm_kernel<BLOCK_N, false, kEpiGroups, kSplit><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(
        tm.a, tm.b, tm.c, tm.a_lo, tm.b_lo, tm.c_lo, kp);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SB_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BLOCK_N, true, kEpiGroups, kSplit>, tm.a, tm.b, tm.c, tm.a_lo,
                               tm.b_lo, tm.c_lo, kp));
  }
  SB_LAUNCHED();
  return SEMSEG_OK;
}

template <int BLOCK_N>
static int launch_conv(const ConvMaps& tm, const ConvKParams& kp, bool split, bool clustered, int grid,
                       cudaStream_t stream) {
  if (split) return launch_conv_g<BLOCK_N, 1, true>(tm, kp, clustered, grid, stream);
  if constexpr (BLOCK_N >= 128) {
    if (epi_groups_wanted() == 2) return launch_conv_g<BLOCK_N, 2, false>(tm, kp, clustered, grid, stream);
  }
  return launch_conv_g<BLOCK_N, 1, false>(tm, kp, clustered, grid, stream);
}

}  // namespace sb

static int conv_block_n(int Cout) {
  if (Cout % 256 == 0 || Cout > 128) return 256;
  return Cout > 64 ? 128 : 64;
}

// Rows of the RAW-epilogue statistics buffer = number of CTAs the kernel will launch for this problem.
// K blocks (64-channel block x tap) of a conv, and the slice count that keeps one accumulation chain <= max_kblocks.
extern "C" int semseg_conv_k_slices(int Cin, int taps, int max_kblocks) {
  if (Cin <= 0 || taps <= 0 || max_kblocks <= 0) return SEMSEG_E_INVALID;
  const int total = taps * sb::cdiv(Cin, sb::kBlockK);
  const int per = sb::cdiv(total, sb::cdiv(total, max_kblocks));
  return sb::cdiv(total, per);
}

extern "C" int semseg_conv_stats_rows(int N, int H, int W, int Cout) {
  if (N <= 0 || H <= 0 || W <= 0 || Cout <= 0) return SEMSEG_E_INVALID;
  int bh, bw;
  sb::choose_box(H, W, sb::kBlockM, &bh, &bw);
  bool clustered;  // four statistics rows per CTA (one per epilogue warp)
  return 4 * sb::conv_grid(N * sb::cdiv(H, bh) * sb::cdiv(W, bw), sb::cdiv(Cout, conv_block_n(Cout)), &clustered);
}

extern "C" int semseg_conv_fprop(const semseg_conv_desc* d, void* stream_) {
  using namespace sb;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(d != nullptr, "conv: null descriptor");
  SB_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: bad sizes");
  SB_CHECK_ARG(d->taps >= 1 && d->taps <= SEMSEG_MAX_TAPS, "conv: taps=%d out of range", d->taps);
  SB_CHECK_ARG(d->x && d->w, "conv: null x/w");
  SB_CHECK_ARG(d->x_pitch % 8 == 0 && d->x_pitch >= d->Cin, "conv: x_pitch %d must be a multiple of 8 and >= Cin",
               d->x_pitch);
  SB_CHECK_ARG(d->w_cols % 8 == 0 && d->w_cols >= d->Cin, "conv: w_cols %d must be a multiple of 8 and >= Cin",
               d->w_cols);
  SB_CHECK_ARG(d->epi_mode >= 0 && d->epi_mode <= 2, "conv: bad epi_mode");

  ConvKParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.N = d->N; kp.H = d->H; kp.W = d->W; kp.Cin = d->Cin; kp.Cout = d->Cout; kp.taps = d->taps;
  choose_box(d->H, d->W, kBlockM, &kp.bh, &kp.bw);
  kp.tiles_h = cdiv(d->H, kp.bh);
  kp.tiles_w = cdiv(d->W, kp.bw);
  kp.num_m_tiles = d->N * kp.tiles_h * kp.tiles_w;
  kp.k_chunks = cdiv(d->Cin, kBlockK);
  for (int t = 0; t < d->taps; ++t) {
    kp.dh[t] = d->dh[t]; kp.dw[t] = d->dw[t]; kp.wtap[t] = d->wtap[t]; kp.img_add[t] = d->img_add[t];
    SB_CHECK_ARG(d->wtap[t] >= 0 && d->wtap[t] < d->n_wtaps, "conv: wtap[%d]=%d out of range", t, d->wtap[t]);
  }
  kp.img_mul = d->img_mul;
  kp.epi_mode = d->epi_mode; kp.relu = d->relu;
  kp.scale = d->scale; kp.shift = d->shift;
  kp.residual = static_cast<const __nv_bfloat16*>(d->residual); kp.res_pitch = d->res_pitch;
  kp.residual_lo = static_cast<const __nv_bfloat16*>(d->residual_lo);
  const bool split = d->x_lo != nullptr;   // bf16x3 operand mode: hi/lo planes for x, w, y, residual
  kp.nseg = split ? 3 : 1;
  if (split) {
    SB_CHECK_ARG(d->w_split != 0, "conv: split activations need split weights (w_split)");
    SB_CHECK_ARG(d->epi_mode == SEMSEG_EPI_F32 || d->y_lo != nullptr, "conv: split input needs a split output (y_lo)");
    SB_CHECK_ARG(!d->residual || d->residual_lo, "conv: split input needs a split residual (residual_lo)");
  } else {
    SB_CHECK_ARG(!d->y_lo && !d->residual_lo, "conv: lo planes given without x_lo");
  }
  kp.out_f32 = d->out_f32; kp.out_pitch = d->out_pitch;
  kp.stats_partial = d->stats_partial;

  const int block_n = conv_block_n(d->Cout);
  kp.n_tiles = cdiv(d->Cout, block_n);
  kp.k_slices = 1;
  kp.kb_per_slice = kp.taps * kp.k_chunks;
  kp.slice_stride = 0;
  if (d->k_slices > 1) {
    SB_CHECK_ARG(d->epi_mode == SEMSEG_EPI_F32, "conv: K slicing needs the F32 epilogue (partials are fp32)");
    SB_CHECK_ARG(d->slice_stride >= static_cast<long long>(d->N) * d->H * d->W * d->out_pitch,
                 "conv: slice_stride too small for an [N,H,W,out_pitch] partial");
    kp.kb_per_slice = cdiv(kp.taps * kp.k_chunks, d->k_slices);
    kp.k_slices = cdiv(kp.taps * kp.k_chunks, kp.kb_per_slice);   // no empty slices
    SB_CHECK_ARG(kp.k_slices == d->k_slices, "conv: k_slices %d not realisable for %d K blocks (use %d)", d->k_slices,
                 kp.taps * kp.k_chunks, kp.k_slices);
    kp.slice_stride = d->slice_stride;
  }
  bool clustered = false;
  const int grid = conv_grid(kp.num_m_tiles, kp.n_tiles * kp.k_slices, &clustered);

  // A: input activations [Nin][Hin][Win][x_pitch] viewed as (C, W, H, N)
  ConvMaps tm;
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Win, (uint64_t)d->Hin, (uint64_t)d->Nin};
    uint64_t str[3] = {(uint64_t)d->x_pitch * 2, (uint64_t)d->x_pitch * 2 * d->Win,
                       (uint64_t)d->x_pitch * 2 * d->Win * d->Hin};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)kp.bw, (uint32_t)kp.bh, 1};
    int r = encode_tmap_bf16(&tm.a, d->x, 4, dims, str, box);
    if (r) return r;
    tm.a_lo = tm.a;
    if (split && (r = encode_tmap_bf16(&tm.a_lo, d->x_lo, 4, dims, str, box))) return r;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->w_cols, (uint64_t)d->w_rows, (uint64_t)d->n_wtaps};
    uint64_t str[2] = {(uint64_t)d->w_cols * 2, (uint64_t)d->w_cols * 2 * d->w_rows};
    // clustered: each CTA loads (and multicasts) half of the weight rows of the tile
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)(clustered ? block_n / 2 : block_n), 1};
    int r = encode_tmap_bf16(&tm.b, d->w, 3, dims, str, box);
    if (r) return r;
    tm.b_lo = tm.b;
    if (split) {   // the lo slab follows the hi slab (semseg_pack_weights with split != 0)
      const __nv_bfloat16* w_lo =
          static_cast<const __nv_bfloat16*>(d->w) + static_cast<size_t>(d->n_wtaps) * d->w_rows * d->w_cols;
      if ((r = encode_tmap_bf16(&tm.b_lo, w_lo, 3, dims, str, box))) return r;
    }
  }
  if (d->epi_mode == SEMSEG_EPI_F32) {
    SB_CHECK_ARG(d->out_f32 != nullptr && d->out_pitch >= d->Cout, "conv: F32 epilogue needs out_f32/out_pitch");
    tm.c = tm.a;  // unused
    tm.c_lo = tm.a;
  } else {
    SB_CHECK_ARG(d->y != nullptr, "conv: null y");
    SB_CHECK_ARG(d->Cout % 64 == 0, "conv: bf16 epilogue needs Cout %% 64 == 0 (got %d)", d->Cout);
    SB_CHECK_ARG(d->y_pitch % 8 == 0 && d->y_pitch >= d->Cout, "conv: bad y_pitch %d", d->y_pitch);
    if (d->residual) SB_CHECK_ARG(d->res_pitch % 8 == 0, "conv: res_pitch must be a multiple of 8");
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)d->y_pitch * 2, (uint64_t)d->y_pitch * 2 * d->W,
                       (uint64_t)d->y_pitch * 2 * d->W * d->H};
    uint32_t box[4] = {64u, (uint32_t)kp.bw, (uint32_t)kp.bh, 1};
    int r = encode_tmap_bf16(&tm.c, d->y, 4, dims, str, box);
    if (r) return r;
    tm.c_lo = tm.c;
    if (split && (r = encode_tmap_bf16(&tm.c_lo, d->y_lo, 4, dims, str, box))) return r;
  }
  switch (block_n) {
    case 256: return launch_conv<256>(tm, kp, split, clustered, grid, stream);
    case 128: return launch_conv<128>(tm, kp, split, clustered, grid, stream);
    default: return launch_conv<64>(tm, kp, split, clustered, grid, stream);
  }
}
