// Fused pseudo-self-attention step for greedy decoding: c_attn GEMM + KV-cache append + single-query attention in ONE
// kernel, with no cross-CTA dependency (language_model.py:124-180 for a 1-token query; replaces K17 + K18 + K19 of
// SURVEY.md §2c).
//
// Tiling is head-aligned: CTA (m, h) computes the 128 x 192 tile [q_h | k_h | v_h] of M tile m — the three 64-row groups
// of the c_attn weight that belong to head h are fetched by one 3-D TMA box into one 192-row operand B — so everything
// the attention of (row, head h) needs is produced by the CTA that consumes it:
//   1. tcgen05 main loop (K = 1024: 16 k-blocks, accumulator 128 x 192 fp32 in TMEM), fed by TMA through a 4-stage ring;
//   2. epilogue: + bias -> bf16; q / k_new / v_new -> swizzled shared-memory tiles (k_new, v_new are appended in place into
//      the KV cache at slot t+1 by the attention warps, one 256-byte store per item; the reference regrows the cache with
//      torch.cat, language_model.py:169-170);
//   3. attention: AW warps, one (row, head) item per warp at a time; the cached keys / values of an item (slots 0..t: ONE
//      contiguous block of L x 256 B, key row and value row of a slot adjacent) are streamed in 16-key chunks, one
//      cp.async.bulk per chunk, into a per-warp ring of NSLOT shared-memory slots (mbarrier complete_tx); the new key /
//      value comes from the tiles of step 2.  One copy per chunk matters: the timeline (tools/decode_timeline.py) showed
//      the phase taking 44 ns per bulk copy whatever its size (1 - 2 KB) at every cache length — bound by the SM's copy
//      issue rate, not by HBM — when keys and values were separate blocks and needed two copies per chunk.  Optional:
//      cp.async.bulk.prefetch.L2 of the next items' blocks, issued before the main loop finishes, so HBM streams while
//      the tensor pipe runs.
// Arithmetic and reduction order are those of dec::attention_dev (decoder_kernels.cuh): 16-key chunks, 4 key subgroups x
// 8 dim segments per warp, online softmax in fp32, q / k / v rounded to bf16 exactly where the two-kernel path rounds
// them — the fused and the two-kernel path produce bit-identical attention outputs.
//
// Optional LayerNorm head (the 16 CTAs = 16 heads of one M tile): instead of a separate LayerNorm kernel, each of the 16
// CTAs reduces the preceding projection's split-K partial sums into the residual stream and normalises 8 rows of the M
// tile, the 16 meet at a per-M-tile arrival counter (group_barrier), and the TMA loads of operand A follow.
#pragma once
#include "decoder_kernels.cuh"
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace rgrg {
namespace fa {

constexpr int BN = 192;      // q | k | v columns of one head
#ifndef RGRG_FA_STAGES
#define RGRG_FA_STAGES 4
#endif
// TMA ring depth of the c_attn phase (16 k-blocks).  Measured step at 928 rows: 3 stages 2.013 ms, 4 stages 1.985, 5 stages 1.990;
// 4 x 40 KB still fits under the attention phase's footprint (48 KB of tiles + 128 KB of staging rings), which aliases it.
constexpr int STAGES = RGRG_FA_STAGES;
constexpr int A_BYTES = tc::BM * tc::BK * 2;   // 16 KB
constexpr int B_BYTES = BN * tc::BK * 2;       // 24 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int K_ITERS = 1024 / tc::BK;
constexpr int TILE_BYTES = tc::BM * 128;       // one 128 x 64 bf16 tile (q, k_new or v_new)
constexpr int CHUNK_KEYS = 16;
constexpr int SLOT_BYTES = 2 * CHUNK_KEYS * 128;  // K chunk + V chunk
constexpr int TMEM_COLS = 256;

// AW = attention / epilogue warps per CTA (a multiple of 4: AW / 4 warps share a TMEM lane quarter); block = 64 + 32 AW threads.
// The per-chunk work of a warp is a chain of dependent shared-memory loads, shuffles and exponentials, so the attention
// phase scales with the number of resident warps until HBM saturates (measured: 8 warps -> 45 us per layer at 928 rows).
// ALG 1 (the only one left): 16-key chunks, arithmetic order of dec::attention_dev (bit-identical to the two-kernel path).
//        A lane-per-key variant (ALG 2: K staged, V rows read straight from global memory) was measured slower
//        (profiles/r02_decode.md) and removed when the cache layout became slot-interleaved.
template <int AW, int NSLOT, int ALG = 1>
struct Smem {
  static constexpr int ATT_WARPS = AW;
  static constexpr int GEMM_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STG_OFFSET = 3 * TILE_BYTES;  // attention phase: tiles at [0, 48 KB), staging rings after them
  static constexpr int ATT_BYTES = STG_OFFSET + ATT_WARPS * NSLOT * SLOT_BYTES;
  static constexpr int BAR_OFFSET = GEMM_BYTES > ATT_BYTES ? GEMM_BYTES : ATT_BYTES;
  static constexpr int NBARS = 2 * STAGES + 1 + ATT_WARPS * NSLOT;
  static constexpr int TOTAL = BAR_OFFSET + NBARS * 8 + 16 + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
  static_assert(TOTAL > 116 * 1024, "one CTA per SM");
};

struct Params {
  const float* bias;     // c_attn bias [3072]
  KvGeom kv;
  int layer;
  const int* step_ptr;   // device-side decode step t (word t is cached at slot t + 1)
  bf16* attn_o;          // [M, 1024]
  int M;
  int rows_per_tile;     // rows OWNED by one M tile (<= 128; the MMA still covers 128 rows, the surplus belongs to the next tile):
                         // rows are spread evenly over as many tiles as there are SMs / 16, so every CTA streams the same amount
  int mc;                // 1: head pairs share operand A by TMA multicast (cluster of 2; needs the 64-row map)
  int early_kv;          // 1: the first K / V chunks are requested before the epilogue instead of after it
  int l2_ahead;          // items whose K / V blocks are prefetched into L2 ahead of the consumer (0 = off; measured: no gain —
                         // neither this nor prefetching the NEXT layer's cache during the GEMM kernels, profiles/r02_decode.md)
  // optional LayerNorm head (null h = off; needs a 16-CTA cluster launch)
  float* h;              // fp32 residual stream [M, 1024]
  bf16* x;               // normalised rows (the GEMM's own operand A)
  const float* gamma;
  const float* beta;
  const float* parts;    // split-K partial sums of the preceding projection (null: plain LayerNorm of h)
  size_t part_stride;
  const float* res_bias;
  long long* trace;      // tuning: [2][8] timestamps of the first / last CTA (entry, setup done, predecessor done, tiles written,
                         // first K / V chunk landed, accumulator ready, attention done, exit)
  unsigned* counters;    // [m_tiles] arrival counters of the LayerNorm-head group barrier (common.cuh)
  int launch_idx, launches_per_step;
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
// c_attn weight viewed as [3 (q | k | v)][1024 rows][1024 K]: the three 64-row groups of one head in ONE request
inline CUtensorMap make_tmap_qkv(const void* w) {
  CUtensorMap m;
  cuuint64_t dims[3] = {1024, 1024, 3};
  cuuint64_t strides[2] = {1024 * 2, 1024ull * 1024 * 2};
  cuuint32_t box[3] = {tc::BK, 64, 3};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = tc::encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(qkv) failed: " + std::to_string((int)r));
  return m;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// A-operand sharing between the two heads of a CTA pair (Params::mc): each CTA fetches 64 of the 128 rows and multicasts
// them into both CTAs' rings; the bytes are credited to the full barrier at the same offset in both CTAs
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          tc::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// "these MMAs have read their operands": arrives on the barrier at this offset in every CTA of the mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int AW, int NSLOT, bool LN_HEAD, int ALG>
__global__ void __launch_bounds__(64 + 32 * AW, 1) attn_fused_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                        const __grid_constant__ CUtensorMap tmA64,
                                                                        const __grid_constant__ CUtensorMap tmW, const Params p) {
  using L = Smem<AW, NSLOT, ALG>;
  constexpr int ATT_WARPS = AW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* att_bar = tmem_full_bar + 1;  // [ATT_WARPS][NSLOT]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(att_bar + ATT_WARPS * NSLOT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) trace_mark(p.trace, 0);
  const int head = blockIdx.x & 15, m_blk = blockIdx.x >> 4;  // the 16 heads of an M tile are consecutive CTAs (one cluster)
  // mc: CTAs (2i, 2i + 1) — two heads of the same M tile — form a cluster and share operand A through TMA multicast
  const bool mc = !LN_HEAD && p.mc;
  const uint32_t rank = mc ? cluster_ctarank() : 0;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmW);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(&full_bar[i], 1);
      tc::mbar_init(&empty_bar[i], mc ? 2 : 1);  // mc: a ring slot is written by both CTAs, so both MMA issuers release it
    }
    tc::mbar_init(tmem_full_bar, 1);
    for (int i = 0; i < ATT_WARPS * NSLOT; ++i) tc::mbar_init(&att_bar[i], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (mc) {  // the peer's barriers must be initialised before anything of ours can signal them
    cluster_arrive();
    cluster_wait();
  }
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_launch_dependents();  // dependents may be scheduled; they block at their own griddep_wait until this grid completes
  if (threadIdx.x == 0) trace_mark(p.trace, 1);

  // the weight tiles never depend on the predecessor kernel: start streaming them before waiting for it
  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_expect_tx(&full_bar[i], STAGE_BYTES);
      tma_load_3d(smem + i * STAGE_BYTES + A_BYTES, &tmW, &full_bar[i], i * tc::BK, head * 64, 0);
    }
  }
  griddep_wait();
  if (threadIdx.x == 0) trace_mark(p.trace, 2);

  if constexpr (LN_HEAD) {
    // rows [m_blk*128 + head*8, +8) of the M tile: h += bias + split-K partial sums of the preceding projection,
    // x = LayerNorm(h); one row per attention warp.  Peers' rows become visible at the cluster barrier; operand A is read
    // through TMA (async proxy), hence the proxy fences on both sides.
    if (warp >= 2 && warp < 10) {
      const int row = m_blk * tc::BM + head * 8 + (warp - 2);
      if (row < p.M) {
        if (p.parts) dec::ln_row_dev<4>(p.h, p.gamma, p.beta, p.x, row, lane, p.parts, p.part_stride, p.res_bias);
        else dec::ln_row_dev<0>(p.h, p.gamma, p.beta, p.x, row, lane, nullptr, 0, nullptr);
      }
    }
    group_barrier(p.counters + m_blk, static_cast<unsigned>(*p.step_ptr * p.launches_per_step + p.launch_idx + 1) * 16u);
  }

  const int t = *p.step_ptr;
  const int Lc = t + 1;  // cached keys: slots 0..t (slot 0 = image key); the word of this step goes to slot t + 1
  const int Ltot = Lc + 1;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < K_ITERS; ++kb) {
        const int st = kb % STAGES;
        uint8_t* a_dst = smem + st * STAGE_BYTES;
        if (kb >= STAGES) {
          tc::mbar_wait(&empty_bar[st], ((kb / STAGES) & 1) ^ 1);
          tc::mbar_expect_tx(&full_bar[st], STAGE_BYTES);
          tma_load_3d(a_dst + A_BYTES, &tmW, &full_bar[st], kb * tc::BK, head * 64, 0);
        }
        if (mc)  // rows [64 rank, +64) of the tile, into both CTAs' slots
          tma_load_2d_mc(a_dst + rank * (A_BYTES / 2), &tmA64, &full_bar[st], kb * tc::BK, m_blk * p.rows_per_tile + static_cast<int>(rank) * 64, 3);
        else
          tc::tma_load_2d(a_dst, &tmA, &full_bar[st], kb * tc::BK, m_blk * p.rows_per_tile);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(tc::BM, BN);
      for (int kb = 0; kb < K_ITERS; ++kb) {
        const int st = kb % STAGES;
        tc::mbar_wait(&full_bar[st], (kb / STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t a_addr = tc::smem_u32(smem + st * STAGE_BYTES);
        const uint64_t a_desc = tc::make_sw128_kmajor_desc(a_addr);
        const uint64_t b_desc = tc::make_sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < tc::BK / tc::UMMA_K; ++k) tc::umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        if (mc) umma_commit_mc(&empty_bar[st], 3);
        else tc::umma_commit(&empty_bar[st]);
      }
      tc::umma_commit(tmem_full_bar);
    }
  } else {
    // ---------------------------------------------------------------- epilogue + attention (warps 2..9)
    const int aw = warp - 2;
    const int row0 = m_blk * p.rows_per_tile;
    const int rows_left = min(p.rows_per_tile, p.M - row0);  // rows this M tile owns (>= 1)
    // items of this warp: local rows aw, aw + AW, ... (a short last M tile still spreads over all warps)
    const int n_items = rows_left > aw ? (rows_left - aw + AW - 1) / AW : 0;
    const int nsc = (Lc + CHUNK_KEYS - 1) / CHUNK_KEYS;       // staged chunks per item (cached keys)
    const int nchunks = (Ltot + CHUNK_KEYS - 1) / CHUNK_KEYS;  // compute chunks per item (cached + new key)
    const uint32_t blk_bytes = static_cast<uint32_t>(Lc) * 256;  // the item's cached keys and values, one contiguous block
    auto item_row = [&](int j) { return row0 + aw + AW * j; };
    if (p.l2_ahead > 0 && lane == 0) {
      for (int j = 0; j < p.l2_ahead && j < n_items; ++j) {
        bulk_prefetch_l2(p.kv.cache + p.kv.offset(p.layer, 0, item_row(j), head, 0), blk_bytes);
      }
    }

    // staging ring of this warp: NSLOT slots of 16 slots x (key row + value row)
    uint8_t* stg = smem + L::STG_OFFSET + aw * NSLOT * SLOT_BYTES;
    uint64_t* bars = att_bar + aw * NSLOT;
    const int total = n_items * nsc;
    auto issue = [&](int seq) {
      if (lane == 0) {
        const int j = seq / nsc, c = seq - j * nsc;
        const int keys = min(CHUNK_KEYS, Lc - c * CHUNK_KEYS);
        const uint32_t bytes = static_cast<uint32_t>(keys) * 256;  // key and value rows of a slot are adjacent: ONE copy per chunk
        const int s = seq % NSLOT;
        tc::mbar_expect_tx(&bars[s], bytes);
        bulk_g2s(stg + s * SLOT_BYTES, p.kv.cache + p.kv.offset(p.layer, 0, item_row(j), head, c * CHUNK_KEYS), bytes, &bars[s]);
      }
    };

    // ---- epilogue: TMEM -> (+bias, bf16) -> q / k_new / v_new tiles (+ KV-cache append)
    tc::mbar_wait(tmem_full_bar, 0);
    tc::tc_fence_after();
    if (warp == 2 && lane == 0) trace_mark(p.trace, 5);
    // Every MMA has completed, so the operand ring is idle; the tiles at [0, 48 KB) and the staging rings behind them do not
    // overlap: with early_kv the first cached chunks (written by earlier steps) start streaming under the epilogue.
    if (p.early_kv)
      for (int s = 0; s < NSLOT && s < total; ++s) issue(s);
    {
      const int q4 = warp & 3;           // TMEM lane quarter this warp may read
      const int grp = aw >> 2;           // which of the AW / 4 warps of that quarter: takes 16-column chunks grp, grp + AW/4, ...
      const int r = q4 * 32 + lane;      // local row
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
#pragma unroll 1
      for (int c = grp * 16; c < BN; c += 4 * AW) {
        uint32_t v[16];
        tc::tmem_ld_32x32b_x16(t_addr + c, v);
        tc::tmem_ld_wait();
        const int which = c >> 6, cc = c & 63;  // 0 = q, 1 = k, 2 = v; first of 16 columns inside the head
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
        add_vec_f32<16>(o, p.bias + which * 1024 + head * 64 + cc);
        const uint4 lo = pack8(o), hi = pack8(o + 8);
        uint8_t* tile_row = smem + which * TILE_BYTES + r * 128;
        const int ch = cc >> 3;  // 16-byte chunk index within the 128-byte row
        *reinterpret_cast<uint4*>(tile_row + ((ch ^ (r & 7)) << 4)) = lo;
        *reinterpret_cast<uint4*>(tile_row + (((ch + 1) ^ (r & 7)) << 4)) = hi;
      }
      tc::tc_fence_before();
    }
    if (warp == 2 && lane == 0) trace_mark(p.trace, 3);
    asm volatile("bar.sync 1, %0;" ::"r"(32 * AW) : "memory");  // all attention warps: tiles complete, the GEMM ring is free for staging

    static_assert(ALG == 1, "the lane-per-key variant (ALG 2) was removed with the interleaved cache layout");
    {
      // ---- attention
      const int sub = lane >> 3, dseg = lane & 7;
      if (!p.early_kv)
        for (int s = 0; s < NSLOT && s < total; ++s) issue(s);
      int seq = 0;
      for (int j = 0; j < n_items; ++j) {
        const int rl = aw + AW * j;
        if (p.l2_ahead > 0 && lane == 0 && j + p.l2_ahead < n_items) {
          bulk_prefetch_l2(p.kv.cache + p.kv.offset(p.layer, 0, item_row(j + p.l2_ahead), head, 0), blk_bytes);
        }
        const int sw = (dseg ^ (rl & 7)) << 4;
        float qv[8];
        unpack8(*reinterpret_cast<const uint4*>(smem + rl * 128 + sw), qv);
        const uint4 knew = *reinterpret_cast<const uint4*>(smem + TILE_BYTES + rl * 128 + sw);
        const uint4 vnew = *reinterpret_cast<const uint4*>(smem + 2 * TILE_BYTES + rl * 128 + sw);
        // KV-cache append of this item (slot t + 1): key row and value row are adjacent, so 16 lanes write one contiguous 256-byte
        // block — off the epilogue's critical path, where it was 32 scattered 16-byte stores per warp instruction
        if (sub < 2)
          *reinterpret_cast<uint4*>(p.kv.cache + p.kv.offset(p.layer, sub, item_row(j), head, t + 1) + dseg * 8) = sub == 0 ? knew : vnew;
        float m = -INFINITY, den = 0.0f;
        float acc[8];
  #pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
        for (int c = 0; c < nchunks; ++c) {
          const int c0 = c * CHUNK_KEYS;
          const bool staged = c < nsc;
          const int s = seq % NSLOT;
          if (staged) tc::mbar_wait(&bars[s], (seq / NSLOT) & 1);
          if (seq == 0 && warp == 2 && lane == 0) trace_mark(p.trace, 4);
          const uint8_t* Kb = stg + s * SLOT_BYTES;  // [key][k | v][64] as in the cache
          const uint8_t* Vb = Kb + 128;
          float sc[4];
          uint4 vr[4];
  #pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int key = c0 + i * 4 + sub;
            uint4 kraw;
            if (key < Lc) {
              kraw = *reinterpret_cast<const uint4*>(Kb + (key - c0) * 256 + dseg * 16);
              vr[i] = *reinterpret_cast<const uint4*>(Vb + (key - c0) * 256 + dseg * 16);
            } else {  // key == Lc: the word of this step; beyond: masked below
              kraw = knew;
              vr[i] = vnew;
            }
            float kf[8];
            unpack8(kraw, kf);
            float part = 0.0f;
  #pragma unroll
            for (int e = 0; e < 8; ++e) part = fmaf(qv[e], kf[e], part);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            sc[i] = (key < Ltot) ? part * 0.125f : -INFINITY;  // / sqrt(64)   (language_model.py:88)
          }
          const float m_new = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
          if (m_new > -INFINITY) {
            const float corr = __expf(m - m_new);
            den *= corr;
  #pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] *= corr;
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float pr = __expf(sc[i] - m_new);  // exp(-inf) = 0 for masked keys
              den += pr;
              float vf[8];
              unpack8(vr[i], vf);
  #pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, vf[e], acc[e]);
            }
            m = m_new;
          }
          if (staged) {
            __syncwarp();  // every lane is done reading this slot before it is refilled
            if (seq + NSLOT < total) issue(seq + NSLOT);
            ++seq;
          }
        }
        // merge the 4 key subgroups (lanes differing in bits 3 and 4), normalise, store
        float Mx = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        Mx = fmaxf(Mx, __shfl_xor_sync(0xffffffffu, Mx, 16));
        const float sc_merge = (m == -INFINITY) ? 0.0f : __expf(m - Mx);
        den *= sc_merge;
  #pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= sc_merge;
        den += __shfl_xor_sync(0xffffffffu, den, 8);
        den += __shfl_xor_sync(0xffffffffu, den, 16);
  #pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
          acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        }
        if (sub == 0) {
          const float inv = 1.0f / den;
  #pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] *= inv;
          *reinterpret_cast<uint4*>(p.attn_o + static_cast<size_t>(item_row(j)) * dec::D + head * dec::HD + dseg * 8) = pack8(acc);
        }
      }
    }
  }
  if (warp == 2 && lane == 0) trace_mark(p.trace, 6);
  tc::tc_fence_before();
  __syncthreads();
  if (mc) {  // no CTA of the pair exits while the other may still signal its barriers
    cluster_arrive();
    cluster_wait();
  }
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) trace_mark(p.trace, 7);
}

template <int AW, int NSLOT, bool LN_HEAD, int ALG>
inline void launch(const CUtensorMap& tmA, const CUtensorMap& tmA64, const CUtensorMap& tmW, const Params& p, cudaStream_t stream,
                   bool pdl) {
  using L = Smem<AW, NSLOT, ALG>;
  auto kern = attn_fused_kernel<AW, NSLOT, LN_HEAD, ALG>;
  static bool configured = false;  // one engine device per process (rgrg_create enforces it)
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(p.M, p.rows_per_tile) * 16);
  cfg.blockDim = dim3(64 + 32 * AW);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (!LN_HEAD && p.mc) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 2;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmA64, tmW, p));
}

}  // namespace fa
}  // namespace rgrg
