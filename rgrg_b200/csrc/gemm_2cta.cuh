// 2-CTA tcgen05 GEMM for the decode-step projections:  D[M,N] = A[M,K] * W[N,K]^T  (bf16 operands, fp32 accumulate).
//
// Why: at 928 decode rows the 1-CTA kernel (gemm_tc.cuh, 128 x 256 tiles) re-reads 16 KB of A and 32 KB of W per k-block
// and CTA — 100 MB of L2 -> SM traffic per c_fc launch — and its main loop runs at the chip-wide L2 delivery cap
// (0.525 us per k-block against 0.27 us of tensor-pipe time, profiles/r02_decode_gemms_ncu_full.md).  A CTA PAIR
// (cluster of 2 = one TPC) computes a 256 x 256 tile with `tcgen05.mma.cta_group::2`: each CTA stages its own 128 rows of
// A and only HALF of the W tile (128 of the 256 N rows); the tensor core reads both halves out of the two CTAs' shared
// memories.  32 KB instead of 48 KB per k-block and CTA.
//
// One pair-tile per CTA pair (grid = 2 x pair-tiles; the decode GEMMs have 64 pair-tiles), split-K optional.
//   warp 0  TMA producer (both CTAs; loads signal the LEADER's full barrier: cp.async.bulk.tensor...cta_group::2)
//   warp 1  TMEM allocator (both CTAs, cta_group::2) + MMA issuer (leader CTA only); tcgen05.commit multicasts the
//           "stage free" / "accumulator ready" arrivals to both CTAs
//   warps 2..9 epilogue: each CTA drains its own 128 accumulator rows.  Two variants:
//           TMA_OUT = false: the scheme of gemm_tc.cuh (transpose through shared memory, 16-byte global stores);
//           TMA_OUT = true (plain bias / activation epilogues): TMEM -> registers -> 128B-swizzled 32-row slabs in the (by then
//           idle) operand ring -> `cp.async.bulk.tensor` stores.  A timeline of the decode step (tools/decode_timeline.py)
//           showed the first variant spending 5-7 us per full tile on its global stores (1 us on a tile whose rows are
//           all past M, i.e. without the stores; no faster with 16 warps) — as long as the 16-k-block main loop.
#pragma once
#include "gemm_tc.cuh"

namespace rgrg {
namespace tc2 {

constexpr int BN = 256;          // pair-tile N
constexpr int HALF_N = BN / 2;   // W rows each CTA stages
constexpr int A_BYTES = tc::BM * tc::BK * 2;      // 16 KB
constexpr int B_BYTES = HALF_N * tc::BK * 2;      // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // per CTA
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int SLAB_BYTES = 32 * 128;             // TMA-store slab: 32 rows x 128 B (32 fp32 or 64 bf16 columns)
constexpr int STG_BYTES = EPI_WARPS * tc::STG_FLOATS * 4;
// STAGES = 6: 209 KB, one CTA per SM.  STAGES = 3: 113 KB — two CTAs (2 x 256 TMEM columns) fit on an SM, so the
// prologue of the next projection kernel overlaps the epilogue of the current one under PDL.
template <int STAGES>
struct Layout {
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;
};
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;       // clears the CTA-rank bit of a shared::cluster address: the leader's copy

struct Shape {
  int M, N;
  int k_iters;            // 64-wide K blocks in total
  int m_pairs, n_tiles;   // pair-tiles: m_pairs x n_tiles (x k_splits)
  int k_splits;           // 0/1 = off
  long long* trace;       // tuning: [2][8] timestamps of the first / last CTA (entry, setup done, predecessor done, first MMA,
                          // last MMA issued, accumulator ready, epilogue done, exit)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// every thread of both CTAs; the non-.aligned forms: the lanes of the producer / MMA warps arrive at different times
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// both CTAs call it; the transaction bytes are credited to the LEADER CTA's barrier at the same shared-memory offset
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared memory (128B-swizzled slab of 32 rows) -> global through the output tensor map {N, M, splits}; rows past M are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(tc::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

template <class Epi, int STAGES, bool TMA_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const Shape s, const Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int BAR_OFFSET = Layout<STAGES>::BAR_OFFSET, STG_OFFSET = Layout<STAGES>::STG_OFFSET;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) trace_mark(s.trace, 0);

  // pair-tile of this cluster
  int tile = blockIdx.x >> 1;
  int split = 0;
  if (s.k_splits > 1) {
    const int per = s.m_pairs * s.n_tiles;
    split = tile / per;
    tile -= split * per;
  }
  const int m_pair = tile % s.m_pairs, n_blk = tile / s.m_pairs;  // consecutive pairs share the W tile
  const int m_blk = 2 * m_pair + static_cast<int>(rank);          // this CTA's 128-row block
  const int kpt = s.k_splits > 1 ? s.k_iters / s.k_splits : s.k_iters;
  const int kb0 = split * kpt;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
#pragma unroll
    for (int i = 0; i < STAGES; ++i) {
      tc::mbar_init(&full_bar[i], 1);
      tc::mbar_init(&empty_bar[i], 1);
    }
    tc::mbar_init(tmem_full_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_ptr_smem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  cluster_sync();  // both CTAs: barriers initialised and TMEM allocated before any cross-CTA signal
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_launch_dependents();
  if (threadIdx.x == 0) trace_mark(s.trace, 1);

  if (warp == 0) {
    if (lane == 0) {
      // W never depends on the predecessor kernel: start the first stages' W halves before waiting for it
      const int pre = kpt < STAGES ? kpt : STAGES;
      for (int i = 0; i < pre; ++i) {
        if (leader) tc::mbar_expect_tx(&full_bar[i], 2 * STAGE_BYTES);
        tma_load_2d_2sm(smem + i * STAGE_BYTES + A_BYTES, &tmB, &full_bar[i], (kb0 + i) * tc::BK, n_blk * BN + static_cast<int>(rank) * HALF_N);
      }
      griddep_wait();
      trace_mark(s.trace, 2);
      for (int kb = 0; kb < kpt; ++kb) {
        const int st = kb % STAGES;
        uint8_t* a_dst = smem + st * STAGE_BYTES;
        if (kb >= pre) {
          tc::mbar_wait(&empty_bar[st], ((kb / STAGES) & 1) ^ 1);
          if (leader) tc::mbar_expect_tx(&full_bar[st], 2 * STAGE_BYTES);
          tma_load_2d_2sm(a_dst + A_BYTES, &tmB, &full_bar[st], (kb0 + kb) * tc::BK, n_blk * BN + static_cast<int>(rank) * HALF_N);
        }
        tma_load_2d_2sm(a_dst, &tmA, &full_bar[st], (kb0 + kb) * tc::BK, m_blk * tc::BM);
      }
    }
  } else if (warp == 1) {
    griddep_wait();
    if (leader && lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16(2 * tc::BM, BN);
      for (int kb = 0; kb < kpt; ++kb) {
        const int st = kb % STAGES;
        tc::mbar_wait(&full_bar[st], (kb / STAGES) & 1);
        tc::tc_fence_after();
        if (kb == 0) trace_mark(s.trace, 3);
        const uint32_t a_addr = tc::smem_u32(smem + st * STAGE_BYTES);
        const uint64_t a_desc = tc::make_sw128_kmajor_desc(a_addr);
        const uint64_t b_desc = tc::make_sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < tc::BK / tc::UMMA_K; ++k) umma_bf16_2sm(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_2sm(&empty_bar[st]);  // both CTAs' ring slots reusable once these MMAs have read them
      }
      umma_commit_2sm(tmem_full_bar);  // both CTAs' accumulator halves complete
      trace_mark(s.trace, 4);
    }
  } else {
    griddep_wait();
    // ---- epilogue: this CTA's 128 accumulator rows x 256 columns (same scheme as tc::Pipe::epilogue, one tile)
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;  // which of the warps of this lane quarter
    float* stg = reinterpret_cast<float*>(smem + STG_OFFSET) + (warp - 2) * tc::STG_FLOATS;
    tc::mbar_wait(tmem_full_bar, 0);
    tc::tc_fence_after();
    if (warp == 2 && lane == 0) trace_mark(s.trace, 5);
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if constexpr (TMA_OUT) {
      // All MMAs have completed (tmem_full), so every TMA load has landed and been consumed: the operand ring is free.
      // This warp owns rows [q * 32, +32) x columns [grp * COLS, +COLS) of the CTA's tile, in slabs of 128 B per row.
      constexpr int COLS = BN / (EPI_WARPS / 4);                      // 128
      constexpr int SLAB_COLS = Epi::kOutBf16 ? 64 : 32;
      constexpr int SLABS = COLS / SLAB_COLS;
      static_assert(EPI_WARPS * SLABS * SLAB_BYTES <= STAGES * STAGE_BYTES, "slabs must fit in the operand ring");
      const int m0 = m_blk * tc::BM + q * 32;
      uint8_t* my_slabs = smem + (warp - 2) * SLABS * SLAB_BYTES;
      const int sw = lane & 7;
      if (m0 < s.M) {
#pragma unroll 1
        for (int b = 0; b < SLABS; ++b) {
          uint8_t* slab_row = my_slabs + b * SLAB_BYTES + lane * 128;
#pragma unroll
          for (int h2 = 0; h2 < SLAB_COLS / 32; ++h2) {
            const int c = grp * COLS + b * SLAB_COLS + h2 * 32;
            uint32_t v[32];
            tc::tmem_ld_32x32b_x16(t_addr + c, v);
            tc::tmem_ld_32x32b_x16(t_addr + c + 16, v + 16);
            tc::tmem_ld_wait();
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
            epi.template transform<32>(n_blk * BN + c, o);
            if constexpr (Epi::kOutBf16) {
#pragma unroll
              for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(slab_row + (((h2 * 4 + j) ^ sw) << 4)) = pack8(o + 8 * j);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(slab_row + ((j ^ sw) << 4)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmC, my_slabs + b * SLAB_BYTES, n_blk * BN + grp * COLS + b * SLAB_COLS, m0, split);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the reads
      }
    } else {
    typename Epi::State est;
    epi.init(est, split);
    int rows[2];
    bool rows_ok[2];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      rows[pass] = m_blk * tc::BM + q * 32 + pass * 16 + (lane >> 1);
      rows_ok[pass] = rows[pass] < s.M;
    }
    const int sw_w = (lane >> 1) & 3;
#pragma unroll 1
    for (int c = grp * 16; c < BN; c += 16 * (EPI_WARPS / 4)) {
      uint32_t v[16];
      tc::tmem_ld_32x32b_x16(t_addr + c, v);
      tc::tmem_ld_wait();
      float4* wr = reinterpret_cast<float4*>(stg) + lane * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        wr[j ^ sw_w] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
      __syncwarp();
      const int col0 = n_blk * BN + c + (lane & 1) * 8;
      const bool col_ok = col0 < s.N;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int rl = pass * 16 + (lane >> 1);
        const int sw_r = (rl >> 1) & 3;
        const float4* rd = reinterpret_cast<const float4*>(stg) + rl * 4;
        const int j0 = (lane & 1) * 2;
        const float4 x0 = rd[j0 ^ sw_r], x1 = rd[(j0 + 1) ^ sw_r];
        const float vals[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        if (rows_ok[pass] && col_ok) epi.template apply<8>(est, rows[pass], col0, vals, s.N);
      }
      __syncwarp();
    }
    }
  }
  if (warp == 2 && lane == 0) trace_mark(s.trace, 6);
  tc::tc_fence_before();
  cluster_sync();  // nobody frees TMEM or exits while the peer may still signal / read
  if (threadIdx.x == 0) trace_mark(s.trace, 7);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

// output map for the TMA-store epilogue: row-major [splits][M][N] (fp32 partial sums) or [M][N] (splits = 1); slabs of 32 rows x 128 B
inline CUtensorMap make_tmap_out(const void* ptr, uint64_t M, uint64_t N, uint64_t splits, bool bf16_out) {
  CUtensorMap m;
  const uint64_t es = bf16_out ? 2 : 4;
  cuuint64_t dims[3] = {N, M, splits};
  cuuint64_t strides[2] = {N * es, M * N * es};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / es), 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = tc::encode_fn()(&m, bf16_out ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr),
                               dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(out) failed: " + std::to_string((int)r));
  return m;
}

// tmC: only read when TMA_OUT (pass tmA otherwise)
template <class Epi, int STAGES, bool TMA_OUT = false>
inline void launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const Shape& s, const Epi& epi,
                   cudaStream_t stream, bool pdl) {
  auto kern = gemm_2cta_kernel<Epi, STAGES, TMA_OUT>;
  constexpr int SMEM_TOTAL = Layout<STAGES>::TOTAL;
  static bool configured = false;  // one engine device per process (rgrg_create enforces it)
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured = true;
  }
  const int pairs = s.m_pairs * s.n_tiles * (s.k_splits > 1 ? s.k_splits : 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, s, epi));
}

}  // namespace tc2
}  // namespace rgrg
