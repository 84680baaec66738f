// tcgen05 + TMA GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T   (bf16 operands, fp32 accumulate in TMEM)
//
// Persistent: one CTA per SM walks 128 x BN output tiles; two TMEM accumulator stages overlap the epilogue of one
// tile with the main loop of the next.  Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA
// issuer, warps 2..9 epilogue (see tc::Pipe).
//
// Operand A is fetched either through a 2-D tensor map over a row-major [M,K] matrix (plain GEMM: every Linear /
// Conv1D / 1x1 conv of the path) or through a 4-D tensor map over an NHWC activation (implicit-GEMM 3x3 conv,
// stride 1, pad 1: the K loop walks the 9 taps, TMA's out-of-bounds zero fill supplies the padding).
// W is always a K-major [N,K] bf16 matrix (weights are repacked once at load time).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace rgrg {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                        // two warps per TMEM lane quarter, alternating 16-column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct GemmShape {
  int M, N;
  int k_iters;    // total 64-wide K blocks (taps * kc_blocks in conv mode)
  int m_tiles, n_tiles;
  int m_fastest;  // raster order: 1 = consecutive CTAs walk M (share a W tile), 0 = walk N (share an A tile)
  // implicit-GEMM conv mode (A through a 4-D NHWC tensor map, 3x3 / stride 1 / pad 1)
  int conv;
  int kc_blocks;         // Cin / 64
  int H, W;              // activation height / width
  int tiles_w, tiles_h;  // output tiles per image: (W/16) x (H/8); one tile = 8 rows x 16 cols = 128 pixels
  long long* trace;      // optional [grid, 8] clock64 timeline of each CTA (bring-up / tuning only)
  int k_splits;          // split-K factor (0/1 = off): tile index -> (split, m, n); each split covers k_iters / k_splits blocks
  // Optional LayerNorm head (16 N tiles per M tile, one tile per CTA, m_fastest = 0 so the 16 CTAs of an M tile are
  // consecutive): before the main loop CTA `n_blk` reduces the preceding projection's split-K partial sums into the residual
  // stream and normalises rows [m_blk*128 + n_blk*8, +8) into `x` (this GEMM's own operand A); the 16 CTAs then meet at a
  // per-M-tile arrival counter (group_barrier, common.cuh).  (A 16-CTA cluster with barrier.cluster was measured first:
  // only 7 such clusters fit on a B200 at this shared-memory size, so 8 M tiles ran as two waves.)
  struct LnHead {
    float* h;             // fp32 residual stream [M, 1024]; null = no head
    bf16* x;              // normalised bf16 rows [M, 1024]
    const float* gamma;
    const float* beta;
    const float* parts;   // split-K partial sums [4][M, 1024] of the preceding projection (null: plain LayerNorm)
    size_t part_stride;
    const float* res_bias;
    unsigned* counters;   // [m_tiles] arrival counters, zeroed at the start of a generate()
    const int* step_ptr;  // device decode step: barrier targets derive from it, so a captured graph can be replayed
    int launch_idx, launches_per_step;  // this launch's position among the launches of a step that share the counters
  } lnh;
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long start = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - start > 4000000000LL) {  // ~2 s at 2 GHz
      printf("rgrg_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 inputs, fp32 accumulate), single-CTA group
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile in shared memory (rows of 64 bf16 = 128 B; 8-row groups 1024 B apart).
// cute::UMMA::SmemDescriptor bit layout: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO: ignored for swizzled K-major, canonical value 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), both K-major, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int STG_FLOATS = 32 * 16;  // per-warp transpose buffer: 32 rows x 16 columns fp32, XOR-swizzled 16-byte chunks

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;
  // per-warp staging: 2 KB transpose buffers (register epilogue); BN = 128 / 256 keep 4 KB per warp so that the TMA-store epilogue
  // (32-row x 128-byte slabs) fits as well — their rings are 192 KB, the other widths have no room for it
  static constexpr int SLAB_BYTES = 32 * 128;
  static constexpr bool kSlabs = (BN == 128 || BN == 256);
  static constexpr int STG_BYTES = EPI_WARPS * (kSlabs ? SLAB_BYTES : STG_FLOATS * 4);
  static constexpr int BAR_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + EPI_WARPS * 8 + 1024;  // + per-warp residual barriers + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget exceeded");
  static_assert(TOTAL > 116 * 1024, "must stay above half an SM's smem: exactly one CTA per SM may own TMEM");
};

template <int BN>
constexpr int tmem_cols() {  // two accumulator stages, power of two >= 32
  return 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
}

struct TileCoord {
  int m_blk, n_blk, img, h0, w0, split;
};
__device__ __forceinline__ TileCoord tile_coord(const GemmShape& s, int tile) {
  TileCoord c;
  c.split = 0;
  if (s.k_splits > 1) {
    const int per = s.m_tiles * s.n_tiles;
    c.split = tile / per;
    tile -= c.split * per;
  }
  if (s.m_fastest) {
    c.m_blk = tile % s.m_tiles;
    c.n_blk = tile / s.m_tiles;
  } else {
    c.n_blk = tile % s.n_tiles;
    c.m_blk = tile / s.n_tiles;
  }
  c.img = c.h0 = c.w0 = 0;
  if (s.conv) {  // decompose the M tile into (image, 8x16 pixel patch)
    const int per_img = s.tiles_w * s.tiles_h;
    c.img = c.m_blk / per_img;
    const int t = c.m_blk - c.img * per_img;
    c.h0 = (t / s.tiles_w) * 8;
    c.w0 = (t % s.tiles_w) * 16;
  }
  return c;
}

// ---------------------------------------------------------------------------------------------------------------
// The tile pipeline.  A CTA is persistent: it walks tiles blockIdx.x, +gridDim.x,
// ...; two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : TMEM allocator + MMA issuer (one lane issues tcgen05.mma; tcgen05.commit frees ring slots)
//   warps 2..9  : epilogue, two warps per TMEM lane quarter alternating 16-column chunks: tcgen05.ld 32x32b.x16
//                 (lane = row) -> swizzled smem transpose -> 8 consecutive columns per lane -> fused epilogue functor
//                 with 16-byte coalesced global accesses.  (One warp per scheduler cannot hide its own latency: the
//                 4-warp version of this epilogue needed 14 us for a 128x192 tile, 3x the tile's MMA time.)
//                 TMA_OUT variant (plain bf16-output GEMMs and implicit convs at BN = 128 / 256): lane = row all the way —
//                 bias / residual / activation in registers, bf16 rows into the warp's 128B-swizzled 4 KB slab, one
//                 cp.async.bulk.tensor store per 32 rows x 64 columns; a bf16 residual is fetched by TMA into the same slab
//                 before the accumulator wait.  The register variant is bound by store issue where K is short.
// Ring-slot and accumulator-stage counters (kbg, it) are carried by the caller.
// ---------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct Pipe {
  using L = SmemLayout<BN, STAGES>;
  uint8_t* smem;
  uint64_t *full_bar, *empty_bar, *tmem_full_bar, *tmem_empty_bar, *res_bar;
  uint32_t tmem_base;

  // barrier init + TMEM allocation; ends with a CTA-wide sync
  __device__ __forceinline__ void setup(uint8_t* smem_raw) {
    smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    empty_bar = full_bar + STAGES;
    tmem_full_bar = empty_bar + STAGES;   // [2]
    tmem_empty_bar = tmem_full_bar + 2;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    res_bar = tmem_empty_bar + 4;         // [EPI_WARPS] residual slab landed (TMA-store epilogue with a bf16 residual)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
#pragma unroll
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
#pragma unroll
      for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&res_bar[i], 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], EPI_WARPS);
      }
      fence_barrier_init();
    }
    if (warp == 1) {
      tmem_alloc(tmem_ptr_smem, tmem_cols<BN>());
      tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem_base = *tmem_ptr_smem;
  }
  __device__ __forceinline__ void teardown() {
    tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) tmem_dealloc(tmem_base, tmem_cols<BN>());
  }

  static __device__ __forceinline__ int num_tiles(const GemmShape& s) { return s.m_tiles * s.n_tiles * (s.k_splits > 1 ? s.k_splits : 1); }
  static __device__ __forceinline__ int k_per_tile(const GemmShape& s) { return s.k_splits > 1 ? s.k_iters / s.k_splits : s.k_iters; }

  // producer, optional head start: arm the first ring slots of this CTA's first tile and start streaming their W
  // (operand B) tiles, which never depend on a predecessor; returns how many k-blocks were started.
  __device__ __forceinline__ int prefetch_w(const CUtensorMap* tmB, const GemmShape& s, int kbg) const {
    if (static_cast<int>(blockIdx.x) >= num_tiles(s)) return 0;
    const TileCoord t0 = tile_coord(s, blockIdx.x);
    const int kpt = k_per_tile(s);
    const int kb0 = t0.split * kpt;
    const int n = kpt < STAGES ? kpt : STAGES;
    for (int i = 0; i < n; ++i) {
      const int st = (kbg + i) % STAGES;
      const uint32_t ph = ((kbg + i) / STAGES) & 1;
      mbar_wait(&empty_bar[st], ph ^ 1);
      mbar_expect_tx(&full_bar[st], L::STAGE_BYTES);
      tma_load_2d(smem + st * L::STAGE_BYTES + L::A_BYTES, tmB, &full_bar[st], (kb0 + i) * BK, t0.n_blk * BN);
    }
    return n;
  }

  // producer main loop (one lane).  `prefetched` k-blocks (from prefetch_w) already have their barrier armed and W in flight.
  __device__ __forceinline__ void produce(const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmShape& s, int& kbg,
                                          int prefetched) const {
    const int nt = num_tiles(s), kpt = k_per_tile(s);
    const int kbg0 = kbg;
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) {
      const TileCoord tc_ = tile_coord(s, tile);
      const int kb_end = (tc_.split + 1) * kpt;
      for (int kb = tc_.split * kpt; kb < kb_end; ++kb, ++kbg) {
        const int st = kbg % STAGES;
        const uint32_t ph = (kbg / STAGES) & 1;
        const bool early_b = (kbg - kbg0) < prefetched;
        if (!early_b) {
          mbar_wait(&empty_bar[st], ph ^ 1);
          mbar_expect_tx(&full_bar[st], L::STAGE_BYTES);
        }
        uint8_t* a_dst = smem + st * L::STAGE_BYTES;
        uint8_t* b_dst = a_dst + L::A_BYTES;
        if (s.conv) {
          const int tap = kb / s.kc_blocks;
          const int kc = kb - tap * s.kc_blocks;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(a_dst, tmA, &full_bar[st], kc * BK, tc_.w0 + dx, tc_.h0 + dy, tc_.img);
        } else {
          tma_load_2d(a_dst, tmA, &full_bar[st], kb * BK, tc_.m_blk * BM);
        }
        if (!early_b) tma_load_2d(b_dst, tmB, &full_bar[st], kb * BK, tc_.n_blk * BN);
      }
    }
  }

  // MMA issuer main loop (one lane)
  __device__ __forceinline__ void mma(const GemmShape& s, int& kbg, int& it, long long* trace) const {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    const int nt = num_tiles(s), kpt = k_per_tile(s);
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < kpt; ++kb, ++kbg) {
        const int st = kbg % STAGES;
        const uint32_t ph = (kbg / STAGES) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        if (trace && kb == 0 && tile == static_cast<int>(blockIdx.x)) trace[2] = clock64();
        const uint32_t a_addr = smem_u32(smem + st * L::STAGE_BYTES);
        const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
        const uint64_t b_desc = make_sw128_kmajor_desc(a_addr + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[st]);  // ring slot reusable once these MMAs have read it
      }
      umma_commit(&tmem_full_bar[acc]);  // accumulator stage complete
      if (trace) trace[3] = clock64();
    }
  }

  // TMA-store epilogue with a bf16 residual: request slab `b` of the warp's part of tile `tc_` from the residual tensor ([M, N] bf16,
  // same boxes as the output map) into the warp's slab, once the slab's previous store has been read out of it
  __device__ __forceinline__ void res_fetch(const GemmShape& s, const TileCoord& tc_, const CUtensorMap* tmR, uint8_t* slab, int q, int half,
                                            int b, int warp, int lane) const {
    const int m0 = tc_.m_blk * BM + q * 32;
    const int cs = tc_.n_blk * BN + half * (BN / 2) + b * 64;
    if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (m0 < s.M && cs < s.N) {
        mbar_expect_tx(&res_bar[warp - 2], L::SLAB_BYTES);  // rows / columns out of bounds are zero-filled and still counted
        tma_load_2d(slab, tmR, &res_bar[warp - 2], cs, m0);
      }
    }
    __syncwarp();
  }
  // epilogue main loop (warps 2..9, all lanes)
  template <class Epi, bool TMA_OUT>
  __device__ __forceinline__ void epilogue(const GemmShape& s, const Epi& epi, int& it, long long* trace, const CUtensorMap* tmC,
                                           const CUtensorMap* tmR) const {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;            // TMEM lane quarter this warp may read: lanes [32q, 32q+32)
    const int half = (warp - 2) >> 2;  // which of the two warps of that quarter: takes 16-column chunks half, half+2, ...
    float* stg = reinterpret_cast<float*>(smem + L::STG_OFFSET) + (warp - 2) * STG_FLOATS;
    uint8_t* slab = smem + L::STG_OFFSET + (warp - 2) * L::SLAB_BYTES;  // TMA_OUT: this warp's 32-row x 128-byte store slab
    uint32_t res_phase = 0;  // parity of the warp's residual barrier (one phase per fetched slab)
    const int nt = num_tiles(s);
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x, ++it) {
      const TileCoord tc_ = tile_coord(s, tile);
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      if constexpr (TMA_OUT) {
        if constexpr (Epi::kResBf16) res_fetch(s, tc_, tmR, slab, q, half, 0, warp, lane);  // lands under the tile's main loop
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      if (trace && warp == 2 && lane == 0 && tile == static_cast<int>(blockIdx.x)) trace[4] = clock64();
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      typename Epi::State st;
      epi.init(st, tc_.split);
      if constexpr (Epi::kDirect) {
        // one row per thread, straight from registers (reductions along N: the arg-max epilogue)
        const int r = q * 32 + lane;
        const int row = tc_.m_blk * BM + r;
        const bool row_ok = row < s.M;
#pragma unroll 1
        for (int c = half * 16; c < BN; c += 32) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + c, v);
          tmem_ld_wait();
          const int col0 = tc_.n_blk * BN + c;
          if (row_ok && col0 < s.N) epi.template apply<16>(st, row, col0, reinterpret_cast<const float*>(v), s.N);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
        if (row_ok) epi.finish(st, row, tc_.n_blk * 2 + half);
      } else if constexpr (TMA_OUT) {
        // Plain GEMM, bf16 output (optionally + bf16 residual): TMEM -> registers (lane = row) -> bias / residual / activation ->
        // bf16 -> the warp's 128B-swizzled slab -> one cp.async.bulk.tensor store per 32 rows x 64 columns.  The register
        // epilogue below issues 16 rows x 32 B per warp store instruction and is bound by store issue on the K <= 512
        // convolutions (tensor pipe 2-8 % busy, profiles/r02_detector_gemms_ncu_full.md); rows past M and columns past N are
        // clipped by the TMA unit.
        constexpr int COLS = BN / 2;        // two warps per TMEM lane quarter, half of the tile's columns each
        constexpr int SLABS = COLS / 64;
        static_assert(L::kSlabs && COLS % 64 == 0, "TMA-store epilogue needs BN = 128 or 256");
        // implicit-conv tiles are 8 x 16 pixel patches: the warp's 32 rows are 2 image rows x 16 pixels -> one 4-D box of the NHWC
        // output map (every conv tile is full); plain GEMM: 32 consecutive rows of [M, N]
        const int r = q * 32 + lane;
        const int row = tc_.m_blk * BM + r;
        const bool row_ok = !s.conv && row < s.M;
        const int m0 = s.conv ? 0 : tc_.m_blk * BM + q * 32;
        const int sw = lane & 7;
#pragma unroll 1
        for (int b = 0; b < SLABS; ++b) {
          const int cs = tc_.n_blk * BN + half * COLS + b * 64;  // first output column of this slab
          const bool live = m0 < s.M && cs < s.N;
          if constexpr (Epi::kResBf16) {
            if (b > 0) res_fetch(s, tc_, tmR, slab, q, half, b, warp, lane);  // slab 0's residual was requested before the accumulator wait
          } else {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the slab's previous store has read it
            __syncwarp();
          }
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int c = half * COLS + b * 64 + h2 * 32;
            uint32_t v[32];
            tmem_ld_32x32b_x16(t_addr + c, v);
            tmem_ld_32x32b_x16(t_addr + c + 16, v + 16);
            tmem_ld_wait();
            if (b == SLABS - 1 && h2 == 1) {  // this warp's last read of the accumulator stage: hand it back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            }
            if (cs < s.N) {
              float o[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
              if constexpr (Epi::kResBf16) {
                // acc + bias, + the residual row segment that TMA put into this lane's slab row, then the activation: apply()'s order
                epi.template add_bias<32>(cs + h2 * 32, o);
                if (live) {
                  if (h2 == 0) {  // first read of the slab: its residual has landed (one phase per fetched slab)
                    mbar_wait(&res_bar[warp - 2], res_phase);
                    res_phase ^= 1;
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float rr[8];
                    unpack8(*reinterpret_cast<const uint4*>(slab + lane * 128 + (((h2 * 4 + j) ^ sw) << 4)), rr);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[8 * j + k] += rr[k];
                  }
                }
                epi.template activate<32>(o);
              } else {
                epi.template transform_row<32>(row, row_ok, cs + h2 * 32, o);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(slab + lane * 128 + (((h2 * 4 + j) ^ sw) << 4)) = pack8(o + 8 * j);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && live) {
            if (s.conv)
              asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                               reinterpret_cast<uint64_t>(tmC)),
                           "r"(smem_u32(slab)), "r"(cs), "r"(tc_.w0), "r"(tc_.h0 + 2 * q), "r"(tc_.img)
                           : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmC)),
                           "r"(smem_u32(slab)), "r"(cs), "r"(m0)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
        // rows of the two transposed passes this lane stores: lane>>1 and 16 + lane>>1 of the warp's 32
        int rows[2];
        bool rows_ok[2];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int r = q * 32 + pass * 16 + (lane >> 1);
          if (s.conv) {
            rows[pass] = (tc_.img * s.H + tc_.h0 + (r >> 4)) * s.W + tc_.w0 + (r & 15);
            rows_ok[pass] = true;
          } else {
            rows[pass] = tc_.m_blk * BM + r;
            rows_ok[pass] = rows[pass] < s.M;
          }
        }
        const int sw_w = (lane >> 1) & 3;  // XOR swizzle of this lane's own row when writing (row = lane)
#pragma unroll 1
        for (int c = half * 16; c < BN; c += 32) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + c, v);
          tmem_ld_wait();
          if (c + 32 >= BN) {  // this warp's last read of the accumulator stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          }
          // transpose through shared memory: thread = row  ->  lane = 8 consecutive columns of 2 rows.
          // 16-byte chunk j of row r lives at slot r*4 + (j ^ ((r>>1)&3)): conflict-free for both phases.
          float4* wr = reinterpret_cast<float4*>(stg) + lane * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            wr[j ^ sw_w] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]));
          __syncwarp();
          const int col0 = tc_.n_blk * BN + c + (lane & 1) * 8;
          const bool col_ok = col0 < s.N;
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const int rl = pass * 16 + (lane >> 1);
            const int sw_r = (rl >> 1) & 3;
            const float4* rd = reinterpret_cast<const float4*>(stg) + rl * 4;
            const int j0 = (lane & 1) * 2;
            const float4 x0 = rd[j0 ^ sw_r], x1 = rd[(j0 + 1) ^ sw_r];
            const float vals[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            if (rows_ok[pass] && col_ok) epi.template apply<8>(st, rows[pass], col0, vals, s.N);
          }
          __syncwarp();
        }
      }
    }
    if constexpr (TMA_OUT) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the stores' reads
    }
  }
};

// one warp: h[row] += bias + split-K partial sums (when parts != null); x[row] = LayerNorm(h[row])   (decoder_kernels.cuh)
__device__ __forceinline__ void ln_head_row(float* h, const float* gamma, const float* beta, bf16* x, int row, int lane,
                                            const float* parts, size_t part_stride, const float* res_bias);

template <int BN, int STAGES, class Epi, bool LN_HEAD = false, bool TMA_OUT = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmC,
                                                               const __grid_constant__ CUtensorMap tmR,
                                                               const GemmShape s, const Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* trace = s.trace ? s.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = clock64();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  Pipe<BN, STAGES> pipe;
  pipe.setup(smem_raw);
  if (trace && threadIdx.x == 0) trace[1] = clock64();
  // PDL: everything above overlapped the predecessor's tail; let our own successor get scheduled early as well
  griddep_launch_dependents();
  int kbg = 0, it = 0;
  // weights never depend on the predecessor kernel: start streaming them before waiting for it
  int prefetched = 0;
  if (warp == 0 && lane == 0) prefetched = pipe.prefetch_w(&tmB, s, 0);
  griddep_wait();
  if constexpr (LN_HEAD) {
    // peers' rows become visible at the cluster barrier; operand A is then read through TMA (async proxy): proxy fences
    const TileCoord t0 = tile_coord(s, blockIdx.x);
    if (warp >= 2) {
      const int row = t0.m_blk * BM + t0.n_blk * 8 + (warp - 2);
      if (row < s.M) ln_head_row(s.lnh.h, s.lnh.gamma, s.lnh.beta, s.lnh.x, row, lane, s.lnh.parts, s.lnh.part_stride, s.lnh.res_bias);
    }
    group_barrier(s.lnh.counters + t0.m_blk,
                  static_cast<unsigned>(*s.lnh.step_ptr * s.lnh.launches_per_step + s.lnh.launch_idx + 1) * 16u);
  }
  if (warp == 0) {
    if (lane == 0) pipe.produce(&tmA, &tmB, s, kbg, prefetched);
  } else if (warp == 1) {
    if (lane == 0) pipe.mma(s, kbg, it, trace);
  } else {
    pipe.template epilogue<Epi, TMA_OUT>(s, epi, it, trace, &tmC, &tmR);
  }
  if (trace && warp == 2 && lane == 0) trace[5] = clock64();
  pipe.teardown();
  if (trace && threadIdx.x == 0) trace[6] = clock64();
}

// ---------------------------------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, K] bf16 matrix; box = 64 (K) x box_rows; 128B swizzle; OOB reads are zero-filled
inline CUtensorMap make_tmap_2d(const void* ptr, uint64_t rows, uint64_t K, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {K * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r));
  return m;
}
// NHWC bf16 activation [N,H,W,C]; box = 64 channels x 16 (W) x 8 (H) x 1 image -> 128 rows of 128 B
inline CUtensorMap make_tmap_nhwc(const void* ptr, uint64_t N, uint64_t H, uint64_t W, uint64_t C) {
  CUtensorMap m;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {BK, 16, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(4d) failed: " + std::to_string((int)r));
  return m;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

template <int BN, int STAGES, class Epi, bool LN_HEAD = false, bool TMA_OUT = false>
inline void launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& s, const Epi& epi,
                   cudaStream_t stream, bool pdl = false, const CUtensorMap* tmC = nullptr, const CUtensorMap* tmR = nullptr) {
  using L = SmemLayout<BN, STAGES>;
  auto kern = gemm_tc_kernel<BN, STAGES, Epi, LN_HEAD, TMA_OUT>;
  if (TMA_OUT && !tmC) throw std::runtime_error("TMA-store epilogue needs the output tensor map");
  constexpr int cluster = 1;
  static bool configured = false;  // one static per template instantiation; one engine device per process (rgrg_create)
  if (!configured) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int tiles = s.m_tiles * s.n_tiles * (s.k_splits > 1 ? s.k_splits : 1);
  // LayerNorm head: one tile per CTA (CTA index == tile index; CTAs are scheduled in index order, so the 16 CTAs of an M
  // tile that meet at the group barrier become resident together)
  const int grid = LN_HEAD ? tiles : (tiles < num_sms() ? tiles : num_sms());
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC ? *tmC : tmA, tmR ? *tmR : tmA, s, epi));
}

// output map of the TMA-store epilogue in implicit-conv mode: NHWC bf16; a warp's slab = 64 channels x 16 (W) x 2 (H) x 1 image
inline CUtensorMap make_tmap_out_nhwc(const void* ptr, uint64_t N, uint64_t H, uint64_t W, uint64_t C) {
  CUtensorMap m;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {64, 16, 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(out nhwc) failed: " + std::to_string((int)r));
  return m;
}
// output map of the TMA-store epilogue: row-major bf16 [M, N]; slabs of 32 rows x 64 columns, 128B swizzle
inline CUtensorMap make_tmap_out_bf16(const void* ptr, uint64_t M, uint64_t N) {
  CUtensorMap m;
  cuuint64_t dims[2] = {N, M};
  cuuint64_t strides[1] = {N * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(out bf16) failed: " + std::to_string((int)r));
  return m;
}

}  // namespace tc
}  // namespace rgrg
