// EXPERIMENTAL (opt-in, `rgrg_set_option("megakernel", 1)`; token-checked against the default path).
// One persistent kernel per decode step (greedy mode): the whole 24-layer transformer body, the lm_head with its fused
// arg-max and the greedy bookkeeping run as PHASES of a single cooperative launch, separated by grid barriers
// (2.0 us each, measured) instead of kernel boundaries (172 per step in the default multi-kernel path).
// Measured slower than the default (3.1 vs 2.4 ms / step): with ten warps per SM the row-parallel phases are
// latency-bound and the GEMM phases keep their fill / drain cost; see profiles/r01_decode_experiments.md.
//
//   grid  = one CTA per SM (cooperative launch: all CTAs are co-resident, which the grid barrier requires)
//   block = 320 threads: the tile pipeline of gemm_tc.cuh (warp 0 TMA, warp 1 MMA, warps 2..9 epilogue) during GEMM
//           phases; all ten warps as plain workers (one warp = one row / one (row, head)) during the other phases.
//
// Cross-CTA visibility between phases: producers write with st.global (generic proxy) and the barrier carries a
// gpu-scope release/acquire (__threadfence) plus fence.proxy.async on both sides, because the next phase may read the
// same bytes through TMA (async proxy).  Data produced by other CTAs is read with ld.global.cg (L2) or TMA, never
// through a possibly stale L1 line; h rows are always owned by the same (CTA, warp).
#pragma once
#include "decoder_kernels.cuh"
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace rgrg {
namespace mega {

constexpr int BN = 256, STAGES = 4, NLAYER = 24, SPLITS = 4;

struct Layer {
  CUtensorMap tm_attn, tm_proj, tm_fc, tm_mproj;  // weight (operand B) maps, N tile 256
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const float *b_attn, *b_proj, *b_fc, *b_mproj;
};

struct Params {
  Layer layer[NLAYER];
  CUtensorMap tm_x, tm_attn_o, tm_mid, tm_lm_head;  // activation (operand A) maps + the tied lm_head weight
  tc::GemmShape s_attn, s_proj, s_fc, s_mproj, s_head;
  const float *lnf_g, *lnf_b, *wte;
  float* h;
  bf16 *x, *q, *attn_o, *mid;
  float* parts;
  KvGeom kv;
  dec::GreedyState g;
  float* part_val;
  int* part_idx;
  int n_parts;
  int rows;
  unsigned* sync_counter;  // [64] (arrival counter + release flag on separate lines), zeroed by the host before every launch
  long long* trace;        // optional [256]: clock64 of CTA 0 after every grid barrier (tuning)
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Barrier over all CTAs of the (co-resident) grid.  counter[0] counts arrivals (monotonic within a launch), counter[32]
// (a different 128-byte line) is the release flag: the last arriver of generation g publishes g there, everybody else
// polls that read-mostly line instead of hammering the line the atomics go to.
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& gen) {
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    unsigned* flag = counter + 32;
    unsigned old;
    // release: everything this CTA wrote (ordered before by bar.sync) is visible before the arrival is
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    if (old == gen * gridDim.x - 1) {
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(gen) : "memory");
    } else {
      long long start = clock64();
      unsigned v;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= gen) break;
        if (clock64() - start > 4000000000LL) {
          printf("rgrg_b200: grid barrier timed out (block %d, gen %u, flag %u)\n", blockIdx.x, gen, v);
          __trap();
        }
      }
    }
  }
  __syncthreads();
  fence_proxy_async();
}
// tuning: cost of one barrier
__global__ void __launch_bounds__(tc::NUM_THREADS, 1) grid_sync_bench_kernel(unsigned* counter, int iters, long long* out) {
  unsigned gen = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) grid_sync(counter, gen);
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = clock64() - t0;
}
__device__ __forceinline__ void mark(long long* trace, int& idx) {
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[idx] = clock64();
  ++idx;
}

using PipeT = tc::Pipe<BN, STAGES>;

// one GEMM phase: the three pipeline roles over this CTA's tiles.  `pre` W k-blocks were already started by prefetch().
template <class Epi>
__device__ __forceinline__ void gemm_phase(const PipeT& pipe, const CUtensorMap* tmA, const CUtensorMap* tmB, const tc::GemmShape& s,
                                           const Epi& epi, int& kbg, int& it, int pre) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) pipe.produce(tmA, tmB, s, kbg, pre);
  } else if (warp == 1) {
    if (lane == 0) pipe.mma(s, kbg, it, nullptr);
  } else {
    pipe.epilogue(s, epi, it, nullptr);
  }
}

__global__ void __launch_bounds__(tc::NUM_THREADS, 1) decoder_step_kernel(const Params* __restrict__ P) {
  extern __shared__ uint8_t smem_raw[];
  PipeT pipe;
  pipe.setup(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * (tc::NUM_THREADS / 32) + warp;  // global warp id for the row-parallel phases
  const int nw = gridDim.x * (tc::NUM_THREADS / 32);
  const bool producer = (warp == 0 && lane == 0);
  const int rows = P->rows;
  const int t = *P->g.step_ptr;  // every CTA reads the step before anybody can advance it
  const int L = t + 2;
  unsigned gen = 0;
  int kbg = 0, it = 0;  // ring-slot / accumulator-stage counters of this thread's pipeline role
  int pre = 0;
  unsigned* sync = P->sync_counter;
  long long* trace = P->trace;
  int ti = 0;
  mark(trace, ti);
  const size_t pstride = static_cast<size_t>(rows) * dec::D;

  // ---- embedding: h = wte[token] + wte[position]
  for (int r = wg; r < rows; r += nw)
    dec::embed_row_dev(P->wte, P->g.ids[static_cast<size_t>(r) * P->g.ids_ld + t], t, P->h + static_cast<size_t>(r) * dec::D, lane);
  // (h rows are produced and consumed by the same warp: no barrier needed before the first LayerNorm)

  for (int l = 0; l < NLAYER; ++l) {
    const Layer& Ly = P->layer[l];
    // ---- LN1 (+ residual update from the previous layer's mlp c_proj)
    if (l == 0) {
      for (int r = wg; r < rows; r += nw) dec::ln_row_dev<0>(P->h, Ly.ln1_g, Ly.ln1_b, P->x, r, lane, nullptr, 0, nullptr);
    } else {
      const float* pb = P->layer[l - 1].b_mproj;
      for (int r = wg; r < rows; r += nw) dec::ln_row_dev<SPLITS>(P->h, Ly.ln1_g, Ly.ln1_b, P->x, r, lane, P->parts, pstride, pb);
    }
    if (producer) pre = pipe.prefetch_w(&Ly.tm_attn, P->s_attn, kbg);  // weights do not depend on the barrier
    grid_sync(sync, gen);
    mark(trace, ti);
    // ---- c_attn: q -> buffer, k / v appended in place into the KV cache
    {
      EpiQkvAppend e{P->q, Ly.b_attn, P->kv, l, P->g.step_ptr};
      gemm_phase(pipe, &P->tm_x, &Ly.tm_attn, P->s_attn, e, kbg, it, pre);
    }
    if (producer) pre = pipe.prefetch_w(&Ly.tm_proj, P->s_proj, kbg);
    grid_sync(sync, gen);
    mark(trace, ti);
    // ---- attention over the cache
    for (int item = wg; item < rows * dec::HEADS; item += nw)
      dec::attention_dev<true>(P->q, P->kv, l, L, P->attn_o, item / dec::HEADS, item % dec::HEADS, lane, nullptr, 0, 1);
    grid_sync(sync, gen);
    mark(trace, ti);
    // ---- attention c_proj, split-K partial sums
    {
      EpiStoreT<false, ACT_NONE, RES_NONE, false> e{};
      e.out = P->parts;
      e.ldc = dec::D;
      e.split_stride = pstride;
      gemm_phase(pipe, &P->tm_attn_o, &Ly.tm_proj, P->s_proj, e, kbg, it, pre);
    }
    grid_sync(sync, gen);
    mark(trace, ti);
    // ---- LN2 (+ residual update from c_proj)
    for (int r = wg; r < rows; r += nw) dec::ln_row_dev<SPLITS>(P->h, Ly.ln2_g, Ly.ln2_b, P->x, r, lane, P->parts, pstride, Ly.b_proj);
    if (producer) pre = pipe.prefetch_w(&Ly.tm_fc, P->s_fc, kbg);
    grid_sync(sync, gen);
    mark(trace, ti);
    // ---- mlp c_fc + gelu_new
    {
      EpiStoreT<true, ACT_GELU_NEW, RES_NONE, true> e{};
      e.out = P->mid;
      e.bias = Ly.b_fc;
      e.ldc = 4 * dec::D;
      e.split_stride = 0;
      gemm_phase(pipe, &P->tm_x, &Ly.tm_fc, P->s_fc, e, kbg, it, pre);
    }
    if (producer) pre = pipe.prefetch_w(&Ly.tm_mproj, P->s_mproj, kbg);
    grid_sync(sync, gen);
    mark(trace, ti);
    // ---- mlp c_proj, split-K partial sums
    {
      EpiStoreT<false, ACT_NONE, RES_NONE, false> e{};
      e.out = P->parts;
      e.ldc = dec::D;
      e.split_stride = pstride;
      gemm_phase(pipe, &P->tm_mid, &Ly.tm_mproj, P->s_mproj, e, kbg, it, pre);
    }
    grid_sync(sync, gen);
    mark(trace, ti);
  }
  // ---- final LayerNorm (+ last residual update)
  for (int r = wg; r < rows; r += nw)
    dec::ln_row_dev<SPLITS>(P->h, P->lnf_g, P->lnf_b, P->x, r, lane, P->parts, pstride, P->layer[NLAYER - 1].b_mproj);
  if (producer) pre = pipe.prefetch_w(&P->tm_lm_head, P->s_head, kbg);
  grid_sync(sync, gen);
  mark(trace, ti);
  // ---- lm_head with fused arg-max partials
  {
    EpiArgmaxPartial e{P->part_val, P->part_idx, P->n_parts};
    gemm_phase(pipe, &P->tm_x, &P->tm_lm_head, P->s_head, e, kbg, it, pre);
  }
  grid_sync(sync, gen);
  mark(trace, ti);
  // ---- greedy bookkeeping
  for (int r = wg; r < rows; r += nw) dec::greedy_row_dev(P->part_val, P->part_idx, P->n_parts, nullptr, P->g, rows, r, t, lane);
  grid_sync(sync, gen);
  mark(trace, ti);
  if (blockIdx.x == 0 && threadIdx.x == 0) *P->g.step_ptr = t + 1;
  pipe.teardown();
}

}  // namespace mega
}  // namespace rgrg
