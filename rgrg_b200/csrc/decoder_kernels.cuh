// HBM-bound kernels of the GPT-2 pseudo-self-attention decoder (everything between the GEMMs).
// All of them read the decode step t from device memory so that one captured CUDA graph replays every step.
#pragma once
#include "common.cuh"
#include "epilogues.cuh"

namespace rgrg {
namespace dec {

constexpr int D = 1024;
constexpr int HEADS = 16;
constexpr int HD = 64;
constexpr int VOCAB = 50257;
constexpr int EOS_ID = 50256;  // bos == eos == pad (language_model.py:200-202)

// order-preserving float <-> uint32 key (ascending)
__device__ __forceinline__ unsigned float_key_dec(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_key_dec(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// K15  h[r] = wte[token] + wte[position]   — positions are embedded through wte, not wpe (language_model.py:307)
// token of row r at step t = ids[r, t]; position = t.
// one warp embeds one row (32 floats per lane)
__device__ __forceinline__ void embed_row_dev(const float* __restrict__ wte, int tok, int pos, float* __restrict__ hrow, int lane) {
  const float4* a = reinterpret_cast<const float4*>(wte + static_cast<size_t>(tok) * D);
  const float4* p = reinterpret_cast<const float4*>(wte + static_cast<size_t>(pos) * D);
  float4* o = reinterpret_cast<float4*>(hrow);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 x = a[i * 32 + lane], y = p[i * 32 + lane];
    o[i * 32 + lane] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ wte, const int* __restrict__ ids, int ids_ld,
                                                    const int* __restrict__ step_ptr, float* __restrict__ h, int rows) {
  griddep_launch_dependents();  // dependents may be scheduled now; they block at their own griddep_wait until this grid completes
  griddep_wait();
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int t = *step_ptr;
  embed_row_dev(wte, ids[static_cast<size_t>(r) * ids_ld + t], t, h + static_cast<size_t>(r) * D, threadIdx.x & 31);
}

// K16  LayerNorm(eps 1e-5) over 1024 features: fp32 residual stream -> bf16 GEMM operand.  One warp per row.
// Fused with the residual update of the GEMM that precedes it (language_model.py:350, :357): when `parts` is given,
// h <- h + bias + sum of the GEMM's split-K partial sums (written by its epilogue), stored back, then normalised.
template <int NPARTS>
__device__ __forceinline__ void ln_row_dev(float* __restrict__ h, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           bf16* __restrict__ out, int row, int lane, const float* __restrict__ parts,
                                           size_t part_stride, const float* __restrict__ res_bias) {
  float4* src = reinterpret_cast<float4*>(h + static_cast<size_t>(row) * D);
  float v[32];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 x = src[i * 32 + lane];
    if constexpr (NPARTS > 0) {
      const float4 b = *reinterpret_cast<const float4*>(res_bias + (i * 32 + lane) * 4);
      x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
#pragma unroll
      for (int p = 0; p < NPARTS; ++p) {
        // partial sums come from other CTAs' epilogues: read through L2 (__ldcg), never from a possibly stale L1 line
        const float4 t = __ldcg(reinterpret_cast<const float4*>(parts + p * part_stride + static_cast<size_t>(row) * D + (i * 32 + lane) * 4));
        x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
      }
      src[i * 32 + lane] = x;
    }
    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
    sum += x.x + x.y + x.z + x.w;
  }
  const float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float d = v[i] - mean;
    sq = fmaf(d, d, sq);
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    __nv_bfloat162 lo = __floats2bfloat162_rn((v[4 * i] - mean) * rstd * g.x + b.x, (v[4 * i + 1] - mean) * rstd * g.y + b.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn((v[4 * i + 2] - mean) * rstd * g.z + b.z, (v[4 * i + 3] - mean) * rstd * g.w + b.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * D + c) = pk;
  }
}
// stand-alone kernel: one row per 128-thread CTA (4 warps x 8 elements per lane): four times the parallelism and a
// quarter of the per-lane load chain of the warp-per-row form, at the price of two block-level reductions
template <int NPARTS>
__global__ void __launch_bounds__(128) layernorm_kernel(float* __restrict__ h, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ out, int rows,
                                                        const float* __restrict__ parts, size_t part_stride,
                                                        const float* __restrict__ res_bias, long long* trace) {
  __shared__ float s_red[2][4];
  if (threadIdx.x == 0) trace_mark(trace, 0);
  // parameters do not depend on the predecessor kernel: fetch them before waiting for it (PDL)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = tid * 8;
  float g[8], b[8], rb[8];
  *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(gamma + c0);
  *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(gamma + c0 + 4);
  *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(beta + c0);
  *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(beta + c0 + 4);
  if constexpr (NPARTS > 0) {
    *reinterpret_cast<float4*>(rb) = *reinterpret_cast<const float4*>(res_bias + c0);
    *reinterpret_cast<float4*>(rb + 4) = *reinterpret_cast<const float4*>(res_bias + c0 + 4);
  }
  griddep_launch_dependents();  // dependents may be scheduled now; they block at their own griddep_wait until this grid completes
  griddep_wait();
  if (threadIdx.x == 0) trace_mark(trace, 2);
  const int row = blockIdx.x;
  float* hp = h + static_cast<size_t>(row) * D + c0;
  float v[8];
  *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(hp);
  *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(hp + 4);
  if constexpr (NPARTS > 0) {
    float4 t[NPARTS][2];
#pragma unroll
    for (int p = 0; p < NPARTS; ++p) {
      const float* pp = parts + p * part_stride + static_cast<size_t>(row) * D + c0;
      t[p][0] = __ldcg(reinterpret_cast<const float4*>(pp));
      t[p][1] = __ldcg(reinterpret_cast<const float4*>(pp + 4));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += rb[e];
#pragma unroll
    for (int p = 0; p < NPARTS; ++p) {
      v[0] += t[p][0].x; v[1] += t[p][0].y; v[2] += t[p][0].z; v[3] += t[p][0].w;
      v[4] += t[p][1].x; v[5] += t[p][1].y; v[6] += t[p][1].z; v[7] += t[p][1].w;
    }
    *reinterpret_cast<float4*>(hp) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(hp + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int e = 0; e < 8; ++e) sum += v[e];
  sum = warp_sum(sum);
  if (lane == 0) s_red[0][warp] = sum;
  __syncthreads();
  const float mean = (s_red[0][0] + s_red[0][1] + s_red[0][2] + s_red[0][3]) * (1.0f / D);
  float sq = 0.0f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float d = v[e] - mean;
    sq = fmaf(d, d, sq);
  }
  sq = warp_sum(sq);
  if (lane == 0) s_red[1][warp] = sq;
  __syncthreads();
  const float rstd = rsqrtf((s_red[1][0] + s_red[1][1] + s_red[1][2] + s_red[1][3]) * (1.0f / D) + 1e-5f);
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = (v[e] - mean) * rstd * g[e] + b[e];
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * D + c0) = pack8(o);
  if (threadIdx.x == 0) trace_mark(trace, 7);
}

// K18  single-query attention over the in-place KV cache (language_model.py:84-114 for a 1-token query):
// scores = q.K^T / 8 over slots [0, t+2) (slot 0 = image key), softmax in fp32, out = P.V.  The causal / padding
// masks are all-pass in generate().  One warp per (row, head); K and V rows are read as contiguous 128-byte lines:
// lane l holds dims (l%8)*8..+8 of key 4*i + l/8.
// Beam search: `anc` (optional) maps (row, slot) to the beam of the same sentence whose physical cache row holds that
// slot, so the reference's per-step index_select of the whole cache (language_model.py:492-496) becomes a table lookup.
template <bool DOUBLE_BUFFER>
__device__ __forceinline__ void attention_dev(const bf16* __restrict__ q, const KvGeom& kv, int layer, int L, bf16* __restrict__ out,
                                              int row, int head, int lane, const unsigned char* __restrict__ anc, int anc_ld, int nb) {
  const int sub = lane >> 3, dseg = lane & 7;

  float qv[8];
  unpack8(__ldcg(reinterpret_cast<const uint4*>(q + static_cast<size_t>(row) * D + head * HD + dseg * 8)), qv);
  const bf16* Kp = kv.cache + kv.offset(layer, 0, row, head, 0);
  const bf16* Vp = kv.cache + kv.offset(layer, 1, row, head, 0);
  const int sent_row0 = anc ? (row / nb) * nb : 0;
  const unsigned char* arow = anc ? anc + static_cast<size_t>(row) * anc_ld : nullptr;

  // Online softmax; each of the 4 key subgroups keeps its own running (max, denominator, accumulator).
  // Keys are processed 16 at a time (4 per subgroup): all eight 16-byte loads of a chunk are issued before any of the
  // dependent math, and the next chunk's loads are issued before the current chunk is reduced (register double buffer),
  // so a warp keeps 8-16 independent 512-byte requests in flight.
  float m = -INFINITY, den = 0.0f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  auto load_chunk = [&](int c0, uint4* kr, uint4* vr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int key = c0 + i * 4 + sub;
      key = key < L ? key : L - 1;  // clamp: in-bounds address, masked out below
      const bf16* kp = Kp;
      const bf16* vp = Vp;
      if (anc) {
        const int prow = sent_row0 + arow[key];
        kp = kv.cache + kv.offset(layer, 0, prow, head, 0);
        vp = kv.cache + kv.offset(layer, 1, prow, head, 0);
      }
      kr[i] = __ldcg(reinterpret_cast<const uint4*>(kp + static_cast<size_t>(key) * KvGeom::SLOT_STRIDE + dseg * 8));  // streamed once: skip L1
      vr[i] = __ldcg(reinterpret_cast<const uint4*>(vp + static_cast<size_t>(key) * KvGeom::SLOT_STRIDE + dseg * 8));
    }
  };
  uint4 kcur[4], vcur[4], knext[DOUBLE_BUFFER ? 4 : 1], vnext[DOUBLE_BUFFER ? 4 : 1];
  load_chunk(0, kcur, vcur);
  for (int c0 = 0; c0 < L; c0 += 16) {
    const bool more = c0 + 16 < L;
    if constexpr (DOUBLE_BUFFER) {
      if (more) load_chunk(c0 + 16, knext, vnext);
    }
    float sc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float kf[8];
      unpack8(kcur[i], kf);
      float part = 0.0f;
#pragma unroll
      for (int e = 0; e < 8; ++e) part = fmaf(qv[e], kf[e], part);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      sc[i] = (c0 + i * 4 + sub < L) ? part * 0.125f : -INFINITY;  // / sqrt(64)   (language_model.py:88)
    }
    const float m_new = fmaxf(fmaxf(m, fmaxf(sc[0], sc[1])), fmaxf(sc[2], sc[3]));
    if (m_new > -INFINITY) {
      const float corr = __expf(m - m_new);
      den *= corr;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = __expf(sc[i] - m_new);  // exp(-inf) = 0 for masked keys
        den += p;
        float vf[8];
        unpack8(vcur[i], vf);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, vf[e], acc[e]);
      }
      m = m_new;
    }
    if constexpr (DOUBLE_BUFFER) {
      if (more) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          kcur[i] = knext[i];
          vcur[i] = vnext[i];
        }
      }
    } else {
      if (more) load_chunk(c0 + 16, kcur, vcur);
    }
  }
  // merge the 4 key subgroups (lanes differing in bits 3 and 4)
  float M = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
  M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 16));
  const float sc_merge = (m == -INFINITY) ? 0.0f : __expf(m - M);
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
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * D + head * HD + dseg * 8) = pack8(acc);
  }
}
// MIN_CTAS CTAs per SM (7 -> 28 warps, 72 registers, measured best of 5..8): a warp's lifetime is three dependent memory hops, so at small L the kernel is bound by how
// many warps are resident, not by loads in flight per warp (single 16-key chunk buffer)
template <int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS) attention_kernel(const bf16* __restrict__ q, KvGeom kv, int layer,
                                                                  const int* __restrict__ step_ptr, bf16* __restrict__ out, int rows,
                                                                  const unsigned char* __restrict__ anc, int anc_ld, int nb) {
  griddep_launch_dependents();  // dependents may be scheduled now; they block at their own griddep_wait until this grid completes
  griddep_wait();
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (gw >= rows * HEADS) return;
  attention_dev<false>(q, kv, layer, *step_ptr + 2, out, gw / HEADS, gw % HEADS, threadIdx.x & 31, anc, anc_ld, nb);
}

// K22/K23  arg-max over the vocabulary + greedy bookkeeping (language_model.py:629-650).  One warp per row.
//   next = argmax(logits) (lowest index on ties); finished rows emit pad; ids[:, t+1] = next; a row finishes when it
//   emits EOS; unfinished_count[t] lets the host stop early without a per-step sync; the last CTA does step += 1.
// Source of the arg-max: tile partials of the fused lm_head epilogue (part_*), or a full fp32 logits matrix.
struct GreedyState {
  int* ids;          // [rows, ids_ld]
  int ids_ld;
  int* unfinished;   // [rows] 1 = still generating
  int* unfinished_count;  // [max_steps], zeroed by greedy_init_kernel
  int* ticket;       // CTA arrival counter of greedy_update_kernel
  int live_rows;     // rows >= live_rows are padding (row count rounded up for graph reuse): born finished, never counted
  int* step_ptr;
  const int* forced; // optional [rows, ids_ld]: teacher forcing — ids[:, t+1] = forced[:, t+1], arg-max is only recorded
  int* argmax_out;   // optional [max_steps, rows] raw arg-max per step (tests)
};

// one warp: arg-max of one row + greedy bookkeeping
__device__ __forceinline__ void greedy_row_dev(const float* __restrict__ part_val, const int* __restrict__ part_idx, int n_tiles,
                                               const float* __restrict__ logits, const GreedyState& g, int rows, int r, int t,
                                               int lane) {
  float best = -INFINITY;
  int idx = 0x7fffffff;
  if (logits) {
    const float* l = logits + static_cast<size_t>(r) * VOCAB;
    for (int c = lane; c < VOCAB; c += 32) {
      const float v = __ldcg(l + c);
      if (v > best) { best = v; idx = c; }
    }
  } else {
    for (int c = lane; c < n_tiles; c += 32) {
      const float v = __ldcg(part_val + static_cast<size_t>(r) * n_tiles + c);
      const int i = __ldcg(part_idx + static_cast<size_t>(r) * n_tiles + c);
      if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
  }
  if (lane == 0) {
    if (g.argmax_out) g.argmax_out[static_cast<size_t>(t) * rows + r] = idx;
    int nxt = idx;
    int unf = g.unfinished[r];
    const bool in_range = t + 1 < g.ids_ld;
    if (g.forced) {
      nxt = in_range ? g.forced[static_cast<size_t>(r) * g.ids_ld + t + 1] : EOS_ID;
    } else {
      if (!unf) nxt = EOS_ID;
      if (nxt == EOS_ID) unf = 0;
      g.unfinished[r] = unf;
    }
    if (in_range) g.ids[static_cast<size_t>(r) * g.ids_ld + t + 1] = nxt;
    if (unf) atomicAdd(&g.unfinished_count[t], 1);
  }
}
__global__ void __launch_bounds__(256) greedy_update_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx,
                                                            int n_tiles, const float* __restrict__ logits /*or null*/,
                                                            GreedyState g, int rows) {
  // one warp per row, 8 rows per CTA; the last CTA to finish publishes step + 1 (all CTAs read the step first)
  griddep_wait();
  const int t = *g.step_ptr;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r < rows) greedy_row_dev(part_val, part_idx, n_tiles, logits, g, rows, r, t, threadIdx.x & 31);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(g.ticket, 1);
    if (ticket == static_cast<int>(gridDim.x) - 1) {
      *g.ticket = 0;
      *g.step_ptr = t + 1;
    }
  }
}

// start of a generate() call: ids[:, 0] = BOS, unfinished = 1, step = 0
__global__ void greedy_init_kernel(GreedyState g, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < g.ids_ld) g.unfinished_count[r] = 0;
  if (r == 0) g.ticket[0] = 0;
  if (r < rows) {
    g.ids[static_cast<size_t>(r) * g.ids_ld] = g.forced ? g.forced[static_cast<size_t>(r) * g.ids_ld] : EOS_ID;
    g.unfinished[r] = r < g.live_rows ? 1 : 0;
  }
  if (r == 0) *g.step_ptr = 0;
}

}  // namespace dec
}  // namespace rgrg

// =====================================================================================================================
// Beam search (language_model.py:529-607 + transformers==4.19.2 BeamSearchScorer, restated in oracle/beam_scorer.py)
// Rows = sentences x beams.  All bookkeeping stays on the device; only finalize() runs on the host, once.
// =====================================================================================================================
namespace rgrg {
namespace dec {

constexpr int MAX_BEAMS = 8;

struct BeamState {
  int nb;             // beams per sentence
  int ids_ld;         // max_length
  int* ids[2];        // [rows, ids_ld] double-buffered token matrix (reordered every step)
  unsigned char* anc[2];  // [rows, slots] physical beam (within the sentence) holding each cache slot of a row
  int slots;
  float* beam_scores; // [sentences, nb]
  float* cand_score;  // [sentences, 2*nb]  top-2nb candidates of the step, sorted descending
  int* cand_token;    // [sentences, 2*nb]
  int* cand_beam;     // [sentences, 2*nb]  beam index within the sentence
  // finished hypotheses (BeamHypotheses): at most nb kept per sentence
  float* hyp_score;   // [sentences, nb]
  int* hyp_len;       // [sentences, nb]
  int* hyp_tok;       // [sentences, nb, ids_ld]
  int* hyp_count;     // [sentences]
  float* worst;       // [sentences]
  int* done;          // [sentences]
  int* not_done_count;  // [max_steps]
  int* step_ptr;
  int early_stopping;
};

__global__ void beam_init_kernel(BeamState s, int sentences) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = sentences * s.nb;
  if (i < rows) {
    s.ids[0][static_cast<size_t>(i) * s.ids_ld] = EOS_ID;  // BOS == EOS (language_model.py:200-202)
    for (int k = 0; k < s.slots; ++k) s.anc[0][static_cast<size_t>(i) * s.slots + k] = static_cast<unsigned char>(i % s.nb);
    s.beam_scores[i] = (i % s.nb == 0) ? 0.0f : -1e9f;  // language_model.py:545-547
  }
  if (i < sentences) {
    s.hyp_count[i] = 0;
    s.worst[i] = 1e9f;
    s.done[i] = 0;
  }
  if (i == 0) *s.step_ptr = 0;
}

// log_softmax over the vocabulary + beam score, then the 2*nb best of the nb*V candidates of a sentence
// (language_model.py:556-568).  One CTA per sentence.  Ties -> lowest flat index.
__global__ void __launch_bounds__(1024) beam_topk_kernel(const float* __restrict__ logits, BeamState s) {
  __shared__ float s_red[32];
  __shared__ float s_lse_m[MAX_BEAMS], s_lse_l[MAX_BEAMS];
  __shared__ unsigned long long s_best[32];
  __shared__ unsigned long long s_pick;
  const int sent = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = s.nb;
  // ---- per beam: max and log-sum-exp
  for (int b = 0; b < nb; ++b) {
    const float* l = logits + static_cast<size_t>(sent * nb + b) * VOCAB;
    float m = -INFINITY;
    for (int c = tid; c < VOCAB; c += 1024) m = fmaxf(m, l[c]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = s_red[lane];
    m = warp_max(m);
    __syncthreads();
    float sum = 0.0f;
    for (int c = tid; c < VOCAB; c += 1024) sum += expf(l[c] - m);
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = warp_sum(s_red[lane]);
    if (tid == 0) {
      s_lse_m[b] = m;
      s_lse_l[b] = logf(sum);
    }
    __syncthreads();
  }
  // ---- thread-local top-K over its strided share of the nb*V candidates, K = 2*nb
  const int K = 2 * nb;
  unsigned long long loc[2 * MAX_BEAMS];
#pragma unroll
  for (int i = 0; i < 2 * MAX_BEAMS; ++i) loc[i] = 0ull;
  const int total = nb * VOCAB;
  for (int j = tid; j < total; j += 1024) {
    const int b = j / VOCAB, c = j - b * VOCAB;
    const float x = logits[static_cast<size_t>(sent * nb + b) * VOCAB + c];
    const float sc = __fadd_rn(__fsub_rn(__fsub_rn(x, s_lse_m[b]), s_lse_l[b]), s.beam_scores[sent * nb + b]);
    unsigned long long key = (static_cast<unsigned long long>(float_key_dec(sc)) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(j));
    if (key > loc[K - 1]) {
      int p = K - 1;
      while (p > 0 && loc[p - 1] < key) {
        loc[p] = loc[p - 1];
        --p;
      }
      loc[p] = key;
    }
  }
  // ---- K rounds of block arg-max over the heads of the per-thread lists
  int head = 0;
  for (int r = 0; r < K; ++r) {
    unsigned long long v = head < K ? loc[head] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ov = __shfl_xor_sync(0xffffffffu, v, o);
      v = ov > v ? ov : v;
    }
    if (lane == 0) s_best[warp] = v;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = s_best[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, w, o);
        w = ov > w ? ov : w;
      }
      if (lane == 0) s_pick = w;
    }
    __syncthreads();
    const unsigned long long pick = s_pick;
    if (head < K && loc[head] == pick) ++head;  // keys are unique (they embed the flat index)
    if (tid == 0) {
      const unsigned j = 0xFFFFFFFFu - static_cast<unsigned>(pick & 0xFFFFFFFFull);
      s.cand_score[sent * K + r] = float_from_key_dec(static_cast<unsigned>(pick >> 32));
      s.cand_token[sent * K + r] = static_cast<int>(j % VOCAB);
      s.cand_beam[sent * K + r] = static_cast<int>(j / VOCAB);
    }
    __syncthreads();
  }
}

// The same selection from the fused lm_head epilogue's per-part summaries (EpiBeamPartial) instead of full logits:
// per beam row the log-softmax denominator is merged from the (max, sum-exp) pairs, then the 2*nb best of the
// nb * n_parts * K candidate (logit, token) pairs are chosen exactly like above (score = logit - lse + beam score,
// ties -> lowest flat index beam * V + token).  One CTA per sentence.
template <int K>
__global__ void __launch_bounds__(1024) beam_merge_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                          const float* __restrict__ part_val, const int* __restrict__ part_idx,
                                                          int n_parts, BeamState s) {
  __shared__ float s_red[32];
  __shared__ float s_lse[MAX_BEAMS];
  __shared__ unsigned long long s_best[32];
  __shared__ unsigned long long s_pick;
  const int sent = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = s.nb;
  for (int b = 0; b < nb; ++b) {
    const size_t base = static_cast<size_t>(sent * nb + b) * n_parts;
    float m = -INFINITY;
    for (int p = tid; p < n_parts; p += 1024) m = fmaxf(m, part_m[base + p]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = warp_max(s_red[lane]);
    __syncthreads();
    float sum = 0.0f;
    for (int p = tid; p < n_parts; p += 1024) sum += part_l[base + p] * expf(part_m[base + p] - m);
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = warp_sum(s_red[lane]);
    if (tid == 0) s_lse[b] = m + logf(sum);
    __syncthreads();
  }
  const int KK = 2 * nb;  // <= K
  unsigned long long loc[2 * MAX_BEAMS];
#pragma unroll
  for (int i = 0; i < 2 * MAX_BEAMS; ++i) loc[i] = 0ull;
  const int total = nb * n_parts * K;
  for (int j = tid; j < total; j += 1024) {
    const int b = j / (n_parts * K);
    const int rest = j - b * (n_parts * K);
    const size_t o = static_cast<size_t>(sent * nb + b) * n_parts * K + rest;
    const int tok = part_idx[o];
    if (tok >= VOCAB) continue;  // empty list slot
    const float sc = __fadd_rn(__fsub_rn(part_val[o], s_lse[b]), s.beam_scores[sent * nb + b]);
    const unsigned flat = static_cast<unsigned>(b) * VOCAB + static_cast<unsigned>(tok);
    const unsigned long long key = (static_cast<unsigned long long>(float_key_dec(sc)) << 32) | (0xFFFFFFFFu - flat);
    if (key > loc[KK - 1]) {
      int p = KK - 1;
      while (p > 0 && loc[p - 1] < key) {
        loc[p] = loc[p - 1];
        --p;
      }
      loc[p] = key;
    }
  }
  int head = 0;
  for (int r = 0; r < KK; ++r) {
    unsigned long long v = head < KK ? loc[head] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ov = __shfl_xor_sync(0xffffffffu, v, o);
      v = ov > v ? ov : v;
    }
    if (lane == 0) s_best[warp] = v;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = s_best[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ov = __shfl_xor_sync(0xffffffffu, w, o);
        w = ov > w ? ov : w;
      }
      if (lane == 0) s_pick = w;
    }
    __syncthreads();
    const unsigned long long pick = s_pick;
    if (head < KK && loc[head] == pick) ++head;  // keys are unique (they embed the flat index)
    if (tid == 0) {
      const unsigned j = 0xFFFFFFFFu - static_cast<unsigned>(pick & 0xFFFFFFFFull);
      s.cand_score[sent * KK + r] = float_from_key_dec(static_cast<unsigned>(pick >> 32));
      s.cand_token[sent * KK + r] = static_cast<int>(j % VOCAB);
      s.cand_beam[sent * KK + r] = static_cast<int>(j / VOCAB);
    }
    __syncthreads();
  }
}

// BeamSearchScorer.process (one thread per sentence) + the reorder of token rows and cache ancestry
// (language_model.py:570-589).  next token matrix goes to ids[dst], ancestry to anc[dst].
__global__ void beam_process_kernel(BeamState s, int sentences, int src, int dst) {
  const int sent = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = *s.step_ptr;       // tokens per row before this step = t + 1 (cur_len in the scorer)
  const int cur_len = t + 1;
  const int nb = s.nb, K = 2 * nb;
  if (sent < sentences) {
    int next_beam[MAX_BEAMS], next_tok[MAX_BEAMS];
    float next_score[MAX_BEAMS];
    if (s.done[sent]) {
      for (int b = 0; b < nb; ++b) {
        next_beam[b] = 0;  // the scorer pads finished sentences with (score 0, pad token, beam 0)
        next_tok[b] = EOS_ID;
        next_score[b] = 0.0f;
      }
    } else {
      int filled = 0;
      for (int r = 0; r < K && filled < nb; ++r) {
        const float sc = s.cand_score[sent * K + r];
        const int tok = s.cand_token[sent * K + r], bm = s.cand_beam[sent * K + r];
        if (tok == EOS_ID) {
          if (r >= nb) continue;
          // BeamHypotheses.add(input_ids[beam].clone(), sum_logprobs): score = sum_logprobs / len ** 1.0
          const float hs = sc / static_cast<float>(cur_len);
          int cnt = s.hyp_count[sent];
          if (cnt < nb || hs > s.worst[sent]) {
            int slot = cnt;
            if (cnt == nb) {
              // over capacity after the append: drop the lowest (score, insertion index).  hs > worst == min of the held
              // scores, so the dropped one is always an existing hypothesis; insertion order of the rest is preserved
              int lo = 0;
              for (int i = 1; i < nb; ++i)
                if (s.hyp_score[sent * nb + i] < s.hyp_score[sent * nb + lo]) lo = i;
              for (int i = lo; i + 1 < nb; ++i) {
                s.hyp_score[sent * nb + i] = s.hyp_score[sent * nb + i + 1];
                s.hyp_len[sent * nb + i] = s.hyp_len[sent * nb + i + 1];
                for (int k = 0; k < s.ids_ld; ++k)
                  s.hyp_tok[(static_cast<size_t>(sent) * nb + i) * s.ids_ld + k] =
                      s.hyp_tok[(static_cast<size_t>(sent) * nb + i + 1) * s.ids_ld + k];
              }
              slot = nb - 1;
            }
            if (slot >= 0) {
              s.hyp_score[sent * nb + slot] = hs;
              s.hyp_len[sent * nb + slot] = cur_len;
              const int* row = s.ids[src] + static_cast<size_t>(sent * nb + bm) * s.ids_ld;
              for (int k = 0; k < cur_len; ++k) s.hyp_tok[(static_cast<size_t>(sent) * nb + slot) * s.ids_ld + k] = row[k];
              if (cnt < nb) {
                s.hyp_count[sent] = cnt + 1;
                s.worst[sent] = fminf(hs, s.worst[sent]);
              } else {
                float w2 = s.hyp_score[sent * nb];
                for (int i = 1; i < nb; ++i) w2 = fminf(w2, s.hyp_score[sent * nb + i]);
                s.worst[sent] = w2;
              }
            }
          }
        } else {
          next_beam[filled] = bm;
          next_tok[filled] = tok;
          next_score[filled] = sc;
          ++filled;
        }
      }
      // is_done(best_sum_logprobs = max candidate score, cur_len)
      bool d = false;
      if (s.hyp_count[sent] >= nb) {
        if (s.early_stopping) d = true;
        else d = s.worst[sent] >= s.cand_score[sent * K] / static_cast<float>(cur_len);
      }
      if (d) s.done[sent] = 1;
    }
    for (int b = 0; b < nb; ++b) {
      const int row = sent * nb + b;
      const int from = sent * nb + next_beam[b];
      s.beam_scores[row] = next_score[b];
      const int* srow = s.ids[src] + static_cast<size_t>(from) * s.ids_ld;
      int* drow = s.ids[dst] + static_cast<size_t>(row) * s.ids_ld;
      for (int k = 0; k < cur_len; ++k) drow[k] = srow[k];
      if (cur_len < s.ids_ld) drow[cur_len] = next_tok[b];
      const unsigned char* sa = s.anc[src] + static_cast<size_t>(from) * s.slots;
      unsigned char* da = s.anc[dst] + static_cast<size_t>(row) * s.slots;
      for (int k = 0; k <= cur_len && k < s.slots; ++k) da[k] = sa[k];  // slots 0 .. t+1 follow the parent beam
      for (int k = cur_len + 1; k < s.slots; ++k) da[k] = static_cast<unsigned char>(b);  // future slots: written by the row itself
    }
  }
}

__global__ void beam_step_end_kernel(BeamState s, int sentences) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < sentences; i += blockDim.x) c += s.done[i] ? 0 : 1;
  if (c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = *s.step_ptr;
    s.not_done_count[t] = s_cnt;
    *s.step_ptr = t + 1;
  }
}

}  // namespace dec
}  // namespace rgrg

// definition of the LayerNorm-head hook declared in gemm_tc.cuh (split-K factor is always 4 on this path)
namespace rgrg {
namespace tc {
__device__ __forceinline__ void ln_head_row(float* h, const float* gamma, const float* beta, bf16* x, int row, int lane,
                                            const float* parts, size_t part_stride, const float* res_bias) {
  if (parts) dec::ln_row_dev<4>(h, gamma, beta, x, row, lane, parts, part_stride, res_bias);
  else dec::ln_row_dev<0>(h, gamma, beta, x, row, lane, nullptr, 0, nullptr);
}
}  // namespace tc
}  // namespace rgrg
