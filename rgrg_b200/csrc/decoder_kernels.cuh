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

// K15  h[r] = wte[token] + wte[position]   — positions are embedded through wte, not wpe (language_model.py:307)
// token of row r at step t = ids[r, t]; position = t.
__global__ void __launch_bounds__(256) embed_kernel(const float* __restrict__ wte, const int* __restrict__ ids, int ids_ld,
                                                    const int* __restrict__ step_ptr, float* __restrict__ h) {
  const int r = blockIdx.x;
  const int t = *step_ptr;
  const int tok = ids[static_cast<size_t>(r) * ids_ld + t];
  const float4* a = reinterpret_cast<const float4*>(wte + static_cast<size_t>(tok) * D);
  const float4* p = reinterpret_cast<const float4*>(wte + static_cast<size_t>(t) * D);
  float4* o = reinterpret_cast<float4*>(h + static_cast<size_t>(r) * D);
  const float4 x = a[threadIdx.x], y = p[threadIdx.x];
  o[threadIdx.x] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

// K16  LayerNorm(eps 1e-5) over 1024 features: fp32 residual stream -> bf16 GEMM operand.  One warp per row.
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ h, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ out, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(h + static_cast<size_t>(row) * D);
  float v[32];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 x = src[i * 32 + lane];
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

// K18  single-query attention over the in-place KV cache (language_model.py:84-114 for a 1-token query):
// scores = q.K^T / 8 over slots [0, t+2) (slot 0 = image key), softmax in fp32, out = P.V.  The causal / padding
// masks are all-pass in generate().  One warp per (row, head); K and V rows are read as contiguous 128-byte lines:
// lane l holds dims (l%8)*8..+8 of key 4*i + l/8.
__global__ void __launch_bounds__(128) attention_kernel(const bf16* __restrict__ q, KvGeom kv, int layer,
                                                        const int* __restrict__ step_ptr, bf16* __restrict__ out, int rows) {
  const int gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= rows * HEADS) return;
  const int row = gw / HEADS, head = gw % HEADS;
  const int L = *step_ptr + 2;
  const int sub = lane >> 3, dseg = lane & 7;

  float qv[8];
  unpack8(*reinterpret_cast<const uint4*>(q + static_cast<size_t>(row) * D + head * HD + dseg * 8), qv);
  const bf16* Kp = kv.cache + kv.offset(layer, 0, row, head, 0);
  const bf16* Vp = kv.cache + kv.offset(layer, 1, row, head, 0);

  // online softmax: each of the 4 key subgroups keeps its own running (max, denominator, accumulator); K and V of a
  // key are fetched together so both 16-byte loads are in flight before the dependent math
  float m = -INFINITY, den = 0.0f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  const int iters = (L + 3) >> 2;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    const int key = i * 4 + sub;
    const bool ok = key < L;
    uint4 kraw = make_uint4(0, 0, 0, 0), vraw = make_uint4(0, 0, 0, 0);
    if (ok) {
      kraw = *reinterpret_cast<const uint4*>(Kp + static_cast<size_t>(key) * HD + dseg * 8);
      vraw = *reinterpret_cast<const uint4*>(Vp + static_cast<size_t>(key) * HD + dseg * 8);
    }
    float kf[8], vf[8];
    unpack8(kraw, kf);
    unpack8(vraw, vf);
    float part = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) part = fmaf(qv[e], kf[e], part);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (ok) {
      const float sc = part * 0.125f;  // / sqrt(64)   (language_model.py:88)
      const float m_new = fmaxf(m, sc);
      const float corr = __expf(m - m_new);
      const float p = __expf(sc - m_new);
      den = fmaf(den, corr, p);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(acc[e], corr, p * vf[e]);
      m = m_new;
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

// K22/K23  arg-max over the vocabulary + greedy bookkeeping (language_model.py:629-650).  One CTA.
//   next = argmax(logits) (lowest index on ties); finished rows emit pad; ids[:, t+1] = next; a row finishes when it
//   emits EOS; unfinished_count[t] lets the host stop early without a per-step sync; finally step += 1.
// Source of the arg-max: tile partials of the fused lm_head epilogue (part_*), or a full fp32 logits matrix.
struct GreedyState {
  int* ids;          // [rows, ids_ld]
  int ids_ld;
  int* unfinished;   // [rows] 1 = still generating
  int* unfinished_count;  // [max_steps]
  int* step_ptr;
  const int* forced; // optional [rows, ids_ld]: teacher forcing — ids[:, t+1] = forced[:, t+1], arg-max is only recorded
  int* argmax_out;   // optional [max_steps, rows] raw arg-max per step (tests)
};

__global__ void __launch_bounds__(1024) greedy_update_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx,
                                                             int n_tiles, const float* __restrict__ logits /*or null*/,
                                                             GreedyState g, int rows) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int t = *g.step_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int local_unfinished = 0;
  for (int r = warp; r < rows; r += 32) {
    float best = -INFINITY;
    int idx = 0x7fffffff;
    if (logits) {
      const float* l = logits + static_cast<size_t>(r) * VOCAB;
      for (int c = lane; c < VOCAB; c += 32) {
        const float v = l[c];
        if (v > best) { best = v; idx = c; }
      }
    } else {
      for (int c = lane; c < n_tiles; c += 32) {
        const float v = part_val[static_cast<size_t>(r) * n_tiles + c];
        const int i = part_idx[static_cast<size_t>(r) * n_tiles + c];
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
      local_unfinished += unf;
    }
  }
  if (lane == 0 && local_unfinished) atomicAdd(&s_cnt, local_unfinished);
  __syncthreads();
  if (threadIdx.x == 0) {
    g.unfinished_count[t] = s_cnt;
    *g.step_ptr = t + 1;
  }
}

// start of a generate() call: ids[:, 0] = BOS, unfinished = 1, step = 0
__global__ void greedy_init_kernel(GreedyState g, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) {
    g.ids[static_cast<size_t>(r) * g.ids_ld] = g.forced ? g.forced[static_cast<size_t>(r) * g.ids_ld] : EOS_ID;
    g.unfinished[r] = 1;
  }
  if (r == 0) *g.step_ptr = 0;
}

}  // namespace dec
}  // namespace rgrg
