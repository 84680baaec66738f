// Fused GEMM epilogues.  A functor sees one output row and NV consecutive accumulator columns at a time:
//   init(State&)                                   once per thread and tile
//   apply<NV>(State&, row, col0, const float* v, N)   v[j] = D[row, col0 + j]; columns >= N are padding
//   finish(State&, row, n_blk)                     kDirect functors only: once per row after the tile's last chunk
// The tcgen05 kernel calls apply<8> after its shared-memory transpose (a lane owns 8 consecutive columns, 4 lanes cover
// a 128-byte fp32 row segment) or, for kDirect functors (reductions along N), apply<32> with one row per thread;
// the CUDA-core kernel calls apply<4>.
#pragma once
#include "common.cuh"

namespace rgrg {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU_NEW = 2 };

// out = act(acc + bias [+ residual]); bf16 or fp32 destination; residual bf16 (bottleneck identity) or fp32
// (decoder residual stream; out_f32 may alias res_f32 — each element is read then written by the same thread).
struct EpiStore {
  struct State {};
  static constexpr bool kDirect = false;
  float* out_f32;
  bf16* out_bf16;
  const float* bias;     // [N] or null
  const bf16* res_bf16;  // [M, ldc] or null
  const float* res_f32;  // [M, ldc] or null
  int ldc;
  int act;

  __device__ __forceinline__ void init(State&) const {}
  __device__ __forceinline__ void finish(State&, int, int) const {}

  template <int NV>
  __device__ __forceinline__ void apply(State&, int row, int col0, const float* v, int N) const {
    const size_t base = static_cast<size_t>(row) * ldc + col0;
    const bool full = (col0 + NV <= N);
    float o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float x = v[j];
      if (full || col0 + j < N) {
        if (bias) x += bias[col0 + j];
        if (res_bf16) x += bf2f(res_bf16[base + j]);
        if (res_f32) x += res_f32[base + j];
      }
      if (act == ACT_RELU) x = fmaxf(x, 0.0f);
      else if (act == ACT_GELU_NEW) x = gelu_new(x);
      o[j] = x;
    }
    if (out_bf16) {
      if (full && NV % 8 == 0 && (ldc & 7) == 0) {
#pragma unroll
        for (int j = 0; j < NV; j += 8) *reinterpret_cast<uint4*>(out_bf16 + base + j) = pack8(o + j);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (col0 + j < N) out_bf16[base + j] = f2bf(o[j]);
      }
    }
    if (out_f32) {
      if (full && NV % 4 == 0 && (ldc & 3) == 0) {
#pragma unroll
        for (int j = 0; j < NV; j += 4)
          *reinterpret_cast<float4*>(out_f32 + base + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (col0 + j < N) out_f32[base + j] = o[j];
      }
    }
  }
};

// KV-cache geometry: cache[layer][kv][row][head][slot][64] bf16 (slot 0 = image key/value, slot 1+t = word t).
struct KvGeom {
  bf16* cache;
  int rows_cap;   // row capacity of the allocation
  int slots_cap;  // slot capacity (max_length + 1)
  __device__ __forceinline__ size_t offset(int layer, int kv, int row, int head, int slot) const {
    return ((((static_cast<size_t>(layer) * 2 + kv) * rows_cap + row) * 16 + head) * slots_cap + slot) * 64;
  }
};

// c_attn epilogue (language_model.py:132 + :169-170 without the torch.cat): columns [0,1024) -> q buffer,
// [1024,2048) -> K cache, [2048,3072) -> V cache, appended in place at slot *step_ptr + 1.
struct EpiQkvAppend {
  struct State {};
  static constexpr bool kDirect = false;
  bf16* q_out;        // [M, 1024]
  const float* bias;  // [3072]
  KvGeom kv;
  int layer;
  const int* step_ptr;  // device-side decode step t (word t is cached at slot t + 1)

  __device__ __forceinline__ void init(State&) const {}
  __device__ __forceinline__ void finish(State&, int, int) const {}

  template <int NV>
  __device__ __forceinline__ void apply(State&, int row, int col0, const float* v, int N) const {
    float o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) o[j] = v[j] + bias[col0 + j];
    bf16* dst;
    if (col0 < 1024) {
      dst = q_out + static_cast<size_t>(row) * 1024 + col0;
    } else {
      const int c = col0 - 1024;
      const int which = c >> 10;  // 0 = K, 1 = V
      const int head = (c & 1023) >> 6;
      const int d0 = c & 63;
      dst = kv.cache + kv.offset(layer, which, row, head, *step_ptr + 1) + d0;
    }
    if (NV % 8 == 0) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) *reinterpret_cast<uint4*>(dst + j) = pack8(o + j);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) dst[j] = f2bf(o[j]);
    }
  }
};

// Image key/value epilogue (language_model.py:140-147): one GEMM over all 24 layers' uk/uv, N = 24*2*1024;
// column n -> layer n/2048, k/v (n/1024)&1, head (n&1023)/64; written to cache slot 0 of every beam of the row.
struct EpiImageKv {
  struct State {};
  static constexpr bool kDirect = false;
  const float* bias;  // [49152]
  KvGeom kv;
  int beams;

  __device__ __forceinline__ void init(State&) const {}
  __device__ __forceinline__ void finish(State&, int, int) const {}

  template <int NV>
  __device__ __forceinline__ void apply(State&, int row, int col0, const float* v, int N) const {
    float o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) o[j] = v[j] + bias[col0 + j];
    const int layer = col0 >> 11;
    const int which = (col0 >> 10) & 1;
    const int head = (col0 & 1023) >> 6;
    const int d0 = col0 & 63;
    for (int b = 0; b < beams; ++b) {
      bf16* dst = kv.cache + kv.offset(layer, which, row * beams + b, head, 0) + d0;
      if (NV % 8 == 0) {
#pragma unroll
        for (int j = 0; j < NV; j += 8) *reinterpret_cast<uint4*>(dst + j) = pack8(o + j);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) dst[j] = f2bf(o[j]);
      }
    }
  }
};

// lm_head epilogue for greedy decoding (language_model.py:366 + :632): the [rows, 50257] logits are never
// materialised; every CTA emits the (max, first arg-max) of its 128 x BN tile per row.
struct EpiArgmaxPartial {
  struct State {
    float best;
    int idx;
  };
  static constexpr bool kDirect = true;  // reduction along N: one row per thread, no transpose
  float* part_val;  // [M, n_tiles]
  int* part_idx;    // [M, n_tiles]
  int n_tiles;

  __device__ __forceinline__ void init(State& s) const {
    s.best = -INFINITY;
    s.idx = 0x7fffffff;
  }
  template <int NV>
  __device__ __forceinline__ void apply(State& s, int row, int col0, const float* v, int N) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (col0 + j < N && v[j] > s.best) {  // strict '>' keeps the lowest index on ties (torch.argmax on CPU)
        s.best = v[j];
        s.idx = col0 + j;
      }
    }
  }
  __device__ __forceinline__ void finish(State& s, int row, int n_blk) const {
    part_val[static_cast<size_t>(row) * n_tiles + n_blk] = s.best;
    part_idx[static_cast<size_t>(row) * n_tiles + n_blk] = s.idx;
  }
};

}  // namespace rgrg
