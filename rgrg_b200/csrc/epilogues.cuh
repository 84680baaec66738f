// Fused GEMM epilogues.  A functor sees one output row and NV consecutive accumulator columns at a time:
//   init(State&, split)                               once per thread and tile (split = split-K slice of the tile)
//   apply<NV>(State&, row, col0, const float* v, N)   v[j] = D[row, col0 + j]; columns >= N are padding
//   finish(State&, row, part)                         kDirect functors only: once per row after the thread's last chunk
// The tcgen05 kernel calls apply<8> after its shared-memory transpose (a lane owns 8 consecutive columns of a row, so
// global accesses are 16-byte vectors, 2 lanes per 64-byte bf16 / 4 lanes per 128-byte fp32 segment) or, for kDirect
// functors (reductions along N), apply<16> with one row per thread; the CUDA-core kernel calls apply<4>.
// Everything that selects behaviour is a template parameter: the hot loop carries no runtime branches on options.
#pragma once
#include "common.cuh"

namespace rgrg {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU_NEW = 2 };
enum Res { RES_NONE = 0, RES_BF16 = 1, RES_F32 = 2 };

template <int NV>
__device__ __forceinline__ void add_vec_f32(float* o, const float* __restrict__ src) {
  if constexpr (NV % 4 == 0) {
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(src + j);
      o[j] += t.x; o[j + 1] += t.y; o[j + 2] += t.z; o[j + 3] += t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) o[j] += src[j];
  }
}

// out = act(acc + bias [+ residual]) -> bf16 or fp32.  Residual: bf16 (bottleneck identity) or fp32 (decoder residual
// stream; `out` may alias `res` — each element is read then written by the same thread).
template <bool OUT_BF16, int ACT, int RES, bool BIAS>
struct EpiStoreT {
  struct State {
    size_t split_off;
  };
  static constexpr bool kDirect = false;
  void* out;            // bf16* or float*
  const float* bias;    // [N]
  const void* res;      // bf16* or float*, [M, ldc]
  int ldc;
  size_t split_stride;  // split-K: slice s of the partial sums goes to out + s * split_stride (elements); 0 otherwise

  __device__ __forceinline__ void init(State& s, int split) const { s.split_off = split * split_stride; }
  __device__ __forceinline__ void finish(State&, int, int) const {}

  // the value part alone (bias + activation), for kernels that store the tile themselves (gemm_2cta.cuh's TMA-store epilogue)
  static constexpr bool kTmaOut = (RES == RES_NONE);
  static constexpr bool kOutBf16 = OUT_BF16;
  template <int NV>
  __device__ __forceinline__ void transform(int col0, float* o) const {
    if constexpr (BIAS) add_vec_f32<NV>(o, bias + col0);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if constexpr (ACT == ACT_RELU) o[j] = fmaxf(o[j], 0.0f);
      if constexpr (ACT == ACT_GELU_NEW) o[j] = gelu_new(o[j]);
    }
  }

  // the same with the bf16 residual of a full-width row segment added between bias and activation (the order of apply());
  // for the persistent kernel's TMA-store epilogue (gemm_tc.cuh), bf16 outputs only
  static constexpr bool kTmaOutRow = OUT_BF16 && (RES == RES_NONE || RES == RES_BF16);
  static constexpr bool kResBf16 = (RES == RES_BF16);
  template <int NV>
  __device__ __forceinline__ void add_bias(int col0, float* o) const {
    if constexpr (BIAS) add_vec_f32<NV>(o, bias + col0);
  }
  template <int NV>
  __device__ __forceinline__ void activate(float* o) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if constexpr (ACT == ACT_RELU) o[j] = fmaxf(o[j], 0.0f);
      if constexpr (ACT == ACT_GELU_NEW) o[j] = gelu_new(o[j]);
    }
  }
  template <int NV>
  __device__ __forceinline__ void transform_row(int row, bool row_ok, int col0, float* o) const {
    if constexpr (BIAS) add_vec_f32<NV>(o, bias + col0);
    if constexpr (RES == RES_BF16) {
      if (row_ok) {
        const bf16* rp = static_cast<const bf16*>(res) + static_cast<size_t>(row) * ldc + col0;
#pragma unroll
        for (int j = 0; j < NV; j += 8) {
          float r[8];
          unpack8(*reinterpret_cast<const uint4*>(rp + j), r);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[j + k] += r[k];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if constexpr (ACT == ACT_RELU) o[j] = fmaxf(o[j], 0.0f);
      if constexpr (ACT == ACT_GELU_NEW) o[j] = gelu_new(o[j]);
    }
  }

  template <int NV>
  __device__ __forceinline__ void apply(State& st, int row, int col0, const float* v, int N) const {
    const size_t base = static_cast<size_t>(row) * ldc + col0 + st.split_off;
    float o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) o[j] = v[j];
    const bool full = (col0 + NV <= N);
    const bool vec = full && ((ldc & 7) == 0);  // row starts stay 16-byte aligned for both fp32 and bf16
    if (vec) {
      if constexpr (BIAS) add_vec_f32<NV>(o, bias + col0);
      if constexpr (RES == RES_F32) add_vec_f32<NV>(o, static_cast<const float*>(res) + base);
      if constexpr (RES == RES_BF16) {
        if constexpr (NV % 8 == 0) {
#pragma unroll
          for (int j = 0; j < NV; j += 8) {
            float r[8];
            unpack8(*reinterpret_cast<const uint4*>(static_cast<const bf16*>(res) + base + j), r);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[j + k] += r[k];
          }
        } else {
#pragma unroll
          for (int j = 0; j < NV; ++j) o[j] += bf2f(static_cast<const bf16*>(res)[base + j]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (col0 + j < N) {
          if constexpr (BIAS) o[j] += bias[col0 + j];
          if constexpr (RES == RES_F32) o[j] += static_cast<const float*>(res)[base + j];
          if constexpr (RES == RES_BF16) o[j] += bf2f(static_cast<const bf16*>(res)[base + j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if constexpr (ACT == ACT_RELU) o[j] = fmaxf(o[j], 0.0f);
      if constexpr (ACT == ACT_GELU_NEW) o[j] = gelu_new(o[j]);
    }
    if constexpr (OUT_BF16) {
      bf16* dst = static_cast<bf16*>(out) + base;
      if (vec && NV % 8 == 0) {
#pragma unroll
        for (int j = 0; j < NV; j += 8) *reinterpret_cast<uint4*>(dst + j) = pack8(o + j);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (col0 + j < N) dst[j] = f2bf(o[j]);
      }
    } else {
      float* dst = static_cast<float*>(out) + base;
      if (vec && NV % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (col0 + j < N) dst[j] = o[j];
      }
    }
  }
};

// KV-cache geometry: cache[layer][row][head][slot][kv][64] bf16 (slot 0 = image key/value, slot 1+t = word t).  The key row
// and the value row of a slot are adjacent (256 B), so the cached history of one (row, head) is ONE contiguous block that
// the fused attention kernel streams with one bulk copy per 16-slot chunk.
struct KvGeom {
  bf16* cache;
  int rows_cap;   // row capacity of the allocation
  int slots_cap;  // slot capacity (max_length + 1)
  static constexpr int SLOT_STRIDE = 128;  // elements between consecutive slots of the same (row, head, kv)
  __host__ __device__ __forceinline__ size_t offset(int layer, int kv, int row, int head, int slot) const {
    return (((((static_cast<size_t>(layer) * rows_cap + row) * 16 + head) * slots_cap + slot) * 2) + kv) * 64;
  }
};

template <int NV>
__device__ __forceinline__ void store_bf16_vec(bf16* dst, const float* o) {
  if constexpr (NV % 8 == 0) {
#pragma unroll
    for (int j = 0; j < NV; j += 8) *reinterpret_cast<uint4*>(dst + j) = pack8(o + j);
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) dst[j] = f2bf(o[j]);
  }
}

// c_attn epilogue (language_model.py:132 + :169-170 without the torch.cat): columns [0,1024) -> q buffer,
// [1024,2048) -> K cache, [2048,3072) -> V cache, appended in place at slot *step_ptr + 1.
struct EpiQkvAppend {
  struct State {
    int slot;
  };
  static constexpr bool kDirect = false;
  bf16* q_out;        // [M, 1024]
  const float* bias;  // [3072]
  KvGeom kv;
  int layer;
  const int* step_ptr;  // device-side decode step t (word t is cached at slot t + 1)

  __device__ __forceinline__ void init(State& s, int) const { s.slot = *step_ptr + 1; }
  __device__ __forceinline__ void finish(State&, int, int) const {}

  template <int NV>
  __device__ __forceinline__ void apply(State& s, int row, int col0, const float* v, int N) const {
    float o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) o[j] = v[j];
    add_vec_f32<NV>(o, bias + col0);
    bf16* dst;
    if (col0 < 1024) {
      dst = q_out + static_cast<size_t>(row) * 1024 + col0;
    } else {
      const int c = col0 - 1024;
      dst = kv.cache + kv.offset(layer, c >> 10, row, (c & 1023) >> 6, s.slot) + (c & 63);
    }
    store_bf16_vec<NV>(dst, o);
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

  __device__ __forceinline__ void init(State&, int) const {}
  __device__ __forceinline__ void finish(State&, int, int) const {}

  template <int NV>
  __device__ __forceinline__ void apply(State&, int row, int col0, const float* v, int N) const {
    float o[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) o[j] = v[j];
    add_vec_f32<NV>(o, bias + col0);
    for (int b = 0; b < beams; ++b)
      store_bf16_vec<NV>(kv.cache + kv.offset(col0 >> 11, (col0 >> 10) & 1, row * beams + b, (col0 & 1023) >> 6, 0) + (col0 & 63), o);
  }
};

// lm_head epilogue for greedy decoding (language_model.py:366 + :632): the [rows, 50257] logits are never
// materialised; every epilogue thread emits the (max, first arg-max) of the columns it saw for its row.
struct EpiArgmaxPartial {
  struct State {
    float best;
    int idx;
  };
  static constexpr bool kDirect = true;  // reduction along N: one row per thread, no transpose
  float* part_val;  // [M, n_parts]
  int* part_idx;    // [M, n_parts]
  int n_parts;

  __device__ __forceinline__ void init(State& s, int) const {
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
  __device__ __forceinline__ void finish(State& s, int row, int part) const {
    part_val[static_cast<size_t>(row) * n_parts + part] = s.best;
    part_idx[static_cast<size_t>(row) * n_parts + part] = s.idx;
  }
};

// lm_head epilogue for beam search (language_model.py:556-568): instead of materialising [rows, 50257] fp32 logits
// (746 MB per step at 3712 rows) every epilogue thread emits, for the 128 columns it saw of its row, the online
// (max, sum of exp) pair of the log-softmax denominator and its K best (logit, token) pairs, K = 2 * num_beams.
// log_softmax(x) + beam_score is monotone in x within a row, so the sentence's top 2 * num_beams over num_beams * V
// candidates are among these per-part lists; dec::beam_merge_kernel combines them.
template <int K>
struct EpiBeamPartial {
  struct State {
    float m, l;
    float val[K];
    int idx[K];
  };
  static constexpr bool kDirect = true;
  float* part_m;    // [M, n_parts]
  float* part_l;    // [M, n_parts]
  float* part_val;  // [M, n_parts, K] descending
  int* part_idx;    // [M, n_parts, K]
  int n_parts;

  __device__ __forceinline__ void init(State& s, int) const {
    s.m = -INFINITY;
    s.l = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      s.val[k] = -INFINITY;
      s.idx[k] = 0x7fffffff;
    }
  }
  template <int NV>
  __device__ __forceinline__ void apply(State& s, int row, int col0, const float* v, int N) const {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (col0 + j < N) {
        const float x = v[j];
        if (x > s.m) {
          s.l = s.l * __expf(s.m - x) + 1.0f;
          s.m = x;
        } else {
          s.l += __expf(x - s.m);
        }
        if (x > s.val[K - 1]) {  // strict: on equal logits the lower token index (seen first) stays ahead
          float cv = x;
          int ci = col0 + j;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            if (cv > s.val[k]) {
              const float tv = s.val[k];
              const int ti = s.idx[k];
              s.val[k] = cv;
              s.idx[k] = ci;
              cv = tv;
              ci = ti;
            }
          }
        }
      }
    }
  }
  __device__ __forceinline__ void finish(State& s, int row, int part) const {
    const size_t o = static_cast<size_t>(row) * n_parts + part;
    part_m[o] = s.m;
    part_l[o] = s.l;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      part_val[o * K + k] = s.val[k];
      part_idx[o * K + k] = s.idx[k];
    }
  }
};

}  // namespace rgrg
