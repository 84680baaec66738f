// Shared helpers for the rgrg_b200 engine (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>
#include <utility>

typedef __nv_bfloat16 bf16;

namespace rgrg {

struct CudaError : std::runtime_error {
  cudaError_t code;
  CudaError(cudaError_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void cuda_check(cudaError_t e, const char* expr, const char* file, int line) {
  if (e != cudaSuccess) {
    std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " (" + expr + ") at " + file + ":" +
                      std::to_string(line);
    // callers of the reference string-match "out of memory" (evaluate_language_model.py:1208)
    if (e == cudaErrorMemoryAllocation && msg.find("out of memory") == std::string::npos) msg += " [out of memory]";
    cudaGetLastError();
    throw CudaError(e, msg);
  }
}
#define CUDA_CHECK(x) ::rgrg::cuda_check((x), #x, __FILE__, __LINE__)
#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor is still running; it must not touch the predecessor's outputs before griddep_wait().
// Both are no-ops for ordinary launches.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch with (optionally) the PDL attribute; works inside stream capture (becomes a programmatic graph edge)
template <class... KArgs, class... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

// Barrier among the `group` CTAs that share one arrival counter (the CTAs of one M tile of a LayerNorm-head kernel), all of
// them co-resident or scheduled in order.  Counters are monotonic within a generate(): launch number `seq` (0, 1, ...) of
// the kernels that share the counter waits for (seq + 1) * group arrivals.  Called by every thread of the CTA.
// Writers' global stores are released by the arrival; the TMA reads that follow go through the async proxy, hence the
// proxy fences on both sides.
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target) {
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
    unsigned v = old + 1;
    const long long start = clock64();
    while (v < target) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (clock64() - start > 4000000000LL) {  // ~2 s: a scheduling bug must trap, not hang the GPU
        printf("rgrg_b200: LayerNorm-head group barrier timed out (block %d, count %u, target %u)\n", blockIdx.x, v, target);
        __trap();
      }
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
}

// tuning: nanosecond timestamps comparable across SMs (decode-step timeline, tools/decode_timeline.py)
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// slot `i` of this CTA's record when tracing is on and this is the first or the last CTA of the grid
__device__ __forceinline__ void trace_mark(long long* trace, int i) {
  if (trace && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) trace[(blockIdx.x == 0 ? 0 : 8) + i] = global_ns();
}

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

// 8 bf16 <-> 8 floats through one 16-byte vector
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// transformers activations.py NewGELUActivation: 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_new(float x) {
  const float k = 0.7978845608028654f;
  const float inner = k * (x + 0.044715f * x * x * x);
  float t;
  // MUFU.TANH: max abs error ~5e-4, below the bf16 rounding (2^-9 relative) applied to the result right after
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.0f + t);
}

}  // namespace rgrg
