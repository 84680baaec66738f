// HBM-/latency-bound kernels of the detector half of the path (everything that is not a GEMM):
// stem conv, im2col for the strided 3x3 convs, RPN top-k + box decode + NMS, RoIAlign, per-class region selection.
// Reference semantics are cited per kernel; integer / boolean outputs are bit-exact given identical fp32 inputs.
#pragma once
#include "common.cuh"

namespace rgrg {
namespace det {

// ---------------------------------------------------------------------------------------------------------------
// K1  conv1 7x7 s2 p3 (1 -> 64) + folded BN + ReLU + maxpool 3x3 s2 p1      (object_detector.py:54; resnet.py)
// in : fp32 [B, S, S]                 out: bf16 NHWC [B, S/4, S/4, 64]
// One CTA = one 8x8 tile of pooled pixels; the 17x17 conv tile it needs is built in shared memory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ img, const float* __restrict__ w /*[49][64]*/,
                                                   const float* __restrict__ bias /*[64]*/, bf16* __restrict__ out, int S) {
  __shared__ float s_in[39][40];
  __shared__ bf16 s_conv[17 * 17][64];
  const int b = blockIdx.z;
  const int py0 = blockIdx.y * 8, px0 = blockIdx.x * 8;
  const int C2 = S / 2, P = S / 4;
  const int iy0 = 4 * py0 - 5, ix0 = 4 * px0 - 5;
  const float* src = img + static_cast<size_t>(b) * S * S;
  for (int i = threadIdx.x; i < 39 * 39; i += 256) {
    const int r = i / 39, c = i % 39;
    const int y = iy0 + r, x = ix0 + c;
    s_in[r][c] = (y >= 0 && y < S && x >= 0 && x < S) ? src[static_cast<size_t>(y) * S + x] : 0.0f;
  }
  const int ch = threadIdx.x & 63, grp = threadIdx.x >> 6;
  float wr[49];
#pragma unroll
  for (int k = 0; k < 49; ++k) wr[k] = w[k * 64 + ch];
  const float bs = bias[ch];
  __syncthreads();
  for (int p = grp; p < 17 * 17; p += 4) {
    const int ly = p / 17, lx = p % 17;
    const int cy = 2 * py0 - 1 + ly, cx = 2 * px0 - 1 + lx;
    float acc = 0.0f;
    if (cy >= 0 && cy < C2 && cx >= 0 && cx < C2) {
      // input row of tap (0,0): 2*cy - 3 -> local 2*ly
      acc = bs;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) acc = fmaf(s_in[2 * ly + ky][2 * lx + kx], wr[ky * 7 + kx], acc);
      acc = fmaxf(acc, 0.0f);
    }
    // conv pixels outside the map act as -inf padding of the max-pool; post-ReLU values are >= 0, so 0 is equivalent
    s_conv[p][ch] = f2bf(acc);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 64 * 64; o += 256) {
    const int c = o & 63, pix = o >> 6;
    const int py = pix >> 3, px = pix & 7;
    float m = 0.0f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) m = fmaxf(m, bf2f(s_conv[(2 * py + dy) * 17 + 2 * px + dx][c]));
    out[((static_cast<size_t>(b) * P + py0 + py) * P + px0 + px) * 64 + c] = f2bf(m);
  }
}

// 8 consecutive channels of an activation row, bf16 (fast path) or fp32 (detector_precise parity path)
__device__ __forceinline__ void load8(const bf16* p, float* f) { unpack8(*reinterpret_cast<const uint4*>(p), f); }
__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float* f) { *reinterpret_cast<uint4*>(p) = pack8(f); }
__device__ __forceinline__ void store8(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// fp32 stem for the parity path: conv1 7x7 s2 p3 + folded BN + ReLU (one thread per output element), then maxpool 3x3 s2 p1
__global__ void stem_conv_f32_kernel(const float* __restrict__ img, const float* __restrict__ w /*[49][64]*/,
                                     const float* __restrict__ bias, float* __restrict__ out /*[B,S/2,S/2,64]*/, int B, int S) {
  const int C2 = S / 2;
  const size_t total = static_cast<size_t>(B) * C2 * C2 * 64;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i & 63);
    size_t t = i >> 6;
    const int x = static_cast<int>(t % C2);
    t /= C2;
    const int y = static_cast<int>(t % C2);
    const int b = static_cast<int>(t / C2);
    const float* src = img + static_cast<size_t>(b) * S * S;
    float acc = 0.0f;
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = 2 * y - 3 + ky;
      if (iy < 0 || iy >= S) continue;
      for (int kx = 0; kx < 7; ++kx) {
        const int ix = 2 * x - 3 + kx;
        if (ix < 0 || ix >= S) continue;
        acc = fmaf(src[static_cast<size_t>(iy) * S + ix], w[(ky * 7 + kx) * 64 + c], acc);
      }
    }
    out[i] = fmaxf(acc + bias[c], 0.0f);
  }
}
__global__ void maxpool3x3s2_f32_kernel(const float* __restrict__ in /*[B,H,H,64]*/, float* __restrict__ out /*[B,H/2,H/2,64]*/, int B, int H) {
  const int P = H / 2;
  const size_t total = static_cast<size_t>(B) * P * P * 64;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i & 63);
    size_t t = i >> 6;
    const int x = static_cast<int>(t % P);
    t /= P;
    const int y = static_cast<int>(t % P);
    const int b = static_cast<int>(t / P);
    float m = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = 2 * y + dy, xx = 2 * x + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < H) m = fmaxf(m, in[((static_cast<size_t>(b) * H + yy) * H + xx) * 64 + c]);
      }
    out[i] = m;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// im2col for 3x3 / pad 1 convolutions (stride 1 or 2), NHWC bf16 -> [B*Ho*Wo, 9*C] with K index (tap, c).
// Used for the three stride-2 3x3 convs of ResNet-50 v1.5 (and as the bring-up path of the stride-1 ones).
// ---------------------------------------------------------------------------------------------------------------
template <class TAct>
__global__ void im2col3x3_kernel(const TAct* __restrict__ in, TAct* __restrict__ col, int B, int H, int W, int C,
                                 int stride, int Ho, int Wo) {
  const int cv = C / 8;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * 9 * cv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % cv);
    size_t t = i / cv;
    const int tap = static_cast<int>(t % 9);
    t /= 9;
    const int wo = static_cast<int>(t % Wo);
    t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    const int y = ho * stride + tap / 3 - 1, x = wo * stride + tap % 3 - 1;
    float v[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (y >= 0 && y < H && x >= 0 && x < W) load8(in + ((static_cast<size_t>(b) * H + y) * W + x) * C + c8 * 8, v);
    store8(col + i * 8, v);
  }
}

// 1x1 stride-2 sampling (downsample branch of the first block of layers 2-4): out[b,ho,wo,:] = in[b,2ho,2wo,:]
template <class TAct>
__global__ void subsample2_kernel(const TAct* __restrict__ in, TAct* __restrict__ out, int B, int H, int W, int C) {
  const int cv = C / 8, Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * cv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % cv);
    size_t t = i / cv;
    const int wo = static_cast<int>(t % Wo);
    t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float v[8];
    load8(in + ((static_cast<size_t>(b) * H + 2 * ho) * W + 2 * wo) * C + c8 * 8, v);
    store8(out + i * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// n3  pre-processing in front of the path (generate_reports_for_images.py:129-147 `get_image_tensor`):
//   uint8 grayscale [H, W] -> cv2.resize(INTER_AREA) to longest side 512 -> centre zero-pad to 512 x 512 ->
//   (x - 0.471*255) * (1 / (0.302*255)) -> fp32 [512, 512].  One thread per output pixel; bit-exact against OpenCV:
//   general (fractional scale) path = computeResizeAreaTab + resizeArea_: per source row the horizontal weighted sum in
//   table order, then the vertical accumulation in table order, both in fp32 WITHOUT fma contraction, cvRound (half to
//   even); integer scales = ResizeAreaFast: integer block sum * (1.f / area) -> cvRound, except 2 x 2 (SIMD path):
//   (sum + 2) >> 2.  Tables are built on the host in double precision exactly like OpenCV (engine.cu).
// ---------------------------------------------------------------------------------------------------------------
struct PreprocTab {
  const int* xoff;    // [nw + 1] CSR offsets into xsi / xal
  const int* xsi;
  const float* xal;
  const int* yoff;    // [nh + 1]
  const int* ysi;
  const float* yal;
  int H, W, nh, nw, top, left;
  int mode;           // 0 general, 1 integer scale (ix, iy), 2 no resize
  int ix, iy;
  float inv_area;     // 1.f / (ix * iy)
  float mean255, denom;
};

__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ src, PreprocTab t, float* __restrict__ out, int S) {
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= S || oy >= S) return;
  const int rx = ox - t.left, ry = oy - t.top;
  int v = 0;  // BORDER_CONSTANT, value 0
  if (rx >= 0 && rx < t.nw && ry >= 0 && ry < t.nh) {
    if (t.mode == 2) {
      v = src[static_cast<size_t>(ry) * t.W + rx];
    } else if (t.mode == 1) {
      int s = 0;
      for (int dy = 0; dy < t.iy; ++dy) {
        const uint8_t* row = src + static_cast<size_t>(ry * t.iy + dy) * t.W + rx * t.ix;
        for (int dx = 0; dx < t.ix; ++dx) s += row[dx];
      }
      v = (t.ix == 2 && t.iy == 2) ? ((s + 2) >> 2) : __float2int_rn(__fmul_rn(static_cast<float>(s), t.inv_area));
    } else {
      const int x0 = t.xoff[rx], x1 = t.xoff[rx + 1];
      float sum = 0.0f;
      bool first = true;
      for (int j = t.yoff[ry]; j < t.yoff[ry + 1]; ++j) {
        const uint8_t* row = src + static_cast<size_t>(t.ysi[j]) * t.W;
        float buf = 0.0f;
        for (int k = x0; k < x1; ++k) buf = __fadd_rn(buf, __fmul_rn(static_cast<float>(row[t.xsi[k]]), t.xal[k]));
        const float term = __fmul_rn(t.yal[j], buf);
        sum = first ? term : __fadd_rn(sum, term);
        first = false;
      }
      v = __float2int_rn(sum);
    }
    v = min(max(v, 0), 255);
  }
  out[static_cast<size_t>(oy) * S + ox] = __fmul_rn(__fsub_rn(static_cast<float>(v), t.mean255), t.denom);
}

// ---------------------------------------------------------------------------------------------------------------
// K4-K6  RPN proposal filtering, one CTA per image, no host sync        (torchvision rpn.py:242-297 filter_proposals)
//   top-k(1000) of the raw objectness (sorted descending, ties -> lowest index)   rpn.py:231-240
//   analytic anchors (anchor_utils.py:58-133) + BoxCoder.decode weights (1,1,1,1), dw/dh clamp ln(1000/16)
//   clip to the image, drop w or h < 1e-3, (score >= 0 always holds)               rpn.py:272-286
//   greedy NMS IoU > 0.7 on the score-ordered boxes, keep first 1000               rpn.py:289-293, boxes.py:20-48
// ---------------------------------------------------------------------------------------------------------------
constexpr int NUM_ANCHORS = 160;
constexpr int TOPK = 1000;
__constant__ float c_base_anchors[NUM_ANCHORS * 4];

struct RpnIn {
  const float* obj;     // objectness logit of (b, pixel, a) at obj[b*obj_bs + pixel*obj_ps + a]
  const float* deltas;  // delta k of (b, pixel, a)        at deltas[b*del_bs + pixel*del_ps + a*4 + k]
  long long obj_bs, del_bs;
  int obj_ps, del_ps;
  const float* decoded;  // optional [B, N, 4]: already-decoded boxes (teacher-forced tests); else null
};
struct RpnOut {
  float* boxes;     // [B, 1000, 4] kept proposals in score order
  float* scores;    // [B, 1000] sigmoid(objectness) of the kept proposals
  int* count;       // [B]
  int* topk_idx;    // optional [B, 1000] anchor indices of the sorted top-k (tests), -1 padded
  int* keep_idx;    // optional [B, 1000] for each kept proposal its rank in the top-k list (tests)
};

__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int RPN_THREADS = 1024;
constexpr size_t RPN_SMEM = 1024 * 8 /*sort*/ + 1024 * 16 /*boxes*/ + 1024 * 4 /*scores*/ + 1024 * 4 /*rank*/ +
                            static_cast<size_t>(TOPK) * 16 * 8 /*nms mask*/ + 64;

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* s_warp, int& total) {
  // v in {0,1}; returns exclusive prefix over the 1024-thread block and the block total
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, v != 0);
  const int in_warp = __popc(bal & ((1u << lane) - 1));
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int x = s_warp[lane];
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    s_warp[lane] = incl - x;
    if (lane == 31) s_warp[32] = incl;
  }
  __syncthreads();
  const int res = s_warp[warp] + in_warp;
  total = s_warp[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(RPN_THREADS) rpn_proposals_kernel(RpnIn in, RpnOut out, int N /*anchors per image*/,
                                                                    int feat /*feature side*/, int image_size,
                                                                    float nms_thresh) {
  extern __shared__ __align__(16) uint8_t smem_rpn[];
  unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(smem_rpn);  // [1024]
  float4* s_box = reinterpret_cast<float4*>(s_sort + 1024);                       // [1024]
  float* s_score = reinterpret_cast<float*>(s_box + 1024);                        // [1024]
  int* s_rank = reinterpret_cast<int*>(s_score + 1024);                           // [1024]
  unsigned long long* s_mask = reinterpret_cast<unsigned long long*>(s_rank + 1024);  // [1000][16]
  __shared__ int s_hist[256];
  __shared__ int s_warp[33];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need;
  __shared__ int s_keep_count;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* obj = in.obj + static_cast<size_t>(b) * in.obj_bs;
  auto obj_at = [&](int n) -> float { return obj[static_cast<size_t>(n / NUM_ANCHORS) * in.obj_ps + n % NUM_ANCHORS]; };
  const int k = N < TOPK ? N : TOPK;

  // ---- radix select: the k-th largest key, 8 bits per pass from the top
  if (tid == 0) {
    s_prefix = 0;
    s_need = k;
  }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += RPN_THREADS) s_hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int n = tid; n < N; n += RPN_THREADS) {
      const uint32_t key = float_key(obj_at(n));
      if ((key & pmask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, bucket = 255;
      for (; bucket > 0; --bucket) {
        if (s_hist[bucket] >= need) break;
        need -= s_hist[bucket];
      }
      s_prefix = prefix | (static_cast<uint32_t>(bucket) << shift);
      s_need = need;
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;  // key of the k-th largest element
  const int need_eq = s_need;     // how many elements equal to thr belong to the top-k (lowest indices first)

  // ---- ordered gather of the k winners (index order), then bitonic sort (key desc, index asc)
  s_sort[tid] = 0ull;
  __syncthreads();
  int run_sel = 0, run_eq = 0;
  for (int n0 = 0; n0 < N; n0 += RPN_THREADS) {
    const int n = n0 + tid;
    uint32_t key = 0;
    int gt = 0, eq = 0;
    if (n < N) {
      key = float_key(obj_at(n));
      gt = key > thr;
      eq = key == thr;
    }
    int tot_eq, tot_sel;
    const int eq_rank = block_exclusive_scan_1024(eq, s_warp, tot_eq);
    const int take = gt | (eq && (run_eq + eq_rank) < need_eq);
    const int pos = block_exclusive_scan_1024(take, s_warp, tot_sel);
    if (take) s_sort[run_sel + pos] = (static_cast<unsigned long long>(key) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(n));
    run_sel += tot_sel;
    run_eq += tot_eq;
  }
  __syncthreads();
  for (int size = 2; size <= 1024; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int partner = tid ^ stride;
      if (partner > tid) {
        const unsigned long long a = s_sort[tid], c = s_sort[partner];
        const bool desc = (tid & size) == 0;
        if (desc ? (a < c) : (a > c)) {
          s_sort[tid] = c;
          s_sort[partner] = a;
        }
      }
      __syncthreads();
    }
  }

  // ---- decode + clip + small-box filter, order-preserving compaction
  int valid = 0;
  float4 box = make_float4(0, 0, 0, 0);
  float score = 0.0f;
  int anchor_idx = -1;
  if (tid < k) {
    const unsigned long long e = s_sort[tid];
    anchor_idx = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(e & 0xFFFFFFFFull));
    const float logit = obj_at(anchor_idx);
    score = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-logit)));
    float x1, y1, x2, y2;
    if (in.decoded) {
      const float* d = in.decoded + (static_cast<size_t>(b) * N + anchor_idx) * 4;
      x1 = d[0]; y1 = d[1]; x2 = d[2]; y2 = d[3];
    } else {
      const int pix = anchor_idx / NUM_ANCHORS, a = anchor_idx % NUM_ANCHORS;
      const float stride = static_cast<float>(image_size / feat);
      const float sx = static_cast<float>(pix % feat) * stride, sy = static_cast<float>(pix / feat) * stride;
      const float ax1 = c_base_anchors[a * 4 + 0] + sx, ay1 = c_base_anchors[a * 4 + 1] + sy;
      const float ax2 = c_base_anchors[a * 4 + 2] + sx, ay2 = c_base_anchors[a * 4 + 3] + sy;
      const float* d = in.deltas + static_cast<size_t>(b) * in.del_bs + static_cast<size_t>(pix) * in.del_ps + a * 4;
      const float clipv = 4.135166556742356f;  // ln(1000/16)
      const float wdt = __fsub_rn(ax2, ax1), hgt = __fsub_rn(ay2, ay1);
      const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, wdt)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, hgt));
      const float dx = d[0], dy = d[1], dw = fminf(d[2], clipv), dh = fminf(d[3], clipv);
      const float pcx = __fadd_rn(__fmul_rn(dx, wdt), cx), pcy = __fadd_rn(__fmul_rn(dy, hgt), cy);
      const float pw = __fmul_rn(expf(dw), wdt), ph = __fmul_rn(expf(dh), hgt);
      const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
      x1 = __fsub_rn(pcx, hw); y1 = __fsub_rn(pcy, hh); x2 = __fadd_rn(pcx, hw); y2 = __fadd_rn(pcy, hh);
    }
    const float lim = static_cast<float>(image_size);
    x1 = fminf(fmaxf(x1, 0.0f), lim); x2 = fminf(fmaxf(x2, 0.0f), lim);
    y1 = fminf(fmaxf(y1, 0.0f), lim); y2 = fminf(fmaxf(y2, 0.0f), lim);
    box = make_float4(x1, y1, x2, y2);
    valid = (__fsub_rn(x2, x1) >= 1e-3f) && (__fsub_rn(y2, y1) >= 1e-3f) && (score >= 0.0f);
  }
  if (out.topk_idx && tid < TOPK) out.topk_idx[static_cast<size_t>(b) * TOPK + tid] = anchor_idx;
  int n_valid;
  const int vpos = block_exclusive_scan_1024(valid, s_warp, n_valid);
  if (valid) {
    s_box[vpos] = box;
    s_score[vpos] = score;
    s_rank[vpos] = tid;
  }
  __syncthreads();

  // ---- NMS suppression bitmask: row i, bit j (j > i) set when IoU(i, j) > thresh
  for (int i = tid; i < n_valid; i += RPN_THREADS) {
    const float4 a = s_box[i];
    const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    for (int w = 0; w < 16; ++w) {
      unsigned long long bits = 0ull;
      const int j0 = w * 64;
      if (j0 + 63 > i) {
        for (int jj = 0; jj < 64; ++jj) {
          const int j = j0 + jj;
          if (j > i && j < n_valid) {
            const float4 c = s_box[j];
            const float xx1 = fmaxf(a.x, c.x), yy1 = fmaxf(a.y, c.y), xx2 = fminf(a.z, c.z), yy2 = fminf(a.w, c.w);
            const float iw = fmaxf(0.0f, __fsub_rn(xx2, xx1)), ih = fmaxf(0.0f, __fsub_rn(yy2, yy1));
            const float inter = __fmul_rn(iw, ih);
            const float area_c = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_c), inter));
            if (ovr > nms_thresh) bits |= (1ull << jj);
          }
        }
      }
      s_mask[static_cast<size_t>(i) * 16 + w] = bits;
    }
  }
  __syncthreads();

  // ---- sequential greedy scan by warp 0 (lanes 0..15 each own one 64-bit word of the removed set)
  if (tid < 32) {
    unsigned long long remv = 0ull;
    int cnt = 0;
    for (int i = 0; i < n_valid; ++i) {
      const unsigned long long word = __shfl_sync(0xffffffffu, remv, i >> 6);
      if (!((word >> (i & 63)) & 1ull)) {
        if (tid == 0 && cnt < TOPK) {
          const float4 bx = s_box[i];
          float* dst = out.boxes + (static_cast<size_t>(b) * TOPK + cnt) * 4;
          dst[0] = bx.x; dst[1] = bx.y; dst[2] = bx.z; dst[3] = bx.w;
          out.scores[static_cast<size_t>(b) * TOPK + cnt] = s_score[i];
          if (out.keep_idx) out.keep_idx[static_cast<size_t>(b) * TOPK + cnt] = s_rank[i];
        }
        ++cnt;
        if (tid < 16) remv |= s_mask[static_cast<size_t>(i) * 16 + tid];
      }
    }
    if (tid == 0) s_keep_count = cnt < TOPK ? cnt : TOPK;
  }
  __syncthreads();
  if (tid == 0) out.count[b] = s_keep_count;
  // pad the unused tail so that downstream kernels never read garbage
  for (int i = s_keep_count + tid; i < TOPK; i += RPN_THREADS) {
    float* dst = out.boxes + (static_cast<size_t>(b) * TOPK + i) * 4;
    dst[0] = dst[1] = dst[2] = dst[3] = 0.0f;
    out.scores[static_cast<size_t>(b) * TOPK + i] = 0.0f;
    if (out.keep_idx) out.keep_idx[static_cast<size_t>(b) * TOPK + i] = -1;
  }
}

// exclusive prefix sum of the per-image proposal counts -> compact RoI row offsets; offsets[B] = total
__global__ void roi_offsets_kernel(const int* __restrict__ count, int* __restrict__ offsets, int B) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < B; ++b) {
      offsets[b] = run;
      run += count[b];
    }
    offsets[B] = run;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K7  RoIAlign 8x8, sampling_ratio 2, aligned=False            (torchvision/ops/roi_align.py:110-190; poolers.py)
// feats bf16 NHWC [B, f, f, C]; out bf16 [rows, 64 bins, C] (the A operand of fc6, K index = bin*C + c).
// The bilinear sum over a bin's 2x2 samples is separable: sum_y sum_x wy*wx*f[y][x]; per RoI the <=4 (row, weight)
// pairs per bin row / column are tabulated once, duplicates merged, and every thread owns 8 channels (16-byte loads).
// ---------------------------------------------------------------------------------------------------------------
struct AxisTaps {
  int n;
  int idx[4];
  float w[4];
};

__device__ __forceinline__ void axis_taps(float start, float bin, int p, int size, AxisTaps& t) {
  t.n = 0;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    float c = start + static_cast<float>(p) * bin + (static_cast<float>(s) + 0.5f) * bin / 2.0f;
    if (c < -1.0f || c > static_cast<float>(size)) continue;  // sample contributes nothing
    if (c <= 0.0f) c = 0.0f;
    int lo = static_cast<int>(c), hi;
    if (lo >= size - 1) {
      hi = lo = size - 1;
      c = static_cast<float>(lo);
    } else {
      hi = lo + 1;
    }
    const float l = c - static_cast<float>(lo), h = 1.0f - l;
    const int ids[2] = {lo, hi};
    const float ws[2] = {h, l};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (ws[e] == 0.0f) continue;
      int f = -1;
      for (int q = 0; q < t.n; ++q)
        if (t.idx[q] == ids[e]) f = q;
      if (f >= 0) t.w[f] += ws[e];
      else {
        t.idx[t.n] = ids[e];
        t.w[t.n] = ws[e];
        ++t.n;
      }
    }
  }
}

template <class TAct>
__global__ void __launch_bounds__(256) roi_align_kernel(const TAct* __restrict__ feats, const float* __restrict__ boxes /*[B,1000,4]*/,
                                                        const int* __restrict__ count, const int* __restrict__ offsets,
                                                        TAct* __restrict__ out, int f, int C, float scale) {
  const int b = blockIdx.y, j = blockIdx.x;
  if (j >= count[b]) return;
  __shared__ AxisTaps s_y[8], s_x[8];
  const float* bx = boxes + (static_cast<size_t>(b) * TOPK + j) * 4;
  if (threadIdx.x < 16) {
    const float x1 = bx[0] * scale, y1 = bx[1] * scale, x2 = bx[2] * scale, y2 = bx[3] * scale;
    const float rw = fmaxf(x2 - x1, 1.0f), rh = fmaxf(y2 - y1, 1.0f);
    if (threadIdx.x < 8) axis_taps(y1, rh / 8.0f, threadIdx.x, f, s_y[threadIdx.x]);
    else axis_taps(x1, rw / 8.0f, threadIdx.x - 8, f, s_x[threadIdx.x - 8]);
  }
  __syncthreads();
  const TAct* fm = feats + static_cast<size_t>(b) * f * f * C;
  TAct* dst = out + static_cast<size_t>(offsets[b] + j) * 64 * C;
  for (int c0 = threadIdx.x * 8; c0 < C; c0 += 256 * 8) {
    for (int bin = 0; bin < 64; ++bin) {
      const AxisTaps& ty = s_y[bin >> 3];
      const AxisTaps& tx = s_x[bin & 7];
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
      for (int iy = 0; iy < ty.n; ++iy) {
        for (int ix = 0; ix < tx.n; ++ix) {
          const float w = ty.w[iy] * tx.w[ix];
          float fv[8];
          load8(fm + (static_cast<size_t>(ty.idx[iy]) * f + tx.idx[ix]) * C + c0, fv);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, fv[e], acc[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= 0.25f;
      store8(dst + static_cast<size_t>(bin) * C + c0, acc);
    }
  }
}

// Separable form of the same RoIAlign (bf16 fast path, f <= 32).  The kernel above re-reads every feature cell once per
// bin that touches it: ~16 L2 -> SM bytes per output byte, which makes it L2-throughput-bound (28.8 k RoIs x 1.5 MB).
// Bilinear weights factor into (row weight) x (column weight), so per bin ROW the vertical interpolation
//   r[x] = sum_iy wy[iy] * feat[y_iy][x]
// is computed ONCE per feature column x the row's eight bins touch (a 4-cell circular window per thread in shared memory:
// a bin's two samples span at most 4 consecutive cells for f <= 32, and the window only moves right), and each bin is
//   out[py][px] = 0.25 * sum_ix wx[ix] * r[x_ix].
// ~3x fewer feature loads; same taps and weights, fp32 throughout, summed in (x outer, y inner) order.
__global__ void __launch_bounds__(256) roi_align_sep_kernel(const bf16* __restrict__ feats, const float* __restrict__ boxes /*[B,1000,4]*/,
                                                            const int* __restrict__ count, const int* __restrict__ offsets,
                                                            bf16* __restrict__ out, int f, int C, float scale) {
  const int b = blockIdx.y, j = blockIdx.x;
  if (j >= count[b]) return;
  __shared__ AxisTaps s_y[8], s_x[8];
  __shared__ float4 s_win[4][2][256];  // [cell & 3][half of the 8 channels][thread]
  const float* bx = boxes + (static_cast<size_t>(b) * TOPK + j) * 4;
  if (threadIdx.x < 16) {
    const float x1 = bx[0] * scale, y1 = bx[1] * scale, x2 = bx[2] * scale, y2 = bx[3] * scale;
    const float rw = fmaxf(x2 - x1, 1.0f), rh = fmaxf(y2 - y1, 1.0f);
    if (threadIdx.x < 8) axis_taps(y1, rh / 8.0f, threadIdx.x, f, s_y[threadIdx.x]);
    else axis_taps(x1, rw / 8.0f, threadIdx.x - 8, f, s_x[threadIdx.x - 8]);
  }
  __syncthreads();
  const bf16* fm = feats + static_cast<size_t>(b) * f * f * C;
  bf16* dst = out + static_cast<size_t>(offsets[b] + j) * 64 * C;
  const int tid = threadIdx.x;
  for (int c0 = tid * 8; c0 < C; c0 += 256 * 8) {
    for (int py = 0; py < 8; ++py) {
      const AxisTaps& ty = s_y[py];
      int have = -1;  // cells <= have (and > have - 4) of this row are in the window
      auto fill_to = [&](int upto) {  // vertical interpolation of cells have+1 .. upto
        for (int x = have + 1; x <= upto; ++x) {
          float r[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] = 0.0f;
          for (int iy = 0; iy < ty.n; ++iy) {
            float fv[8];
            load8(fm + (static_cast<size_t>(ty.idx[iy]) * f + x) * C + c0, fv);
            const float w = ty.w[iy];
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = fmaf(w, fv[e], r[e]);
          }
          s_win[x & 3][0][tid] = make_float4(r[0], r[1], r[2], r[3]);
          s_win[x & 3][1][tid] = make_float4(r[4], r[5], r[6], r[7]);
        }
        have = upto;
      };
      for (int px = 0; px < 8; ++px) {
        const AxisTaps& tx = s_x[px];
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
        if (tx.n > 0 && ty.n > 0) {
          int lo = tx.idx[0], hi = tx.idx[0];
          for (int q = 1; q < tx.n; ++q) {
            lo = min(lo, tx.idx[q]);
            hi = max(hi, tx.idx[q]);
          }
          if (have < lo - 1) have = lo - 1;  // cells left of this bin are never needed again
          if (hi > have) fill_to(hi);
          for (int q = 0; q < tx.n; ++q) {
            const float4 a = s_win[tx.idx[q] & 3][0][tid], c = s_win[tx.idx[q] & 3][1][tid];
            const float w = tx.w[q];
            acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
            acc[4] = fmaf(w, c.x, acc[4]); acc[5] = fmaf(w, c.y, acc[5]); acc[6] = fmaf(w, c.z, acc[6]); acc[7] = fmaf(w, c.w, acc[7]);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= 0.25f;
        store8(dst + static_cast<size_t>(py * 8 + px) * C + c0, acc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K10  per-class top-1 region selection, one CTA per image                          (custom_roi_heads.py:63-208)
//   softmax(30) -> drop background -> argmax over 29 (first max) -> per class: max score over the RoIs that predict
//   it, lowest RoI index on ties; class_detected = (#RoIs predicting the class) > 0; undetected class -> index 0, score 0
//   boxes: BoxCoder.decode weights (10,10,5,5) of the class' own deltas, clipped (roi_heads.py:542-544)
// ---------------------------------------------------------------------------------------------------------------
struct RoiTailOut {
  uint8_t* detected;  // [B,29]
  int* top_idx;       // [B,29] index into the image's proposal list
  float* top_scores;  // [B,29]
  float* top_boxes;   // [B,29,4]
};

__global__ void __launch_bounds__(256) roi_tail_kernel(const float* __restrict__ cls, int cls_ld, const float* __restrict__ reg,
                                                       int reg_ld, const float* __restrict__ boxes /*[B,1000,4]*/,
                                                       const int* __restrict__ count, const int* __restrict__ offsets,
                                                       RoiTailOut out, int image_size) {
  __shared__ unsigned long long s_best[29];
  const int b = blockIdx.x;
  const int P = count[b], off = offsets[b];
  if (threadIdx.x < 29) s_best[threadIdx.x] = 0ull;
  __syncthreads();
  for (int j = threadIdx.x; j < P; j += 256) {
    const float* l = cls + static_cast<size_t>(off + j) * cls_ld;
    float v[30];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 30; ++c) {
      v[c] = l[c];
      m = fmaxf(m, v[c]);
    }
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < 30; ++c) {
      v[c] = expf(v[c] - m);
      sum += v[c];
    }
    int best = 1;
    float bv = __fdiv_rn(v[1], sum);
#pragma unroll
    for (int c = 2; c < 30; ++c) {
      const float p = __fdiv_rn(v[c], sum);
      if (p > bv) {
        bv = p;
        best = c;
      }
    }
    // score >= 0 -> the IEEE bit pattern is order-preserving; the low word prefers the lowest RoI index on ties
    // and is never 0 (j < 2^32 - 1), so key != 0 <=> at least one RoI predicts the class.
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(bv)) << 32) |
                                   static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<uint32_t>(j));
    atomicMax(&s_best[best - 1], key);
  }
  __syncthreads();
  if (threadIdx.x < 29) {
    const int c = threadIdx.x;
    const unsigned long long key = s_best[c];
    const bool det = key != 0ull;
    const int idx = det ? static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull)) : 0;
    const float score = det ? __uint_as_float(static_cast<uint32_t>(key >> 32)) : 0.0f;
    out.detected[b * 29 + c] = det ? 1 : 0;
    out.top_idx[b * 29 + c] = idx;
    out.top_scores[b * 29 + c] = score;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    if (P > 0) {
      const float* pb = boxes + (static_cast<size_t>(b) * TOPK + idx) * 4;
      const float* d = reg + static_cast<size_t>(off + idx) * reg_ld + (c + 1) * 4;
      const float clipv = 4.135166556742356f;
      const float wdt = __fsub_rn(pb[2], pb[0]), hgt = __fsub_rn(pb[3], pb[1]);
      const float cx = __fadd_rn(pb[0], __fmul_rn(0.5f, wdt)), cy = __fadd_rn(pb[1], __fmul_rn(0.5f, hgt));
      const float dx = __fdiv_rn(d[0], 10.0f), dy = __fdiv_rn(d[1], 10.0f);
      const float dw = fminf(__fdiv_rn(d[2], 5.0f), clipv), dh = fminf(__fdiv_rn(d[3], 5.0f), clipv);
      const float pcx = __fadd_rn(__fmul_rn(dx, wdt), cx), pcy = __fadd_rn(__fmul_rn(dy, hgt), cy);
      const float pw = __fmul_rn(expf(dw), wdt), ph = __fmul_rn(expf(dh), hgt);
      const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
      const float lim = static_cast<float>(image_size);
      x1 = fminf(fmaxf(__fsub_rn(pcx, hw), 0.0f), lim);
      y1 = fminf(fmaxf(__fsub_rn(pcy, hh), 0.0f), lim);
      x2 = fminf(fmaxf(__fadd_rn(pcx, hw), 0.0f), lim);
      y2 = fminf(fmaxf(__fadd_rn(pcy, hh), 0.0f), lim);
    }
    float* ob = out.top_boxes + (b * 29 + c) * 4;
    ob[0] = x1; ob[1] = y1; ob[2] = x2; ob[3] = y2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K11  AvgPool2d(8) of the RoIAlign map of the 29 winning proposals        (custom_roi_heads.py:253-258, :141-159)
// RoIAlign is linear, so mean over the 8x8 bins == weighted sum over feature cells with separable weights
// WY[row] * WX[col] (each axis: 16 sample points x 2 taps / 16).  fp32 output [B*29, C].
// ---------------------------------------------------------------------------------------------------------------
template <class TAct>
__global__ void __launch_bounds__(256) roi_mean_kernel(const TAct* __restrict__ feats, const float* __restrict__ boxes,
                                                       const int* __restrict__ count, const int* __restrict__ top_idx,
                                                       float* __restrict__ out, int f, int C, float scale) {
  const int c = blockIdx.x, b = blockIdx.y;
  __shared__ float s_wy[64], s_wx[64];  // f <= 64
  float* dst = out + static_cast<size_t>(b * 29 + c) * C;
  if (count[b] == 0) {
    for (int i = threadIdx.x; i < C; i += 256) dst[i] = 0.0f;
    return;
  }
  const float* bx = boxes + (static_cast<size_t>(b) * TOPK + top_idx[b * 29 + c]) * 4;
  if (threadIdx.x < 128) {
    const int axis = threadIdx.x >> 6, cell = threadIdx.x & 63;  // axis 0 = y, 1 = x
    const float lo = bx[axis == 0 ? 1 : 0] * scale, hi = bx[axis == 0 ? 3 : 2] * scale;
    const float bin = fmaxf(hi - lo, 1.0f) / 8.0f;
    float w = 0.0f;
    if (cell < f) {
      for (int p = 0; p < 8; ++p) {
        AxisTaps t;
        axis_taps(lo, bin, p, f, t);
        for (int q = 0; q < t.n; ++q)
          if (t.idx[q] == cell) w += t.w[q];
      }
    }
    (axis == 0 ? s_wy : s_wx)[cell] = w * (1.0f / 16.0f);
  }
  __syncthreads();
  const TAct* fm = feats + static_cast<size_t>(b) * f * f * C;
  for (int c0 = threadIdx.x * 8; c0 < C; c0 += 256 * 8) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    for (int y = 0; y < f; ++y) {
      const float wy = s_wy[y];
      if (wy == 0.0f) continue;
      for (int x = 0; x < f; ++x) {
        const float w = wy * s_wx[x];
        if (w == 0.0f) continue;
        float fv[8];
        load8(fm + (static_cast<size_t>(y) * f + x) * C + c0, fv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, fv[e], acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[c0 + e] = acc[e];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// K12  tail of the selection classifier: Linear(128 -> 1), `logit > -1`, AND class_detected, row compaction
//      (binary_classifier_region_selection.py:32, :53-61).  One CTA; rows = B*29 (image-major, region-minor).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) selection_tail_kernel(const float* __restrict__ hidden /*[rows,128]*/,
                                                              const float* __restrict__ w /*[128]*/, const float* __restrict__ bias,
                                                              const uint8_t* __restrict__ detected, float* __restrict__ logits,
                                                              uint8_t* __restrict__ selected, int* __restrict__ sel_rows,
                                                              int* __restrict__ num_selected, int rows) {
  __shared__ int s_warp[33];
  int run = 0;
  for (int r0 = 0; r0 < rows; r0 += 1024) {
    const int r = r0 + threadIdx.x;
    int sel = 0;
    if (r < rows) {
      float acc = 0.0f;
      const float* h = hidden + static_cast<size_t>(r) * 128;
      for (int k = 0; k < 128; ++k) acc = fmaf(h[k], w[k], acc);
      acc += bias[0];
      logits[r] = acc;
      // selection: logit > -1 AND class_detected (binary_classifier_region_selection.py:53-57); abnormal (detected == null):
      // logit > -1 only, undetected regions are masked later by the caller (binary_classifier_region_abnormal.py:53-57)
      sel = (acc > -1.0f) && (detected == nullptr || detected[r]);
      selected[r] = static_cast<uint8_t>(sel);
    }
    int tot;
    const int pos = block_exclusive_scan_1024(sel, s_warp, tot);
    if (sel && sel_rows) sel_rows[run + pos] = r;
    run += tot;
  }
  if (threadIdx.x == 0 && num_selected) *num_selected = run;
}

// gather the selected rows' fp32 features as the bf16 A operand of the decoder's first GEMM; rows [n, round_up(n, 32))
// are zero-filled (the decoder runs on a padded row count so that its step graph is reused across batches)
__global__ void gather_rows_bf16_kernel(const float* __restrict__ src, const int* __restrict__ rows, const int* __restrict__ n_rows,
                                        bf16* __restrict__ dst, int D) {
  const int r = blockIdx.x;
  const int n = *n_rows;
  if (r >= ((n + 31) & ~31)) return;
  if (r >= n) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[static_cast<size_t>(r) * D + i] = f2bf(0.0f);
    return;
  }
  const float* s = src + static_cast<size_t>(rows[r]) * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) dst[static_cast<size_t>(r) * D + i] = f2bf(s[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// weight repacking (load time): fp32 checkpoint layouts -> bf16 / fp32 K-major GEMM operands
// ---------------------------------------------------------------------------------------------------------------
// out[n, r*Cin + c] = in[n, c*R + r] * scale[n]     (conv OIHW -> O(HW)I with folded BN scale; fc6 (c,bin) -> (bin,c))
template <class TOut>
__global__ void repack_oihw_kernel(const float* __restrict__ in, TOut* __restrict__ out, const float* __restrict__ scale,
                                   long long N, int Cin, int R) {
  const long long K = static_cast<long long>(Cin) * R;
  const long long total = N * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / K;
    const int kk = static_cast<int>(i % K);
    const int r = kk / Cin, c = kk % Cin;
    float v = in[n * K + static_cast<long long>(c) * R + r];
    if (scale) v *= scale[n];
    if constexpr (sizeof(TOut) == 2) out[i] = f2bf(v);
    else out[i] = v;
  }
}
// out[n, k] = in[k, n]   (HF Conv1D weight [K, N] -> K-major [N, K])
__global__ void repack_transpose_kernel(const float* __restrict__ in, bf16* __restrict__ out, int K, int N) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? in[static_cast<size_t>(k) * N + n] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) out[static_cast<size_t>(n) * K + k] = f2bf(tile[threadIdx.x][i]);
  }
}
__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = f2bf(in[i]);
}
// Result blob of one rank (layout of rgrg_b200/parallel.py pack_result): [R, width] int32 | ids int32 [rows, T] padded with
// EOS | selected u8 [rows] | detected u8 [rows] | boxes f32 [rows, 4] | scores f32 [rows]; rows = B * 29.  Byte-wise copies:
// the tail segments are not 4-byte aligned when rows is odd.
__global__ void pack_blob_kernel(uint8_t* __restrict__ blob, const int* __restrict__ ids, int ids_ld, int R, int width, int rows, int T,
                                 const uint8_t* __restrict__ selected, const uint8_t* __restrict__ detected,
                                 const float* __restrict__ boxes, const float* __restrict__ scores) {
  const size_t n_ids = static_cast<size_t>(rows) * T;
  const size_t o_sel = 8 + n_ids * 4, o_det = o_sel + rows, o_box = o_det + rows, o_sc = o_box + static_cast<size_t>(rows) * 16;
  const size_t total = o_sc + static_cast<size_t>(rows) * 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint8_t v;
    if (i < 8) {
      const int head = i < 4 ? R : width;
      v = static_cast<uint8_t>(static_cast<unsigned>(head) >> (8 * (i & 3)));
    } else if (i < o_sel) {
      const size_t e = (i - 8) >> 2;
      const int r = static_cast<int>(e / T), c = static_cast<int>(e % T);
      const int tok = (r < R && c < width) ? ids[static_cast<size_t>(r) * ids_ld + c] : 50256;
      v = static_cast<uint8_t>(static_cast<unsigned>(tok) >> (8 * ((i - 8) & 3)));
    } else if (i < o_det) {
      v = selected[i - o_sel];
    } else if (i < o_box) {
      v = detected[i - o_det];
    } else if (i < o_sc) {
      v = reinterpret_cast<const uint8_t*>(boxes)[i - o_box];
    } else {
      v = reinterpret_cast<const uint8_t*>(scores)[i - o_sc];
    }
    blob[i] = v;
  }
}

// out[i] = bias[i % N] + sum_s parts[s][i]   (test harness of the split-K GEMM form)
__global__ void sum_parts_kernel(const float* __restrict__ parts, int nparts, long long total, const float* __restrict__ bias, int N,
                                 float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v = bias ? bias[i % N] : 0.0f;
    for (int s = 0; s < nparts; ++s) v += parts[s * total + i];
    out[i] = v;
  }
}
// folded BN: scale = gamma / sqrt(var + eps), bias = beta - mean * scale
__global__ void bn_fold_kernel(const float* g, const float* b, const float* mean, const float* var, float* scale, float* bias,
                               int C, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float s = g[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    bias[i] = b[i] - mean[i] * s;
  }
}
// conv1 weights [64,1,7,7] -> [49][64] with folded scale
__global__ void repack_stem_kernel(const float* in, const float* scale, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 49 * 64) {
    const int k = i / 64, c = i % 64;
    out[i] = in[c * 49 + k] * scale[c];
  }
}

}  // namespace det
}  // namespace rgrg
