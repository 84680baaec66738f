// rgrg_b200 engine: weights, workspace, orchestration of the generate() path and the C ABI (include/rgrg_b200.h).
// Reference call stack being replaced: SURVEY.md §3b-§3c (report_generation_model.py:212-276 ->
// object_detector.py:184-261 -> custom_rpn.py:53-85 / custom_roi_heads.py:210-269 ->
// binary_classifier_region_selection.py:24-68 -> language_model.py:401-479, :609-652).
#include <cuda.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rgrg_b200.h"
#include "common.cuh"
#include "attn_fused.cuh"
#include "decoder_kernels.cuh"
#include "detector_kernels.cuh"
#include "epilogues.cuh"
#include "gemm_simt.cuh"
#include "gemm_2cta.cuh"
#include "gemm_tc.cuh"

using namespace rgrg;

namespace {

constexpr int NLAYER = 24;
constexpr int DM = 1024;
constexpr int VOCAB = 50257;
constexpr int NREG = 29;
constexpr int TOPK = det::TOPK;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  // grow-only; the new block is allocated BEFORE the old one is freed, so a failed growth leaves the buffer usable
  void ensure(size_t n) {
    if (n <= bytes) return;
    void* np = nullptr;
    CUDA_CHECK(cudaMalloc(&np, n));
    if (p) cudaFree(p);
    p = np;
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostRef {
  const float* p;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// K-major bf16 weight [N, K] + fp32 bias, with TMA maps for both N-tile widths
// epilogue types that can leave through gemm_2cta.cuh's TMA-store path (EpiStoreT without a residual)
template <class E, class = void>
struct TmaOutOk : std::false_type {};
template <class E>
struct TmaOutOk<E, std::void_t<decltype(E::kTmaOut)>> : std::bool_constant<E::kTmaOut> {};

template <class E, class = void>
struct TmaOutRowOk : std::false_type {};
template <class E>
struct TmaOutRowOk<E, std::void_t<decltype(E::kTmaOutRow)>> : std::bool_constant<E::kTmaOutRow> {};

struct Linear {
  bf16* w = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0;
  CUtensorMap tm[4];  // TMA maps for N-tile widths 64 / 128 / 192 / 256
  float* w32 = nullptr;  // fp32 twin [N, K] (detector_precise parity mode only)
  void make_maps() {
    for (int i = 0; i < 4; ++i) tm[i] = tc::make_tmap_2d(w, N, K, 64 * (i + 1));
  }
};
struct LinearF32 {
  float* w = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0;
};

struct BlockW {
  Linear c1, c2, c3, ds;
  int cin, width, cout, stride;
  bool has_ds;
};

struct LayerW {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  Linear attn, proj, fc, mproj;
  CUtensorMap tm_qkv;  // c_attn weight as [3][1024][1024]: q | k | v rows of one head in one box (attn_fused.cuh)
};

}  // namespace

// ---- NCCL, resolved at run time from the library already loaded into the process (torch's bundled libnccl.so.2) or the
// system one: the engine has no link-time dependency on NCCL and builds on a box without it.
namespace nccl {
struct UniqueId {
  char internal[128];
};
typedef void* Comm;
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(Comm*, int, UniqueId, int);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, Comm, cudaStream_t);
typedef int (*CommDestroyFn)(Comm);
typedef const char* (*GetErrorStringFn)(int);
struct Api {
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllGatherFn all_gather = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
};
inline const Api& api() {
  static Api a;
  static bool loaded = false;
  if (!loaded) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy torch.distributed already loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) throw std::runtime_error("NCCL not found: libnccl.so.2 is neither loaded nor on the library path");
    a.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
    a.all_gather = reinterpret_cast<AllGatherFn>(dlsym(h, "ncclAllGather"));
    a.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
    a.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
    if (!a.get_unique_id || !a.comm_init_rank || !a.all_gather || !a.comm_destroy) throw std::runtime_error("NCCL symbols missing");
    loaded = true;
  }
  return a;
}
inline void check(int rc, const char* what) {
  if (rc != 0) {
    const Api& a = api();
    throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (a.error_string ? a.error_string(rc) : std::to_string(rc).c_str()));
  }
}
}  // namespace nccl

struct rgrg_engine;
struct ProfScope {
  rgrg_engine* e;
  cudaStream_t st;
  int rec;
  ProfScope(rgrg_engine* e, const char* tag, cudaStream_t st);
  ~ProfScope();
};

struct rgrg_engine {
  // ---- optional per-category device timing (CUDA events on the launch stream; bench.py's roofline source)
  struct ProfRec {
    int cat;
    cudaEvent_t a, b;
  };
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_pool;
  size_t prof_used = 0;
  std::vector<ProfRec> prof_recs;
  std::vector<std::string> prof_names;
  std::map<std::string, int> prof_ids;
  cudaEvent_t prof_event() {
    if (prof_used == prof_pool.size()) {
      cudaEvent_t ev;
      CUDA_CHECK(cudaEventCreate(&ev));
      prof_pool.push_back(ev);
    }
    return prof_pool[prof_used++];
  }
  int prof_cat(const char* tag) {
    auto it = prof_ids.find(tag);
    if (it != prof_ids.end()) return it->second;
    const int id = static_cast<int>(prof_names.size());
    prof_names.push_back(tag);
    prof_ids[tag] = id;
    return id;
  }
  std::string prof_report() {
    CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<double> ms(prof_names.size(), 0.0);
    std::vector<long long> cnt(prof_names.size(), 0);
    for (const ProfRec& r : prof_recs) {
      float t = 0.0f;
      CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
      ms[r.cat] += t;
      cnt[r.cat] += 1;
    }
    std::string out;
    for (size_t i = 0; i < prof_names.size(); ++i) {
      char line[256];
      snprintf(line, sizeof(line), "%s %.6f %lld\n", prof_names[i].c_str(), ms[i], cnt[i]);
      out += line;
    }
    return out;
  }
  void prof_reset() {
    prof_recs.clear();
    prof_used = 0;
  }

  // ---- result gather across ranks: one ncclAllGather of fixed-size result blobs, device to device
  nccl::Comm comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  DevBuf blob_dev, gather_dev;
  int last_R = 0, last_width = 0, last_T = 0;  // geometry of the ids the last generate() left in `ids` (device)

  int device = 0;
  // all work runs on an engine-owned non-blocking stream (the legacy default stream cannot be captured into a CUDA
  // graph); it is ordered after the caller's stream on entry, and every entry point host-synchronises before returning
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_enter = nullptr;
  cudaStream_t enter(void* caller_stream) {
    if (!own_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_enter, cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventRecord(ev_enter, static_cast<cudaStream_t>(caller_stream)));
    CUDA_CHECK(cudaStreamWaitEvent(own_stream, ev_enter, 0));
    return own_stream;
  }
  std::string err;
  int64_t launches = 0;
  bool weights_ready = false;
  bool precise_weights = false;  // fp32 twins of the detector weights were built (detector_precise was set before finalize)
  int opt_implicit_conv = 1;
  int opt_cuda_graph = 1;
  int opt_gemm_impl = 0;
  int opt_pdl = 1;
  int opt_detector_precise = 0;  // fp32 detector (parity mode, see run_detect)
  int opt_fused_attn = 1;  // greedy decode: c_attn + KV append + attention as ONE head-aligned kernel (attn_fused.cuh)
  int opt_ln_head = 0;     // LayerNorm (+ split-K reduce + residual) as the cluster-cooperative head of the consumer GEMM
  int opt_trace = 0;            // tuning: per-kernel timestamps of the decode step (first / last CTA), read with debug "decode_trace"
  DevBuf trace_buf;             // [slots][2][8] int64
  int trace_slot = 0;
  long long* trace_ptr() {
    if (!opt_trace) return nullptr;
    trace_buf.ensure(256 * 16 * 8);
    if (trace_slot >= 256) return nullptr;
    return trace_buf.as<long long>() + static_cast<size_t>(trace_slot++) * 16;
  }
  int opt_attn_balance = 1;     // fused attention: rows spread evenly over (#SMs / 16) M tiles instead of 128-row tiles
  int opt_gemm_2cta_waves = 2;  // the pair kernel is used while its grid fits in this many waves (else the persistent 1-CTA kernel)
  int opt_attn_mc = 0;          // fused attention: head pairs (clusters of 2) share operand A through TMA multicast
  int opt_attn_early = 0;       // fused attention: request the first K / V chunks before the epilogue
  int opt_epi_tma_conv = 1;     // ... also for the implicit 3x3 convolutions (4-D output boxes)
  int opt_epi_tma = 1;          // CTA-pair kernel: plain epilogues leave through shared-memory slabs + TMA stores
  int opt_gemm_2cta_stages = 6; // its TMA ring depth: 6 (one CTA per SM) / 4 / 3 (two CTAs per SM: prologue overlaps the predecessor's epilogue)
  int opt_gemm_2cta = 1;        // decode projections (c_proj / c_fc / mlp c_proj) through the CTA-pair kernel (gemm_2cta.cuh)
  int opt_dual = 0;             // greedy decode step as two row halves half a layer out of phase (decode_forward_dual)
  int opt_roi_align_sep = 1;    // RoIAlign in separable form (vertical interpolation once per feature column of a bin row)
  int opt_beam_fused_head = 1;  // beam search: log-softmax + per-part top-k fused into the lm_head epilogue (0: fp32 logits in HBM)
  int opt_attn_warps = 16;  // fused attention: attention / epilogue warps per CTA
  int opt_attn_slots = 2;  // fused attention: shared-memory K/V ring slots per warp
  int opt_l2_ahead = 0;    // fused attention: items whose K/V blocks are prefetched into L2 ahead of the consumer
  int opt_cattn_bn = 0;    // tuning: force the N tile of c_attn (0 = pick_bn); two-kernel attention path only
  int opt_attn_occ = 7;    // CTAs per SM the stand-alone attention kernel is compiled for (5: 88 regs, 6: 78, 7: 72, 8: 64)
  int opt_ablate = 0;  // tuning only: bit mask of decode-step kernels to skip (results become meaningless, timing attributes cost)
  bool pdl_now = false;  // set while the decode step is being issued: its kernels carry the PDL launch attribute
  bool force_2cta = false;    // GEMM test entry: the pair kernel whatever the problem size
  bool use_2cta_now = false;  // set while the decode projections are issued (and by the GEMM test entry): CTA-pair tiles
  std::unordered_map<std::string, HostRef> host;
  std::vector<void*> weight_allocs;

  // ---- weights
  float *stem_w = nullptr, *stem_b = nullptr;
  std::vector<BlockW> blocks;
  Linear rpn_conv, rpn_heads, fc6, fc7, pred;
  LinearF32 dimred, sel0, sel2, sel4, abn0, abn2, abn4;
  bool has_abnormal = false;
  Linear fst0, fst2, ukv, lm_head;
  float* wte_f32 = nullptr;
  LayerW layers[NLAYER];
  float *lnf_g = nullptr, *lnf_b = nullptr;

  // ---- workspace (detector)
  int ws_B = 0, ws_S = 0;
  DevBuf images, act[2], t1, t2, idb, sub, col, feats, rpn_t, rpn_out;
  DevBuf prop_boxes, prop_scores, prop_count, roi_off, pooled, f6, f7, pred_out;
  DevBuf detected, top_idx, top_scores, top_boxes, mean2048, trf, s0, s1, sel_logits, selected, sel_rows, num_sel, abn_logits, abnormal;
  DevBuf lm_in;
  // ---- workspace (decoder)
  int ws_rows = 0, ws_slots = 0;
  DevBuf splitk_parts, ln_counters;
  DevBuf kv_cache, h, x, q, attn_o, mlp_mid, a1, img, part_val, part_idx, ids, unfinished, unf_count, step, logits_tmp;
  int last_B = 0, last_S = 0, last_P = 0;
  // beam search: cache-slot ancestry of the current step (null in greedy mode)
  const unsigned char* beam_anc = nullptr;
  int beam_slots = 0, beam_nb = 1;

  // ---- CUDA graph of one decode step, keyed by row count
  std::map<int, cudaGraphExec_t> step_graphs;
  std::map<int, int> step_graph_nodes;
  std::map<int, long long> beam_graph_sig;  // buffer addresses a beam graph captured

  ~rgrg_engine() {
    if (comm) nccl::api().comm_destroy(comm);
    for (auto& g : step_graphs) cudaGraphExecDestroy(g.second);
    for (cudaEvent_t ev : prof_pool) cudaEventDestroy(ev);
    if (ev_enter) cudaEventDestroy(ev_enter);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (cudaEvent_t ev : ev_a) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ev_b) cudaEventDestroy(ev);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (own_stream) cudaStreamDestroy(own_stream);
    for (void* p : weight_allocs) cudaFree(p);
    for (auto& kv : preproc_tabs) kv.second.buf.release();
    DevBuf* all[] = {&images, &act[0], &act[1], &t1, &t2, &idb, &sub, &col, &feats, &rpn_t, &rpn_out, &prop_boxes,
                     &prop_scores, &prop_count, &roi_off, &pooled, &f6, &f7, &pred_out, &detected, &top_idx,
                     &top_scores, &top_boxes, &mean2048, &trf, &s0, &s1, &sel_logits, &selected, &sel_rows, &num_sel, &abn_logits, &abnormal,
                     &lm_in, &kv_cache, &h, &x, &q, &attn_o, &mlp_mid, &a1, &img, &part_val, &part_idx, &ids,
                     &unfinished, &unf_count, &step, &logits_tmp, &splitk_parts, &ln_counters, &trace_buf, &blob_dev, &gather_dev, &preproc_src, &preproc_out, &p_act[0], &p_act[1], &p_t1, &p_t2, &p_idb, &p_sub, &p_col, &p_c1, &p_feats, &p_rpn_t, &p_pooled, &p_f6, &p_f7, &b_ids2, &b_anc[0], &b_anc[1], &b_scores, &b_cand_score,
                     &b_cand_token, &b_cand_beam, &b_hyp_score, &b_hyp_len, &b_hyp_tok, &b_hyp_count, &b_worst, &b_done, &b_not_done, &b_part_m, &b_part_l, &b_part_val, &b_part_idx};
    for (DevBuf* b : all) b->release();
  }

  template <class T>
  T* walloc(size_t n) {
    void* p = nullptr;
    CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
    weight_allocs.push_back(p);
    return reinterpret_cast<T*>(p);
  }

  // ================================================================================================================
  // GEMM dispatch
  // ================================================================================================================
  template <class Epi>
  void simt_f32(const float* A, const float* Wt, int M, int N, int K, const Epi& ep, cudaStream_t st) {
    simt::launch<float, float, Epi>(A, Wt, M, N, K, ep, st);
  }

  // CTA-pair kernel; plain bias / activation epilogues (c_fc, the split-K partial sums) leave through TMA stores
  template <class Epi>
  void launch_pair(const CUtensorMap& tmA2, const Linear& W, const tc2::Shape& s2, const Epi& epi, cudaStream_t st) {
    if constexpr (TmaOutOk<Epi>::value) {
      const int splits = s2.k_splits > 1 ? s2.k_splits : 1;
      const bool layout_ok = epi.ldc == W.N && (splits == 1 || epi.split_stride == static_cast<size_t>(s2.M) * W.N);
      if (opt_epi_tma && opt_gemm_2cta_stages >= 4 && layout_ok) {
        const CUtensorMap tmC = tc2::make_tmap_out(epi.out, s2.M, W.N, splits, Epi::kOutBf16);
        if (opt_gemm_2cta_stages == 4) tc2::launch<Epi, 4, true>(tmA2, W.tm[1], tmC, s2, epi, st, pdl_now);
        else tc2::launch<Epi, 6, true>(tmA2, W.tm[1], tmC, s2, epi, st, pdl_now);
        return;
      }
    }
    if (opt_gemm_2cta_stages == 3) tc2::launch<Epi, 3>(tmA2, W.tm[1], tmA2, s2, epi, st, pdl_now);
    else if (opt_gemm_2cta_stages == 4) tc2::launch<Epi, 4>(tmA2, W.tm[1], tmA2, s2, epi, st, pdl_now);
    else tc2::launch<Epi, 6>(tmA2, W.tm[1], tmA2, s2, epi, st, pdl_now);
  }

  template <class Epi>
  void gemm(const char* tag, const bf16* A, int M, const Linear& W, const Epi& epi, cudaStream_t st, bool m_fastest,
            int force_bn = 0) {
    if (M <= 0) return;
    ProfScope ps(this, tag, st);
    if (opt_gemm_impl == 2) {
      if constexpr (std::is_same<Epi, EpiArgmaxPartial>::value) {
        throw std::runtime_error("arg-max epilogue has no CUDA-core variant");
      } else {
        simt::launch<bf16, bf16, Epi>(A, W.w, M, W.N, W.K, epi, st);
        ++launches;
        return;
      }
    }
    if (W.K % 64 != 0) throw std::runtime_error("GEMM K must be a multiple of 64");
    const int mt = ceil_div(M, tc::BM);
    if constexpr (!Epi::kDirect) {
      // decode projections: CTA pairs on 256 x 256 tiles (gemm_2cta.cuh) while one wave of pairs covers the problem; beyond that
      // the persistent 1-CTA kernel (epilogue of tile i overlapped with the main loop of tile i+1) is the better shape
      if (use_2cta_now && W.N % tc2::BN == 0 && (force_2cta || 2 * ceil_div(mt, 2) * (W.N / tc2::BN) <= opt_gemm_2cta_waves * tc::num_sms())) {
        tc2::Shape s2{};
        s2.M = M;
        s2.N = W.N;
        s2.k_iters = W.K / 64;
        s2.m_pairs = ceil_div(mt, 2);
        s2.n_tiles = W.N / tc2::BN;
        s2.trace = trace_ptr();
        CUtensorMap tmA2 = tc::make_tmap_2d(A, M, W.K, 128);
        launch_pair(tmA2, W, s2, epi, st);
        ++launches;
        return;
      }
    }
    const int bn = force_bn ? force_bn : pick_bn(mt, W.N);
    tc::GemmShape s{};
    s.M = M;
    s.N = W.N;
    s.k_iters = W.K / 64;
    s.m_tiles = mt;
    s.n_tiles = ceil_div(W.N, bn);
    s.m_fastest = m_fastest ? 1 : 0;
    CUtensorMap tmA = tc::make_tmap_2d(A, M, W.K, 128);
    launch_bn(bn, tmA, W, s, epi, st);
    ++launches;
  }

  // split-K GEMM for the two residual projections of a decoder layer (M = rows is only ~8 tiles tall): partial sums of
  // slice s go to parts[s] (fp32 [M, N]); the LayerNorm that follows adds them (+ bias) into the residual stream.
  void gemm_splitk(const char* tag, const bf16* A, int M, const Linear& W, float* parts, int splits, cudaStream_t st) {
    if (M <= 0) return;
    ProfScope ps(this, tag, st);
    auto ep = epi<false, ACT_NONE, RES_NONE, false>(parts, nullptr, W.N);
    ep.split_stride = static_cast<size_t>(M) * W.N;
    if (opt_gemm_impl == 2) {  // CUDA-core cross-check: one full-K pass into slice 0, zeros elsewhere
      CUDA_CHECK(cudaMemsetAsync(parts, 0, static_cast<size_t>(splits) * M * W.N * 4, st));
      simt::launch<bf16, bf16, decltype(ep)>(A, W.w, M, W.N, W.K, ep, st);
      ++launches;
      return;
    }
    if ((W.K / 64) % splits) throw std::runtime_error("split-K factor must divide K / 64");
    if (use_2cta_now && W.N % tc2::BN == 0 && (force_2cta || 2 * ceil_div(ceil_div(M, tc::BM), 2) * (W.N / tc2::BN) * splits <= opt_gemm_2cta_waves * tc::num_sms())) {
      tc2::Shape s2{};
      s2.M = M;
      s2.N = W.N;
      s2.k_iters = W.K / 64;
      s2.k_splits = splits;
      s2.m_pairs = ceil_div(ceil_div(M, tc::BM), 2);
      s2.n_tiles = W.N / tc2::BN;
      s2.trace = trace_ptr();
      CUtensorMap tmA2 = tc::make_tmap_2d(A, M, W.K, 128);
      launch_pair(tmA2, W, s2, ep, st);
      ++launches;
      return;
    }
    tc::GemmShape s{};
    s.M = M;
    s.N = W.N;
    s.k_iters = W.K / 64;
    s.k_splits = splits;
    s.m_tiles = ceil_div(M, tc::BM);
    s.n_tiles = ceil_div(W.N, 256);
    s.m_fastest = 1;
    CUtensorMap tmA = tc::make_tmap_2d(A, M, W.K, 128);
    launch_bn(256, tmA, W, s, ep, st);
    ++launches;
  }

  // GEMM with a LayerNorm head (decoder c_fc): N must be 16 tiles of 256; launched as clusters of 16 CTAs
  template <class Epi>
  void gemm_ln_head(const char* tag, const bf16* A, int M, const Linear& W, const Epi& epi, const tc::GemmShape::LnHead& lnh,
                    cudaStream_t st) {
    if (M <= 0) return;
    ProfScope ps(this, tag, st);
    if (W.N != 16 * 256 || W.K % 64) throw std::runtime_error("LayerNorm-head GEMM needs N = 4096");
    tc::GemmShape s{};
    s.M = M;
    s.N = W.N;
    s.k_iters = W.K / 64;
    s.m_tiles = ceil_div(M, tc::BM);
    s.n_tiles = 16;
    s.m_fastest = 0;  // tile -> (m_blk = tile / 16, n_blk = tile % 16): a cluster is one M tile
    s.lnh = lnh;
    CUtensorMap tmA = tc::make_tmap_2d(A, M, W.K, 128);
    tc::launch<256, 4, Epi, true>(tmA, W.tm[3], s, epi, st, pdl_now);
    ++launches;
  }

  // N-tile width: fewest waves x per-k-block cycles.  Measured on B200 (profiles/r01_v2_gemm_timeline.md): a
  // 128 x N x 16 tcgen05.mma occupies the tensor pipe ~128 cycles for every N in {64..256}, so a k-block costs
  // ~520 cycles regardless of the tile width and narrow tiles only help by shortening the tail wave.
  static int pick_bn(int m_tiles, int N) {
    int best = 128;
    long long best_cost = -1;
    for (int bn = 64; bn <= 256; bn += 64) {
      const long long tiles = static_cast<long long>(m_tiles) * ceil_div(N, bn);
      const long long waves = (tiles + tc::num_sms() - 1) / tc::num_sms();
      const long long cyc = 520 + bn / 2;             // + epilogue share, which grows with the tile
      const long long cost = waves * (cyc * 16 + 600);  // + fixed per-tile overhead (pipeline fill / drain)
      if (best_cost < 0 || cost <= best_cost) {
        best_cost = cost;
        best = bn;
      }
    }
    return best;
  }
  template <class Epi>
  void launch_bn(int bn, const CUtensorMap& tmA, const Linear& W, const tc::GemmShape& s, const Epi& epi, cudaStream_t st) {
    // plain GEMMs with a bf16 output (1x1 convolutions, fc6 / fc7): TMA-store epilogue (gemm_tc.cuh)
    if constexpr (TmaOutRowOk<Epi>::value) {
      const bool conv_ok = !s.conv || (!Epi::kResBf16 && opt_epi_tma_conv);  // implicit conv: 4-D NHWC output boxes, no residual variant
      if (opt_epi_tma && conv_ok && s.k_splits <= 1 && (bn == 128 || bn == 256) && epi.ldc == s.N && s.N % 64 == 0) {
        const CUtensorMap tmC = s.conv ? tc::make_tmap_out_nhwc(epi.out, s.M / (static_cast<uint64_t>(s.H) * s.W), s.H, s.W, s.N)
                                       : tc::make_tmap_out_bf16(epi.out, s.M, s.N);
        CUtensorMap tmR = tmC;  // the bf16 residual, if any, comes in through the same boxes
        if constexpr (Epi::kResBf16) tmR = tc::make_tmap_out_bf16(epi.res, s.M, s.N);
        if (bn == 128) tc::launch<128, 6, Epi, false, true>(tmA, W.tm[1], s, epi, st, pdl_now, &tmC, &tmR);
        else tc::launch<256, 4, Epi, false, true>(tmA, W.tm[3], s, epi, st, pdl_now, &tmC, &tmR);
        return;
      }
    }
    switch (bn) {
      case 64: tc::launch<64, 8, Epi>(tmA, W.tm[0], s, epi, st, pdl_now); break;
      case 128: tc::launch<128, 6, Epi>(tmA, W.tm[1], s, epi, st, pdl_now); break;
      case 192: tc::launch<192, 5, Epi>(tmA, W.tm[2], s, epi, st, pdl_now); break;
      case 256: tc::launch<256, 4, Epi>(tmA, W.tm[3], s, epi, st, pdl_now); break;
      default: throw std::runtime_error("unsupported N tile");
    }
  }

  // 3x3 / stride 1 / pad 1 conv as implicit GEMM through a 4-D tensor map (K loop = 9 taps x Cin/64)
  template <class Epi>
  void conv3x3_implicit(const char* tag, const bf16* in, int B, int H, int Wd, int Cin, const Linear& W, const Epi& epi,
                        cudaStream_t st) {
    ProfScope ps(this, tag, st);
    if (H % 8 || Wd % 16 || Cin % 64) throw std::runtime_error("implicit conv needs H%8==0, W%16==0, Cin%64==0");
    tc::GemmShape s{};
    s.M = B * H * Wd;
    s.N = W.N;
    s.conv = 1;
    s.kc_blocks = Cin / 64;
    s.k_iters = 9 * s.kc_blocks;
    s.H = H;
    s.W = Wd;
    s.tiles_w = Wd / 16;
    s.tiles_h = H / 8;
    s.m_tiles = B * s.tiles_w * s.tiles_h;
    const int bn = pick_bn(s.m_tiles, W.N);
    s.n_tiles = ceil_div(W.N, bn);
    s.m_fastest = 0;
    CUtensorMap tmA = tc::make_tmap_nhwc(in, B, H, Wd, Cin);
    launch_bn(bn, tmA, W, s, epi, st);
    ++launches;
  }

  void im2col(const bf16* in, bf16* colbuf, int B, int H, int Wd, int C, int stride, cudaStream_t st) {
    ProfScope ps(this, "im2col", st);
    const int Ho = H / stride, Wo = Wd / stride;
    const size_t total = static_cast<size_t>(B) * Ho * Wo * 9 * (C / 8);
    const int grid = static_cast<int>(std::min<size_t>(ceil_div64(total, 256), 148 * 16));
    det::im2col3x3_kernel<bf16><<<grid, 256, 0, st>>>(in, colbuf, B, H, Wd, C, stride, Ho, Wo);
    KERNEL_CHECK();
    ++launches;
  }

  template <class Epi>
  void conv3x3(const char* tag, const bf16* in, int B, int H, int Wd, int Cin, int stride, const Linear& W, const Epi& epi,
               cudaStream_t st) {
    if (stride == 1 && opt_implicit_conv && opt_gemm_impl != 2) {
      conv3x3_implicit(tag, in, B, H, Wd, Cin, W, epi, st);
    } else {
      const int Ho = H / stride, Wo = Wd / stride;
      col.ensure(static_cast<size_t>(B) * Ho * Wo * 9 * Cin * 2);
      im2col(in, col.as<bf16>(), B, H, Wd, Cin, stride, st);
      gemm(tag, col.as<bf16>(), B * Ho * Wo, W, epi, st, false);
    }
  }

  template <bool OUT_BF16, int ACT, int RES, bool BIAS>
  static EpiStoreT<OUT_BF16, ACT, RES, BIAS> epi(void* out, const float* bias, int ldc, const void* res = nullptr) {
    EpiStoreT<OUT_BF16, ACT, RES, BIAS> e{};
    e.out = out;
    e.bias = bias;
    e.res = res;
    e.ldc = ldc;
    e.split_stride = 0;
    return e;
  }

  // ================================================================================================================
  // weights
  // ================================================================================================================
  const HostRef& need(const std::string& name) {
    auto it = host.find(name);
    if (it == host.end()) throw std::runtime_error("missing weight: " + name);
    return it->second;
  }
  const HostRef& need_any(const std::string& a, const std::string& b) {
    auto it = host.find(a);
    if (it != host.end()) return it->second;
    return need(b);
  }

  DevBuf stage, stage2;
  const float* upload(const HostRef& r, DevBuf& buf) {
    buf.ensure(static_cast<size_t>(r.numel()) * 4);
    CUDA_CHECK(cudaMemcpy(buf.p, r.p, static_cast<size_t>(r.numel()) * 4, cudaMemcpyHostToDevice));
    return buf.as<float>();
  }
  float* upload_keep(const HostRef& r) {
    float* d = walloc<float>(r.numel());
    CUDA_CHECK(cudaMemcpy(d, r.p, static_cast<size_t>(r.numel()) * 4, cudaMemcpyHostToDevice));
    return d;
  }
  static int grid_for(long long n) { return static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 32)); }

  // conv weight [Cout, Cin, k, k] (+ optional BN fold) -> bf16 [Cout, k*k*Cin]; returns Linear with bias
  Linear make_conv(const std::string& conv, const std::string& bn, bool has_conv_bias) {
    const HostRef& w = need(conv + ".weight");
    const int cout = static_cast<int>(w.shape[0]), cin = static_cast<int>(w.shape[1]);
    const int R = static_cast<int>(w.shape[2] * w.shape[3]);
    Linear L;
    L.N = cout;
    L.K = cin * R;
    L.w = walloc<bf16>(static_cast<size_t>(L.N) * L.K);
    L.bias = walloc<float>(cout);
    const float* dw = upload(w, stage);
    float* scale = nullptr;
    if (!bn.empty()) {
      float* g = upload_tmp(need(bn + ".weight"));
      float* b = upload_tmp(need(bn + ".bias"));
      float* m = upload_tmp(need(bn + ".running_mean"));
      float* v = upload_tmp(need(bn + ".running_var"));
      scale = tmp_alloc(cout);
      det::bn_fold_kernel<<<ceil_div(cout, 256), 256>>>(g, b, m, v, scale, L.bias, cout, 1e-5f);
      KERNEL_CHECK();
    } else if (has_conv_bias) {
      CUDA_CHECK(cudaMemcpy(L.bias, need(conv + ".bias").p, cout * 4, cudaMemcpyHostToDevice));
    } else {
      CUDA_CHECK(cudaMemset(L.bias, 0, cout * 4));
    }
    det::repack_oihw_kernel<bf16><<<grid_for(static_cast<long long>(L.N) * L.K), 256>>>(dw, L.w, scale, L.N, cin, R);
    KERNEL_CHECK();
    if (opt_detector_precise) {
      L.w32 = walloc<float>(static_cast<size_t>(L.N) * L.K);
      det::repack_oihw_kernel<float><<<grid_for(static_cast<long long>(L.N) * L.K), 256>>>(dw, L.w32, scale, L.N, cin, R);
      KERNEL_CHECK();
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    free_tmps();
    L.make_maps();
    return L;
  }
  std::vector<void*> tmps;
  float* tmp_alloc(size_t n) {
    void* p;
    CUDA_CHECK(cudaMalloc(&p, n * 4));
    tmps.push_back(p);
    return reinterpret_cast<float*>(p);
  }
  float* upload_tmp(const HostRef& r) {
    float* d = tmp_alloc(r.numel());
    CUDA_CHECK(cudaMemcpy(d, r.p, static_cast<size_t>(r.numel()) * 4, cudaMemcpyHostToDevice));
    return d;
  }
  void free_tmps() {
    for (void* p : tmps) cudaFree(p);
    tmps.clear();
  }

  // nn.Linear weight [N, K] (already K-major) -> bf16 rows [row0, row0+N) of dst
  void put_linear_rows(const std::string& name, bf16* dst_w, float* dst_b, int row0, int K, float* dst_w32 = nullptr) {
    const HostRef& w = need(name + ".weight");
    const int N = static_cast<int>(w.shape[0]);
    if (w.numel() != static_cast<int64_t>(N) * K) throw std::runtime_error("bad shape for " + name);
    const float* dw = upload(w, stage);
    det::cast_bf16_kernel<<<grid_for(w.numel()), 256>>>(dw, dst_w + static_cast<size_t>(row0) * K, w.numel());
    KERNEL_CHECK();
    if (dst_w32) CUDA_CHECK(cudaMemcpy(dst_w32 + static_cast<size_t>(row0) * K, dw, static_cast<size_t>(w.numel()) * 4, cudaMemcpyDeviceToDevice));
    CUDA_CHECK(cudaMemcpy(dst_b + row0, need(name + ".bias").p, N * 4, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaDeviceSynchronize());
  }
  Linear make_linear(const std::string& name) {
    const HostRef& w = need(name + ".weight");
    Linear L;
    L.N = static_cast<int>(w.shape[0]);
    L.K = static_cast<int>(w.numel() / w.shape[0]);
    L.w = walloc<bf16>(static_cast<size_t>(L.N) * L.K);
    L.bias = walloc<float>(L.N);
    if (opt_detector_precise && name.find("object_detector") == 0) L.w32 = walloc<float>(static_cast<size_t>(L.N) * L.K);
    put_linear_rows(name, L.w, L.bias, 0, L.K, L.w32);
    L.make_maps();
    return L;
  }
  // HF Conv1D weight [K, N] -> K-major bf16 [N, K]
  Linear make_conv1d(const std::string& name) {
    const HostRef& w = need(name + ".weight");
    Linear L;
    L.K = static_cast<int>(w.shape[0]);
    L.N = static_cast<int>(w.shape[1]);
    L.w = walloc<bf16>(static_cast<size_t>(L.N) * L.K);
    L.bias = upload_keep(need(name + ".bias"));
    const float* dw = upload(w, stage);
    dim3 grid(ceil_div(L.N, 32), ceil_div(L.K, 32)), block(32, 8);
    det::repack_transpose_kernel<<<grid, block>>>(dw, L.w, L.K, L.N);
    KERNEL_CHECK();
    CUDA_CHECK(cudaDeviceSynchronize());
    L.make_maps();
    return L;
  }
  LinearF32 make_linear_f32(const std::string& name) {
    const HostRef& w = need(name + ".weight");
    LinearF32 L;
    L.N = static_cast<int>(w.shape[0]);
    L.K = static_cast<int>(w.numel() / w.shape[0]);
    L.w = upload_keep(w);
    L.bias = upload_keep(need(name + ".bias"));
    return L;
  }

  void finalize() {
    CUDA_CHECK(cudaSetDevice(device));
    const std::string bb = "object_detector.backbone";
    // ---- stem (object_detector.py:54: conv1 replaced by a 1-channel 7x7)
    {
      const HostRef& w = need(bb + ".0.weight");
      if (w.numel() != 64 * 49) throw std::runtime_error("conv1 must be [64,1,7,7]");
      float* dw = upload_tmp(w);
      float* g = upload_tmp(need(bb + ".1.weight"));
      float* b = upload_tmp(need(bb + ".1.bias"));
      float* m = upload_tmp(need(bb + ".1.running_mean"));
      float* v = upload_tmp(need(bb + ".1.running_var"));
      float* scale = tmp_alloc(64);
      stem_w = walloc<float>(49 * 64);
      stem_b = walloc<float>(64);
      det::bn_fold_kernel<<<1, 64>>>(g, b, m, v, scale, stem_b, 64, 1e-5f);
      det::repack_stem_kernel<<<ceil_div(49 * 64, 256), 256>>>(dw, scale, stem_w);
      KERNEL_CHECK();
      CUDA_CHECK(cudaDeviceSynchronize());
      free_tmps();
    }
    // ---- 16 bottlenecks (torchvision resnet.py Bottleneck, v1.5: stride on conv2)
    blocks.clear();
    const int nblk[4] = {3, 4, 6, 3}, widths[4] = {64, 128, 256, 512};
    int cin = 64;
    for (int li = 0; li < 4; ++li) {
      for (int bi = 0; bi < nblk[li]; ++bi) {
        const std::string p = bb + "." + std::to_string(4 + li) + "." + std::to_string(bi);
        BlockW bw;
        bw.cin = cin;
        bw.width = widths[li];
        bw.cout = widths[li] * 4;
        bw.stride = (bi == 0 && li > 0) ? 2 : 1;
        bw.has_ds = bi == 0;
        bw.c1 = make_conv(p + ".conv1", p + ".bn1", false);
        bw.c2 = make_conv(p + ".conv2", p + ".bn2", false);
        bw.c3 = make_conv(p + ".conv3", p + ".bn3", false);
        if (bw.has_ds) bw.ds = make_conv(p + ".downsample.0", p + ".downsample.1", false);
        blocks.push_back(bw);
        cin = bw.cout;
      }
    }
    // ---- RPN head (torchvision rpn.py RPNHead); old checkpoints name the conv "conv" instead of "conv.0.0"
    const std::string rp = "object_detector.rpn.head";
    if (host.count(rp + ".conv.0.0.weight")) rpn_conv = make_conv(rp + ".conv.0.0", "", true);
    else rpn_conv = make_conv(rp + ".conv", "", true);
    {
      rpn_heads.N = 800;
      rpn_heads.K = 2048;
      rpn_heads.w = walloc<bf16>(800 * 2048);
      rpn_heads.bias = walloc<float>(800);
      if (opt_detector_precise) rpn_heads.w32 = walloc<float>(800 * 2048);
      put_linear_rows(rp + ".cls_logits", rpn_heads.w, rpn_heads.bias, 0, 2048, rpn_heads.w32);
      put_linear_rows(rp + ".bbox_pred", rpn_heads.w, rpn_heads.bias, 160, 2048, rpn_heads.w32);
      rpn_heads.make_maps();
    }
    // ---- RoI heads: fc6 input is flattened (c, ph, pw) in the reference; our RoIAlign emits (bin, c)
    const std::string rh = "object_detector.roi_heads";
    {
      const HostRef& w = need(rh + ".box_head.fc6.weight");
      fc6.N = 1024;
      fc6.K = 2048 * 64;
      if (w.numel() != static_cast<int64_t>(fc6.N) * fc6.K) throw std::runtime_error("fc6 must be [1024, 131072]");
      fc6.w = walloc<bf16>(static_cast<size_t>(fc6.N) * fc6.K);
      fc6.bias = upload_keep(need(rh + ".box_head.fc6.bias"));
      const float* dw = upload(w, stage);
      det::repack_oihw_kernel<bf16><<<grid_for(w.numel()), 256>>>(dw, fc6.w, nullptr, fc6.N, 2048, 64);
      KERNEL_CHECK();
      if (opt_detector_precise) {
        fc6.w32 = walloc<float>(static_cast<size_t>(fc6.N) * fc6.K);
        det::repack_oihw_kernel<float><<<grid_for(w.numel()), 256>>>(dw, fc6.w32, nullptr, fc6.N, 2048, 64);
        KERNEL_CHECK();
      }
      CUDA_CHECK(cudaDeviceSynchronize());
      fc6.make_maps();
    }
    fc7 = make_linear(rh + ".box_head.fc7");
    {
      pred.N = 150;
      pred.K = 1024;
      pred.w = walloc<bf16>(150 * 1024);
      pred.bias = walloc<float>(150);
      if (opt_detector_precise) pred.w32 = walloc<float>(150 * 1024);
      put_linear_rows(rh + ".box_predictor.cls_score", pred.w, pred.bias, 0, 1024, pred.w32);
      put_linear_rows(rh + ".box_predictor.bbox_pred", pred.w, pred.bias, 30, 1024, pred.w32);
      pred.make_maps();
    }
    dimred = make_linear_f32(rh + ".dim_reduction");
    sel0 = make_linear_f32("binary_classifier_region_selection.classifier.0");
    sel2 = make_linear_f32("binary_classifier_region_selection.classifier.2");
    sel4 = make_linear_f32("binary_classifier_region_selection.classifier.4");
    has_abnormal = host.count("binary_classifier_region_abnormal.classifier.0.weight") != 0;
    if (has_abnormal) {
      abn0 = make_linear_f32("binary_classifier_region_abnormal.classifier.0");
      abn2 = make_linear_f32("binary_classifier_region_abnormal.classifier.2");
      abn4 = make_linear_f32("binary_classifier_region_abnormal.classifier.4");
    }
    // ---- language model (canonical alias set: language_model.gpt2_blocks.* etc.)
    const std::string lm = "language_model";
    fst0 = make_linear(lm + ".feature_space_transformation_nn.0");
    fst2 = make_linear(lm + ".feature_space_transformation_nn.2");
    ukv.N = NLAYER * 2 * DM;
    ukv.K = DM;
    ukv.w = walloc<bf16>(static_cast<size_t>(ukv.N) * DM);
    ukv.bias = walloc<float>(ukv.N);
    for (int l = 0; l < NLAYER; ++l) {
      const std::string p = lm + ".gpt2_blocks." + std::to_string(l);
      LayerW& L = layers[l];
      L.ln1_g = upload_keep(need(p + ".0.weight"));
      L.ln1_b = upload_keep(need(p + ".0.bias"));
      L.ln2_g = upload_keep(need(p + ".2.weight"));
      L.ln2_b = upload_keep(need(p + ".2.bias"));
      L.attn = make_conv1d(p + ".1.c_attn");
      if (L.attn.N != 3 * DM || L.attn.K != DM) throw std::runtime_error("c_attn must be [1024, 3072]");
      L.tm_qkv = fa::make_tmap_qkv(L.attn.w);
      L.proj = make_conv1d(p + ".1.c_proj");
      L.fc = make_conv1d(p + ".3.c_fc");
      L.mproj = make_conv1d(p + ".3.c_proj");
      put_linear_rows(p + ".1.uk", ukv.w, ukv.bias, l * 2048, DM);
      put_linear_rows(p + ".1.uv", ukv.w, ukv.bias, l * 2048 + 1024, DM);
    }
    ukv.make_maps();
    lnf_g = upload_keep(need(lm + ".final_layernorm.weight"));
    lnf_b = upload_keep(need(lm + ".final_layernorm.bias"));
    {
      const HostRef& w = need_any(lm + ".wte.weight", lm + ".lm_head.weight");
      if (w.numel() != static_cast<int64_t>(VOCAB) * DM) throw std::runtime_error("wte must be [50257,1024]");
      wte_f32 = upload_keep(w);
      lm_head.N = VOCAB;
      lm_head.K = DM;
      lm_head.w = walloc<bf16>(static_cast<size_t>(VOCAB) * DM);
      lm_head.bias = nullptr;
      det::cast_bf16_kernel<<<grid_for(w.numel()), 256>>>(wte_f32, lm_head.w, w.numel());
      KERNEL_CHECK();
      lm_head.make_maps();
    }
    // ---- base anchors (torchvision anchor_utils.py generate_anchors: ratio-major, size-minor, round half-even)
    {
      const float sizes[10] = {20, 40, 60, 80, 100, 120, 140, 160, 180, 300};
      const float ratios[16] = {0.2f, 0.25f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f, 0.9f, 1.0f, 1.3f, 1.5f, 2.1f, 2.6f, 3.0f, 5.0f, 8.0f};
      float base[160 * 4];
      for (int r = 0; r < 16; ++r) {
        const float hr = sqrtf(ratios[r]);
        const float wr = 1.0f / hr;
        for (int s = 0; s < 10; ++s) {
          const float ws = wr * sizes[s], hs = hr * sizes[s];
          float* a = base + (r * 10 + s) * 4;
          a[0] = nearbyintf(-ws / 2.0f);
          a[1] = nearbyintf(-hs / 2.0f);
          a[2] = nearbyintf(ws / 2.0f);
          a[3] = nearbyintf(hs / 2.0f);
        }
      }
      CUDA_CHECK(cudaMemcpyToSymbol(det::c_base_anchors, base, sizeof(base)));
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    stage.release();
    stage2.release();
    host.clear();
    CUDA_CHECK(cudaFuncSetAttribute(det::rpn_proposals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(det::RPN_SMEM)));
    weights_ready = true;
    precise_weights = opt_detector_precise != 0;
  }

  // ================================================================================================================
  // pre-processing (n3): generate_reports_for_images.py:129-147
  // ================================================================================================================
  struct PreprocEntry {
    det::PreprocTab tab;
    DevBuf buf;  // one allocation holding all six tables
  };
  std::map<std::pair<int, int>, PreprocEntry> preproc_tabs;
  DevBuf preproc_src, preproc_out;

  // OpenCV computeResizeAreaTab (modules/imgproc/src/resize.cpp), CSR form: entries of dst index d are [off[d], off[d+1])
  static void area_tab(int ssize, int dsize, std::vector<int>& off, std::vector<int>& si, std::vector<float>& al) {
    const double scale = 1.0 / (static_cast<double>(dsize) / ssize);
    off.assign(1, 0);
    for (int dx = 0; dx < dsize; ++dx) {
      const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
      const double cell = std::min(scale, ssize - fsx1);
      int sx1 = static_cast<int>(std::ceil(fsx1)), sx2 = static_cast<int>(std::floor(fsx2));
      sx2 = std::min(sx2, ssize - 1);
      sx1 = std::min(sx1, sx2);
      if (sx1 - fsx1 > 1e-3) {
        si.push_back(sx1 - 1);
        al.push_back(static_cast<float>((sx1 - fsx1) / cell));
      }
      for (int sx = sx1; sx < sx2; ++sx) {
        si.push_back(sx);
        al.push_back(static_cast<float>(1.0 / cell));
      }
      if (fsx2 - sx2 > 1e-3) {
        si.push_back(sx2);
        al.push_back(static_cast<float>(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell));
      }
      off.push_back(static_cast<int>(si.size()));
    }
  }
  static int py3round(double v) { return static_cast<int>(std::nearbyint(v)); }  // round half to even (default FP mode)

  const det::PreprocTab& preproc_table(int H, int W, int S) {
    auto key = std::make_pair(H, W);
    auto it = preproc_tabs.find(key);
    if (it != preproc_tabs.end()) return it->second.tab;
    if (std::max(H, W) < S) throw std::runtime_error("pre-processing handles down-scaling only (longest side >= 512)");
    PreprocEntry& en = preproc_tabs[key];
    det::PreprocTab& t = en.tab;
    memset(&t, 0, sizeof(t));
    // albumentations 1.1.0 longest_max_size: scale = max_size / max(h, w); dims = py3round(dim * scale); no-op if scale == 1
    const double sc = static_cast<double>(S) / std::max(H, W);
    t.H = H;
    t.W = W;
    t.nh = sc == 1.0 ? H : py3round(H * sc);
    t.nw = sc == 1.0 ? W : py3round(W * sc);
    // PadIfNeeded (position = center): top = int((min - rows) / 2.0)
    t.top = t.nh < S ? static_cast<int>((S - t.nh) / 2.0) : 0;
    t.left = t.nw < S ? static_cast<int>((S - t.nw) / 2.0) : 0;
    // Normalize: mean * 255 and reciprocal(std * 255) in float32
    t.mean255 = 0.471f * 255.0f;
    t.denom = 1.0f / (0.302f * 255.0f);
    std::vector<int> xoff, xsi, yoff, ysi;
    std::vector<float> xal, yal;
    if (t.nh == H && t.nw == W) {
      t.mode = 2;
    } else {
      const double sx = 1.0 / (static_cast<double>(t.nw) / W), sy = 1.0 / (static_cast<double>(t.nh) / H);
      const int ix = static_cast<int>(std::nearbyint(sx)), iy = static_cast<int>(std::nearbyint(sy));
      if (std::abs(sx - ix) < 2.220446049250313e-16 && std::abs(sy - iy) < 2.220446049250313e-16) {
        t.mode = 1;
        t.ix = ix;
        t.iy = iy;
        t.inv_area = 1.0f / static_cast<float>(ix * iy);
      } else {
        t.mode = 0;
        area_tab(W, t.nw, xoff, xsi, xal);
        area_tab(H, t.nh, yoff, ysi, yal);
      }
    }
    if (t.mode == 0) {
      const size_t n_i = xoff.size() + xsi.size() + yoff.size() + ysi.size();
      const size_t n_f = xal.size() + yal.size();
      en.buf.ensure((n_i + n_f) * 4);
      int* di = en.buf.as<int>();
      auto put_i = [&](const std::vector<int>& v) {
        CUDA_CHECK(cudaMemcpy(di, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
        int* r = di;
        di += v.size();
        return r;
      };
      t.xoff = put_i(xoff);
      t.xsi = put_i(xsi);
      t.yoff = put_i(yoff);
      t.ysi = put_i(ysi);
      float* df = reinterpret_cast<float*>(di);
      CUDA_CHECK(cudaMemcpy(df, xal.data(), xal.size() * 4, cudaMemcpyHostToDevice));
      t.xal = df;
      df += xal.size();
      CUDA_CHECK(cudaMemcpy(df, yal.data(), yal.size() * 4, cudaMemcpyHostToDevice));
      t.yal = df;
    }
    return t;
  }

  // ================================================================================================================
  // detector
  // ================================================================================================================
  void ensure_detector_ws(int B, int S) {
    if (B <= ws_B && S <= ws_S) return;
    B = std::max(B, ws_B);
    S = std::max(S, ws_S);
    const size_t P = S / 4, f = S / 32;
    images.ensure(static_cast<size_t>(B) * S * S * 4);
    const size_t big = static_cast<size_t>(B) * P * P * 256 * 2;  // layer1 output is the largest activation
    act[0].ensure(big);
    act[1].ensure(big);
    t1.ensure(static_cast<size_t>(B) * P * P * 128 * 2);  // layer2.0.conv1 runs at full layer-1 resolution
    t2.ensure(static_cast<size_t>(B) * P * P * 64 * 2);
    idb.ensure(big);
    sub.ensure(big / 4);
    feats.ensure(static_cast<size_t>(B) * f * f * 2048 * 2);
    rpn_t.ensure(static_cast<size_t>(B) * f * f * 2048 * 2);
    rpn_out.ensure(static_cast<size_t>(B) * f * f * 800 * 4);
    prop_boxes.ensure(static_cast<size_t>(B) * TOPK * 4 * 4);
    prop_scores.ensure(static_cast<size_t>(B) * TOPK * 4);
    prop_count.ensure(static_cast<size_t>(B) * 4);
    roi_off.ensure(static_cast<size_t>(B + 1) * 4);
    const size_t rows = static_cast<size_t>(B) * NREG;
    detected.ensure(rows);
    top_idx.ensure(rows * 4);
    top_scores.ensure(rows * 4);
    top_boxes.ensure(rows * 16);
    mean2048.ensure(rows * 2048 * 4);
    trf.ensure(rows * 1024 * 4);
    s0.ensure(rows * 512 * 4);
    s1.ensure(rows * 128 * 4);
    sel_logits.ensure(rows * 4);
    abn_logits.ensure(rows * 4);
    abnormal.ensure(rows);
    selected.ensure(rows);
    sel_rows.ensure(rows * 4);
    num_sel.ensure(4);
    lm_in.ensure((rows + 32) * 1024 * 2);
    ws_B = B;
    ws_S = S;
  }
  void ensure_roi_ws(int P_total) {
    const size_t rows = static_cast<size_t>(std::max(P_total, 1));
    pooled.ensure(rows * 131072 * 2);
    f6.ensure(rows * 1024 * 2);
    f7.ensure(rows * 1024 * 2);
    pred_out.ensure(rows * 150 * 4);
  }

  // images (device fp32 [B,1,S,S]) -> feats bf16 NHWC
  void run_backbone(const float* img_dev, int B, int S, bf16* out_feats, cudaStream_t st) {
    if (S % 128 != 0) throw std::runtime_error("image size must be a multiple of 128");
    const int P = S / 4;
    {
      ProfScope ps(this, "stem", st);
      det::stem_kernel<<<dim3(P / 8, P / 8, B), 256, 0, st>>>(img_dev, stem_w, stem_b, act[0].as<bf16>(), S);
      KERNEL_CHECK();
      ++launches;
    }
    int cur = 0, H = P;
    for (size_t i = 0; i < blocks.size(); ++i) {
      const BlockW& bw = blocks[i];
      const bf16* xin = act[cur].as<bf16>();
      const bool last = (i + 1 == blocks.size());
      bf16* yout = last ? out_feats : act[cur ^ 1].as<bf16>();
      const int Ho = H / bw.stride;
      const int M_in = B * H * H, M_out = B * Ho * Ho;
      // conv1 1x1 + BN + ReLU
      gemm("conv1x1", xin, M_in, bw.c1, epi<true, ACT_RELU, RES_NONE, true>(t1.p, bw.c1.bias, bw.width), st, false);
      // conv2 3x3 (stride here, v1.5) + BN + ReLU
      conv3x3("conv3x3", t1.as<bf16>(), B, H, H, bw.width, bw.stride, bw.c2, epi<true, ACT_RELU, RES_NONE, true>(t2.p, bw.c2.bias, bw.width), st);
      // identity / downsample branch
      const bf16* identity = xin;
      if (bw.has_ds) {
        const bf16* ds_in = xin;
        if (bw.stride == 2) {
          ProfScope ps(this, "subsample", st);
          const size_t total = static_cast<size_t>(M_out) * (bw.cin / 8);
          det::subsample2_kernel<bf16><<<static_cast<int>(std::min<size_t>(ceil_div64(total, 256), 148 * 16)), 256, 0, st>>>(
              xin, sub.as<bf16>(), B, H, H, bw.cin);
          KERNEL_CHECK();
          ++launches;
          ds_in = sub.as<bf16>();
        }
        gemm("conv1x1", ds_in, M_out, bw.ds, epi<true, ACT_NONE, RES_NONE, true>(idb.p, bw.ds.bias, bw.cout), st, false);
        identity = idb.as<bf16>();
      }
      // conv3 1x1 + BN, + identity, ReLU
      gemm("conv1x1", t2.as<bf16>(), M_out, bw.c3, epi<true, ACT_RELU, RES_BF16, true>(yout, bw.c3.bias, bw.cout, identity), st, false);
      cur ^= 1;
      H = Ho;
    }
  }

  void run_rpn_filter(const det::RpnIn& in, const det::RpnOut& out, int B, int feat, int S, cudaStream_t st) {
    ProfScope ps(this, "rpn_topk_nms", st);
    det::rpn_proposals_kernel<<<B, det::RPN_THREADS, det::RPN_SMEM, st>>>(in, out, feat * feat * det::NUM_ANCHORS, feat, S, 0.7f);
    KERNEL_CHECK();
    ++launches;
  }

  // ---- fp32 parity path of the detector ("detector_precise"): the same layer sequence with fp32 operands, fp32 NHWC
  // activations and fp32 accumulation on CUDA cores, so that the decision tensors (objectness, deltas, class logits) carry
  // reference-grade arithmetic and the per-class top-1 region INDICES can be compared with the oracle (SURVEY.md §7/§8(d):
  // the bf16 tensor-core path flips arg-maxes that are decided below bf16 resolution).  Not a product path: ~100x slower.
  DevBuf p_act[2], p_t1, p_t2, p_idb, p_sub, p_col, p_c1, p_feats, p_rpn_t, p_pooled, p_f6, p_f7;
  template <int ACT, int RES>
  void gemm_f32(const float* A, int M, const Linear& W, float* out, const float* res, cudaStream_t st) {
    if (!W.w32) throw std::runtime_error("detector_precise must be set before the weights are finalized");
    simt::launch<float, float, EpiStoreT<false, ACT, RES, true>>(A, W.w32, M, W.N, W.K, epi<false, ACT, RES, true>(out, W.bias, W.N, res), st);
    ++launches;
  }
  const float* conv3x3_f32_cols(const float* in, int B, int H, int C, int stride, cudaStream_t st) {
    const int Ho = H / stride;
    p_col.ensure(static_cast<size_t>(B) * Ho * Ho * 9 * C * 4);
    const size_t total = static_cast<size_t>(B) * Ho * Ho * 9 * (C / 8);
    det::im2col3x3_kernel<float><<<static_cast<int>(std::min<size_t>(ceil_div64(total, 256), 148 * 16)), 256, 0, st>>>(
        in, p_col.as<float>(), B, H, H, C, stride, Ho, Ho);
    KERNEL_CHECK();
    ++launches;
    return p_col.as<float>();
  }
  void run_backbone_rpn_precise(const float* img_dev, int B, int S, cudaStream_t st) {
    const int P = S / 4, f = S / 32;
    const size_t big = static_cast<size_t>(B) * P * P * 256 * 4;
    p_c1.ensure(static_cast<size_t>(B) * (S / 2) * (S / 2) * 64 * 4);
    p_act[0].ensure(big);
    p_act[1].ensure(big);
    p_t1.ensure(big / 2);
    p_t2.ensure(big / 4);
    p_idb.ensure(big);
    p_sub.ensure(big / 4);
    p_feats.ensure(static_cast<size_t>(B) * f * f * 2048 * 4);
    p_rpn_t.ensure(static_cast<size_t>(B) * f * f * 2048 * 4);
    det::stem_conv_f32_kernel<<<148 * 8, 256, 0, st>>>(img_dev, stem_w, stem_b, p_c1.as<float>(), B, S);
    det::maxpool3x3s2_f32_kernel<<<148 * 8, 256, 0, st>>>(p_c1.as<float>(), p_act[0].as<float>(), B, S / 2);
    KERNEL_CHECK();
    launches += 2;
    int cur = 0, H = P;
    for (size_t i = 0; i < blocks.size(); ++i) {
      const BlockW& bw = blocks[i];
      const float* xin = p_act[cur].as<float>();
      float* yout = (i + 1 == blocks.size()) ? p_feats.as<float>() : p_act[cur ^ 1].as<float>();
      const int Ho = H / bw.stride;
      const int M_in = B * H * H, M_out = B * Ho * Ho;
      gemm_f32<ACT_RELU, RES_NONE>(xin, M_in, bw.c1, p_t1.as<float>(), nullptr, st);
      gemm_f32<ACT_RELU, RES_NONE>(conv3x3_f32_cols(p_t1.as<float>(), B, H, bw.width, bw.stride, st), M_out, bw.c2, p_t2.as<float>(), nullptr, st);
      const float* identity = xin;
      if (bw.has_ds) {
        const float* ds_in = xin;
        if (bw.stride == 2) {
          const size_t total = static_cast<size_t>(M_out) * (bw.cin / 8);
          det::subsample2_kernel<float><<<static_cast<int>(std::min<size_t>(ceil_div64(total, 256), 148 * 16)), 256, 0, st>>>(
              xin, p_sub.as<float>(), B, H, H, bw.cin);
          KERNEL_CHECK();
          ++launches;
          ds_in = p_sub.as<float>();
        }
        gemm_f32<ACT_NONE, RES_NONE>(ds_in, M_out, bw.ds, p_idb.as<float>(), nullptr, st);
        identity = p_idb.as<float>();
      }
      gemm_f32<ACT_RELU, RES_F32>(p_t2.as<float>(), M_out, bw.c3, yout, identity, st);
      cur ^= 1;
      H = Ho;
    }
    gemm_f32<ACT_RELU, RES_NONE>(conv3x3_f32_cols(p_feats.as<float>(), B, f, 2048, 1, st), B * f * f, rpn_conv, p_rpn_t.as<float>(), nullptr, st);
    gemm_f32<ACT_NONE, RES_NONE>(p_rpn_t.as<float>(), B * f * f, rpn_heads, rpn_out.as<float>(), nullptr, st);
  }

  // full detector + selection; leaves lm_in [R,1024] bf16 on the device; returns R
  bool want_abnormal = false;  // set by rgrg_detect when the caller asks for the abnormal-region predictions
  int run_detect(const float* img_dev, int B, int S, cudaStream_t st) {
    ensure_detector_ws(B, S);
    const int f = S / 32;
    const bool precise = opt_detector_precise != 0;
    if (precise && !precise_weights) throw std::runtime_error("detector_precise must be set before the weights are finalized");
    if (precise) {
      run_backbone_rpn_precise(img_dev, B, S, st);
    } else {
      run_backbone(img_dev, B, S, feats.as<bf16>(), st);
      // RPN head: 3x3 conv + ReLU, then both 1x1 heads as one N = 160 + 640 GEMM with fp32 (decision-critical) output
      conv3x3("rpn_conv", feats.as<bf16>(), B, f, f, 2048, 1, rpn_conv, epi<true, ACT_RELU, RES_NONE, true>(rpn_t.p, rpn_conv.bias, 2048), st);
      gemm("rpn_heads", rpn_t.as<bf16>(), B * f * f, rpn_heads, epi<false, ACT_NONE, RES_NONE, true>(rpn_out.p, rpn_heads.bias, 800), st, false);
    }
    det::RpnIn in{};
    in.obj = rpn_out.as<float>();
    in.deltas = rpn_out.as<float>() + 160;
    in.obj_bs = in.del_bs = static_cast<long long>(f) * f * 800;
    in.obj_ps = in.del_ps = 800;
    det::RpnOut out{};
    out.boxes = prop_boxes.as<float>();
    out.scores = prop_scores.as<float>();
    out.count = prop_count.as<int>();
    run_rpn_filter(in, out, B, f, S, st);
    det::roi_offsets_kernel<<<1, 32, 0, st>>>(prop_count.as<int>(), roi_off.as<int>(), B);
    KERNEL_CHECK();
    ++launches;
    int P_total = 0;
    CUDA_CHECK(cudaMemcpyAsync(&P_total, roi_off.as<int>() + B, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));  // the RoI GEMMs are sized by the number of surviving proposals
    last_P = P_total;
    const size_t prow = static_cast<size_t>(std::max(P_total, 1));
    const float scale = exp2f(roundf(log2f(static_cast<float>(f) / static_cast<float>(S))));  // poolers.py _infer_scale
    if (precise) {
      p_pooled.ensure(prow * 131072 * 4);
      p_f6.ensure(prow * 1024 * 4);
      p_f7.ensure(prow * 1024 * 4);
      pred_out.ensure(prow * 150 * 4);
      if (P_total > 0) {
        det::roi_align_kernel<float><<<dim3(TOPK, B), 256, 0, st>>>(p_feats.as<float>(), prop_boxes.as<float>(), prop_count.as<int>(),
                                                                    roi_off.as<int>(), p_pooled.as<float>(), f, 2048, scale);
        KERNEL_CHECK();
        ++launches;
        gemm_f32<ACT_RELU, RES_NONE>(p_pooled.as<float>(), P_total, fc6, p_f6.as<float>(), nullptr, st);
        gemm_f32<ACT_RELU, RES_NONE>(p_f6.as<float>(), P_total, fc7, p_f7.as<float>(), nullptr, st);
        gemm_f32<ACT_NONE, RES_NONE>(p_f7.as<float>(), P_total, pred, pred_out.as<float>(), nullptr, st);
      }
    } else {
      ensure_roi_ws(P_total);
      if (P_total > 0) {
        {
          ProfScope ps(this, "roi_align", st);
          if (opt_roi_align_sep && f <= 32)
            det::roi_align_sep_kernel<<<dim3(TOPK, B), 256, 0, st>>>(feats.as<bf16>(), prop_boxes.as<float>(), prop_count.as<int>(),
                                                                     roi_off.as<int>(), pooled.as<bf16>(), f, 2048, scale);
          else
            det::roi_align_kernel<bf16><<<dim3(TOPK, B), 256, 0, st>>>(feats.as<bf16>(), prop_boxes.as<float>(), prop_count.as<int>(),
                                                                       roi_off.as<int>(), pooled.as<bf16>(), f, 2048, scale);
          KERNEL_CHECK();
          ++launches;
        }
        gemm("fc6", pooled.as<bf16>(), P_total, fc6, epi<true, ACT_RELU, RES_NONE, true>(f6.p, fc6.bias, 1024), st, false);
        gemm("fc7", f6.as<bf16>(), P_total, fc7, epi<true, ACT_RELU, RES_NONE, true>(f7.p, fc7.bias, 1024), st, true);
        gemm("box_predictor", f7.as<bf16>(), P_total, pred, epi<false, ACT_NONE, RES_NONE, true>(pred_out.p, pred.bias, 150), st, true);
      }
    }
    ProfScope ps_tail(this, "region_tail", st);
    det::RoiTailOut to{detected.as<uint8_t>(), top_idx.as<int>(), top_scores.as<float>(), top_boxes.as<float>()};
    det::roi_tail_kernel<<<B, 256, 0, st>>>(pred_out.as<float>(), 150, pred_out.as<float>() + 30, 150, prop_boxes.as<float>(),
                                            prop_count.as<int>(), roi_off.as<int>(), to, S);
    KERNEL_CHECK();
    if (precise)
      det::roi_mean_kernel<float><<<dim3(NREG, B), 256, 0, st>>>(p_feats.as<float>(), prop_boxes.as<float>(), prop_count.as<int>(),
                                                                 top_idx.as<int>(), mean2048.as<float>(), f, 2048, scale);
    else
      det::roi_mean_kernel<bf16><<<dim3(NREG, B), 256, 0, st>>>(feats.as<bf16>(), prop_boxes.as<float>(), prop_count.as<int>(),
                                                                top_idx.as<int>(), mean2048.as<float>(), f, 2048, scale);
    KERNEL_CHECK();
    launches += 2;
    const int rows = B * NREG;
    // dim_reduction + selection MLP (+ abnormal MLP) in fp32 on CUDA cores (decision-critical, 0.01 % of the FLOPs)
    simt_f32(mean2048.as<float>(), dimred.w, rows, 1024, 2048,
             epi<false, ACT_NONE, RES_NONE, true>(trf.p, dimred.bias, 1024), st);
    simt_f32(trf.as<float>(), sel0.w, rows, 512, 1024, epi<false, ACT_RELU, RES_NONE, true>(s0.p, sel0.bias, 512), st);
    simt_f32(s0.as<float>(), sel2.w, rows, 128, 512, epi<false, ACT_RELU, RES_NONE, true>(s1.p, sel2.bias, 128), st);
    det::selection_tail_kernel<<<1, 1024, 0, st>>>(s1.as<float>(), sel4.w, sel4.bias, detected.as<uint8_t>(), sel_logits.as<float>(),
                                                   selected.as<uint8_t>(), sel_rows.as<int>(), num_sel.as<int>(), rows);
    KERNEL_CHECK();
    det::gather_rows_bf16_kernel<<<rows + 31, 256, 0, st>>>(trf.as<float>(), sel_rows.as<int>(), num_sel.as<int>(), lm_in.as<bf16>(), 1024);
    KERNEL_CHECK();
    launches += 5;
    if (want_abnormal && has_abnormal) {
      // a9': abnormal classifier (report_generation_model.py:103-106; not evaluated by generate()): same MLP shape
      simt_f32(trf.as<float>(), abn0.w, rows, 512, 1024, epi<false, ACT_RELU, RES_NONE, true>(s0.p, abn0.bias, 512), st);
      simt_f32(s0.as<float>(), abn2.w, rows, 128, 512, epi<false, ACT_RELU, RES_NONE, true>(s1.p, abn2.bias, 128), st);
      det::selection_tail_kernel<<<1, 1024, 0, st>>>(s1.as<float>(), abn4.w, abn4.bias, static_cast<const uint8_t*>(nullptr), abn_logits.as<float>(),
                                                     abnormal.as<uint8_t>(), static_cast<int*>(nullptr), static_cast<int*>(nullptr), rows);
      KERNEL_CHECK();
      launches += 3;
    }
    int R = 0;
    CUDA_CHECK(cudaMemcpyAsync(&R, num_sel.as<int>(), 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));  // report_generation_model.py:260: R == 0 -> return -1
    last_B = B;
    last_S = S;
    return R;
  }

  // ================================================================================================================
  // decoder
  // ================================================================================================================
  void drop_step_graphs() {
    for (auto& g : step_graphs) cudaGraphExecDestroy(g.second);
    step_graphs.clear();
    step_graph_nodes.clear();
  }
  // Growth is transactional: capacities are published only after every allocation succeeded; on failure the
  // whole decoder workspace is released and the capacities reset, so the caller's next (smaller) batch starts clean —
  // the reference's callers catch "out of memory" and carry on (evaluate_language_model.py:1207-1222).
  void ensure_decoder_ws(int rows, int max_length) {
    const int slots = max_length + 1;
    if (rows <= ws_rows && slots <= ws_slots) return;
    const int r = std::max(rows, ws_rows), sl = std::max(slots, ws_slots);
    drop_step_graphs();  // buffers move: captured pointers would be stale
    try {
      kv_cache.ensure(static_cast<size_t>(NLAYER) * 2 * r * 16 * sl * 64 * 2);
      const size_t rr = static_cast<size_t>(r);
      h.ensure(rr * DM * 4);
      x.ensure(rr * DM * 2);
      q.ensure(rr * DM * 2);
      attn_o.ensure(rr * DM * 2);
      mlp_mid.ensure(rr * 4 * DM * 2);
      splitk_parts.ensure(rr * DM * 4 * 4);
      ln_counters.ensure(1024 * 4);
      a1.ensure(rr * DM * 2);
      img.ensure(rr * DM * 2);
      part_val.ensure(rr * 2048 * 4);
      part_idx.ensure(rr * 2048 * 4);
      ids.ensure(rr * (sl + 1) * 4);
      unfinished.ensure(rr * 4);
      unf_count.ensure(static_cast<size_t>(sl + 1) * 4);
      step.ensure(16);
    } catch (...) {
      DevBuf* all[] = {&kv_cache, &h, &x, &q, &attn_o, &mlp_mid, &splitk_parts, &ln_counters, &a1, &img, &part_val, &part_idx, &ids,
                       &unfinished, &unf_count, &step, &logits_tmp};
      for (DevBuf* b : all) b->release();
      ws_rows = ws_slots = 0;
      throw;
    }
    ws_rows = r;
    ws_slots = sl;
  }
  KvGeom kv_geom() const { return KvGeom{kv_cache.as<bf16>(), ws_rows, ws_slots}; }

  // language_model.py:284 (once instead of every step) + :140-147 for all 24 layers in one GEMM
  void lm_prologue(const bf16* feats_bf16, int R, int beams, cudaStream_t st) {
    gemm("lm_prologue", feats_bf16, R, fst0, epi<true, ACT_RELU, RES_NONE, true>(a1.p, fst0.bias, DM), st, true);
    gemm("lm_prologue", a1.as<bf16>(), R, fst2, epi<true, ACT_NONE, RES_NONE, true>(img.p, fst2.bias, DM), st, true);
    EpiImageKv e{ukv.bias, kv_geom(), beams};
    gemm("lm_image_kv", img.as<bf16>(), R, ukv, e, st, true);
  }

  // is the 16-CTA cluster shape of the LayerNorm-head kernels launchable on this device?  (queried once)
  int cluster16_ok = -1;
  bool ln_head_available() {
    if (cluster16_ok < 0) {
      auto kern = fa::attn_fused_kernel<8, 4, true, 1>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::Smem<8, 4>::TOTAL);
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16);
      cfg.blockDim = dim3(tc::NUM_THREADS);
      cfg.dynamicSmemBytes = fa::Smem<8, 4>::TOTAL;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 16;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      cudaGetLastError();
      cluster16_ok = (e == cudaSuccess && n >= 1) ? n : 0;
    }
    return cluster16_ok > 0;
  }

  // (attention warps, ring slots per warp): 48 KB of q/k/v tiles + AW * NSLOT * 4 KB of K/V staging must fit in 227 KB
  template <bool LN_HEAD>
  void launch_attn_fused(const CUtensorMap& tmA, const CUtensorMap& tmA64, const CUtensorMap& tmW, const fa::Params& fp, cudaStream_t st) {
    switch (opt_attn_warps * 10 + opt_attn_slots) {
      case 84: fa::launch<8, 4, LN_HEAD, 1>(tmA, tmA64, tmW, fp, st, pdl_now); break;
      case 162: fa::launch<16, 2, LN_HEAD, 1>(tmA, tmA64, tmW, fp, st, pdl_now); break;
      case 241: fa::launch<24, 1, LN_HEAD, 1>(tmA, tmA64, tmW, fp, st, pdl_now); break;
      default: throw std::runtime_error("unsupported (attn_warps, attn_slots)");
    }
  }

  // ---- transformer body of one decode step (embedding .. final LayerNorm) for a contiguous row range ("view").  Every
  // kernel reads the step index from device memory, so a captured CUDA graph of the step is replayed for every t.
  //
  // Greedy path, per layer: [c_attn + KV append + attention] (attn_fused.cuh) -> attn c_proj (split-K) -> LayerNorm ->
  // c_fc + gelu_new -> mlp c_proj (split-K) -> LayerNorm; the LayerNorm kernels also fold the preceding projection's
  // partial sums + bias into the residual stream.  Optional "ln_head": the LayerNorm runs as the head of its consumer
  // kernel instead (measured slower, profiles/r02_decode.md).  Beam search (cache rows are gathered through the ancestry
  // table) and the cross-check GEMM use c_attn (KV-append epilogue) + the stand-alone attention kernel.
  struct DecView {
    int row0, rows;
    float* h;
    bf16 *x, *q, *attn_o, *mid;
    float* parts;
    size_t pstride;
    KvGeom kv;
    const int* ids;  // first row of the view
    int ids_ld;
    unsigned* counters;  // LayerNorm-head counters of this view
    CUtensorMap tm_x, tm_x64;  // operand A of the GEMMs: 128-row boxes; 64-row boxes for the multicast halves of attn_fused
    const float* pending_bias;  // bias of the split-K projection whose partial sums wait in `parts`
  };
  DecView dec_view(int row0, int rows, const int* ids_all, int ids_ld) {
    DecView v{};
    v.row0 = row0;
    v.rows = rows;
    v.h = h.as<float>() + static_cast<size_t>(row0) * DM;
    v.x = x.as<bf16>() + static_cast<size_t>(row0) * DM;
    v.q = q.as<bf16>() + static_cast<size_t>(row0) * DM;
    v.attn_o = attn_o.as<bf16>() + static_cast<size_t>(row0) * DM;
    v.mid = mlp_mid.as<bf16>() + static_cast<size_t>(row0) * 4 * DM;
    v.parts = splitk_parts.as<float>() + static_cast<size_t>(row0) * DM * 4;  // each view keeps its 4 slices contiguous
    v.pstride = static_cast<size_t>(rows) * DM;
    v.kv = kv_geom();
    v.kv.cache += v.kv.offset(0, 0, row0, 0, 0);  // KvGeom::offset is linear in the row index
    v.ids = ids_all + static_cast<size_t>(row0) * ids_ld;
    v.ids_ld = ids_ld;
    v.counters = ln_counters.as<unsigned>() + (row0 ? 256 : 0);
    v.tm_x = tc::make_tmap_2d(v.x, rows, DM, 128);
    v.tm_x64 = tc::make_tmap_2d(v.x, rows, DM, 64);
    v.pending_bias = nullptr;
    return v;
  }
  bool decode_tensor_path() const { return opt_gemm_impl != 2; }
  bool decode_use_fused() const { return opt_fused_attn && !beam_anc && decode_tensor_path(); }
  bool decode_use_head(const DecView& v) const { return opt_ln_head && decode_tensor_path() && ceil_div(v.rows, tc::BM) <= 128; }

  void view_embed(DecView& v, cudaStream_t st) {
    ProfScope ps(this, "embed", st);
    launch_kernel(dec::embed_kernel, dim3(ceil_div(v.rows, 8)), dim3(256), 0, st, false, wte_f32, v.ids, v.ids_ld, step.as<int>(), v.h, v.rows);
    ++launches;
  }
  // LayerNorm fused with the residual update of the preceding split-K projection (pending_bias != null)
  void view_ln(DecView& v, const float* g, const float* b, cudaStream_t st) {
    ProfScope ps(this, "layernorm", st);
    if (!(opt_ablate & 2)) {
      if (v.pending_bias)
        launch_kernel(dec::layernorm_kernel<4>, dim3(v.rows), dim3(128), 0, st, pdl_now, v.h, g, b, v.x, v.rows, v.parts, v.pstride, v.pending_bias, trace_ptr());
      else
        launch_kernel(dec::layernorm_kernel<0>, dim3(v.rows), dim3(128), 0, st, pdl_now, v.h, g, b, v.x, v.rows, v.parts, v.pstride, v.pending_bias, trace_ptr());
      ++launches;
    }
    v.pending_bias = nullptr;
  }
  // LN1 + c_attn + KV append + attention of layer l; `before_attn` runs right before the attention kernel is launched
  // (the two-halves schedule hangs its cross-stream dependency there)
  template <class F>
  void view_attn(DecView& v, int l, cudaStream_t st, F&& before_attn) {
    const LayerW& L = layers[l];
    const int* sp = step.as<int>();
    if (decode_use_fused()) {
      const bool head = decode_use_head(v);
      fa::Params fp{};
      fp.bias = L.attn.bias;
      fp.kv = v.kv;
      fp.layer = l;
      fp.step_ptr = sp;
      fp.attn_o = v.attn_o;
      fp.M = v.rows;
      // spread the rows evenly over as many M tiles as fit the SMs (16 CTAs per tile): 928 rows -> 9 tiles of 104 rows on 144
      // SMs instead of 7 full tiles + one quarter tile on 128; the LayerNorm head needs the 128-row tiling
      {
        const int min_tiles = ceil_div(v.rows, tc::BM);
        const int sm_tiles = std::max(1, tc::num_sms() / 16);
        const int tiles = (opt_attn_balance && !head) ? std::max(min_tiles, std::min(sm_tiles, ceil_div(v.rows, 32))) : min_tiles;
        fp.rows_per_tile = (opt_attn_balance && !head) ? ceil_div(v.rows, tiles) : tc::BM;
      }
      fp.l2_ahead = opt_l2_ahead;
      fp.trace = trace_ptr();
      fp.early_kv = opt_attn_early;
      fp.mc = opt_attn_mc;
      if (head) {
        fp.h = v.h;
        fp.x = v.x;
        fp.gamma = L.ln1_g;
        fp.beta = L.ln1_b;
        fp.parts = v.pending_bias ? v.parts : nullptr;
        fp.part_stride = v.pstride;
        fp.res_bias = v.pending_bias;
        fp.counters = v.counters;  // first half of the view's counters: the attention kernels
        fp.launch_idx = l;
        fp.launches_per_step = NLAYER;
        v.pending_bias = nullptr;
      } else {
        view_ln(v, L.ln1_g, L.ln1_b, st);
      }
      before_attn();
      if (!(opt_ablate & 64)) {
        ProfScope ps(this, "attn_fused", st);
        if (head) launch_attn_fused<true>(v.tm_x, v.tm_x64, L.tm_qkv, fp, st);
        else launch_attn_fused<false>(v.tm_x, v.tm_x64, L.tm_qkv, fp, st);
        ++launches;
      }
    } else {
      view_ln(v, L.ln1_g, L.ln1_b, st);
      EpiQkvAppend eq{v.q, L.attn.bias, v.kv, l, sp};
      if (!(opt_ablate & 4)) gemm("c_attn", v.x, v.rows, L.attn, eq, st, true, opt_cattn_bn);
      before_attn();
      if (!(opt_ablate & 1)) {
        ProfScope ps(this, "attention", st);
        const dim3 grid(ceil_div(v.rows * 16, 4));
        const unsigned char* anc = beam_anc ? beam_anc + static_cast<size_t>(v.row0) * beam_slots : nullptr;
        auto go = [&](auto kern) {
          launch_kernel(kern, grid, dim3(128), 0, st, pdl_now, v.q, v.kv, l, sp, v.attn_o, v.rows, anc, beam_slots, beam_nb);
        };
        if (opt_attn_occ == 6) go(dec::attention_kernel<6>);
        else if (opt_attn_occ == 5) go(dec::attention_kernel<5>);
        else if (opt_attn_occ == 8) go(dec::attention_kernel<8>);
        else go(dec::attention_kernel<7>);
        ++launches;
      }
    }
  }
  // attn c_proj (split-K) -> LN2 -> c_fc + gelu_new -> mlp c_proj (split-K) of layer l
  void view_mlp(DecView& v, int l, cudaStream_t st) {
    const LayerW& L = layers[l];
    constexpr int SPLITS = 4;
    use_2cta_now = opt_gemm_2cta && decode_tensor_path();
    struct Reset {
      bool& f;
      ~Reset() { f = false; }
    } reset{use_2cta_now};
    if (!(opt_ablate & 8)) gemm_splitk("attn_c_proj", v.attn_o, v.rows, L.proj, v.parts, SPLITS, st);
    v.pending_bias = L.proj.bias;
    auto ep_fc = epi<true, ACT_GELU_NEW, RES_NONE, true>(v.mid, L.fc.bias, 4 * DM);
    if (decode_use_head(v)) {
      tc::GemmShape::LnHead lh{v.h, v.x, L.ln2_g, L.ln2_b, v.parts, v.pstride, v.pending_bias, v.counters + 128, step.as<int>(),
                               l, NLAYER};  // second half of the view's counters: the c_fc kernels
      v.pending_bias = nullptr;
      if (!(opt_ablate & 16)) gemm_ln_head("mlp_c_fc", v.x, v.rows, L.fc, ep_fc, lh, st);
    } else {
      view_ln(v, L.ln2_g, L.ln2_b, st);
      if (!(opt_ablate & 16)) gemm("mlp_c_fc", v.x, v.rows, L.fc, ep_fc, st, true);
    }
    if (!(opt_ablate & 32)) gemm_splitk("mlp_c_proj", v.mid, v.rows, L.mproj, v.parts, SPLITS, st);
    v.pending_bias = L.mproj.bias;
  }

  // one chain over all rows; leaves ln_f(h) as bf16 in x
  void decode_forward(int rows, const int* ids_ptr, int ids_ld, cudaStream_t st) {
    DecView v = dec_view(0, rows, ids_ptr, ids_ld);
    view_embed(v, st);
    pdl_now = opt_pdl != 0;
    for (int l = 0; l < NLAYER; ++l) {
      view_attn(v, l, st, [] {});
      view_mlp(v, l, st);
    }
    view_ln(v, lnf_g, lnf_b, st);
  }

  // The same as TWO row halves on two streams (inside the step graph: two branches).  Rows never interact.  The halves are
  // kept half a layer out of phase by cross-stream dependencies — half B's attention kernel of layer l starts when half
  // A's has finished, half A's of layer l+1 when half B's of layer l has — so that while one half streams its KV cache
  // (HBM-bound, 64 CTAs) the other half runs its GEMM / LayerNorm kernels (L2- and tensor-bound, 64 CTAs) on the other SMs:
  // the fused attention kernel and the decode GEMMs each take one SM per CTA, so half-size launches co-reside.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_a, ev_b;
  void decode_forward_dual(int rows, const int* ids_ptr, int ids_ld, cudaStream_t st) {
    if (!side_stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      ev_a.resize(NLAYER);
      ev_b.resize(NLAYER);
      for (int l = 0; l < NLAYER; ++l) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_a[l], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_b[l], cudaEventDisableTiming));
      }
    }
    int rows_a = ((rows / 2 + tc::BM - 1) / tc::BM) * tc::BM;  // split on an M-tile boundary: no extra tile padding
    if (rows_a >= rows) rows_a = rows / 2;
    DecView va = dec_view(0, rows_a, ids_ptr, ids_ld), vb = dec_view(rows_a, rows - rows_a, ids_ptr, ids_ld);
    cudaStream_t sa = st, sb = side_stream;
    CUDA_CHECK(cudaEventRecord(ev_fork, sa));
    CUDA_CHECK(cudaStreamWaitEvent(sb, ev_fork, 0));
    view_embed(va, sa);
    view_embed(vb, sb);
    pdl_now = opt_pdl != 0;
    for (int l = 0; l < NLAYER; ++l) {
      view_attn(va, l, sa, [&] {
        if (l > 0) CUDA_CHECK(cudaStreamWaitEvent(sa, ev_b[l - 1], 0));
      });
      CUDA_CHECK(cudaEventRecord(ev_a[l], sa));
      view_attn(vb, l, sb, [&] { CUDA_CHECK(cudaStreamWaitEvent(sb, ev_a[l], 0)); });
      CUDA_CHECK(cudaEventRecord(ev_b[l], sb));
      view_mlp(va, l, sa);
      view_mlp(vb, l, sb);
    }
    view_ln(va, lnf_g, lnf_b, sa);
    view_ln(vb, lnf_g, lnf_b, sb);
    CUDA_CHECK(cudaEventRecord(ev_join, sb));
    CUDA_CHECK(cudaStreamWaitEvent(sa, ev_join, 0));
  }
  void end_pdl() { pdl_now = false; }

  // lm_head + greedy bookkeeping.  logits_out != null: lm_head stores fp32 logits there instead of the fused arg-max.
  void decode_head(int rows, const dec::GreedyState& g, float* logits_out, cudaStream_t st) {
    if (logits_out || opt_gemm_impl == 2) {
      float* dst = logits_out ? logits_out : logits_tmp.as<float>();
      gemm("lm_head", x.as<bf16>(), rows, lm_head, epi<false, ACT_NONE, RES_NONE, false>(dst, nullptr, VOCAB), st, true);
      ProfScope ps(this, "greedy_update", st);
      launch_kernel(dec::greedy_update_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, st, pdl_now, static_cast<const float*>(nullptr),
                    static_cast<const int*>(nullptr), 0, static_cast<const float*>(dst), g, rows);
    } else {
      const int bn = pick_bn(ceil_div(rows, tc::BM), VOCAB);
      const int n_tiles = 2 * ceil_div(VOCAB, bn);  // two partials per tile: one per epilogue warp of a lane quarter
      EpiArgmaxPartial ea{part_val.as<float>(), part_idx.as<int>(), n_tiles};
      gemm("lm_head", x.as<bf16>(), rows, lm_head, ea, st, true, bn);
      ProfScope ps(this, "greedy_update", st);
      launch_kernel(dec::greedy_update_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, st, pdl_now,
                    static_cast<const float*>(part_val.as<float>()), static_cast<const int*>(part_idx.as<int>()), n_tiles,
                    static_cast<const float*>(nullptr), g, rows);
    }
    end_pdl();
    ++launches;
  }

  // one greedy decode step over all rows
  int decode_step(int rows, const dec::GreedyState& g, float* logits_out, cudaStream_t st) {
    const int before = static_cast<int>(launches);
    trace_slot = 0;
    if (opt_dual && rows >= 2 * tc::BM && decode_use_fused() && !prof_on) decode_forward_dual(rows, g.ids, g.ids_ld, st);
    else decode_forward(rows, g.ids, g.ids_ld, st);
    decode_head(rows, g, logits_out, st);
    return static_cast<int>(launches) - before;
  }

  // ================================================================================================================
  // beam search (language_model.py:529-607; BeamSearchScorer restated from transformers 4.19.2)
  // ================================================================================================================
  DevBuf b_ids2, b_anc[2], b_scores, b_cand_score, b_cand_token, b_cand_beam, b_hyp_score, b_hyp_len, b_hyp_tok, b_hyp_count,
      b_worst, b_done, b_not_done;

  dec::BeamState beam_state(int sentences, int nb, int max_length, bool early) {
    const size_t rows = static_cast<size_t>(sentences) * nb;
    const int slots = max_length + 1;
    b_ids2.ensure(rows * max_length * 4);
    b_anc[0].ensure(rows * slots);
    b_anc[1].ensure(rows * slots);
    b_scores.ensure(rows * 4);
    b_cand_score.ensure(rows * 2 * 4);
    b_cand_token.ensure(rows * 2 * 4);
    b_cand_beam.ensure(rows * 2 * 4);
    b_hyp_score.ensure(rows * 4);
    b_hyp_len.ensure(rows * 4);
    b_hyp_tok.ensure(rows * max_length * 4);
    b_hyp_count.ensure(static_cast<size_t>(sentences) * 4);
    b_worst.ensure(static_cast<size_t>(sentences) * 4);
    b_done.ensure(static_cast<size_t>(sentences) * 4);
    b_not_done.ensure(static_cast<size_t>(max_length + 1) * 4);
    dec::BeamState s{};
    s.nb = nb;
    s.ids_ld = max_length;
    s.ids[0] = ids.as<int>();
    s.ids[1] = b_ids2.as<int>();
    s.anc[0] = b_anc[0].as<unsigned char>();
    s.anc[1] = b_anc[1].as<unsigned char>();
    s.slots = slots;
    s.beam_scores = b_scores.as<float>();
    s.cand_score = b_cand_score.as<float>();
    s.cand_token = b_cand_token.as<int>();
    s.cand_beam = b_cand_beam.as<int>();
    s.hyp_score = b_hyp_score.as<float>();
    s.hyp_len = b_hyp_len.as<int>();
    s.hyp_tok = b_hyp_tok.as<int>();
    s.hyp_count = b_hyp_count.as<int>();
    s.worst = b_worst.as<float>();
    s.done = b_done.as<int>();
    s.not_done_count = b_not_done.as<int>();
    s.step_ptr = step.as<int>();
    s.early_stopping = early ? 1 : 0;
    return s;
  }

  void beam_bookkeeping(const dec::BeamState& s, const float* logits, int sentences, int src, cudaStream_t st) {
    ProfScope ps(this, "beam_bookkeeping", st);
    dec::beam_topk_kernel<<<sentences, 1024, 0, st>>>(logits, s);
    KERNEL_CHECK();
    dec::beam_process_kernel<<<ceil_div(sentences, 64), 64, 0, st>>>(s, sentences, src, src ^ 1);
    KERNEL_CHECK();
    dec::beam_step_end_kernel<<<1, 256, 0, st>>>(s, sentences);
    KERNEL_CHECK();
    launches += 3;
  }

  // the same from the fused lm_head epilogue's per-part summaries (generate() path: no [rows, V] logits in HBM)
  DevBuf b_part_m, b_part_l, b_part_val, b_part_idx;
  void beam_parts_ensure(int rows, int nb) {
    const int n_parts = 2 * ceil_div(VOCAB, pick_bn(ceil_div(rows, tc::BM), VOCAB));
    const int K = nb <= 4 ? 8 : 16;
    const size_t rp = static_cast<size_t>(rows) * n_parts;
    b_part_m.ensure(rp * 4);
    b_part_l.ensure(rp * 4);
    b_part_val.ensure(rp * K * 4);
    b_part_idx.ensure(rp * K * 4);
  }
  void beam_head_fused(const dec::BeamState& s, int rows, int sentences, int src, cudaStream_t st) {
    const int bn = pick_bn(ceil_div(rows, tc::BM), VOCAB);
    const int n_parts = 2 * ceil_div(VOCAB, bn);
    const int K = s.nb <= 4 ? 8 : 16;
    if (K == 8) {
      EpiBeamPartial<8> ep{b_part_m.as<float>(), b_part_l.as<float>(), b_part_val.as<float>(), b_part_idx.as<int>(), n_parts};
      gemm("lm_head", x.as<bf16>(), rows, lm_head, ep, st, true, bn);
    } else {
      EpiBeamPartial<16> ep{b_part_m.as<float>(), b_part_l.as<float>(), b_part_val.as<float>(), b_part_idx.as<int>(), n_parts};
      gemm("lm_head", x.as<bf16>(), rows, lm_head, ep, st, true, bn);
    }
    end_pdl();
    ProfScope ps(this, "beam_bookkeeping", st);
    if (K == 8)
      dec::beam_merge_kernel<8><<<sentences, 1024, 0, st>>>(b_part_m.as<float>(), b_part_l.as<float>(), b_part_val.as<float>(),
                                                            b_part_idx.as<int>(), n_parts, s);
    else
      dec::beam_merge_kernel<16><<<sentences, 1024, 0, st>>>(b_part_m.as<float>(), b_part_l.as<float>(), b_part_val.as<float>(),
                                                             b_part_idx.as<int>(), n_parts, s);
    KERNEL_CHECK();
    dec::beam_process_kernel<<<ceil_div(sentences, 64), 64, 0, st>>>(s, sentences, src, src ^ 1);
    KERNEL_CHECK();
    dec::beam_step_end_kernel<<<1, 256, 0, st>>>(s, sentences);
    KERNEL_CHECK();
    launches += 3;
  }

  // BeamSearchScorer.finalize on the host (once per generate); returns the reference width
  int beam_finalize(const dec::BeamState& s, int sentences, int cur_len, int final_buf, int max_length, int32_t* out_ids,
                    cudaStream_t st) {
    const int nb = s.nb;
    const size_t rows = static_cast<size_t>(sentences) * nb;
    std::vector<int> h_ids(rows * max_length), h_hyp_tok(rows * max_length), h_hyp_len(rows), h_count(sentences), h_done(sentences);
    std::vector<float> h_scores(rows), h_hyp_score(rows), h_worst(sentences);
    CUDA_CHECK(cudaMemcpyAsync(h_ids.data(), s.ids[final_buf], h_ids.size() * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_hyp_tok.data(), s.hyp_tok, h_hyp_tok.size() * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_hyp_len.data(), s.hyp_len, rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_hyp_score.data(), s.hyp_score, rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_scores.data(), s.beam_scores, rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_count.data(), s.hyp_count, sentences * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_done.data(), s.done, sentences * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_worst.data(), s.worst, sentences * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    struct Hyp {
      float score;
      std::vector<int> tok;
    };
    std::vector<std::vector<int>> best(sentences);
    int max_len = 0, min_len = 1 << 30;
    for (int i = 0; i < sentences; ++i) {
      std::vector<Hyp> beams;
      for (int k = 0; k < h_count[i]; ++k) {
        const int* t = h_hyp_tok.data() + (static_cast<size_t>(i) * nb + k) * max_length;
        beams.push_back(Hyp{h_hyp_score[i * nb + k], std::vector<int>(t, t + h_hyp_len[i * nb + k])});
      }
      float worst = h_worst[i];
      if (!h_done[i]) {  // all open beams become hypotheses (BeamHypotheses.add)
        for (int b = 0; b < nb; ++b) {
          const int* t = h_ids.data() + (static_cast<size_t>(i) * nb + b) * max_length;
          const float score = h_scores[i * nb + b] / static_cast<float>(cur_len);
          if (static_cast<int>(beams.size()) < nb || score > worst) {
            beams.push_back(Hyp{score, std::vector<int>(t, t + cur_len)});
            if (static_cast<int>(beams.size()) > nb) {
              size_t lo = 0;
              for (size_t k = 1; k < beams.size(); ++k)
                if (beams[k].score < beams[lo].score) lo = k;
              beams.erase(beams.begin() + lo);
              worst = beams[0].score;
              for (const Hyp& hh : beams) worst = std::min(worst, hh.score);
            } else {
              worst = std::min(score, worst);
            }
          }
        }
      }
      // sorted(beams, key=score) is stable; pop() takes the last of the highest score
      size_t pick = 0;
      for (size_t k = 1; k < beams.size(); ++k)
        if (beams[k].score >= beams[pick].score) pick = k;
      best[i] = beams.empty() ? std::vector<int>() : beams[pick].tok;
      max_len = std::max<int>(max_len, static_cast<int>(best[i].size()));
      min_len = std::min<int>(min_len, static_cast<int>(best[i].size()));
    }
    const int width = std::min(max_len + 1, max_length);
    for (int i = 0; i < sentences; ++i) {
      int32_t* row = out_ids + static_cast<size_t>(i) * max_length;
      for (int c = 0; c < max_length; ++c) row[c] = RGRG_EOS;
      const int len = static_cast<int>(best[i].size());
      for (int c = 0; c < len && c < max_length; ++c) row[c] = best[i][c];
      if (len < width) row[len] = RGRG_EOS;
    }
    return width;
  }

  int run_beam(const bf16* feats_bf16, int R, int nb, int max_length, bool early, int32_t* out_ids, cudaStream_t st) {
    if (nb > dec::MAX_BEAMS) throw std::runtime_error("num_beams > 8 is not supported");
    const int rows = R * nb;
    ensure_decoder_ws(rows, max_length);
    const bool fused_head = opt_beam_fused_head && opt_gemm_impl != 2;
    if (fused_head) beam_parts_ensure(rows, nb);  // before the graph signature below reads the addresses
    else logits_tmp.ensure(static_cast<size_t>(rows) * VOCAB * 4);
    dec::BeamState s = beam_state(R, nb, max_length, early);
    lm_prologue(feats_bf16, R, nb, st);
    CUDA_CHECK(cudaMemsetAsync(ln_counters.p, 0, 1024 * 4, st));
    dec::beam_init_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(s, R);
    KERNEL_CHECK();
    ++launches;
    const int steps = max_length - 1;
    int cur = 0, done_steps = 0;
    beam_slots = s.slots;
    beam_nb = nb;
    // one beam step on the buffers of parity `par` (token matrix / ancestry table are double-buffered)
    auto beam_step = [&](int par) {
      beam_anc = s.anc[par];
      decode_forward(rows, s.ids[par], max_length, st);
      if (fused_head) {
        beam_head_fused(s, rows, R, par, st);
      } else {
        gemm("lm_head", x.as<bf16>(), rows, lm_head, epi<false, ACT_NONE, RES_NONE, false>(logits_tmp.p, nullptr, VOCAB), st, true);
        end_pdl();
        beam_bookkeeping(s, logits_tmp.as<float>(), R, par, st);
      }
    };
    // CUDA graph of TWO consecutive steps (parity 0 then 1), replayed from step 2 on; keyed by the geometry and by the
    // buffer addresses it captured
    cudaGraphExec_t exec = nullptr;
    int nodes = 0;
    const long long sig = reinterpret_cast<long long>(s.ids[1]) ^ (reinterpret_cast<long long>(s.anc[0]) << 1) ^
                          (reinterpret_cast<long long>(fused_head ? b_part_val.p : logits_tmp.p) << 2) ^ (reinterpret_cast<long long>(s.hyp_tok) << 3);
    const int graph_key = -(((rows * 4096 + max_length) * 16 + nb) * 2 + (early ? 1 : 0));  // negative: beam graphs
    const bool want_graph = opt_cuda_graph && !prof_on && steps >= 6;
    int t = 0;
    while (t < steps) {
      if (want_graph && !exec && t == 2) {
        auto it = step_graphs.find(graph_key);
        if (it != step_graphs.end() && beam_graph_sig[graph_key] == sig) {
          exec = it->second;
          nodes = step_graph_nodes[graph_key];
        } else {
          if (it != step_graphs.end()) {
            cudaGraphExecDestroy(it->second);
            step_graphs.erase(it);
          }
          cudaGraph_t graph;
          const int64_t saved = launches;
          CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
          try {
            beam_step(0);
            beam_step(1);
          } catch (...) {
            cudaGraph_t dead;
            cudaStreamEndCapture(st, &dead);
            throw;
          }
          CUDA_CHECK(cudaStreamEndCapture(st, &graph));
          nodes = static_cast<int>(launches - saved);
          launches = saved;
          CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
          CUDA_CHECK(cudaGraphDestroy(graph));
          step_graphs[graph_key] = exec;
          step_graph_nodes[graph_key] = nodes;
          beam_graph_sig[graph_key] = sig;
        }
      }
      if (exec && (t & 1) == 0 && t + 1 < steps) {
        CUDA_CHECK(cudaGraphLaunch(exec, st));
        launches += nodes;
        t += 2;
        done_steps += 2;  // cur is unchanged after an even number of steps
      } else {
        beam_step(cur);
        cur ^= 1;
        ++t;
        ++done_steps;
      }
      if (((t - 1) & 7) == 7 && t < steps) {  // beam_scorer.is_done (language_model.py:594) without a per-step sync
        int c = 1;
        CUDA_CHECK(cudaMemcpyAsync(&c, s.not_done_count + (t - 1), 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (c == 0) break;
      }
    }
    beam_anc = nullptr;
    beam_nb = 1;
    // the reference leaves the loop right after the first step at which every sentence is done; later steps only pad
    std::vector<int> nd(done_steps > 0 ? done_steps : 1);
    CUDA_CHECK(cudaMemcpyAsync(nd.data(), s.not_done_count, static_cast<size_t>(done_steps) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    int cur_len = done_steps + 1;
    for (int t = 0; t < done_steps; ++t)
      if (nd[t] == 0) {
        cur_len = t + 2;
        break;
      }
    return beam_finalize(s, R, cur_len, cur, max_length, out_ids, st);
  }

  // Greedy decode of R rows whose bf16 features sit in feats_bf16 (zero-padded to round_up(R, 32) rows); host ids
  // [R, max_length]; returns the reference width.  The step graph is keyed by the PADDED row count, so batches whose
  // number of selected regions differs by a few rows replay the same graph (rows never interact; padded rows start
  // "finished", emit EOS and are not counted).
  static int padded_rows(int R) { return (R + 31) & ~31; }
  // injected_logits != null (tests): the model forward is replaced by given logits [steps, R, V]; everything else — the
  // device-side greedy bookkeeping, the every-8-steps exit check, the width rule — is the code generate() runs.
  int run_greedy(const bf16* feats_bf16, int R, int max_length, int32_t* out_ids, cudaStream_t st,
                 const float* injected_logits = nullptr, int injected_steps = 0) {
    const int Rp = injected_logits ? R : padded_rows(R);
    ensure_decoder_ws(Rp, max_length);
    if (opt_gemm_impl == 2) logits_tmp.ensure(static_cast<size_t>(Rp) * VOCAB * 4);
    if (!injected_logits) lm_prologue(feats_bf16, Rp, 1, st);
    dec::GreedyState g{};
    g.ids = ids.as<int>();
    g.ids_ld = max_length;
    g.unfinished = unfinished.as<int>();
    g.unfinished_count = unf_count.as<int>();
    g.step_ptr = step.as<int>();
    g.ticket = step.as<int>() + 1;
    g.live_rows = R;
    CUDA_CHECK(cudaMemsetAsync(ln_counters.p, 0, 1024 * 4, st));
    dec::greedy_init_kernel<<<ceil_div(std::max(Rp, max_length), 256), 256, 0, st>>>(g, Rp);
    KERNEL_CHECK();
    ++launches;
    const int steps = max_length - 1;
    cudaGraphExec_t exec = nullptr;
    int nodes = 0;
    const int graph_key = Rp * 4096 + max_length;
    int inj_t = 0;
    auto step_fn = [&]() {
      if (!injected_logits) return decode_step(Rp, g, nullptr, st);
      if (inj_t >= injected_steps) throw std::runtime_error("not enough injected logit steps");
      launch_kernel(dec::greedy_update_kernel, dim3(ceil_div(Rp, 8)), dim3(256), 0, st, false, static_cast<const float*>(nullptr),
                    static_cast<const int*>(nullptr), 0, injected_logits + static_cast<size_t>(inj_t++) * R * VOCAB, g, Rp);
      ++launches;
      return 1;
    };
    // step 0 always runs eagerly (it also makes sure every kernel is loaded and configured before a capture)
    if (steps > 0) step_fn();
    if (opt_cuda_graph && !prof_on && steps > 1 && !injected_logits) {
      auto it = step_graphs.find(graph_key);
      if (it == step_graphs.end()) {
        cudaGraph_t graph;
        const int64_t saved = launches;
        CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
          nodes = decode_step(Rp, g, nullptr, st);
        } catch (...) {
          cudaGraph_t dead;
          cudaStreamEndCapture(st, &dead);
          throw;
        }
        CUDA_CHECK(cudaStreamEndCapture(st, &graph));
        launches = saved;
        CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        step_graphs[graph_key] = exec;
        step_graph_nodes[graph_key] = nodes;
      } else {
        exec = it->second;
        nodes = step_graph_nodes[graph_key];
      }
    }
    int done_steps = steps > 0 ? 1 : 0;
    std::vector<int> counts(steps > 0 ? steps : 1);
    for (int t = 1; t < steps; ++t) {
      if (exec) {
        CUDA_CHECK(cudaGraphLaunch(exec, st));
        launches += nodes;
      } else {
        step_fn();
      }
      ++done_steps;
      if ((t & 7) == 7 && t + 1 < steps) {  // early exit without a per-step sync (language_model.py:649)
        int c = 1;
        CUDA_CHECK(cudaMemcpyAsync(&c, unf_count.as<int>() + t, 4, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (c == 0) break;
      }
    }
    std::vector<int32_t> tmp(static_cast<size_t>(R) * max_length);
    CUDA_CHECK(cudaMemcpyAsync(tmp.data(), ids.as<int>(), tmp.size() * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(counts.data(), unf_count.as<int>(), static_cast<size_t>(done_steps) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    // the reference stops right after the first step that leaves no row unfinished (language_model.py:649)
    int width = done_steps + 1;
    for (int t = 0; t < done_steps; ++t) {
      if (counts[t] == 0) {
        width = t + 2;
        break;
      }
    }
    for (int r = 0; r < R; ++r) {
      for (int c = 0; c < max_length; ++c)
        out_ids[static_cast<size_t>(r) * max_length + c] = c < width ? tmp[static_cast<size_t>(r) * max_length + c] : RGRG_EOS;
    }
    return width;
  }
};

ProfScope::ProfScope(rgrg_engine* e_, const char* tag, cudaStream_t st_) : e(e_), st(st_), rec(-1) {
  if (!e->prof_on) return;
  rgrg_engine::ProfRec r;
  r.cat = e->prof_cat(tag);
  r.a = e->prof_event();
  r.b = e->prof_event();
  cudaEventRecord(r.a, st);
  rec = static_cast<int>(e->prof_recs.size());
  e->prof_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (rec >= 0) cudaEventRecord(e->prof_recs[rec].b, st);
}

// ====================================================================================================================
// C ABI
// ====================================================================================================================
static std::string g_create_error;

#define RGRG_TRY(e, ...)                                          \
  try {                                                           \
    CUDA_CHECK(cudaSetDevice((e)->device));                       \
    __VA_ARGS__;                                                  \
    return 0;                                                     \
  } catch (const CudaError& ex) {                                 \
    (e)->err = ex.what();                                         \
    return ex.code == cudaErrorMemoryAllocation ? 2 : 1;          \
  } catch (const std::exception& ex) {                            \
    (e)->err = ex.what();                                         \
    return 1;                                                     \
  }

extern "C" {

const char* rgrg_version(void) { return "rgrg_b200 0.1 (sm_100a)"; }

int rgrg_create(int device, rgrg_engine_t** out) {
  try {
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) throw std::runtime_error("no such CUDA device: " + std::to_string(device));
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) throw std::runtime_error("rgrg_b200 requires a Blackwell (sm_100a) GPU; there is no fallback path");
    // kernel attributes (opt-in shared memory, cluster sizes) and the SM count are cached per process: one engine device
    // per process, which is the deployment model anyway (one process per GPU)
    static int process_device = -1;
    if (process_device >= 0 && process_device != device)
        throw std::runtime_error("rgrg_b200: this process already drives cuda:" + std::to_string(process_device) +
                                 "; use one process per GPU (torch.distributed / torchrun)");
    process_device = device;
    rgrg_engine* e = new rgrg_engine();
    e->device = device;
    const char* ic = getenv("RGRG_IMPLICIT_CONV");
    if (ic) e->opt_implicit_conv = atoi(ic);
    *out = e;
    return 0;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return 1;
  }
}

void rgrg_destroy(rgrg_engine_t* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  delete e;
}

const char* rgrg_last_error(const rgrg_engine_t* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int64_t rgrg_kernel_launches(const rgrg_engine_t* e) { return e->launches; }

int rgrg_profile_read(rgrg_engine_t* e, char* buf, size_t buflen) {
  RGRG_TRY(e, {
    const std::string r = e->prof_report();
    if (r.size() + 1 > buflen) throw std::runtime_error("profile buffer too small");
    memcpy(buf, r.c_str(), r.size() + 1);
    e->prof_reset();
  });
}

int rgrg_set_option(rgrg_engine_t* e, const char* key, int value) {
  const std::string k(key);
  if (k == "profile") {
    e->prof_on = value != 0;
    e->prof_reset();
    return 0;
  }
  if (k == "implicit_conv") e->opt_implicit_conv = value;
  else if (k == "cuda_graph") e->opt_cuda_graph = value;
  else if (k == "gemm_impl") e->opt_gemm_impl = value;
  else if (k == "pdl") e->opt_pdl = value;
  else if (k == "fused_attn") e->opt_fused_attn = value;
  else if (k == "ln_head") e->opt_ln_head = value;
  else if (k == "attn_slots") e->opt_attn_slots = value;
  else if (k == "attn_warps") e->opt_attn_warps = value;
  else if (k == "beam_fused_head") e->opt_beam_fused_head = value;
  else if (k == "roi_align_sep") e->opt_roi_align_sep = value;
  else if (k == "dual") e->opt_dual = value;
  else if (k == "gemm_2cta") e->opt_gemm_2cta = value;
  else if (k == "gemm_2cta_stages") e->opt_gemm_2cta_stages = value;
  else if (k == "epi_tma") e->opt_epi_tma = value;
  else if (k == "epi_tma_conv") e->opt_epi_tma_conv = value;
  else if (k == "attn_early") e->opt_attn_early = value;
  else if (k == "attn_mc") e->opt_attn_mc = value;
  else if (k == "gemm_2cta_waves") e->opt_gemm_2cta_waves = value;
  else if (k == "attn_balance") e->opt_attn_balance = value;
  else if (k == "trace") e->opt_trace = value;
  else if (k == "l2_ahead") e->opt_l2_ahead = value;
  else if (k == "attn_occ") e->opt_attn_occ = value;
  else if (k == "cattn_bn") e->opt_cattn_bn = value;
  else if (k == "ablate") e->opt_ablate = value;
  else if (k == "detector_precise") e->opt_detector_precise = value;
  else {
    e->err = "unknown option: " + k;
    return 1;
  }
  e->drop_step_graphs();  // every option may change what a decode step launches: never replay a stale graph
  return 0;
}

int rgrg_load_weight(rgrg_engine_t* e, const char* name, const float* host_data, const int64_t* shape, int ndim) {
  HostRef r;
  r.p = host_data;
  r.shape.assign(shape, shape + ndim);
  e->host[name] = r;
  return 0;
}

int rgrg_finalize_weights(rgrg_engine_t* e) { RGRG_TRY(e, e->finalize()); }

static const float* stage_images(rgrg_engine* e, const float* images, int on_host, int B, int S, cudaStream_t st) {
  if (!on_host) return images;
  e->ensure_detector_ws(B, S);
  CUDA_CHECK(cudaMemcpyAsync(e->images.p, images, static_cast<size_t>(B) * S * S * 4, cudaMemcpyHostToDevice, st));
  return e->images.as<float>();
}

static void read_detections(rgrg_engine* e, int B, uint8_t* out_selected, uint8_t* out_detected, float* out_boxes,
                            float* out_scores, float* out_trf, int32_t* out_top_idx, int32_t* out_np, cudaStream_t st) {
  const size_t rows = static_cast<size_t>(B) * NREG;
  if (out_selected) CUDA_CHECK(cudaMemcpyAsync(out_selected, e->selected.p, rows, cudaMemcpyDeviceToHost, st));
  if (out_detected) CUDA_CHECK(cudaMemcpyAsync(out_detected, e->detected.p, rows, cudaMemcpyDeviceToHost, st));
  if (out_boxes) CUDA_CHECK(cudaMemcpyAsync(out_boxes, e->top_boxes.p, rows * 16, cudaMemcpyDeviceToHost, st));
  if (out_scores) CUDA_CHECK(cudaMemcpyAsync(out_scores, e->top_scores.p, rows * 4, cudaMemcpyDeviceToHost, st));
  if (out_trf) CUDA_CHECK(cudaMemcpyAsync(out_trf, e->trf.p, rows * 1024 * 4, cudaMemcpyDeviceToHost, st));
  if (out_top_idx) CUDA_CHECK(cudaMemcpyAsync(out_top_idx, e->top_idx.p, rows * 4, cudaMemcpyDeviceToHost, st));
  if (out_np) CUDA_CHECK(cudaMemcpyAsync(out_np, e->prop_count.p, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, st));
  CUDA_CHECK(cudaStreamSynchronize(st));
}

static void check_ready(rgrg_engine* e) {
  if (!e->weights_ready) throw std::runtime_error("weights not loaded: call rgrg_finalize_weights first");
}

int rgrg_detect(rgrg_engine_t* e, const float* images, int images_on_host, int B, int S, uint8_t* out_selected,
                uint8_t* out_detected, float* out_boxes, float* out_scores, float* out_region_features,
                int32_t* out_top_idx, int32_t* out_num_proposals, uint8_t* out_abnormal, int* out_R, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    if (out_abnormal && !e->has_abnormal) throw std::runtime_error("checkpoint has no binary_classifier_region_abnormal weights");
    cudaStream_t st = e->enter(stream);
    const float* img = stage_images(e, images, images_on_host, B, S, st);
    e->want_abnormal = out_abnormal != nullptr;
    int R = 0;
    try {
      R = e->run_detect(img, B, S, st);
    } catch (...) {
      e->want_abnormal = false;
      throw;
    }
    e->want_abnormal = false;
    if (out_abnormal) CUDA_CHECK(cudaMemcpyAsync(out_abnormal, e->abnormal.p, static_cast<size_t>(B) * NREG, cudaMemcpyDeviceToHost, st));
    if (out_R) *out_R = R;
    read_detections(e, B, out_selected, out_detected, out_boxes, out_scores, out_region_features, out_top_idx,
                    out_num_proposals, st);
  });
}

int rgrg_generate(rgrg_engine_t* e, const float* images, int images_on_host, int B, int S, int max_length,
                  int num_beams, int early_stopping, int32_t* out_ids, int* out_width, uint8_t* out_selected,
                  uint8_t* out_detected, float* out_boxes, float* out_scores, int* out_R, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    if (num_beams < 1) throw std::runtime_error("num_beams must be >= 1");
    if (max_length < 2 || max_length > 1024) throw std::runtime_error("max_length must be in [2, 1024]");
    cudaStream_t st = e->enter(stream);
    const float* img = stage_images(e, images, images_on_host, B, S, st);
    const int R = e->run_detect(img, B, S, st);
    *out_R = R;
    int width = 0;
    if (R > 0)
      width = num_beams == 1 ? e->run_greedy(e->lm_in.as<bf16>(), R, max_length, out_ids, st)
                             : e->run_beam(e->lm_in.as<bf16>(), R, num_beams, max_length, early_stopping != 0, out_ids, st);
    if (out_width) *out_width = width;
    // keep the final ids on the device for rgrg_allgather_results (greedy: they already are; beam: finalize ran on the host)
    if (R > 0 && num_beams > 1)
      CUDA_CHECK(cudaMemcpyAsync(e->ids.p, out_ids, static_cast<size_t>(R) * max_length * 4, cudaMemcpyHostToDevice, st));
    e->last_R = R;
    e->last_width = width;
    e->last_T = max_length;
    read_detections(e, B, out_selected, out_detected, out_boxes, out_scores, nullptr, nullptr, nullptr, st);
  });
}

static const bf16* stage_feats(rgrg_engine* e, const float* feats, int on_host, int R, cudaStream_t st) {
  const int Rp = rgrg_engine::padded_rows(R);
  e->trf.ensure(static_cast<size_t>(R) * 1024 * 4);
  e->lm_in.ensure(static_cast<size_t>(Rp) * 1024 * 2);
  const float* src = feats;
  if (on_host) {
    CUDA_CHECK(cudaMemcpyAsync(e->trf.p, feats, static_cast<size_t>(R) * 1024 * 4, cudaMemcpyHostToDevice, st));
    src = e->trf.as<float>();
  }
  det::cast_bf16_kernel<<<rgrg_engine::grid_for(static_cast<long long>(R) * 1024), 256, 0, st>>>(src, e->lm_in.as<bf16>(),
                                                                                             static_cast<long long>(R) * 1024);
  KERNEL_CHECK();
  if (Rp > R) CUDA_CHECK(cudaMemsetAsync(e->lm_in.as<bf16>() + static_cast<size_t>(R) * 1024, 0, static_cast<size_t>(Rp - R) * 1024 * 2, st));
  ++e->launches;
  return e->lm_in.as<bf16>();
}

int rgrg_lm_generate(rgrg_engine_t* e, const float* feats, int feats_on_host, int R, int max_length, int num_beams,
                     int early_stopping, int32_t* out_ids, int* out_width, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    if (num_beams < 1) throw std::runtime_error("num_beams must be >= 1");
    if (max_length < 2 || max_length > 1024) throw std::runtime_error("max_length must be in [2, 1024]");
    if (R <= 0) throw std::runtime_error("R must be positive");
    cudaStream_t st = e->enter(stream);
    const bf16* f = stage_feats(e, feats, feats_on_host, R, st);
    const int width = num_beams == 1 ? e->run_greedy(f, R, max_length, out_ids, st)
                                     : e->run_beam(f, R, num_beams, max_length, early_stopping != 0, out_ids, st);
    if (out_width) *out_width = width;
  });
}

int rgrg_bbox_features(rgrg_engine_t* e, const float* images, int images_on_host, int B, int S, const float* boxes_host,
                       float* out_features_host, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    cudaStream_t st = e->enter(stream);
    const float* img = stage_images(e, images, images_on_host, B, S, st);
    e->ensure_detector_ws(B, S);
    const int f = S / 32;
    e->run_backbone(img, B, S, e->feats.as<bf16>(), st);
    // the 29 user boxes of image b occupy proposal slots 0..28; region c uses slot c
    std::vector<float> slots(static_cast<size_t>(B) * TOPK * 4, 0.0f);
    std::vector<int> cnt(B, NREG), idx(static_cast<size_t>(B) * NREG);
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < NREG; ++c) {
        memcpy(&slots[(static_cast<size_t>(b) * TOPK + c) * 4], boxes_host + (static_cast<size_t>(b) * NREG + c) * 4, 16);
        idx[b * NREG + c] = c;
      }
    CUDA_CHECK(cudaMemcpyAsync(e->prop_boxes.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(e->prop_count.p, cnt.data(), cnt.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(e->top_idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, st));
    const float scale = exp2f(roundf(log2f(static_cast<float>(f) / static_cast<float>(S))));
    det::roi_mean_kernel<bf16><<<dim3(NREG, B), 256, 0, st>>>(e->feats.as<bf16>(), e->prop_boxes.as<float>(), e->prop_count.as<int>(),
                                                        e->top_idx.as<int>(), e->mean2048.as<float>(), f, 2048, scale);
    KERNEL_CHECK();
    const int rows = B * NREG;
    e->simt_f32(e->mean2048.as<float>(), e->dimred.w, rows, 1024, 2048,
                rgrg_engine::epi<false, ACT_NONE, RES_NONE, true>(e->trf.p, e->dimred.bias, 1024), st);
    e->launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(out_features_host, e->trf.p, static_cast<size_t>(rows) * 1024 * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));  // `slots`, `cnt`, `idx` must outlive the copies
  });
}

int rgrg_preprocess(rgrg_engine_t* e, const uint8_t* image, int image_on_host, int H, int W, float* out, int out_on_host,
                    void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    constexpr int S = 512;  // IMAGE_INPUT_SIZE, generate_reports_for_images.py:26
    if (H <= 0 || W <= 0) throw std::runtime_error("bad image size");
    const det::PreprocTab& tab = e->preproc_table(H, W, S);
    const uint8_t* src = image;
    if (image_on_host) {
      e->preproc_src.ensure(static_cast<size_t>(H) * W);
      CUDA_CHECK(cudaMemcpyAsync(e->preproc_src.p, image, static_cast<size_t>(H) * W, cudaMemcpyHostToDevice, st));
      src = e->preproc_src.as<uint8_t>();
    }
    float* dst = out;
    if (out_on_host) {
      e->preproc_out.ensure(static_cast<size_t>(S) * S * 4);
      dst = e->preproc_out.as<float>();
    }
    det::preprocess_kernel<<<dim3(S / 32, S / 8), 256, 0, st>>>(src, tab, dst, S);
    KERNEL_CHECK();
    ++e->launches;
    if (out_on_host) CUDA_CHECK(cudaMemcpyAsync(out, dst, static_cast<size_t>(S) * S * 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_comm_unique_id(void* out_id_128_bytes) {
  try {
    nccl::UniqueId id;
    nccl::check(nccl::api().get_unique_id(&id), "ncclGetUniqueId");
    memcpy(out_id_128_bytes, &id, sizeof(id));
    return 0;
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return 1;
  }
}

int rgrg_comm_init(rgrg_engine_t* e, const void* id_128_bytes, int rank, int world) {
  RGRG_TRY(e, {
    if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("bad rank / world size");
    if (e->comm) {
      nccl::api().comm_destroy(e->comm);
      e->comm = nullptr;
    }
    nccl::UniqueId id;
    memcpy(&id, id_128_bytes, sizeof(id));
    nccl::check(nccl::api().comm_init_rank(&e->comm, world, id, rank), "ncclCommInitRank");
    e->comm_rank = rank;
    e->comm_world = world;
  });
}

int rgrg_allgather_results(rgrg_engine_t* e, int B, int max_length, uint8_t* out_host, size_t blob_bytes, void* stream) {
  RGRG_TRY(e, {
    if (!e->comm) throw std::runtime_error("rgrg_comm_init has not been called");
    const int rows = B * NREG;
    const size_t need = 8 + static_cast<size_t>(rows) * max_length * 4 + 2 * static_cast<size_t>(rows) + static_cast<size_t>(rows) * 20;
    if (blob_bytes != need) throw std::runtime_error("blob size does not match (B, max_length)");
    if (B != e->last_B || max_length != e->last_T) throw std::runtime_error("rgrg_allgather_results must follow rgrg_generate with the same B and max_length");
    cudaStream_t st = e->enter(stream);
    e->blob_dev.ensure(need);
    e->gather_dev.ensure(need * e->comm_world);
    det::pack_blob_kernel<<<rgrg_engine::grid_for(static_cast<long long>(need)), 256, 0, st>>>(
        e->blob_dev.as<uint8_t>(), e->ids.as<int>(), max_length, e->last_R, e->last_width, rows, max_length, e->selected.as<uint8_t>(),
        e->detected.as<uint8_t>(), e->top_boxes.as<float>(), e->top_scores.as<float>());
    KERNEL_CHECK();
    ++e->launches;
    nccl::check(nccl::api().all_gather(e->blob_dev.p, e->gather_dev.p, need, /*ncclUint8*/ 1, e->comm, st), "ncclAllGather");
    CUDA_CHECK(cudaMemcpyAsync(out_host, e->gather_dev.p, need * e->comm_world, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_lm_forced_logits(rgrg_engine_t* e, const float* feats_dev, int R, const int32_t* forced_ids_dev, int n_tokens,
                          float* out_logits_dev, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    cudaStream_t st = e->enter(stream);
    const bf16* f = stage_feats(e, feats_dev, 0, R, st);
    e->ensure_decoder_ws(R, n_tokens + 1);
    e->lm_prologue(f, R, 1, st);
    dec::GreedyState g{};
    g.ids = e->ids.as<int>();
    g.ids_ld = n_tokens;
    g.unfinished = e->unfinished.as<int>();
    g.unfinished_count = e->unf_count.as<int>();
    g.step_ptr = e->step.as<int>();
    g.ticket = e->step.as<int>() + 1;
    g.live_rows = R;
    g.forced = forced_ids_dev;
    CUDA_CHECK(cudaMemsetAsync(e->ln_counters.p, 0, 1024 * 4, st));
    dec::greedy_init_kernel<<<ceil_div(std::max(R, n_tokens), 256), 256, 0, st>>>(g, R);
    KERNEL_CHECK();
    ++e->launches;
    for (int t = 0; t < n_tokens; ++t) e->decode_step(R, g, out_logits_dev + static_cast<size_t>(t) * R * VOCAB, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_greedy_bookkeeping(rgrg_engine_t* e, const float* logits_steps_dev, int n_steps, int R, int max_length, int32_t* out_ids,
                            int* out_width, void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    if (R <= 0 || max_length < 2) throw std::runtime_error("R must be positive and max_length >= 2");
    const int width = e->run_greedy(nullptr, R, max_length, out_ids, st, logits_steps_dev, n_steps);
    if (out_width) *out_width = width;
  });
}

int rgrg_beam_bookkeeping(rgrg_engine_t* e, const float* logits_steps_dev, int n_steps, int sentences, int num_beams,
                          int max_length, int early_stopping, int32_t* out_ids, int* out_width, void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    if (num_beams > dec::MAX_BEAMS || num_beams < 2) throw std::runtime_error("num_beams must be in [2, 8]");
    const int rows = sentences * num_beams;
    e->ensure_decoder_ws(rows, max_length);
    dec::BeamState s = e->beam_state(sentences, num_beams, max_length, early_stopping != 0);
    dec::beam_init_kernel<<<ceil_div(rows, 256), 256, 0, st>>>(s, sentences);
    KERNEL_CHECK();
    int cur = 0;
    int cur_len = 1;
    std::vector<int> nd(n_steps);
    for (int t = 0; t < n_steps && t < max_length - 1; ++t) {
      e->beam_bookkeeping(s, logits_steps_dev + static_cast<size_t>(t) * rows * VOCAB, sentences, cur, st);
      cur ^= 1;
      ++cur_len;
      int c = 1;
      CUDA_CHECK(cudaMemcpyAsync(&c, s.not_done_count + t, 4, cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
      if (c == 0) break;
    }
    const int width = e->beam_finalize(s, sentences, cur_len, cur, max_length, out_ids, st);
    if (out_width) *out_width = width;
  });
}

int rgrg_rpn_filter(rgrg_engine_t* e, const float* objectness_dev, const float* deltas_dev, const float* decoded_dev, int B,
                    int feat, int image_size, float* boxes_dev, float* scores_dev, int32_t* count_dev,
                    int32_t* topk_idx_dev, int32_t* keep_rank_dev, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    cudaStream_t st = e->enter(stream);
    const long long N = static_cast<long long>(feat) * feat * det::NUM_ANCHORS;
    det::RpnIn in{};
    in.obj = objectness_dev;
    in.obj_bs = N;
    in.obj_ps = det::NUM_ANCHORS;
    in.deltas = deltas_dev;
    in.del_bs = N * 4;
    in.del_ps = det::NUM_ANCHORS * 4;
    in.decoded = decoded_dev;
    det::RpnOut out{boxes_dev, scores_dev, count_dev, topk_idx_dev, keep_rank_dev};
    e->run_rpn_filter(in, out, B, feat, image_size, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_roi_align(rgrg_engine_t* e, const void* feats_bf16_dev, const float* boxes_dev, const int32_t* count_dev, int B,
                   int feat, int C, int image_size, void* out_bf16_dev, void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    if (C % 8) throw std::runtime_error("C must be a multiple of 8");
    e->roi_off.ensure(static_cast<size_t>(B + 1) * 4);
    det::roi_offsets_kernel<<<1, 32, 0, st>>>(count_dev, e->roi_off.as<int>(), B);
    const float scale = exp2f(roundf(log2f(static_cast<float>(feat) / static_cast<float>(image_size))));
    if (e->opt_roi_align_sep && feat <= 32)
      det::roi_align_sep_kernel<<<dim3(TOPK, B), 256, 0, st>>>(static_cast<const bf16*>(feats_bf16_dev), boxes_dev, count_dev,
                                                               e->roi_off.as<int>(), static_cast<bf16*>(out_bf16_dev), feat, C, scale);
    else
      det::roi_align_kernel<bf16><<<dim3(TOPK, B), 256, 0, st>>>(static_cast<const bf16*>(feats_bf16_dev), boxes_dev, count_dev,
                                                                 e->roi_off.as<int>(), static_cast<bf16*>(out_bf16_dev), feat, C, scale);
    KERNEL_CHECK();
    e->launches += 2;
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_roi_tail(rgrg_engine_t* e, const float* class_logits_dev, const float* box_regression_dev, const float* boxes_dev,
                  const int32_t* count_dev, int B, int image_size, uint8_t* detected_dev, int32_t* top_idx_dev,
                  float* scores_dev, float* top_boxes_dev, void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    e->roi_off.ensure(static_cast<size_t>(B + 1) * 4);
    det::roi_offsets_kernel<<<1, 32, 0, st>>>(count_dev, e->roi_off.as<int>(), B);
    det::RoiTailOut to{detected_dev, top_idx_dev, scores_dev, top_boxes_dev};
    det::roi_tail_kernel<<<B, 256, 0, st>>>(class_logits_dev, 30, box_regression_dev, 120, boxes_dev, count_dev,
                                            e->roi_off.as<int>(), to, image_size);
    KERNEL_CHECK();
    e->launches += 2;
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_gemm_bf16(rgrg_engine_t* e, const void* A_dev, const void* W_dev, const float* bias_dev, int M, int N, int K, int act,
                   int impl, float* out_dev, void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    Linear L;
    L.w = static_cast<bf16*>(const_cast<void*>(W_dev));
    L.bias = const_cast<float*>(bias_dev);
    L.N = N;
    L.K = K;
    const int saved = e->opt_gemm_impl;
    e->opt_gemm_impl = impl == 2 ? 2 : 0;
    const int fbn = impl == 0 ? 128 : impl == 1 ? 64 : impl == 3 ? 192 : impl == 4 ? 256 : 0;
    const bf16* A = static_cast<const bf16*>(A_dev);
    try {
      if (impl != 2) L.make_maps();
      if (impl == 7 || impl == 8) e->use_2cta_now = e->force_2cta = true;  // CTA-pair kernel: plain (7) / split-K (8)
      struct Reset2 {
        bool &f, &g;
        ~Reset2() { f = g = false; }
      } reset2{e->use_2cta_now, e->force_2cta};
      if (impl == 6 || impl == 8) {  // the decoder's split-K form: 4 K slices -> fp32 partial sums, reduced (+ bias) afterwards
        if (act != ACT_NONE) throw std::runtime_error("split-K test path has no activation");
        DevBuf parts;
        parts.ensure(static_cast<size_t>(4) * M * N * 4);
        e->opt_gemm_impl = 0;
        e->gemm_splitk("test_gemm", A, M, L, parts.as<float>(), 4, st);
        const long long total = static_cast<long long>(M) * N;
        det::sum_parts_kernel<<<rgrg_engine::grid_for(total), 256, 0, st>>>(parts.as<float>(), 4, total, bias_dev, N, out_dev);
        KERNEL_CHECK();
        CUDA_CHECK(cudaStreamSynchronize(st));
        parts.release();
        e->opt_gemm_impl = saved;
        return 0;
      }
      if (bias_dev) {
        if (act == ACT_RELU) e->gemm("test_gemm", A, M, L, rgrg_engine::epi<false, ACT_RELU, RES_NONE, true>(out_dev, bias_dev, N), st, true, fbn);
        else if (act == ACT_GELU_NEW) e->gemm("test_gemm", A, M, L, rgrg_engine::epi<false, ACT_GELU_NEW, RES_NONE, true>(out_dev, bias_dev, N), st, true, fbn);
        else e->gemm("test_gemm", A, M, L, rgrg_engine::epi<false, ACT_NONE, RES_NONE, true>(out_dev, bias_dev, N), st, true, fbn);
      } else {
        if (act == ACT_RELU) e->gemm("test_gemm", A, M, L, rgrg_engine::epi<false, ACT_RELU, RES_NONE, false>(out_dev, nullptr, N), st, true, fbn);
        else if (act == ACT_GELU_NEW) e->gemm("test_gemm", A, M, L, rgrg_engine::epi<false, ACT_GELU_NEW, RES_NONE, false>(out_dev, nullptr, N), st, true, fbn);
        else e->gemm("test_gemm", A, M, L, rgrg_engine::epi<false, ACT_NONE, RES_NONE, false>(out_dev, nullptr, N), st, true, fbn);
      }
    } catch (...) {
      e->opt_gemm_impl = saved;
      throw;
    }
    e->opt_gemm_impl = saved;
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

// tuning harness: `iters` back-to-back launches of one GEMM on the engine stream (no host sync in between), optionally
// interleaved with a LayerNorm launch; returns the average time per iteration and the last launch's CTA timeline
int rgrg_gemm_bench(rgrg_engine_t* e, int M, int N, int K, int bn, int iters, int interleave, float* out_ms,
                    long long* trace_host, int trace_ctas) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(nullptr);
    DevBuf A, W, O, H, G, T;
    A.ensure(static_cast<size_t>(M) * K * 2);
    W.ensure(static_cast<size_t>(N) * K * 2);
    O.ensure(static_cast<size_t>(M) * N * 2);
    H.ensure(static_cast<size_t>(M) * 1024 * 4);
    G.ensure(1024 * 4);
    T.ensure(static_cast<size_t>(160) * 8 * 8);
    CUDA_CHECK(cudaMemsetAsync(A.p, 0, A.bytes, st));
    CUDA_CHECK(cudaMemsetAsync(W.p, 0, W.bytes, st));
    CUDA_CHECK(cudaMemsetAsync(H.p, 0, H.bytes, st));
    CUDA_CHECK(cudaMemsetAsync(G.p, 0, G.bytes, st));
    CUDA_CHECK(cudaMemsetAsync(T.p, 0, T.bytes, st));
    Linear L;
    L.w = W.as<bf16>();
    L.N = N;
    L.K = K;
    L.make_maps();
    auto ep = rgrg_engine::epi<true, ACT_NONE, RES_NONE, false>(O.p, nullptr, N);
    tc::GemmShape s{};
    s.M = M;
    s.N = N;
    s.k_iters = K / 64;
    s.m_tiles = ceil_div(M, 128);
    s.n_tiles = ceil_div(N, bn == 512 ? 256 : bn);
    s.m_fastest = 1;
    CUtensorMap tmA = tc::make_tmap_2d(A.as<bf16>(), M, K, 128);
    cudaEvent_t a;
    cudaEvent_t b;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
      if (rep == 1) CUDA_CHECK(cudaEventRecord(a, st));
      for (int i = 0; i < iters; ++i) {
        s.trace = (rep == 1 && i == iters - 1) ? T.as<long long>() : nullptr;
        if (bn == 512) {  // the CTA-pair kernel (256 x 256 pair-tiles)
          tc2::Shape s2{};
          s2.M = M;
          s2.N = N;
          s2.k_iters = K / 64;
          s2.m_pairs = ceil_div(ceil_div(M, 128), 2);
          s2.n_tiles = N / 256;
          tc2::launch<decltype(ep), 6>(tmA, L.tm[1], tmA, s2, ep, st, false);
        } else {
          e->launch_bn(bn, tmA, L, s, ep, st);
        }
        if (interleave)
          dec::layernorm_kernel<0><<<M, 128, 0, st>>>(H.as<float>(), G.as<float>(), G.as<float>(), A.as<bf16>(), M, nullptr, 0,
                                                                  nullptr, nullptr);
      }
      if (rep == 1) CUDA_CHECK(cudaEventRecord(b, st));
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    *out_ms = ms / iters;
    if (trace_host) CUDA_CHECK(cudaMemcpy(trace_host, T.p, static_cast<size_t>(trace_ctas) * 64, cudaMemcpyDeviceToHost));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    A.release(); W.release(); O.release(); H.release(); G.release(); T.release();
  });
}

int rgrg_conv3x3_bf16(rgrg_engine_t* e, const void* in_dev, const void* w_dev, const float* bias_dev, int B, int H, int W,
                      int Cin, int Cout, int relu, int implicit, float* out_dev, void* stream) {
  RGRG_TRY(e, {
    cudaStream_t st = e->enter(stream);
    Linear L;
    L.w = static_cast<bf16*>(const_cast<void*>(w_dev));
    L.bias = const_cast<float*>(bias_dev);
    L.N = Cout;
    L.K = 9 * Cin;
    L.make_maps();
    if (!bias_dev) throw std::runtime_error("conv test entry needs a bias");
    const int saved = e->opt_implicit_conv;
    e->opt_implicit_conv = implicit;
    try {
      if (relu) e->conv3x3("test_conv", static_cast<const bf16*>(in_dev), B, H, W, Cin, 1, L, rgrg_engine::epi<false, ACT_RELU, RES_NONE, true>(out_dev, bias_dev, Cout), st);
      else e->conv3x3("test_conv", static_cast<const bf16*>(in_dev), B, H, W, Cin, 1, L, rgrg_engine::epi<false, ACT_NONE, RES_NONE, true>(out_dev, bias_dev, Cout), st);
    } catch (...) {
      e->opt_implicit_conv = saved;
      throw;
    }
    e->opt_implicit_conv = saved;
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_backbone(rgrg_engine_t* e, const float* images_dev, int B, int S, void* out_feats_bf16_dev, void* stream) {
  RGRG_TRY(e, {
    check_ready(e);
    cudaStream_t st = e->enter(stream);
    e->ensure_detector_ws(B, S);
    e->run_backbone(images_dev, B, S, static_cast<bf16*>(out_feats_bf16_dev), st);
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int rgrg_debug_read(rgrg_engine_t* e, const char* name, void* host_dst, size_t bytes) {
  RGRG_TRY(e, {
    const std::string n(name);
    const DevBuf* b = nullptr;
    if (n == "rpn_out") b = &e->rpn_out;
    else if (n == "features") b = &e->feats;
    else if (n == "pred_out") b = &e->pred_out;
    else if (n == "proposals") b = &e->prop_boxes;
    else if (n == "proposal_scores") b = &e->prop_scores;
    else if (n == "selection_logits") b = &e->sel_logits;
    else if (n == "abnormal_logits") b = &e->abn_logits;
    else if (n == "decode_trace") b = &e->trace_buf;
    else if (n == "region_features_2048") b = &e->mean2048;
    else if (n == "fc7") b = &e->f7;
    else if (n.rfind("max_clusters_", 0) == 0) {  // tuning: co-resident clusters of the given size for a 216 KB-smem GEMM CTA
      const int cs = atoi(n.c_str() + 13);
      auto kern = tc::gemm_tc_kernel<256, 4, EpiStoreT<true, ACT_GELU_NEW, RES_NONE, true>, true>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SmemLayout<256, 4>::TOTAL);
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 16);
      cfg.blockDim = dim3(tc::NUM_THREADS);
      cfg.dynamicSmemBytes = tc::SmemLayout<256, 4>::TOTAL;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int nc = 0;
      const cudaError_t er = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
      cudaGetLastError();
      if (bytes < 4) throw std::runtime_error("need 4 bytes");
      *static_cast<int*>(host_dst) = er == cudaSuccess ? nc : -1;
      return 0;
    }
    else if (n == "cluster16_max_active") {
      if (bytes < 4) throw std::runtime_error("need 4 bytes");
      e->ln_head_available();
      *static_cast<int*>(host_dst) = e->cluster16_ok;
      return 0;
    }
    else throw std::runtime_error("unknown debug buffer: " + n);
    if (bytes > b->bytes) throw std::runtime_error("debug read larger than buffer: " + n);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(host_dst, b->p, bytes, cudaMemcpyDeviceToHost));
  });
}

}  // extern "C"
