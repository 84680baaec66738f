/*
 * rgrg_b200 — C ABI of the B200-native region-guided report-generation inference engine.
 *
 * Drop-in boundary for the reference's inference path (ttanida/rgrg @ 9520b6d):
 *   ReportGenerationModel.generate()            src/full_model/report_generation_model.py:212-276
 *   LanguageModel.generate()                    src/language_model/language_model.py:401-479
 *   ObjectDetector.forward() (inference)        src/object_detector/object_detector.py:184-261
 * The reference has no FFI of its own (100 % Python); these are the entry points a ctypes / cffi binding on the
 * reference side would bind (INTEGRATION.md shows that binding).  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; rgrg_last_error() then describes the failure
 *     (CUDA allocation failures contain the substring "out of memory", which callers of the reference
 *     string-match: src/full_model/evaluate_full_model/evaluate_language_model.py:1208);
 *   - one engine per GPU, not thread-safe, calls serialised by the caller (as in the reference: one Python thread);
 *   - the engine owns weights (device copies, repacked), workspace and KV cache; callers own inputs and outputs;
 *   - "dev" pointers are CUDA device pointers on the engine's device, "host" pointers are ordinary host memory;
 *   - `stream` is a cudaStream_t (NULL = default stream); work is enqueued on it and the call returns after the
 *     results it hands back on the host are complete.
 */
#ifndef RGRG_B200_H
#define RGRG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rgrg_engine rgrg_engine_t;

#if defined(__GNUC__)
#define RGRG_API __attribute__((visibility("default")))
#else
#define RGRG_API
#endif

#define RGRG_NUM_REGIONS 29
#define RGRG_VOCAB 50257
#define RGRG_EOS 50256
#define RGRG_MAX_PROPOSALS 1000

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */

/* replaces: ReportGenerationModel.__init__ + .to(device) (generate_reports_for_images.py:160-163) */
RGRG_API int rgrg_create(int device, rgrg_engine_t** out);
RGRG_API void rgrg_destroy(rgrg_engine_t* e);
RGRG_API const char* rgrg_last_error(const rgrg_engine_t* e); /* e may be NULL: error of the last failed rgrg_create */
RGRG_API const char* rgrg_version(void);

/* replaces: model.load_state_dict(checkpoint["model"]) (generate_reports_for_images.py:161).
 * `name` is the reference's state_dict key; fp32 host data, borrowed until rgrg_finalize_weights() returns.
 * Unknown / unused keys (wpe, causal_mask, alias trees ...) are accepted and ignored. */
RGRG_API int rgrg_load_weight(rgrg_engine_t* e, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* folds BatchNorm, repacks to K-major bf16, uploads.  Fails listing the first missing key. */
RGRG_API int rgrg_finalize_weights(rgrg_engine_t* e);

/* ---- the hot path ---------------------------------------------------------------------------------------------- */

/* replaces: ReportGenerationModel.generate(images, max_length, num_beams=1) (report_generation_model.py:212-276).
 * num_beams > 1 runs beam search (language_model.py:529-607) with BeamSearchScorer(length_penalty=1, keep 1).
 * images: fp32 [B,1,S,S] (host or device).  Host outputs:
 *   out_ids      int32 [B*29, max_length]  rows 0..R-1 valid (image-major, region-minor order of selected regions),
 *                                          column 0 = BOS, finished rows padded with 50256
 *   out_width    reference width of the id matrix (greedy_search returns [R, out_width], language_model.py:652)
 *   out_selected uint8 [B,29]; out_detected uint8 [B,29]; out_boxes fp32 [B,29,4]; out_scores fp32 [B,29]
 *   out_R        number of selected regions (0 -> the reference returns -1, report_generation_model.py:260-261) */
RGRG_API int rgrg_generate(rgrg_engine_t* e, const float* images, int images_on_host, int B, int S, int max_length,
                  int num_beams, int early_stopping, int32_t* out_ids, int* out_width, uint8_t* out_selected,
                  uint8_t* out_detected, float* out_boxes, float* out_scores, int* out_R, void* stream);

/* replaces: LanguageModel.generate(image_hidden_states[R,1024], max_length) (language_model.py:401-479; the entry the
 * bbox-variation caller uses directly, evaluate_bbox_variations.py:131-136).  feats fp32 [R,1024] host or device. */
RGRG_API int rgrg_lm_generate(rgrg_engine_t* e, const float* feats, int feats_on_host, int R, int max_length, int num_beams,
                     int early_stopping, int32_t* out_ids /* host [R, max_length] */, int* out_width, void* stream);

/* replaces: ObjectDetector.forward(images) + BinaryClassifierRegionSelection.forward (inference branch) and, when
 * out_abnormal is given, BinaryClassifierRegionAbnormal.forward (binary_classifier_region_abnormal.py:31-57, the eval-mode
 * extra of report_generation_model.py:103-106; `logit > -1`, NOT masked by class_detected — the reference masks later).
 * Host outputs as in rgrg_generate; out_region_features fp32 host [B,29,1024] (may be NULL);
 * out_top_idx int32 host [B,29] (may be NULL), out_num_proposals int32 host [B] (may be NULL),
 * out_abnormal uint8 host [B,29] (may be NULL). */
RGRG_API int rgrg_detect(rgrg_engine_t* e, const float* images, int images_on_host, int B, int S, uint8_t* out_selected,
                uint8_t* out_detected, float* out_boxes, float* out_scores, float* out_region_features,
                int32_t* out_top_idx, int32_t* out_num_proposals, uint8_t* out_abnormal, int* out_R, void* stream);

/* replaces: get_bbox_features(model, images, bbox_coordinates) (evaluate_bbox_variations.py:92-110): user boxes -> backbone ->
 * RoIAlign 8x8 -> AvgPool(8) -> dim_reduction.  boxes host fp32 [B,29,4] (x1,y1,x2,y2 in pixels);
 * out_features host fp32 [B*29,1024], the input of rgrg_lm_generate (the "always 29 rows per image" mode). */
RGRG_API int rgrg_bbox_features(rgrg_engine_t* e, const float* images, int images_on_host, int B, int S, const float* boxes_host,
                       float* out_features_host, void* stream);

/* ---- stage-level entry points (teacher-forced parity tests and the roofline harness call these) ---------------- */

/* decoder logits with forced tokens: forced_ids int32 dev [R, n_tokens]; out_logits fp32 dev [n_tokens, R, 50257]
 * (logits after consuming token t, language_model.py:258-399 with the cache of :169-170). */
RGRG_API int rgrg_lm_forced_logits(rgrg_engine_t* e, const float* feats_dev, int R, const int32_t* forced_ids_dev,
                          int n_tokens, float* out_logits_dev, void* stream);

/* Multi-GPU result gather (the reference has no distributed code).  One process per GPU; rank 0 obtains an
 * id with rgrg_comm_unique_id and hands it to the other ranks by any means (bench.py: torch.distributed broadcast); every
 * rank then calls rgrg_comm_init.  rgrg_allgather_results, called right after rgrg_generate with the same B and
 * max_length, packs this rank's results ON THE DEVICE into the fixed-size blob of rgrg_b200/parallel.py (head [R, width] |
 * ids int32 [B*29, max_length] padded with 50256 | selected | detected | boxes | scores), runs ONE ncclAllGather
 * (device to device over NVLink) and copies the world_size blobs to out_host.  NCCL is resolved at run time (dlopen of the
 * libnccl.so.2 already in the process); the library has no link-time dependency on it. */
RGRG_API int rgrg_comm_unique_id(void* out_id_128_bytes);
RGRG_API int rgrg_comm_init(rgrg_engine_t* e, const void* id_128_bytes, int rank, int world);
RGRG_API int rgrg_allgather_results(rgrg_engine_t* e, int B, int max_length, uint8_t* out_host, size_t blob_bytes, void* stream);

/* Pre-processing in front of the path (replaces `get_image_tensor`, src/full_model/generate_reports_for_images.py:129-147:
 * cv2 INTER_AREA resize to longest side 512 -> centre zero-pad to 512x512 -> (x/255 - 0.471)/0.302), bit-exact against
 * cv2.resize + the albumentations 1.1.0 transforms.  image: uint8 grayscale [H, W] (host or device); out: fp32 [512*512]
 * (host or device) = one image of the [B,1,512,512] batch rgrg_generate consumes.  Down-scaling only (max(H,W) >= 512). */
RGRG_API int rgrg_preprocess(rgrg_engine_t* e, const uint8_t* image, int image_on_host, int H, int W, float* out, int out_on_host,
                    void* stream);

/* greedy-search bookkeeping only (language_model.py:629-650: arg-max, pad-if-finished, EOS tracking, stop rule) driven by
 * given logits: logits_steps dev fp32 [n_steps, R, 50257]; out_ids host int32 [R, max_length]; out_width = reference width.
 * Runs the same device kernels and host loop (incl. the every-8-steps exit check) as generate(). */
RGRG_API int rgrg_greedy_bookkeeping(rgrg_engine_t* e, const float* logits_steps_dev, int n_steps, int R, int max_length,
                            int32_t* out_ids, int* out_width, void* stream);

/* beam-search bookkeeping only (language_model.py:556-605 + BeamSearchScorer.process / finalize) driven by given logits:
 * logits_steps dev fp32 [n_steps, sentences*num_beams, 50257] (row b of step t is what beam slot b sees at step t);
 * out_ids host int32 [sentences, max_length]; stops early when every sentence is done, like the reference loop. */
RGRG_API int rgrg_beam_bookkeeping(rgrg_engine_t* e, const float* logits_steps_dev, int n_steps, int sentences, int num_beams,
                          int max_length, int early_stopping, int32_t* out_ids, int* out_width, void* stream);

/* RPN filter_proposals (torchvision rpn.py:242-297) on given fp32 head outputs.
 * objectness dev [B,N]; deltas dev [B,N,4] (or NULL when decoded_boxes dev [B,N,4] is given); N = feat*feat*160.
 * outputs (dev): boxes [B,1000,4], scores [B,1000], count [B], topk_idx [B,1000] (opt), keep_rank [B,1000] (opt) */
RGRG_API int rgrg_rpn_filter(rgrg_engine_t* e, const float* objectness_dev, const float* deltas_dev, const float* decoded_dev,
                    int B, int feat, int image_size, float* boxes_dev, float* scores_dev, int32_t* count_dev,
                    int32_t* topk_idx_dev, int32_t* keep_rank_dev, void* stream);

/* RoIAlign 8x8 / sampling 2 (torchvision roi_align.py) : feats bf16 NHWC dev [B,f,f,C]; boxes dev [B,1000,4];
 * count dev [B]; out bf16 dev [sum(count), 64, C] (row order: image-major) */
RGRG_API int rgrg_roi_align(rgrg_engine_t* e, const void* feats_bf16_dev, const float* boxes_dev, const int32_t* count_dev,
                   int B, int feat, int C, int image_size, void* out_bf16_dev, void* stream);

/* per-class top-1 selection (custom_roi_heads.py:63-208): class_logits dev [sum(count),30], box_regression dev
 * [sum(count),120]; outputs dev: detected uint8 [B,29], top_idx int32 [B,29], scores [B,29], boxes [B,29,4] */
RGRG_API int rgrg_roi_tail(rgrg_engine_t* e, const float* class_logits_dev, const float* box_regression_dev,
                  const float* boxes_dev, const int32_t* count_dev, int B, int image_size, uint8_t* detected_dev,
                  int32_t* top_idx_dev, float* scores_dev, float* top_boxes_dev, void* stream);

/* D[M,N] (fp32 dev) = A[M,K] (bf16 dev) * W[N,K]^T (bf16 dev) + bias[N] (fp32 dev or NULL).
 * impl: 0 / 1 / 3 / 4 = tcgen05 with N tile 128 / 64 / 192 / 256, 5 = tcgen05 with the engine's own tile choice,
 *       2 = CUDA-core cross-check, 6 = the decoder's split-K form (4 K slices -> fp32 partial sums -> reduce, act must be 0),
 *       7 / 8 = the CTA-pair kernel (tcgen05.mma.cta_group::2, N must be a multiple of 256): plain / split-K.
 *       act: 0 none, 1 relu, 2 gelu_new. */
RGRG_API int rgrg_gemm_bf16(rgrg_engine_t* e, const void* A_dev, const void* W_dev, const float* bias_dev, int M, int N, int K,
                   int act, int impl, float* out_dev, void* stream);

/* tuning harness: iters back-to-back launches of one bf16 GEMM (N tile `bn`), optionally interleaved with a LayerNorm launch;
 * out_ms = average device time per iteration; trace_host (optional) = [trace_ctas, 8] clock64 timeline of the last launch */
RGRG_API int rgrg_gemm_bench(rgrg_engine_t* e, int M, int N, int K, int bn, int iters, int interleave, float* out_ms,
                    long long* trace_host, int trace_ctas);

/* 3x3 / pad 1 / stride 1 conv, NHWC: in bf16 dev [B,H,W,Cin]; w bf16 dev [Cout, 9*Cin] (tap-major); out fp32 dev
 * [B,H,W,Cout].  implicit: 1 = TMA implicit GEMM (4-D tensor map, OOB zero fill), 0 = im2col + GEMM. */
RGRG_API int rgrg_conv3x3_bf16(rgrg_engine_t* e, const void* in_dev, const void* w_dev, const float* bias_dev, int B, int H,
                      int W, int Cin, int Cout, int relu, int implicit, float* out_dev, void* stream);

/* backbone only: images fp32 dev [B,1,S,S] -> features bf16 NHWC dev [B,S/32,S/32,2048] */
RGRG_API int rgrg_backbone(rgrg_engine_t* e, const float* images_dev, int B, int S, void* out_feats_bf16_dev, void* stream);

/* copy a named internal buffer of the last call to the host (tests): "rpn_out" fp32 [B*f*f,800],
 * "features" bf16 [B,f,f,2048], "pred_out" fp32 [P,150], "proposals" fp32 [B,1000,4], "selection_logits" fp32 [B*29] */
RGRG_API int rgrg_debug_read(rgrg_engine_t* e, const char* name, void* host_dst, size_t bytes);

/* behaviour switches (every change drops the cached decode-step graphs):
 *   "cuda_graph" (0/1)        replay one captured graph per decode step
 *   "pdl" (0/1)               programmatic dependent launch between the kernels of a decode step
 *   "fused_attn" (0/1)        greedy decode: c_attn + KV append + attention as one head-aligned kernel (attn_fused.cuh);
 *                             0 = c_attn GEMM (KV-append epilogue) + stand-alone attention kernel (always used by beam search)
 *   "gemm_2cta" (0/1)         decode projections through the CTA-pair kernel (256 x 256 tiles, each CTA stages half of W);
 *                             "gemm_2cta_waves" (1/2): used while the pair grid fits in this many waves (default 2)
 *                             "gemm_2cta_stages" (6/4/3): depth of its TMA ring
 *   "epi_tma" (0/1)           tcgen05 GEMMs with a plain bf16 / fp32-partial epilogue (decode projections, 1x1 convolutions, fc6,
 *                             fc7): results leave through shared-memory slabs + TMA stores (1, default) or through registers
 *                             and 16-byte global stores (0); bit-identical
 *                             "epi_tma_conv" (0/1): the same for the implicit 3x3 convolutions (4-D NHWC output boxes)
 *   "attn_balance" (0/1)      fused attention: rows spread evenly over as many M tiles as the SMs hold (1, default) or 128-row tiles
 *   "beam_fused_head" (0/1)   beam search: log-softmax + per-part top-k fused into the lm_head epilogue (1, default) or fp32 logits +
 *                             separate top-k kernels
 *   "roi_align_sep" (0/1)     separable RoIAlign kernel (1, default) or the direct one
 *   "attn_mc", "attn_early" (0/1)  fused attention: operand A shared between head pairs by TMA multicast / first K,V chunks
 *                             requested before the epilogue (both bit-identical, measured without gain, default 0)
 *   "trace" (0/1)             tuning only: %globaltimer stamps of the first / last CTA of the decode kernels
 *                             (tools/decode_timeline.py)
 *   "dual" (0/1)              greedy decode step as two row halves half a layer out of phase (measured slower)
 *   "ln_head" (0/1)           LayerNorm (+ split-K reduce + residual) as the 16-CTA-cluster head of the consumer GEMM;
 *                             0 = separate LayerNorm kernels
 *   "attn_warps" / "attn_slots" (16/2, 8/4, 24/1)   fused attention: warps per CTA and K/V ring slots per warp;
 *                             "l2_ahead" (n): items prefetched into L2
 *   "attn_occ" (5..8), "cattn_bn" (0/128/192/256)   tuning of the two-kernel attention path
 *   "implicit_conv" (0/1)     3x3 convs as TMA implicit GEMM (1) or im2col + GEMM (0)
 *   "gemm_impl" (0/2)         0 tcgen05, 2 CUDA-core cross-check of every bf16 GEMM
 *   "detector_precise" (0/1)  fp32 detector (parity mode: fp32 operands and activations on CUDA cores)
 *   "ablate" (bit mask)       tuning only: skip kernel groups of the decode step to attribute time
 *   "profile" (0/1)           record CUDA events around every kernel category on the launch stream (disables graph replay) */
RGRG_API int rgrg_set_option(rgrg_engine_t* e, const char* key, int value);

/* per-category device time since "profile" was switched on: text lines "<category> <total ms> <launches>\n" */
RGRG_API int rgrg_profile_read(rgrg_engine_t* e, char* buf, size_t buflen);

/* counters since creation: kernels launched by this library (bench.py's gpu_launches claim) */
RGRG_API int64_t rgrg_kernel_launches(const rgrg_engine_t* e);

#ifdef __cplusplus
}
#endif
#endif /* RGRG_B200_H */
