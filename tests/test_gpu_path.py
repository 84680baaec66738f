"""-m gpu: the CUDA path against the CPU oracle on the same seeded weights and inputs, through the reference-facing
API.  Integer / boolean stages are bit-exact given identical inputs (tests/test_gpu_kernels.py); here the whole chain runs
in bf16 on the tensor cores, so stage outputs carry the tolerances stated per assertion."""
import numpy as np
import pytest
import torch

import rgrg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(synth_sd):
    from rgrg_b200 import ReportGenerationModel

    m = ReportGenerationModel(pretrain_without_lm_model=True)
    m.load_state_dict(synth_sd)
    m.to(torch.device("cuda", 0))
    m.eval()
    return m


@pytest.fixture(scope="module")
def images():
    from rgrg_b200 import synth

    return synth.synthetic_images(2, 512, seed=1001)


@pytest.fixture(scope="module")
def oracle_detail(synth_sd, images):
    detail = {}
    with torch.no_grad():
        det = O.detect(synth_sd, images, detail)
        sel, feats, logits = O.region_selection(synth_sd, det["top_region_features"], det["class_detected"])
    detail.update(det=det, selected=sel, sel_feats=feats, sel_logits=logits)
    return detail


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_backbone_features_within_bf16_tolerance(model, images, oracle_detail):
    eng = model._engine()
    feats = eng.backbone(images.cuda()).float().cpu().permute(0, 3, 1, 2)
    assert _rel(feats, oracle_detail["features"]) < 0.05  # 53 chained bf16 convs (~4 %)


def test_tma_store_epilogue_of_the_persistent_gemm_is_bit_identical(model, images):
    """1x1 convolutions, fc6 and fc7 (plain GEMMs with a bf16 output, with and without a bf16 residual) leave through
    shared-memory slabs + TMA stores by default; `epi_tma=0` is the register epilogue: same arithmetic in the same order, so
    the backbone features, the fc7 activations and every detector output are identical bit for bit."""
    eng = model._engine()
    x = images.cuda()
    try:
        a = eng.backbone(x)
        da = eng.detect(x)
        fa = eng.debug_read("fc7", (64, 1024), np.uint16)
        eng.set_option("epi_tma", 0)
        b = eng.backbone(x)
        db = eng.detect(x)
        fb = eng.debug_read("fc7", (64, 1024), np.uint16)
    finally:
        eng.set_option("epi_tma", 1)
    assert torch.equal(a, b)
    assert np.array_equal(fa, fb)
    for k in ("selected", "detected", "boxes", "scores", "region_features", "top_idx", "num_proposals"):
        assert np.array_equal(da[k], db[k]), k


def _oracle_class_scores(synth_sd, oracle_detail, b, boxes):
    """The oracle's 29 per-class scores (softmax over 30, background dropped) of arbitrary boxes of image b."""
    with torch.no_grad():
        pooled = O.box_roi_pool(oracle_detail["features"][b:b + 1], [boxes], 512)
        _, cls, _ = O.box_head_and_predictor(synth_sd, pooled)
    return torch.softmax(cls, -1)[:, 1:]


def test_detector_masks_and_epsilon_optimal_regions(model, images, oracle_detail, synth_sd):
    """bf16 tensor-core detector against the fp32 oracle.  Boolean outputs are exact.  Per-class top-1 proposals are decided
    among ~900 overlapping proposals whose scores differ by less than bf16 resolution, so the index check
    is the epsilon-optimal one: the proposal the engine picked for class c must score, UNDER THE ORACLE, within EPS of the
    oracle's own best score for c (and the fp32 mode below pins the indices themselves)."""
    EPS = 0.001  # the per-class scores live in [0.02, 0.06]; measured gap 2e-4
    eng = model._engine()
    out = eng.detect(images.cuda())
    det = oracle_detail["det"]
    assert np.array_equal(out["detected"], det["class_detected"].numpy())
    assert np.array_equal(out["selected"], oracle_detail["selected"].numpy())
    ref_counts = np.array([p.shape[0] for p in oracle_detail["proposals"]])
    assert np.all(np.abs(out["num_proposals"] - ref_counts) <= 0.05 * ref_counts)
    props = eng.debug_read("proposals", (2, 1000, 4))
    worst = 0.0
    for b in range(2):
        chosen = torch.from_numpy(props[b][out["top_idx"][b]])  # [29, 4] the engine's pick per class
        sc = _oracle_class_scores(synth_sd, oracle_detail, b, chosen)  # [29, 29]
        own = sc[torch.arange(29), torch.arange(29)]
        gap = (det["top_scores"][b] - own).max().item()
        worst = max(worst, gap)
        assert gap <= EPS, "image %d: engine's pick is %.5f below the oracle's best score" % (b, gap)
        assert np.abs(out["scores"][b] - own.numpy()).max() < EPS  # the engine's own score of its pick
    print("epsilon-optimality gap (bf16 detector): %.5f" % worst)
    assert _rel(torch.from_numpy(out["region_features"]), det["top_region_features"]) < 0.35  # different (equally good) picks


@pytest.fixture(scope="module")
def precise_engine(synth_sd):
    from rgrg_b200 import Engine

    e = Engine(0)
    e.load_state_dict(synth_sd, detector_precise=True)
    yield e
    e.close()


def test_detector_fp32_mode_pins_region_indices(precise_engine, images, oracle_detail, synth_sd):
    """`detector_precise`: fp32 operands / activations / accumulation.  Proposals, per-class top-1 proposal INDICES, boxes,
    scores and region features equal the fp32 oracle's (the north-star's "bit-exact region indices" on identical
    arithmetic precision); a pick may differ only where the oracle's own top-1 / top-2 margin is below fp32 noise."""
    eng = precise_engine
    out = eng.detect(images.cuda())
    det = oracle_detail["det"]
    assert np.array_equal(out["detected"], det["class_detected"].numpy())
    assert np.array_equal(out["selected"], oracle_detail["selected"].numpy())
    props = eng.debug_read("proposals", (2, 1000, 4))
    mismatched = 0
    for b in range(2):
        ref_p = oracle_detail["proposals"][b]
        n = int(out["num_proposals"][b])
        assert n == ref_p.shape[0], "image %d: %d proposals vs %d" % (b, n, ref_p.shape[0])
        # the same proposals (px); two proposals whose objectness differs by fp32 noise may swap places in the score order,
        # so indices are compared through the box-to-box correspondence
        d = (torch.from_numpy(props[b][:n])[:, None, :] - ref_p[None, :, :]).abs().amax(-1)  # [n, n]
        dist, perm = d.min(1)
        assert dist.max().item() < 2e-2 and len(set(perm.tolist())) == n
        swapped = int((perm != torch.arange(n)).sum())
        assert swapped <= 0.02 * n, "%d proposals out of order" % swapped
        same = perm.numpy()[out["top_idx"][b]] == det_top_idx(oracle_detail, b)
        mismatched += int((~same).sum())
        if not same.all():
            sc = _oracle_class_scores(synth_sd, oracle_detail, b, torch.from_numpy(props[b][out["top_idx"][b]]))
            own = sc[torch.arange(29), torch.arange(29)]
            assert ((det["top_scores"][b] - own)[torch.from_numpy(~same)] < 2e-6).all()  # only fp32-noise ties may differ
        assert np.abs(out["boxes"][b][same] - det["top_region_boxes"][b].numpy()[same]).max() < 2e-2
        assert np.abs(out["scores"][b][same] - det["top_scores"][b].numpy()[same]).max() < 1e-5
        f_e, f_o = torch.from_numpy(out["region_features"][b][same]), det["top_region_features"][b][torch.from_numpy(same)]
        assert _rel(f_e, f_o) < 1e-3
    print("fp32 detector: %d of 58 per-class picks differ from the oracle" % mismatched)
    assert mismatched <= 2


def det_top_idx(oracle_detail, b):
    roi = oracle_detail["roi"]
    cls, reg = roi["class_logits"], roi["box_regression"]
    out = O.top_regions(cls, reg, oracle_detail["proposals"], 512)
    return out["top_idx"][b].numpy()


def test_abnormal_classifier_output(model, images, oracle_detail, synth_sd):
    """a9': BinaryClassifierRegionAbnormal (eval branch, binary_classifier_region_abnormal.py:53-57) as an extra output of
    rgrg_detect: `logit > -1`, not masked by class_detected."""
    eng = model._engine()
    out = eng.detect(images.cuda(), abnormal=True)
    feats = torch.from_numpy(out["region_features"])
    pred, logits = O.region_abnormal(synth_sd, feats, torch.from_numpy(out["detected"]))  # teacher-forced on the engine's features
    mine = eng.debug_read("abnormal_logits", (2, 29))
    assert np.abs(mine - logits.numpy()).max() < 1e-4
    confident = (logits.abs_() if False else (logits + 1).abs()) > 1e-3
    assert np.array_equal(out["predicted_abnormal_regions"][confident.numpy()], pred.numpy()[confident.numpy()])
    # and end to end against the oracle's own region features (fp32 decision head on bf16-path features)
    pred_o, logits_o = O.region_abnormal(synth_sd, oracle_detail["det"]["top_region_features"], oracle_detail["det"]["class_detected"])
    far = ((logits_o + 1).abs() > 0.2).numpy()
    assert np.array_equal(out["predicted_abnormal_regions"][far], pred_o.numpy()[far])


def test_decoder_logits_teacher_forced(model, synth_sd, oracle_detail):
    """Decoder logits on ORACLE region features and ORACLE tokens: max |dlogit| <= 0.05 (bf16 tolerance,
    SURVEY.md §8(d)); arg-max must agree wherever the oracle's top-1 / top-2 margin exceeds 0.1."""
    feats = oracle_detail["sel_feats"][:12].contiguous()
    rec = {}
    ids = O.lm_generate(synth_sd, feats, max_length=7, record=rec)
    forced = ids[:, :-1].to(torch.int32)
    logits = model._engine().lm_forced_logits(feats.cuda(), forced.cuda()).cpu()  # [n, R, V]
    for t, ref in enumerate(rec["logits"]):
        d = (logits[t] - ref).abs().max().item()
        assert d <= 0.05, "step %d: max |dlogit| = %g" % (t, d)
        top2 = ref.topk(2, dim=-1).values
        confident = (top2[:, 0] - top2[:, 1]) > 0.1
        assert torch.equal(logits[t].argmax(-1)[confident], ref.argmax(-1)[confident])


def test_lm_generate_greedy_shape_and_prefix(model, synth_sd, oracle_detail):
    feats = oracle_detail["sel_feats"][:6].contiguous()
    ref = O.lm_generate(synth_sd, feats, max_length=6)
    ids = model.language_model.generate(feats.cuda(), max_length=6)
    assert ids.shape == ref.shape and ids.dtype == torch.int64
    assert torch.equal(ids[:, 0].cpu(), ref[:, 0])
    assert (ids.cpu() == ref).float().mean().item() > 0.6  # free-running bf16 decode may legitimately branch


def test_generate_end_to_end_contract(model, images, oracle_detail):
    out = model.generate(images, max_length=5)  # host images: H2D inside the call
    ids, selected, detections, class_detected = out
    R = int(oracle_detail["selected"].sum())
    assert ids.shape == (R, 5) and ids.dtype == torch.int64
    assert torch.equal(selected.cpu(), oracle_detail["selected"])
    assert torch.equal(class_detected.cpu(), oracle_detail["det"]["class_detected"])
    assert detections["top_region_boxes"].shape == (2, 29, 4) and detections["top_scores"].shape == (2, 29)
    assert (ids[:, 0] == 50256).all()
    # CUDA-graph replay and eager launches give the same tokens
    model._engine().set_option("cuda_graph", 0)
    ids2 = model.generate(images, max_length=5)[0]
    model._engine().set_option("cuda_graph", 1)
    assert torch.equal(ids, ids2)


def test_gemm_cross_check_path_agrees(model, images):
    """The CUDA-core cross-check GEMM and the tcgen05 GEMM run the same network: outputs agree to bf16 noise."""
    eng = model._engine()
    a = eng.detect(images.cuda())
    eng.set_option("gemm_impl", 2)
    b = eng.detect(images.cuda())
    eng.set_option("gemm_impl", 0)
    assert np.array_equal(a["detected"], b["detected"])
    assert _rel(torch.from_numpy(a["region_features"]), torch.from_numpy(b["region_features"])) < 0.2


def test_beam_search_end_to_end_contract(model, synth_sd, oracle_detail):
    """Beam search through the reference-facing API.  Free-running bf16 vs fp32 beams may branch on near-ties, so the
    bit-exact check of the bookkeeping is teacher-forced (tests/test_gpu_kernels.py); here: contract + the first token,
    whose score gaps at step 0 are decided by the BOS logits alone."""
    feats = oracle_detail["sel_feats"][:3].contiguous()
    ref = O.lm_generate(synth_sd, feats, max_length=6, num_beams=4, early_stopping=True)
    ids = model.language_model.generate(feats.cuda(), max_length=6, num_beams=4, early_stopping=True)
    assert ids.shape == ref.shape and ids.dtype == torch.int64
    assert (ids[:, 0] == 50256).all()
    assert (ids.cpu() == ref).float().mean().item() > 0.5


def test_beam_search_graph_replay_equals_eager(model, oracle_detail):
    """Beam steps replayed as a CUDA graph of two steps (double-buffered token / ancestry tables) vs eager launches, with
    and without early stopping, odd and even step counts."""
    eng = model._engine()
    feats = oracle_detail["sel_feats"][:7].contiguous().cuda()
    for T, es in ((12, True), (13, False), (20, True)):
        eng.set_option("cuda_graph", 1)
        a = eng.lm_generate(feats, T, num_beams=4, early_stopping=es)
        b = eng.lm_generate(feats, T, num_beams=4, early_stopping=es)  # cached graph
        eng.set_option("cuda_graph", 0)
        c = eng.lm_generate(feats, T, num_beams=4, early_stopping=es)
        eng.set_option("cuda_graph", 1)
        assert np.array_equal(a, b) and np.array_equal(a, c), (T, es)


def test_beam_fused_head_matches_logits_path(model, oracle_detail):
    """Beam search with log-softmax + per-part top-k fused into the lm_head epilogue (no [rows, V] logits in HBM) vs the
    path that stores fp32 logits: same candidates; the log-softmax denominator is summed in a different order (1e-7), so
    only a near-tie may flip.  num_beams 4 (K = 8 lists) and 6 (K = 16 lists)."""
    eng = model._engine()
    feats = oracle_detail["sel_feats"][:9].contiguous().cuda()
    for nb in (4, 6):
        eng.set_option("beam_fused_head", 0)
        a = eng.lm_generate(feats, 14, num_beams=nb, early_stopping=True)
        eng.set_option("beam_fused_head", 1)
        b = eng.lm_generate(feats, 14, num_beams=nb, early_stopping=True)
        assert a.shape == b.shape
        assert np.array_equal(a[:, :3], b[:, :3]) and (a == b).mean() > 0.9, nb


def test_out_of_memory_is_recoverable(model, oracle_detail):
    """the reference's callers catch "out of memory" and carry on with the next batch
    (evaluate_language_model.py:1207-1222).  A KV cache that cannot be allocated (4096 rows x 1024 tokens = 412 GB) must raise
    a RuntimeError carrying that substring and leave the engine usable: workspace growth is transactional."""
    eng = model._engine()
    feats = oracle_detail["sel_feats"][:8].contiguous().cuda()
    ref = eng.lm_generate(feats, 6)
    big = torch.zeros(4096, 1024, device="cuda")
    with pytest.raises(RuntimeError, match="out of memory"):
        eng.lm_generate(big, 1024)
    assert np.array_equal(eng.lm_generate(feats, 6), ref)          # same small batch: workspace rebuilt from scratch
    assert eng.lm_generate(feats[:3], 9).shape == (3, 9)


ATTN_MC_DEFAULT = 0  # engine default of attn_mc


def _opts(eng, **kw):
    for k, v in kw.items():
        eng.set_option(k, v)


def test_fused_attention_is_bit_identical_to_two_kernel_attention(model, oracle_detail):
    """attn_fused.cuh (c_attn + KV append + attention in one head-aligned kernel) vs c_attn GEMM + attention kernel:
    same operand rounding and reduction order, so greedy tokens are IDENTICAL — at cache lengths that cross the 16-key
    chunk boundary (L = 17, 33) and for every ring depth."""
    eng = model._engine()
    feats = torch.cat([oracle_detail["sel_feats"]] * 3, 0).contiguous().cuda()  # 174 rows: a full and a partial M tile
    _opts(eng, ln_head=0, fused_attn=0)
    ref = eng.lm_generate(feats, 36)
    try:
        for warps, slots in ((8, 4), (16, 2), (24, 1)):
            for ahead in (0, 2):
                _opts(eng, fused_attn=1, attn_warps=warps, attn_slots=slots, l2_ahead=ahead)
                out = eng.lm_generate(feats, 36)
                assert np.array_equal(ref, out), "warps=%d slots=%d l2_ahead=%d" % (warps, slots, ahead)
        _opts(eng, cuda_graph=0)
        assert np.array_equal(ref, eng.lm_generate(feats, 36))
        # first K / V chunks requested before the epilogue instead of after it: same bits
        _opts(eng, cuda_graph=1, attn_warps=16, attn_slots=2, l2_ahead=0, attn_early=1)
        assert np.array_equal(ref, eng.lm_generate(feats, 36))
        # head pairs sharing operand A through TMA multicast (clusters of 2): same MMAs on the same bytes
        for mc in (1, 0):  # set_option drops the step graph each time
            _opts(eng, attn_early=0, attn_mc=mc)
            assert np.array_equal(ref, eng.lm_generate(feats, 36)), "attn_mc=%d" % mc
        _opts(eng, attn_mc=1, cuda_graph=0)
        assert np.array_equal(ref, eng.lm_generate(feats, 36))
        _opts(eng, cuda_graph=1)
    finally:
        _opts(eng, cuda_graph=1, fused_attn=1, attn_early=0, attn_mc=ATTN_MC_DEFAULT, attn_warps=16, attn_slots=2, l2_ahead=0, ln_head=0)


def test_layernorm_head_matches_separate_layernorm(model, oracle_detail):
    """LayerNorm as the cluster-cooperative head of the consumer GEMM vs separate LayerNorm kernels: same values up to
    the reduction order inside a row (one warp per row vs four warps per row), so tokens agree except at near-ties;
    graph replay and eager launches of the head variant are identical."""
    eng = model._engine()
    feats = torch.cat([oracle_detail["sel_feats"]] * 3, 0).contiguous().cuda()
    try:
        _opts(eng, ln_head=0)
        a = eng.lm_generate(feats, 16)
        _opts(eng, ln_head=1)
        b = eng.lm_generate(feats, 16)
        _opts(eng, cuda_graph=0)
        c = eng.lm_generate(feats, 16)
        _opts(eng, cuda_graph=1, fused_attn=0)  # head on c_fc only, two-kernel attention
        d = eng.lm_generate(feats, 16)
    finally:
        _opts(eng, cuda_graph=1, fused_attn=1, ln_head=0)
    assert np.array_equal(b, c) and np.array_equal(b, d)
    assert np.array_equal(a[:, :4], b[:, :4]) and (a == b).mean() > 0.9


def test_cta_pair_projections_are_bit_identical(model, oracle_detail):
    """Decode projections through the CTA-pair kernel vs the 1-CTA kernel: same k order and fp32 accumulation, so tokens
    are identical (full M tiles, a partial last tile, an odd tile count, more pairs than one wave)."""
    eng = model._engine()
    feats = torch.cat([oracle_detail["sel_feats"]] * 21, 0).contiguous().cuda()  # 1218 rows
    try:
        for rows in (406, 290, 60, 1218):  # 1218 rows: two waves of CTA pairs (gemm_2cta_waves = 2)
            _opts(eng, gemm_2cta=0)
            a = eng.lm_generate(feats[:rows], 16)
            _opts(eng, gemm_2cta=1)
            b = eng.lm_generate(feats[:rows], 16)
            _opts(eng, cuda_graph=0)
            c = eng.lm_generate(feats[:rows], 16)
            _opts(eng, cuda_graph=1, epi_tma=0)  # register epilogue instead of the shared-memory slabs + TMA stores
            d = eng.lm_generate(feats[:rows], 16)
            _opts(eng, epi_tma=1)
            assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d), rows
    finally:
        _opts(eng, gemm_2cta=1, cuda_graph=1, epi_tma=1)


def test_two_halves_schedule_is_bit_identical_to_single_chain(model, oracle_detail):
    """decode_forward_dual: two row halves on two streams, half a layer out of phase (cross-stream dependencies per layer).
    Rows never interact and the split is on an M-tile boundary, so tokens are IDENTICAL to the single chain — under graph
    replay and eagerly, with an even and an uneven split."""
    eng = model._engine()
    feats = torch.cat([oracle_detail["sel_feats"]] * 8, 0).contiguous().cuda()  # 464 rows
    try:
        for rows in (464, 300, 256):
            _opts(eng, dual=0, cuda_graph=1)
            a = eng.lm_generate(feats[:rows], 20)
            _opts(eng, dual=1)
            b = eng.lm_generate(feats[:rows], 20)
            _opts(eng, cuda_graph=0)
            c = eng.lm_generate(feats[:rows], 20)
            assert np.array_equal(a, b) and np.array_equal(a, c), rows
    finally:
        _opts(eng, dual=0, cuda_graph=1)


def test_padded_row_count_does_not_change_rows(model, oracle_detail):
    """The decoder runs on round_up(R, 32) rows (one step graph per padded size): a row's tokens do not depend on how
    many rows share the batch."""
    eng = model._engine()
    feats = torch.cat([oracle_detail["sel_feats"]] * 2, 0).contiguous().cuda()
    a = eng.lm_generate(feats[:33], 12)
    b = eng.lm_generate(feats[:64], 12)
    c = eng.lm_generate(feats[:5], 12)
    assert np.array_equal(a, b[:33]) and np.array_equal(c, b[:5])


def test_bbox_features_entry(model, synth_sd, images, golden):
    """n1: user boxes -> region features -> one sentence per box (always 29 rows per image)."""
    from rgrg_b200 import get_bbox_features

    g = golden("bbox_features.npz")
    boxes = [torch.from_numpy(g["boxes"][0]), torch.from_numpy(g["boxes"][1])]
    ref = O.bbox_features(synth_sd, images, boxes)
    feats = get_bbox_features(model, images.cuda(), [b.cuda() for b in boxes])
    assert feats.shape == (58, 1024)
    assert _rel(feats.cpu(), ref) < 0.05  # bf16 backbone
    ids = model.language_model.generate(feats, max_length=5)
    assert ids.shape == (58, 5)
