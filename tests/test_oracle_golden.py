"""Pins oracle/rgrg_oracle.py (the CPU restatement) against vectors produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import numpy as np
import torch

import rgrg_oracle as O

T = torch.from_numpy


def test_anchors_match_torchvision_anchor_generator(golden):
    g = golden("anchors_512.npz")
    assert torch.equal(O.base_anchors(), T(g["base"]))
    assert torch.equal(O.anchors_for(512, 16), T(g["anchors"]))


def test_rpn_decode_and_filter_match_reference(golden):
    g = golden("rpn_filter.npz")
    obj, deltas = T(g["objectness"]), T(g["deltas"])
    anchors = O.anchors_for(512, 16)
    decoded = O.decode_boxes(deltas.reshape(-1, 4), anchors.repeat(2, 1)).view(2, -1, 4)
    assert torch.equal(decoded, T(g["decoded"]))
    detail = []
    props = O.filter_proposals(obj, decoded, 512, detail=detail)
    assert [p.shape[0] for p in props] == list(g["count"])
    for b in range(2):
        assert torch.equal(props[b], T(g["boxes%d" % b]))
        assert torch.equal(detail[b]["scores"], T(g["scores%d" % b]))


def test_roi_align_restatement_matches_torchvision_kernel(golden):
    g = golden("roi_align.npz")
    feats = T(g["feats"])
    rois = [T(g["rois0"]), T(g["rois1"])]
    mine = O.box_roi_pool(feats, rois, 512, use_torchvision=False)
    assert torch.allclose(mine, T(g["pooled"]), rtol=1e-5, atol=1e-6)
    assert torch.equal(O.box_roi_pool(feats, rois, 512, use_torchvision=True), T(g["pooled"]))


def test_roi_tail_matches_reference(golden):
    g = golden("roi_tail.npz")
    props = [T(g["proposals0"]), T(g["proposals1"])]
    out = O.top_regions(T(g["class_logits"]), T(g["box_regression"]), props, 512)
    assert torch.equal(out["class_detected"], T(g["class_detected"]))
    assert torch.equal(out["top_idx"], T(g["top_idx"]))
    assert torch.equal(out["top_scores"], T(g["top_scores"]))
    assert torch.equal(out["top_region_boxes"], T(g["top_region_boxes"]))


def test_selection_matches_reference(golden, lm_sd):
    g = golden("selection.npz")
    sel, feats, logits = O.region_selection(lm_sd, T(g["top_region_features"]), T(g["class_detected"]))
    assert torch.equal(sel, T(g["selected"]))
    assert torch.equal(feats, T(g["selected_features"]))
    assert torch.allclose(logits, T(g["logits"]), rtol=0, atol=1e-6)


def test_abnormal_classifier_matches_reference(golden, lm_sd):
    g, a = golden("selection.npz"), golden("abnormal.npz")
    pred, logits = O.region_abnormal(lm_sd, T(g["top_region_features"]), T(g["class_detected"]))
    assert torch.equal(pred, T(a["predicted_abnormal_regions"]))
    assert torch.allclose(logits, T(a["logits"]), rtol=0, atol=1e-6)


def test_lm_greedy_matches_reference(golden, lm_sd):
    g = golden("lm_greedy.npz")
    rec = {}
    ids = O.lm_generate(lm_sd, T(g["feats"]), max_length=8, record=rec)
    assert torch.equal(ids, T(g["ids"]))
    for t, logits in enumerate(rec["logits"]):
        v, i = logits.topk(8, dim=-1)
        assert torch.equal(i, T(g["top_idx"][t]))
        assert torch.allclose(v, T(g["top_val"][t]), rtol=0, atol=2e-5)
        assert torch.allclose(torch.logsumexp(logits, -1), T(g["logsumexp"][t]), rtol=0, atol=2e-5)


def test_lm_long_teacher_forced_matches_reference(golden, lm_sd):
    """Decoder at real cache lengths (up to 129 keys): the reference's own 128-step greedy decode, re-run teacher-forced
    through the oracle's cached forward; logits agree to fp32 noise at every step."""
    g = golden("lm_long.npz")
    feats, ids = T(g["feats"]), T(g["ids"]).long()
    rows = feats.shape[0]
    past = None
    with torch.no_grad():
        for t in range(ids.shape[1] - 1):
            pos = torch.full((rows, 1), t, dtype=torch.int64)
            mask = torch.ones(rows, t + 1, dtype=torch.int64)
            logits, past = O.lm_forward(lm_sd, ids[:, t:t + 1], feats, past, pos, mask)
            l = logits[:, -1, :]
            ref_idx = T(g["top_idx"][t]).long()
            assert torch.allclose(l.gather(1, ref_idx), T(g["top_val"][t]), rtol=0, atol=5e-5), t
            assert torch.allclose(torch.logsumexp(l, -1), T(g["logsumexp"][t]), rtol=0, atol=5e-5), t
            assert torch.equal(l.argmax(-1), ids[:, t + 1]), t


def test_search_loops_on_crafted_logits_match_reference(golden):
    """greedy_search / beam_search bookkeeping (pad-if-finished, stop rule, width; beam ties, EOS beyond rank num_beams,
    simultaneous finishes) against ids the UNMODIFIED reference loops produced from the same crafted logits."""
    import crafted
    from rgrg_b200 import synth

    g = golden("lm_crafted.npz")
    for name, (seed, rows, max_length, kind) in crafted.GREEDY_CASES.items():
        mask = crafted.eos_schedule(kind, max_length - 1, rows)
        assert np.array_equal(mask.numpy(), g["greedy_%s_mask" % name])
        logits = synth.crafted_logits(seed, max_length - 1, rows, mask)
        ids = O.greedy_search({}, torch.zeros(rows, 1024), max_length, given_logits=logits)
        assert torch.equal(ids, T(g["greedy_%s_ids" % name]).long()), name
    for name, (seed, sentences, nb, max_length, es, kind) in crafted.BEAM_CASES.items():
        logits = crafted.beam_crafted_logits(seed, sentences, nb, max_length, kind)
        ids = O.beam_search({}, torch.zeros(sentences, 1024), max_length, nb, es, given_logits=logits)
        assert torch.equal(ids, T(g["beam_%s_ids" % name]).long()), name


def test_lm_beam_matches_reference(golden, lm_sd):
    for es in (True, False):
        g = golden("lm_beam_es%d.npz" % int(es))
        ids = O.lm_generate(lm_sd, T(g["feats"]), max_length=7, num_beams=4, early_stopping=es)
        assert torch.equal(ids, T(g["ids"]))


def test_lm_generate_error_behaviour(lm_sd):
    import pytest

    f = torch.zeros(1, 1024)
    with pytest.raises(NotImplementedError):
        O.lm_generate(lm_sd, f, max_length=4, do_sample=True)
    with pytest.raises(NotImplementedError):
        O.lm_generate(lm_sd, f, max_length=4, num_beams=4, num_beam_groups=2)
    with pytest.raises(ValueError):
        O.lm_generate(lm_sd, f, max_length=None, num_beams=4)
    with pytest.raises(ValueError):
        O.lm_generate(lm_sd, f, max_length=4, num_beams=2, num_beam_groups=4)


def test_cached_greedy_equals_uncached_teacher_forced_argmax(lm_sd):
    """SURVEY.md §4 identity: stepwise cached decoding == arg-max of one full-sequence forward."""
    feats = torch.randn(2, 1024, generator=torch.Generator().manual_seed(5))
    ids = O.lm_generate(lm_sd, feats, max_length=5)
    L = ids.shape[1] - 1
    pos = torch.arange(L)[None, :].expand(2, L)
    logits, _ = O.lm_forward(lm_sd, ids[:, :L], feats, None, pos, torch.ones(2, L, dtype=torch.int64))
    assert torch.equal(logits.argmax(-1), ids[:, 1:])


def test_whole_path_matches_reference(golden, synth_sd):
    g = golden("generate_b2.npz")
    from rgrg_b200 import synth

    imgs = synth.synthetic_images(2, 512, seed=1001)
    feats = O.backbone(synth_sd, imgs)
    chk = np.array([float(feats.double().sum()), float(feats.double().abs().sum())])
    same_host_weights = np.allclose(chk, g["backbone_checksum"], rtol=1e-12)
    out = O.generate(synth_sd, imgs, max_length=6)
    ids, selected, det, cd = out
    assert torch.equal(cd, T(g["class_detected"]))
    assert torch.equal(selected, T(g["selected"]))
    if same_host_weights:  # BN calibration rounds identically: everything is bit-exact
        assert torch.equal(ids, T(g["ids"]))
        assert torch.equal(det["top_region_boxes"], T(g["top_region_boxes"]))
        assert torch.equal(det["top_scores"], T(g["top_scores"]))
    else:
        assert ids.shape == g["ids"].shape


def test_bbox_features_match_reference(golden, synth_sd):
    """Selection-based entry (evaluate_bbox_variations.py:92-110).  Depends on the BN-calibrated backbone: exact on the
    host that generated the fixture, tolerance elsewhere."""
    from rgrg_b200 import synth

    g = golden("bbox_features.npz")
    imgs = synth.synthetic_images(2, 512, seed=1001)
    feats = O.bbox_features(synth_sd, imgs, [T(g["boxes"][0]), T(g["boxes"][1])])
    assert feats.shape == (58, 1024)
    assert torch.allclose(feats, T(g["features"]), rtol=1e-3, atol=1e-3)
