"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build container.

    python oracle/make_golden.py            # ~4 min on 8 cores; writes tests/golden/

Each fixture stores the INPUTS handed to a reference function and the OUTPUTS the reference returned, so that
(a) oracle/rgrg_oracle.py can be pinned against them anywhere (`-m "not gpu"` tests) and
(b) the CUDA kernels can be checked against the same vectors on the GPU box, where /root/reference does not exist.
Weights come from rgrg_b200.synth (seed 0); fixtures that depend on BN-calibrated weights carry the tensors they
need, so they stay valid even if a different host rounds the calibration differently.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness  # noqa: E402
from rgrg_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def npz(name, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items()})


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    sd = synth.make_state_dict(0)
    model = ref_harness.build_reference_model(None)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    model.eval()
    det = model.object_detector
    from src.object_detector.image_list import ImageList
    from torchvision.models.detection.rpn import concat_box_prediction_layers

    imgs = synth.synthetic_images(2, 512, seed=1001)
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None

    if only == "abnormal":
        # ---- a9': abnormal classifier on the reference's own region features (binary_classifier_region_abnormal.py:31-57, eval)
        g = np.load(os.path.join(OUT, "selection.npz"))
        feats29, detected = torch.from_numpy(g["top_region_features"]), torch.from_numpy(g["class_detected"])
        loss, pred = model.binary_classifier_region_abnormal(feats29, detected, torch.zeros_like(detected))
        logits = model.binary_classifier_region_abnormal.classifier(feats29).squeeze(-1)
        npz("abnormal.npz", predicted_abnormal_regions=pred, logits=logits)
        return
    if only in ("lm_long", "lm_crafted"):
        lm_only(model, only)
        return

    # ---- selection-based entry: user boxes -> region features (evaluate_bbox_variations.py:92-110), run on the reference's modules
    g = torch.Generator().manual_seed(77)
    ctr = torch.rand(2, 29, 2, generator=g) * 300 + 100
    wh = torch.rand(2, 29, 2, generator=g) * 160 + 40
    user_boxes = [torch.cat([(ctr[b] - wh[b] / 2).clamp(0, 512), (ctr[b] + wh[b] / 2).clamp(0, 512)], dim=1) for b in range(2)]
    f_ = det.backbone(imgs)
    maps = det.roi_heads.box_roi_pool({"0": f_}, user_boxes, [(512, 512)] * 2)
    bbox_feats = det.roi_heads.dim_reduction(torch.squeeze(det.roi_heads.avg_pool(maps)))
    npz("bbox_features.npz", boxes=torch.stack(user_boxes), features=bbox_feats)
    if only == "bbox":
        return

    # ---- anchors (torchvision AnchorGenerator as configured at object_detector.py:78-83)
    feats = det.backbone(imgs)
    il = ImageList(imgs)
    anchors = det.rpn.anchor_generator(il, [feats])
    npz("anchors_512.npz", base=det.rpn.anchor_generator.cell_anchors[0], anchors=anchors[0])

    # ---- RPN: head outputs -> decode -> filter_proposals (custom_rpn.py:61-71)
    obj, deltas = det.rpn.head([feats])
    obj_f, deltas_f = concat_box_prediction_layers(obj, deltas)
    decoded = det.rpn.box_coder.decode(deltas_f.detach(), anchors).view(2, -1, 4)
    boxes, scores = det.rpn.filter_proposals(decoded, obj_f.detach(), il.image_sizes, [obj_f.shape[0] // 2])
    npz("rpn_filter.npz", objectness=obj_f.view(2, -1), deltas=deltas_f.view(2, -1, 4), decoded=decoded,
        count=[b.shape[0] for b in boxes], boxes0=boxes[0], boxes1=boxes[1], scores0=scores[0], scores1=scores[1])

    # ---- RoIAlign (torchvision binary kernel via MultiScaleRoIAlign) on a small slice: 64 channels, 48 proposals / image
    small = feats[:, :64].contiguous()
    small_props = [boxes[0][::19][:48].contiguous(), boxes[1][::17][:48].contiguous()]
    pooled_small = det.roi_heads.box_roi_pool({"0": small}, small_props, il.image_sizes)
    npz("roi_align.npz", feats=small, rois0=small_props[0], rois1=small_props[1], pooled=pooled_small)

    # ---- RoI tail: logits / deltas / proposals -> class_detected, top idx, boxes, scores (custom_roi_heads.py:63-208)
    pooled = det.roi_heads.box_roi_pool({"0": feats}, boxes, il.image_sizes)
    fc7 = det.roi_heads.box_head(pooled)
    cls, reg = det.roi_heads.box_predictor(fc7)
    # index recovery: feed "box_features" = global row index so top_region_features returns the chosen rows
    row_idx = torch.arange(cls.shape[0], dtype=torch.float32)[:, None].repeat(1, 2)
    out = det.roi_heads.get_top_region_features_detections_class_detected(row_idx, reg, cls, boxes, il.image_sizes)
    top_idx_global = out["top_region_features"][:, :, 0].long()
    offs = torch.tensor([0, boxes[0].shape[0]])[:, None]
    npz("roi_tail.npz", class_logits=cls, box_regression=reg, proposals0=boxes[0], proposals1=boxes[1],
        class_detected=out["class_detected"], top_idx=top_idx_global - offs,
        top_region_boxes=out["detections"]["top_region_boxes"], top_scores=out["detections"]["top_scores"])

    # ---- region features + selection (custom_roi_heads.py:253-264; binary_classifier_region_selection.py:24-68)
    rh = det.roi_heads(({"0": feats}), boxes, il.image_sizes)
    sel, sel_feats = model.binary_classifier_region_selection(rh["top_region_features"], rh["class_detected"], return_loss=False)
    logits = model.binary_classifier_region_selection.classifier(rh["top_region_features"]).squeeze(-1)
    npz("selection.npz", top_region_features=rh["top_region_features"], class_detected=rh["class_detected"],
        logits=logits, selected=sel, selected_features=sel_feats)

    # ---- decoder: greedy on 5 rows, 7 steps, teacher-forcing record (language_model.py:609-652)
    lm = model.language_model
    rows = sel_feats[:5].contiguous()
    ids = lm.generate(rows, max_length=8)
    # per-step logits summaries by re-running the cached loop through the reference's own forward
    input_ids = ids[:, :1]
    mask = torch.ones(5, 1, dtype=torch.int64)
    past = None
    top_val, top_idx, lse = [], [], []
    for t in range(ids.shape[1] - 1):
        mi = lm.prepare_inputs_for_generation(input_ids, past=past, attention_mask=mask, use_cache=True)
        logits_t, past = lm.forward(**mi, image_hidden_states=rows, return_loss=False)
        l = logits_t[:, -1, :]
        v, i = l.topk(8, dim=-1)
        top_val.append(v); top_idx.append(i); lse.append(torch.logsumexp(l, -1))
        input_ids = ids[:, : t + 2]
        mask = torch.ones(5, t + 2, dtype=torch.int64)
    npz("lm_greedy.npz", feats=rows, ids=ids, top_val=torch.stack(top_val), top_idx=torch.stack(top_idx),
        logsumexp=torch.stack(lse))

    # ---- decoder: beam search (4 beams) on 3 rows (language_model.py:529-607 + oracle/beam_scorer.py)
    for es in (True, False):
        ids_b = lm.generate(rows[:3], max_length=7, num_beams=4, early_stopping=es)
        npz("lm_beam_es%d.npz" % int(es), feats=rows[:3], ids=ids_b)

    # ---- whole path (report_generation_model.py:212-276): depends on calibrated weights of THIS host
    ids_full, selected, detections, class_detected = model.generate(imgs, max_length=6)
    npz("generate_b2.npz", ids=ids_full, selected=selected, class_detected=class_detected,
        top_region_boxes=detections["top_region_boxes"], top_scores=detections["top_scores"],
        backbone_checksum=[float(feats.double().sum()), float(feats.double().abs().sum())])


from crafted import BEAM_CASES, GREEDY_CASES, beam_crafted_logits, eos_schedule  # noqa: E402


@torch.no_grad()
def lm_only(model, only):
    """Decoder-only fixtures.  Inputs: the selected region features of selection.npz (reference output, committed)."""
    from rgrg_b200 import synth

    lm = model.language_model
    feats_all = torch.from_numpy(np.load(os.path.join(OUT, "selection.npz"))["selected_features"])
    if only == "lm_long":
        # ---- teacher-forced record at real cache lengths: 6 rows x 128 generated tokens (cache length up to 129)
        rows = feats_all[:6].contiguous()
        ids = lm.generate(rows, max_length=129)
        input_ids = ids[:, :1]
        mask = torch.ones(rows.shape[0], 1, dtype=torch.int64)
        past = None
        top_val, top_idx, lse = [], [], []
        for t in range(ids.shape[1] - 1):
            mi = lm.prepare_inputs_for_generation(input_ids, past=past, attention_mask=mask, use_cache=True)
            logits_t, past = lm.forward(**mi, image_hidden_states=rows, return_loss=False)
            l = logits_t[:, -1, :]
            v, i = l.topk(8, dim=-1)
            top_val.append(v); top_idx.append(i.to(torch.int32)); lse.append(torch.logsumexp(l, -1))
            assert torch.equal(l.argmax(-1), ids[:, t + 1])
            input_ids = ids[:, : t + 2]
            mask = torch.ones(rows.shape[0], t + 2, dtype=torch.int64)
        npz("lm_long.npz", feats=rows, ids=ids.to(torch.int32), top_val=torch.stack(top_val), top_idx=torch.stack(top_idx),
            logsumexp=torch.stack(lse))
        return
    # ---- search loops of the reference driven by GIVEN logits (the model forward is replaced, nothing else):
    # greedy_search (language_model.py:609-652) and beam_search (:529-607, with oracle/beam_scorer.py)
    out = {}
    real_forward = lm.forward
    for name, (seed, rows, max_length, kind) in GREEDY_CASES.items():
        steps = max_length - 1
        mask = eos_schedule(kind, steps, rows)
        logits = synth.crafted_logits(seed, steps, rows, mask)
        state = {"t": 0}

        def fake_forward(**kw):
            t = state["t"]
            state["t"] += 1
            return logits[t][:, None, :], ("given",)

        lm.forward = fake_forward
        ids = lm.generate(torch.zeros(rows, 1024), max_length=max_length)
        out["greedy_%s_ids" % name] = ids.to(torch.int32)
        out["greedy_%s_mask" % name] = mask
        out["greedy_%s_meta" % name] = np.array([seed, rows, max_length], dtype=np.int32)
        print(name, tuple(ids.shape), "forward calls", state["t"])
    for name, (seed, sentences, nb, max_length, es, kind) in BEAM_CASES.items():
        logits = beam_crafted_logits(seed, sentences, nb, max_length, kind)
        state = {"t": 0}

        def fake_forward(**kw):
            t = state["t"]
            state["t"] += 1
            return logits[t][:, None, :], tuple((torch.zeros(sentences * nb, 1, 1, 1), torch.zeros(sentences * nb, 1, 1, 1)) for _ in range(1))

        lm.forward = fake_forward
        ids = lm.generate(torch.zeros(sentences, 1024), max_length=max_length, num_beams=nb, early_stopping=es)
        out["beam_%s_ids" % name] = ids.to(torch.int32)
        out["beam_%s_meta" % name] = np.array([seed, sentences, nb, max_length, int(es)], dtype=np.int32)
        print(name, tuple(ids.shape), "forward calls", state["t"])
    lm.forward = real_forward
    npz("lm_crafted.npz", **out)


if __name__ == "__main__":
    main()
