"""CPU: the oracle's restatements of third-party arithmetic against the library kernels the reference actually calls
(torchvision.ops.nms / roi_align), on seeded random inputs including the edge cases the domain has: degenerate and
out-of-image boxes, exact score ties, tiny RoIs, RoIs larger than the feature map."""
import numpy as np
import pytest
import torch
import torchvision

import rgrg_oracle as O


def _random_boxes(n, g, size=512.0, degenerate=0.1):
    c = torch.rand(n, 2, generator=g) * size
    wh = torch.rand(n, 2, generator=g) * 200 + 1
    wh[torch.rand(n, generator=g) < degenerate] = 0.0  # zero-area boxes
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    return b


@pytest.mark.parametrize("seed", range(6))
def test_nms_restatement_matches_torchvision(seed):
    g = torch.Generator().manual_seed(seed)
    n = [5, 64, 257, 1000, 1000, 33][seed]
    boxes = _random_boxes(n, g).clamp(0, 512)
    if seed % 2:  # heavy overlap: many boxes around a few centres
        centres = torch.rand(4, 2, generator=g) * 400 + 50
        pick = torch.randint(0, 4, (n,), generator=g)
        jitter = torch.randn(n, 4, generator=g) * 6
        boxes = (torch.cat([centres[pick] - 60, centres[pick] + 60], 1) + jitter).clamp(0, 512)
    scores = torch.rand(n, generator=g)
    scores, order = scores.sort(descending=True, stable=True)
    boxes = boxes[order]
    mine = O.nms_keep(boxes, 0.7)
    ref = torchvision.ops.nms(boxes, scores, 0.7)
    assert torch.equal(mine, ref)


def test_nms_with_tied_scores_keeps_input_order():
    g = torch.Generator().manual_seed(9)
    boxes = _random_boxes(200, g, degenerate=0.0).clamp(0, 512)
    scores = torch.full((200,), 0.5)  # all tied: a stable descending sort leaves the order unchanged
    ref = torchvision.ops.nms(boxes, scores, 0.7)
    assert torch.equal(O.nms_keep(boxes, 0.7), ref)


@pytest.mark.parametrize("seed", range(4))
def test_roi_align_restatement_matches_torchvision_on_edge_cases(seed):
    g = torch.Generator().manual_seed(100 + seed)
    feat = torch.randn(1, 8, 16, 16, generator=g)
    rois = _random_boxes(40, g, degenerate=0.15)
    rois[0] = torch.tensor([-40.0, -40.0, 30.0, 30.0])     # sticks out of the image (samples < -1 contribute zero)
    rois[1] = torch.tensor([500.0, 500.0, 700.0, 700.0])   # beyond the far edge
    rois[2] = torch.tensor([0.0, 0.0, 512.0, 512.0])       # the whole image: bins of two feature cells
    rois[3] = torch.tensor([100.0, 100.0, 100.5, 100.2])   # smaller than one cell: width / height clamp to 1
    ref = torchvision.ops.roi_align(feat, [rois], (8, 8), 1.0 / 32, 2)
    mine = O.roi_align(feat[0], rois, 1.0 / 32)
    assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-6)


def test_decode_boxes_clamps_exponent_like_boxcoder():
    anchors = torch.tensor([[0.0, 0.0, 32.0, 32.0]])
    deltas = torch.tensor([[0.0, 0.0, 50.0, 50.0]])  # exp(50) would overflow: BoxCoder clamps dw, dh at ln(1000/16)
    out = O.decode_boxes(deltas, anchors)
    w = out[0, 2] - out[0, 0]
    assert torch.isfinite(out).all() and abs(w.item() - 32 * 1000 / 16) < 1e-2


def test_top_regions_undetected_class_falls_back_to_index_zero():
    """custom_roi_heads.py:141-159: a class that is never the arg-max of any RoI is `not detected` and reports RoI 0."""
    logits = torch.full((5, 30), -5.0)
    logits[:, 3] = 5.0  # every RoI predicts class 3 (region index 2)
    reg = torch.zeros(5, 120)
    props = [torch.tensor([[10.0, 10.0, 50.0, 60.0]] * 5)]
    out = O.top_regions(logits, reg, props, 512)
    det = out["class_detected"][0]
    assert det.sum() == 1 and det[2]
    assert (out["top_idx"][0][~det] == 0).all() and (out["top_scores"][0][~det] == 0).all()
