"""CPU oracle: a functional fp32 restatement of the reference's `ReportGenerationModel.generate()` path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this file, and there only as the checker / the CPU arm — never as the thing shipped.  The product
(rgrg_b200/) must not import it.

Every function cites the reference file:line it restates (paths relative to ttanida/rgrg @ 9520b6d).  Third-party
arithmetic the reference reaches (torchvision 0.13.1 detection utilities, HF GPT-2 blocks) is restated from its
published algorithm; torch CPU ops (conv2d, linear, softmax ...) are used as plain fp32 arithmetic.

PINNING.  The reference has no tests, golden vectors or fixtures, so the oracle is pinned against
outputs of the reference itself run in the build container: oracle/make_golden.py imports the unmodified reference
through oracle/ref_harness.py and commits its stage inputs / outputs under tests/golden/; tests/test_oracle_golden.py
checks this file against every one of them.  Beam search rests on oracle/beam_scorer.py, whose upstream
(transformers==4.19.2 BeamSearchScorer) is not vendored: beam parity is "unpinned" beyond the reference's own call
sites.

Deliberately mirrors the reference's *algorithm*, including its inefficiencies that define the CPU baseline:
the feature-space MLP is recomputed every decode step (language_model.py:284), the KV cache is regrown with
torch.cat every step (language_model.py:169-170), and proposals are filtered in a per-image Python loop.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BOS = EOS = PAD = 50256  # language_model.py:200-202
NUM_REGIONS = 29
ANCHOR_SIZES = (20, 40, 60, 80, 100, 120, 140, 160, 180, 300)  # object_detector.py:79
ANCHOR_RATIOS = (0.2, 0.25, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.3, 1.5, 2.1, 2.6, 3.0, 5.0, 8.0)  # :80
BBOX_XFORM_CLIP = math.log(1000.0 / 16)  # torchvision/models/detection/_utils.py BoxCoder.__init__
RESNET_LAYERS = (3, 4, 6, 3)

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------------------------------
# a3: ResNet-50 trunk  (object_detector.py:51-62; torchvision/models/resnet.py Bottleneck.forward, v1.5 stride on 3x3)
# ----------------------------------------------------------------------------------------------------------------------
def _bn(sd: SD, x, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def backbone(sd: SD, images: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """images [B,1,S,S] fp32 -> [B,2048,S/32,S/32].  object_detector.py:219."""
    bb = "object_detector.backbone"
    x = F.conv2d(images, sd[bb + ".0.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(sd, x, bb + ".1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["stem"] = x
    for li, nblocks in enumerate(RESNET_LAYERS):
        for bi in range(nblocks):
            p = "%s.%d.%d" % (bb, 4 + li, bi)
            stride = 2 if (bi == 0 and li > 0) else 1
            identity = x
            out = F.relu(_bn(sd, F.conv2d(x, sd[p + ".conv1.weight"]), p + ".bn1"))
            out = F.relu(_bn(sd, F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), p + ".bn2"))
            out = _bn(sd, F.conv2d(out, sd[p + ".conv3.weight"]), p + ".bn3")
            if bi == 0:
                identity = _bn(sd, F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), p + ".downsample.1")
            x = F.relu(out + identity)
        if taps is not None:
            taps["layer%d" % (li + 1)] = x
    return x


# ----------------------------------------------------------------------------------------------------------------------
# a4: RPN head, anchors, box decoding  (custom_rpn.py:53-71; torchvision rpn.py RPNHead / anchor_utils.py / _utils.py)
# ----------------------------------------------------------------------------------------------------------------------
def rpn_head(sd: SD, feats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> objectness [B, H*W*160] and deltas [B, H*W*160, 4], both in (h, w, a) order
    (custom_rpn.py:61,66 -> rpn.py concat_box_prediction_layers / permute_and_flatten)."""
    p = "object_detector.rpn.head"
    t = F.relu(F.conv2d(feats, sd[p + ".conv.0.0.weight"], sd[p + ".conv.0.0.bias"], padding=1))
    logits = F.conv2d(t, sd[p + ".cls_logits.weight"], sd[p + ".cls_logits.bias"])  # [B, A, H, W]
    deltas = F.conv2d(t, sd[p + ".bbox_pred.weight"], sd[p + ".bbox_pred.bias"])  # [B, A*4, H, W]
    B, A, H, W = logits.shape
    objectness = logits.permute(0, 2, 3, 1).reshape(B, H * W * A)
    deltas = deltas.view(B, A, 4, H, W).permute(0, 3, 4, 1, 2).reshape(B, H * W * A, 4)
    return objectness, deltas


def base_anchors() -> torch.Tensor:
    """anchor_utils.py AnchorGenerator.generate_anchors: ratio-major, size-minor, round(half-even)([-w,-h,w,h]/2)."""
    scales = torch.as_tensor(ANCHOR_SIZES, dtype=torch.float32)
    ratios = torch.as_tensor(ANCHOR_RATIOS, dtype=torch.float32)
    h_ratios = torch.sqrt(ratios)
    w_ratios = 1 / h_ratios
    ws = (w_ratios[:, None] * scales[None, :]).view(-1)
    hs = (h_ratios[:, None] * scales[None, :]).view(-1)
    return (torch.stack([-ws, -hs, ws, hs], dim=1) / 2).round()


def anchors_for(image_size: int, feat_size: int) -> torch.Tensor:
    """anchor_utils.py grid_anchors: stride = image // feat; shifts y-major; [(h, w, a), 4]."""
    stride = image_size // feat_size
    shifts = torch.arange(0, feat_size, dtype=torch.int32) * stride
    sy, sx = torch.meshgrid(shifts, shifts, indexing="ij")
    sx, sy = sx.reshape(-1), sy.reshape(-1)
    shift = torch.stack((sx, sy, sx, sy), dim=1)
    return (shift.view(-1, 1, 4) + base_anchors().view(1, -1, 4)).reshape(-1, 4)


def decode_boxes(deltas: torch.Tensor, boxes: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0)) -> torch.Tensor:
    """_utils.py BoxCoder.decode_single.  deltas [N, 4*k], boxes [N,4] -> [N, 4*k]."""
    boxes = boxes.to(deltas.dtype)
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=BBOX_XFORM_CLIP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=BBOX_XFORM_CLIP)
    pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
    pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
    pred_w = torch.exp(dw) * widths[:, None]
    pred_h = torch.exp(dh) * heights[:, None]
    c_to_c_h = torch.tensor(0.5, dtype=pred_ctr_y.dtype) * pred_h
    c_to_c_w = torch.tensor(0.5, dtype=pred_ctr_x.dtype) * pred_w
    x1 = pred_ctr_x - c_to_c_w
    y1 = pred_ctr_y - c_to_c_h
    x2 = pred_ctr_x + c_to_c_w
    y2 = pred_ctr_y + c_to_c_h
    return torch.stack((x1, y1, x2, y2), dim=2).flatten(1)


def clip_boxes(boxes: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """torchvision/ops/boxes.py clip_boxes_to_image."""
    h, w = size
    bx = boxes[..., 0::2].clamp(min=0, max=w)
    by = boxes[..., 1::2].clamp(min=0, max=h)
    return torch.stack((bx, by), dim=boxes.dim()).reshape(boxes.shape)


# ----------------------------------------------------------------------------------------------------------------------
# a5: filter_proposals  (torchvision rpn.py:242-297 via custom_rpn.py:71; cfg object_detector.py:93-96)
# ----------------------------------------------------------------------------------------------------------------------
def nms_keep(boxes: torch.Tensor, thresh: float) -> torch.Tensor:
    """Greedy NMS over boxes ALREADY sorted by descending score (torchvision csrc/ops/cpu/nms_kernel.cpp: stable
    descending sort, then `ovr > thresh` suppresses; areas (x2-x1)*(y2-y1)).  Plain fp32 loops in numpy."""
    import numpy as np

    b = boxes.detach().cpu().numpy().astype(np.float32)
    n = b.shape[0]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(thresh)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        if i + 1 >= n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):  # 0/0 for two zero-area boxes is NaN and never suppresses, as in the C++ kernel
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= ovr > thr
    return torch.as_tensor(keep, dtype=torch.int64)


def filter_proposals(objectness: torch.Tensor, proposals: torch.Tensor, image_size: int, pre_nms_top_n=1000,
                     post_nms_top_n=1000, nms_thresh=0.7, score_thresh=0.0, min_size=1e-3,
                     detail: Optional[list] = None) -> List[torch.Tensor]:
    """objectness [B, N], proposals [B, N, 4] -> list of [P_i, 4] (score-descending).  Single feature level, so
    batched_nms degenerates to plain nms (boxes.py:87-104 coordinate trick with idxs == 0)."""
    B, N = objectness.shape
    k = min(pre_nms_top_n, N)
    _, top_idx = objectness.topk(k, dim=1)  # sorted descending
    out = []
    for b in range(B):
        idx = top_idx[b]
        boxes = proposals[b, idx]
        scores = torch.sigmoid(objectness[b, idx])
        boxes = clip_boxes(boxes, (image_size, image_size))
        ws, hs = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        keep = torch.where((ws >= min_size) & (hs >= min_size))[0]
        boxes, scores, idx_k = boxes[keep], scores[keep], idx[keep]
        keep = torch.where(scores >= score_thresh)[0]
        boxes, scores, idx_k = boxes[keep], scores[keep], idx_k[keep]
        # scores are sigmoid(sorted logits): a stable descending sort leaves the order unchanged
        keep = nms_keep(boxes, nms_thresh)[:post_nms_top_n]
        out.append(boxes[keep])
        if detail is not None:
            detail.append({"topk_idx": idx, "pre_nms_boxes": boxes, "pre_nms_anchor_idx": idx_k, "keep": keep,
                           "scores": scores[keep]})
    return out


def rpn(sd: SD, feats: torch.Tensor, image_size: int, detail: Optional[dict] = None) -> List[torch.Tensor]:
    """custom_rpn.py:53-85 (inference branch: targets None)."""
    objectness, deltas = rpn_head(sd, feats)
    B, N = objectness.shape
    anchors = anchors_for(image_size, feats.shape[-1])
    proposals = decode_boxes(deltas.reshape(B * N, 4), anchors.repeat(B, 1)).view(B, N, 4)
    per_image = [] if detail is not None else None
    props = filter_proposals(objectness, proposals, image_size, detail=per_image)
    if detail is not None:
        detail.update(objectness=objectness, deltas=deltas, decoded=proposals, per_image=per_image)
    return props


# ----------------------------------------------------------------------------------------------------------------------
# a6: RoIAlign + box head  (custom_roi_heads.py:232-236; torchvision/ops/roi_align.py:110-190; poolers.py:98-107)
# ----------------------------------------------------------------------------------------------------------------------
def roi_align(feat: torch.Tensor, rois: torch.Tensor, spatial_scale: float, out_size: int = 8, sampling_ratio: int = 2
              ) -> torch.Tensor:
    """feat [C,H,W] of ONE image, rois [K,4] (x1,y1,x2,y2) -> [K,C,out,out].  aligned=False: no -0.5 offset, roi
    width/height clamped to >= 1, samples at start + (p + (i+.5)/sr)*bin, bilinear with the torchvision edge rules
    (sample < -1 or > size -> 0; clamp to [0, size-1]); mean over sr*sr samples."""
    C, H, W = feat.shape
    K = rois.shape[0]
    x1, y1, x2, y2 = [rois[:, i] * spatial_scale for i in range(4)]
    roi_w = torch.clamp(x2 - x1, min=1.0)
    roi_h = torch.clamp(y2 - y1, min=1.0)
    bin_h, bin_w = roi_h / out_size, roi_w / out_size
    p = torch.arange(out_size, dtype=feat.dtype)
    s = torch.arange(sampling_ratio, dtype=feat.dtype)
    # sample coordinates [K, out, sr]
    ys = y1[:, None, None] + p[None, :, None] * bin_h[:, None, None] + (s[None, None, :] + 0.5) * bin_h[:, None, None] / sampling_ratio
    xs = x1[:, None, None] + p[None, :, None] * bin_w[:, None, None] + (s[None, None, :] + 0.5) * bin_w[:, None, None] / sampling_ratio

    def prep(c, size):
        valid = (c >= -1.0) & (c <= size)
        c = c.clamp(min=0)
        low = c.floor().long()
        high = low + 1
        at_edge = low >= size - 1
        low = torch.where(at_edge, torch.full_like(low, size - 1), low)
        high = torch.where(at_edge, torch.full_like(high, size - 1), high)
        c = torch.where(at_edge, low.to(c.dtype), c)
        l = c - low.to(c.dtype)
        return valid, low, high, l, 1.0 - l

    vy, yl, yh, ly, hy = prep(ys, H)
    vx, xl, xh, lx, hx = prep(xs, W)
    out = torch.zeros(K, C, out_size, out_size, dtype=feat.dtype)
    f = feat.reshape(C, H * W)
    for iy in range(sampling_ratio):
        for ix in range(sampling_ratio):
            valid = (vy[:, :, iy, None] & vx[:, None, :, ix]).to(feat.dtype)  # [K, ph, pw]
            for (yy, wy) in ((yl, hy), (yh, ly)):
                for (xx, wx) in ((xl, hx), (xh, lx)):
                    idx = yy[:, :, iy, None] * W + xx[:, None, :, ix]  # [K, ph, pw]
                    w = wy[:, :, iy, None] * wx[:, None, :, ix] * valid
                    v = f[:, idx.reshape(-1)].view(C, K, out_size, out_size).permute(1, 0, 2, 3)
                    out += v * w[:, None]
    return out / (sampling_ratio * sampling_ratio)


def spatial_scale_for(image_size: int, feat_size: int) -> float:
    """poolers.py _infer_scale: 2 ** round(log2(feat / image))."""
    return 2.0 ** float(round(math.log2(feat_size / image_size)))


def box_roi_pool(feats: torch.Tensor, proposals: List[torch.Tensor], image_size: int, use_torchvision: bool = True
                 ) -> torch.Tensor:
    """MultiScaleRoIAlign(["0"], 8, sampling_ratio=2) on a single map (object_detector.py:106; poolers.py:174-181).
    torchvision.ops.roi_align is the binary kernel the reference itself reaches; `use_torchvision=False` runs the
    restatement above (tests check the two agree)."""
    scale = spatial_scale_for(image_size, feats.shape[-1])
    if use_torchvision:
        import torchvision

        return torchvision.ops.roi_align(feats, list(proposals), output_size=(8, 8), spatial_scale=scale, sampling_ratio=2)
    return torch.cat([roi_align(feats[b], proposals[b], scale) for b in range(feats.shape[0])], dim=0)


def box_head_and_predictor(sd: SD, pooled: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """TwoMLPHead + FastRCNNPredictor (faster_rcnn.py; custom_roi_heads.py:235-236).  -> fc7, cls logits, box deltas."""
    rh = "object_detector.roi_heads"
    x = pooled.flatten(start_dim=1)  # (c, ph, pw)
    x = F.relu(F.linear(x, sd[rh + ".box_head.fc6.weight"], sd[rh + ".box_head.fc6.bias"]))
    x = F.relu(F.linear(x, sd[rh + ".box_head.fc7.weight"], sd[rh + ".box_head.fc7.bias"]))
    cls = F.linear(x, sd[rh + ".box_predictor.cls_score.weight"], sd[rh + ".box_predictor.cls_score.bias"])
    reg = F.linear(x, sd[rh + ".box_predictor.bbox_pred.weight"], sd[rh + ".box_predictor.bbox_pred.bias"])
    return x, cls, reg


# ----------------------------------------------------------------------------------------------------------------------
# a7: per-class top-1 region selection  (custom_roi_heads.py:63-208)
# ----------------------------------------------------------------------------------------------------------------------
def top_regions(class_logits: torch.Tensor, box_regression: torch.Tensor, proposals: List[torch.Tensor],
                image_size: int) -> dict:
    """-> class_detected bool[B,29], top_idx int64[B,29] (index into the image's proposals), top_scores [B,29],
    top_region_boxes [B,29,4]."""
    scores = F.softmax(class_logits, -1)[:, 1:]
    counts = [p.shape[0] for p in proposals]
    pred_boxes = decode_boxes(box_regression, torch.cat(proposals, 0), weights=(10.0, 10.0, 5.0, 5.0)).view(-1, 30, 4)
    det, idxs, top_s, top_b = [], [], [], []
    for sc, bx in zip(torch.split(scores, counts), torch.split(pred_boxes, counts)):
        pred_classes = torch.argmax(sc, dim=1)
        mask = F.one_hot(pred_classes, num_classes=NUM_REGIONS)
        top_scores, top_idx = torch.max(sc * mask, dim=0)
        det.append(mask.sum(0) > 0)
        idxs.append(top_idx)
        top_s.append(top_scores)
        bx = clip_boxes(bx, (image_size, image_size))[:, 1:]
        top_b.append(bx[top_idx, torch.arange(NUM_REGIONS)])
    return {"class_detected": torch.stack(det), "top_idx": torch.stack(idxs), "top_scores": torch.stack(top_s),
            "top_region_boxes": torch.stack(top_b)}


def roi_heads(sd: SD, feats: torch.Tensor, proposals: List[torch.Tensor], image_size: int,
              detail: Optional[dict] = None) -> dict:
    """custom_roi_heads.py:210-269 (return_feature_vectors=True, eval)."""
    pooled = box_roi_pool(feats, proposals, image_size)
    fc7, cls, reg = box_head_and_predictor(sd, pooled)
    box_features = F.avg_pool2d(pooled, 8).reshape(pooled.shape[0], -1)  # :253-256 (squeeze; assumes > 1 RoI)
    out = top_regions(cls, reg, proposals, image_size)
    counts = [p.shape[0] for p in proposals]
    feats29 = torch.stack([bf[i] for bf, i in zip(torch.split(box_features, counts), out["top_idx"])])  # [B,29,2048]
    rh = "object_detector.roi_heads.dim_reduction"
    out["top_region_features"] = F.linear(feats29, sd[rh + ".weight"], sd[rh + ".bias"])  # :264
    if detail is not None:
        detail.update(pooled=pooled, fc7=fc7, class_logits=cls, box_regression=reg, top_features_2048=feats29)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# a9: region-selection classifier  (binary_classifier_region_selection.py:24-68)
# ----------------------------------------------------------------------------------------------------------------------
def _mlp3(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    x = F.relu(F.linear(x, sd[prefix + ".classifier.0.weight"], sd[prefix + ".classifier.0.bias"]))
    x = F.relu(F.linear(x, sd[prefix + ".classifier.2.weight"], sd[prefix + ".classifier.2.bias"]))
    return F.linear(x, sd[prefix + ".classifier.4.weight"], sd[prefix + ".classifier.4.bias"]).squeeze(-1)


def region_selection(sd: SD, top_region_features: torch.Tensor, class_detected: torch.Tensor):
    logits = _mlp3(sd, "binary_classifier_region_selection", top_region_features)
    selected = logits > -1  # :53
    selected = selected & class_detected  # :57
    return selected, top_region_features[selected], logits  # :61


def region_abnormal(sd: SD, top_region_features: torch.Tensor, class_detected: torch.Tensor):
    """binary_classifier_region_abnormal.py:31-57 — NOT on the generate() path.  Eval branch: returns
    `logits > -1` for ALL 29 regions (:53-57: undetected regions are filtered later by the caller via class_detected)."""
    logits = _mlp3(sd, "binary_classifier_region_abnormal", top_region_features)
    return logits > -1, logits


# ----------------------------------------------------------------------------------------------------------------------
# a11-a13: GPT-2-medium with pseudo self-attention  (language_model.py:32-180, 258-399)
# ----------------------------------------------------------------------------------------------------------------------
def gelu_new(x):  # transformers activations.py NewGELUActivation
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def _conv1d(sd: SD, p: str, x):  # language_model.py:11-29  x @ W[K,N] + b
    return torch.addmm(sd[p + ".bias"], x.reshape(-1, x.shape[-1]), sd[p + ".weight"]).view(*x.shape[:-1], -1)


def _split_heads(t):  # :76-82
    return t.view(t.shape[0], t.shape[1], 16, 64).permute(0, 2, 1, 3)


def lm_forward(sd: SD, input_ids: torch.Tensor, image_hidden_states: torch.Tensor, past, position_ids,
               attention_mask: Optional[torch.Tensor] = None, taps: Optional[dict] = None):
    """language_model.py:258-399 with use_cache=True, return_loss=False.
    input_ids [rows, q]; past = None or list of (K, V) [rows,16,L,64]; returns (logits [rows,q,V], presents)."""
    lm = "language_model"
    rows, q = input_ids.shape
    img = F.linear(image_hidden_states, sd[lm + ".feature_space_transformation_nn.0.weight"],
                   sd[lm + ".feature_space_transformation_nn.0.bias"])
    img = F.linear(F.relu(img), sd[lm + ".feature_space_transformation_nn.2.weight"],
                   sd[lm + ".feature_space_transformation_nn.2.bias"])  # :284 (every step)
    wte = sd[lm + ".wte.weight"]
    h = wte[input_ids] + wte[position_ids]  # :291,:307 — positions go through wte, not wpe
    if attention_mask is None:
        attention_mask = torch.ones(rows, (0 if past is None else past[0][0].shape[-2] - 1) + q)
    am = torch.cat((torch.ones(rows, 1), attention_mask.to(torch.float32)), dim=-1)[:, None, None, :]  # :318-326
    am = (1.0 - am) * -10000.0  # :334
    presents = []
    for i in range(24):
        p = "%s.gpt2_blocks.%d" % (lm, i)
        res = h
        x = F.layer_norm(h, (1024,), sd[p + ".0.weight"], sd[p + ".0.bias"], 1e-5)
        qw, kw, vw = _conv1d(sd, p + ".1.c_attn", x).split(1024, dim=2)  # :132
        if past is None:
            k_img = F.linear(img, sd[p + ".1.uk.weight"], sd[p + ".1.uk.bias"])[:, None, :]  # :140
            v_img = F.linear(img, sd[p + ".1.uv.weight"], sd[p + ".1.uv.bias"])[:, None, :]
            if k_img.shape[0] != kw.shape[0]:  # beam search: :143-147
                nb = kw.shape[0] // k_img.shape[0]
                k_img, v_img = k_img.repeat_interleave(nb, 0), v_img.repeat_interleave(nb, 0)
            K = _split_heads(torch.cat((k_img, kw), dim=1))
            V = _split_heads(torch.cat((v_img, vw), dim=1))
        else:
            K = torch.cat((past[i][0], _split_heads(kw)), dim=-2)  # :169 (full regrow every step)
            V = torch.cat((past[i][1], _split_heads(vw)), dim=-2)
        Q = _split_heads(qw)
        presents.append((K, V))
        w = torch.matmul(Q, K.transpose(-1, -2)) / (64 ** 0.5)  # :85-88
        ql, kl = Q.shape[-2], K.shape[-2]
        causal = torch.tril(torch.ones(1024, 1024, dtype=torch.bool))[kl - ql: kl, :kl]  # :96
        w = torch.where(causal, w, torch.tensor(-1e4))  # :99
        w = F.softmax(w + am, dim=-1)  # :104-106
        a = torch.matmul(w, V).permute(0, 2, 1, 3).reshape(rows, ql, 1024)
        h = _conv1d(sd, p + ".1.c_proj", a) + res  # :177, :350
        res = h
        x = F.layer_norm(h, (1024,), sd[p + ".2.weight"], sd[p + ".2.bias"], 1e-5)
        x = _conv1d(sd, p + ".3.c_proj", gelu_new(_conv1d(sd, p + ".3.c_fc", x)))  # HF GPT2MLP
        h = x + res
        if taps is not None:
            taps.setdefault("hidden", []).append(h)
    h = F.layer_norm(h, (1024,), sd[lm + ".final_layernorm.weight"], sd[lm + ".final_layernorm.bias"], 1e-5)
    return F.linear(h, wte), presents  # lm_head tied to wte (:366)


def greedy_search(sd: SD, feats: torch.Tensor, max_length: Optional[int], record: Optional[dict] = None,
                  given_logits: Optional[torch.Tensor] = None) -> torch.Tensor:
    """language_model.py:609-652 (+ prepare_inputs_for_generation :498-520).  given_logits [steps, rows, V] replaces the
    model forward (bookkeeping tests)."""
    rows = feats.shape[0]
    ids = torch.full((rows, 1), BOS, dtype=torch.int64)
    mask = torch.ones(rows, 1, dtype=torch.int64)
    unfinished = torch.ones(rows, dtype=torch.int64)
    past = None
    cur_len = 1
    while True:
        inp = ids if past is None else ids[:, -1:]
        pos = mask.cumsum(-1) - 1
        pos = pos if past is None else pos[:, -1:]
        if given_logits is not None:
            nxt_logits = given_logits[cur_len - 1]
        else:
            logits, past = lm_forward(sd, inp, feats, past, pos, mask)
            nxt_logits = logits[:, -1, :]
        if record is not None:
            record.setdefault("logits", []).append(nxt_logits.clone())
        nxt = torch.argmax(nxt_logits, dim=-1)
        nxt = nxt * unfinished + PAD * (1 - unfinished)
        ids = torch.cat([ids, nxt[:, None]], dim=-1)
        mask = torch.cat([mask, mask.new_ones(rows, 1)], dim=-1)
        cur_len += 1
        unfinished = unfinished * (nxt != EOS).long()
        if unfinished.max() == 0 or (max_length and cur_len >= max_length):
            break
    return ids


def beam_search(sd: SD, feats: torch.Tensor, max_length: int, num_beams: int, early_stopping: bool,
                given_logits: Optional[torch.Tensor] = None, stable_ties: bool = False) -> torch.Tensor:
    """language_model.py:529-607 with BeamSearchScorer(length_penalty=1.0, num_beam_hyps_to_keep=1) (:457-464).
    given_logits [steps, rows * beams, V] replaces the model forward (bookkeeping tests).
    stable_ties: order exactly tied candidates by lowest flat index (a stable descending sort).  torch.topk leaves the
    order of equal elements unspecified (it differs between the CPU and CUDA kernels), so the reference's behaviour on
    exact ties is implementation-defined; the engine's top-k breaks ties by lowest flat index."""
    from beam_scorer import BeamSearchScorer

    batch = feats.shape[0]
    scorer = BeamSearchScorer(batch_size=batch, num_beams=num_beams, device=torch.device("cpu"), length_penalty=1.0,
                              do_early_stopping=early_stopping, num_beam_hyps_to_keep=1)
    ids = torch.full((batch * num_beams, 1), BOS, dtype=torch.int64)  # _expand_inputs_for_generation :481-490
    mask = torch.ones(batch * num_beams, 1, dtype=torch.int64)
    beam_scores = torch.zeros(batch, num_beams)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    past = None
    cur_len = 1
    while True:
        inp = ids if past is None else ids[:, -1:]
        pos = mask.cumsum(-1) - 1
        pos = pos if past is None else pos[:, -1:]
        if given_logits is not None:
            step_logits = given_logits[cur_len - 1]
        else:
            logits, past = lm_forward(sd, inp, feats, past, pos, mask)
            step_logits = logits[:, -1, :]
        scores = F.log_softmax(step_logits, dim=-1) + beam_scores[:, None]
        V = scores.shape[-1]
        if stable_ties:
            srt, order = torch.sort(scores.view(batch, num_beams * V), dim=1, descending=True, stable=True)
            scores, tokens = srt[:, : 2 * num_beams], order[:, : 2 * num_beams]
        else:
            scores, tokens = torch.topk(scores.view(batch, num_beams * V), 2 * num_beams, dim=1, largest=True, sorted=True)
        indices = torch.div(tokens, V, rounding_mode="floor")
        tokens = tokens % V
        out = scorer.process(ids, scores, tokens, indices, pad_token_id=PAD, eos_token_id=EOS)
        beam_scores, beam_tok, beam_idx = out["next_beam_scores"], out["next_beam_tokens"], out["next_beam_indices"]
        ids = torch.cat([ids[beam_idx, :], beam_tok.unsqueeze(-1)], dim=-1)
        mask = torch.cat([mask, mask.new_ones(mask.shape[0], 1)], dim=-1)
        if given_logits is None:
            past = [(k.index_select(0, beam_idx), v.index_select(0, beam_idx)) for k, v in past]  # _reorder_cache :492-496
        cur_len += 1
        if scorer.is_done or (max_length and cur_len >= max_length):
            break
    return scorer.finalize(ids, beam_scores, tokens, indices, pad_token_id=PAD, eos_token_id=EOS,
                           max_length=max_length)["sequences"]


def lm_generate(sd: SD, feats: torch.Tensor, max_length=None, num_beams=1, num_beam_groups=1, do_sample=False,
                num_return_sequences=1, early_stopping=False, record: Optional[dict] = None) -> torch.Tensor:
    """language_model.py:401-479 mode dispatch and error behaviour."""
    greedy = num_beams == 1 and num_beam_groups == 1 and do_sample is False
    sample = num_beams == 1 and num_beam_groups == 1 and do_sample is True
    beam = num_beams > 1 and num_beam_groups == 1 and do_sample is False
    beam_sample = num_beams > 1 and num_beam_groups == 1 and do_sample is True
    group = num_beams > 1 and num_beam_groups > 1
    if num_beam_groups > num_beams:
        raise ValueError("'num_beam_groups' has to be smaller or equal to 'num_beams'")
    if group and do_sample is True:
        raise ValueError("Diverse beam search cannot be used in sampling mode. Make sure that 'do_sample' is set to 'False'.")
    if greedy:
        if num_return_sequences > 1:
            raise ValueError("num_return_sequences has to be 1, but is %d when doing greedy search." % num_return_sequences)
        return greedy_search(sd, feats, max_length, record)
    if sample:
        raise NotImplementedError("Multinomial sampling is not implemented.")
    if beam:
        if num_return_sequences > num_beams:
            raise ValueError("'num_return_sequences' has to be smaller or equal to 'num_beams'.")
        if max_length is None:
            raise ValueError("max_length has to be set for beam generation.")
        return beam_search(sd, feats, max_length, num_beams, early_stopping)
    if beam_sample:
        raise NotImplementedError("Beam-search multinomial sampling is not implemented.")
    if group:
        raise NotImplementedError("Diverse beam-search decoding is not implemented.")


# ----------------------------------------------------------------------------------------------------------------------
# a1/a2: the whole path  (report_generation_model.py:212-276; object_detector.py:184-261)
# ----------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def detect(sd: SD, images: torch.Tensor, detail: Optional[dict] = None) -> dict:
    S = images.shape[-1]
    feats = backbone(sd, images)
    rpn_detail = {} if detail is not None else None
    proposals = rpn(sd, feats, S, rpn_detail)
    roi_detail = {} if detail is not None else None
    out = roi_heads(sd, feats, proposals, S, roi_detail)
    if detail is not None:
        detail.update(features=feats, proposals=proposals, rpn=rpn_detail, roi=roi_detail)
    return out


@torch.no_grad()
def bbox_features(sd: SD, images: torch.Tensor, bbox_coordinates: List[torch.Tensor]) -> torch.Tensor:
    """evaluate_bbox_variations.py:92-110 `get_bbox_features`: user boxes (29 per image) -> RoIAlign -> AvgPool(8) ->
    dim_reduction -> [(B*29), 1024]; the caller then runs `language_model.generate` on every row (:131-136)."""
    S = images.shape[-1]
    feats = backbone(sd, images)
    pooled = box_roi_pool(feats, bbox_coordinates, S)
    f = F.avg_pool2d(pooled, 8).reshape(pooled.shape[0], -1)
    rh = "object_detector.roi_heads.dim_reduction"
    return F.linear(f, sd[rh + ".weight"], sd[rh + ".bias"])


@torch.no_grad()
def generate(sd: SD, images: torch.Tensor, max_length=None, num_beams=1, num_beam_groups=1, do_sample=False,
             num_return_sequences=1, early_stopping=False, detail: Optional[dict] = None):
    det = detect(sd, images, detail)
    selected, sel_feats, sel_logits = region_selection(sd, det["top_region_features"], det["class_detected"])
    if detail is not None:
        detail.update(selection_logits=sel_logits, selected_region_features=sel_feats)
    if sel_feats.shape[0] == 0:
        return -1  # report_generation_model.py:260-261
    ids = lm_generate(sd, sel_feats, max_length, num_beams, num_beam_groups, do_sample, num_return_sequences,
                      early_stopping)
    detections = {"top_region_boxes": det["top_region_boxes"], "top_scores": det["top_scores"]}
    return ids, selected, detections, det["class_detected"]
