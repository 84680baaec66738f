"""Seeded synthetic checkpoint for the rgrg inference path (no network, no MIMIC weights).

Produces a `state_dict` with the reference's parameter names / shapes / dtypes
(canonical alias set; names as produced by
`ReportGenerationModel.state_dict()`, src/full_model/report_generation_model.py:21-33).
Every tensor is drawn from its own `torch.Generator` seeded by crc32(name) ^ seed, so the result does
not depend on construction order and is identical wherever the same torch build runs.

Plain random init degenerates the path (~25 proposals / image, 2 of 29 regions
detected, chaotic backbone).  The conditioning below restores a realistic workload
(~850-900 proposals / image, 29 / 29 regions detected and selected, full-length decodes):
  1. every bottleneck bn3.weight = 0.2         (residual-branch damping)
  2. BN running statistics calibrated on seeded N(0,1) images (cumulative average, 3 passes)
  3. cls_score.bias = -(W @ mean fc7 feature)  (centres the 30-way RoI classifier)
  4. selection-head final bias = +2            (logit > -1 robustly true)
GPT-2 with N(0, 0.02) weights practically never emits EOS, so every row decodes max_length tokens:
a deterministic amount of work.

This module only uses torch / torchvision library ops; it is fixture + benchmark input, not part of the
engine, and it never touches the oracle.
"""
from __future__ import annotations

import math
import os
import zlib

import torch
import torch.nn.functional as F

NUM_LAYERS = 24
D_MODEL = 1024
VOCAB = 50257
RESNET_LAYERS = (3, 4, 6, 3)  # torchvision resnet50; trunk = children()[:-2] -> indices 0,1,(2,3),4..7


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(name, shape, std, seed):
    return torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32) * std


def _uniform(name, shape, bound, seed):
    return (torch.rand(shape, generator=_gen(name, seed), dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(sd, prefix, out_f, in_f, seed):
    bound = 1.0 / math.sqrt(in_f)
    sd[prefix + ".weight"] = _uniform(prefix + ".weight", (out_f, in_f), bound, seed)
    sd[prefix + ".bias"] = _uniform(prefix + ".bias", (out_f,), bound, seed)


def _conv_bn(sd, conv, bn, cout, cin, k, seed, gamma=1.0):
    std = math.sqrt(2.0 / (cout * k * k))  # kaiming normal, fan_out
    sd[conv + ".weight"] = _normal(conv + ".weight", (cout, cin, k, k), std, seed)
    sd[bn + ".weight"] = torch.full((cout,), gamma)
    sd[bn + ".bias"] = torch.zeros(cout)
    sd[bn + ".running_mean"] = torch.zeros(cout)
    sd[bn + ".running_var"] = torch.ones(cout)
    sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def backbone_block_specs():
    """[(prefix, cin, width, cout, stride, has_downsample)] for the 16 bottlenecks (ResNet-50 v1.5)."""
    specs = []
    cin = 64
    for li, (nblocks, width) in enumerate(zip(RESNET_LAYERS, (64, 128, 256, 512))):
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and li > 0) else 1
            specs.append(("object_detector.backbone.%d.%d" % (4 + li, bi), cin, width, width * 4, stride, bi == 0))
            cin = width * 4
    return specs


def _raw_state_dict(seed: int, parts=("detector", "heads", "lm")) -> dict:
    sd = {}
    if "detector" in parts:
        _raw_detector(sd, seed)
    if "heads" in parts:
        _raw_heads(sd, seed)
    if "lm" in parts:
        _raw_lm(sd, seed)
    return sd


def make_partial_state_dict(seed: int = 0, parts=("heads", "lm")) -> dict:
    """RNG-only subsets (no BN calibration): bit-identical on every host.  "detector" here is uncalibrated."""
    return _raw_state_dict(seed, parts)


def _raw_detector(sd, seed):
    bb = "object_detector.backbone"
    _conv_bn(sd, bb + ".0", bb + ".1", 64, 1, 7, seed)
    for prefix, cin, width, cout, stride, has_ds in backbone_block_specs():
        _conv_bn(sd, prefix + ".conv1", prefix + ".bn1", width, cin, 1, seed)
        _conv_bn(sd, prefix + ".conv2", prefix + ".bn2", width, width, 3, seed)
        _conv_bn(sd, prefix + ".conv3", prefix + ".bn3", cout, width, 1, seed, gamma=0.2)
        if has_ds:
            _conv_bn(sd, prefix + ".downsample.0", prefix + ".downsample.1", cout, cin, 1, seed)
    rpn = "object_detector.rpn.head"
    sd[rpn + ".conv.0.0.weight"] = _normal(rpn + ".conv.0.0.weight", (2048, 2048, 3, 3), 0.01, seed)
    sd[rpn + ".conv.0.0.bias"] = torch.zeros(2048)
    sd[rpn + ".cls_logits.weight"] = _normal(rpn + ".cls_logits.weight", (160, 2048, 1, 1), 0.01, seed)
    sd[rpn + ".cls_logits.bias"] = torch.zeros(160)
    sd[rpn + ".bbox_pred.weight"] = _normal(rpn + ".bbox_pred.weight", (640, 2048, 1, 1), 0.01, seed)
    sd[rpn + ".bbox_pred.bias"] = torch.zeros(640)
    rh = "object_detector.roi_heads"
    _linear(sd, rh + ".box_head.fc6", 1024, 2048 * 64, seed)
    _linear(sd, rh + ".box_head.fc7", 1024, 1024, seed)
    _linear(sd, rh + ".box_predictor.cls_score", 30, 1024, seed)
    _linear(sd, rh + ".box_predictor.bbox_pred", 120, 1024, seed)
    _linear(sd, rh + ".dim_reduction", 1024, 2048, seed)


def _raw_heads(sd, seed):
    for head in ("binary_classifier_region_selection", "binary_classifier_region_abnormal"):
        _linear(sd, head + ".classifier.0", 512, 1024, seed)
        _linear(sd, head + ".classifier.2", 128, 512, seed)
        _linear(sd, head + ".classifier.4", 1, 128, seed)
        sd[head + ".loss_fn.pos_weight"] = torch.tensor([2.2 if "selection" in head else 6.0])
    sd["binary_classifier_region_selection.classifier.4.bias"] += 2.0


def _raw_lm(sd, seed):
    lm = "language_model"
    sd[lm + ".wte.weight"] = _normal(lm + ".wte.weight", (VOCAB, D_MODEL), 0.02, seed)
    sd[lm + ".wpe.weight"] = _normal(lm + ".wpe.weight", (1024, D_MODEL), 0.02, seed)  # dead weight (F3)
    for i in range(NUM_LAYERS):
        p = "%s.gpt2_blocks.%d" % (lm, i)
        for ln in (".0", ".2"):
            sd[p + ln + ".weight"] = 1.0 + _normal(p + ln + ".weight", (D_MODEL,), 0.05, seed)
            sd[p + ln + ".bias"] = _normal(p + ln + ".bias", (D_MODEL,), 0.02, seed)
        sd[p + ".1.c_attn.weight"] = _normal(p + ".1.c_attn.weight", (D_MODEL, 3 * D_MODEL), 0.02, seed)
        sd[p + ".1.c_attn.bias"] = _normal(p + ".1.c_attn.bias", (3 * D_MODEL,), 0.02, seed)
        sd[p + ".1.c_proj.weight"] = _normal(p + ".1.c_proj.weight", (D_MODEL, D_MODEL), 0.02, seed)
        sd[p + ".1.c_proj.bias"] = _normal(p + ".1.c_proj.bias", (D_MODEL,), 0.02, seed)
        _linear(sd, p + ".1.uk", D_MODEL, D_MODEL, seed)
        _linear(sd, p + ".1.uv", D_MODEL, D_MODEL, seed)
        sd[p + ".3.c_fc.weight"] = _normal(p + ".3.c_fc.weight", (D_MODEL, 4 * D_MODEL), 0.02, seed)
        sd[p + ".3.c_fc.bias"] = _normal(p + ".3.c_fc.bias", (4 * D_MODEL,), 0.02, seed)
        sd[p + ".3.c_proj.weight"] = _normal(p + ".3.c_proj.weight", (4 * D_MODEL, D_MODEL), 0.02, seed)
        sd[p + ".3.c_proj.bias"] = _normal(p + ".3.c_proj.bias", (D_MODEL,), 0.02, seed)
    sd[lm + ".final_layernorm.weight"] = 1.0 + _normal(lm + ".final_layernorm.weight", (D_MODEL,), 0.05, seed)
    sd[lm + ".final_layernorm.bias"] = _normal(lm + ".final_layernorm.bias", (D_MODEL,), 0.02, seed)
    _linear(sd, lm + ".feature_space_transformation_nn.0", D_MODEL, D_MODEL, seed)
    _linear(sd, lm + ".feature_space_transformation_nn.2", D_MODEL, D_MODEL, seed)


def _calibration_images(n=4, size=512, seed=123):
    return torch.randn(n, 1, size, size, generator=torch.Generator().manual_seed(seed))


@torch.no_grad()
def _calibrate_bn_and_trunk_features(sd: dict) -> torch.Tensor:
    """Train-mode passes through a torchvision ResNet-50 trunk carrying our conv weights; BN running stats
    become the cumulative average over 3 passes (momentum=None).  Returns eval-mode features of the
    calibration batch [4, 2048, 16, 16]."""
    import torchvision

    net = torchvision.models.resnet50(weights=None)
    net.conv1 = torch.nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
    trunk = torch.nn.Sequential(*list(net.children())[:-2])
    own = {k[len("object_detector.backbone."):]: v for k, v in sd.items() if k.startswith("object_detector.backbone.")}
    trunk.load_state_dict(own, strict=True)
    for m in trunk.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = None
            m.reset_running_stats()
    trunk.train()
    x = _calibration_images()
    for _ in range(3):
        trunk(x)
    trunk.eval()
    for k, v in trunk.state_dict().items():
        sd["object_detector.backbone." + k] = v.clone()
    return trunk(x)


def _calibration_boxes(size=512):
    """Deterministic assortment of boxes (anchor-like sizes / ratios on a coarse grid) used only to estimate
    the mean fc7 feature for step 3."""
    boxes = []
    for s in (60, 120, 180, 300):
        for r in (0.4, 1.0, 2.6):
            h, w = s * math.sqrt(r), s / math.sqrt(r)
            for cy in (128, 256, 384):
                for cx in (128, 256, 384):
                    boxes.append([max(cx - w / 2, 0.0), max(cy - h / 2, 0.0), min(cx + w / 2, size), min(cy + h / 2, size)])
    return torch.tensor(boxes, dtype=torch.float32)


@torch.no_grad()
def _centre_cls_score(sd: dict, feats: torch.Tensor):
    import torchvision

    rh = "object_detector.roi_heads"
    boxes = _calibration_boxes()
    rois = [boxes for _ in range(feats.shape[0])]
    pooled = torchvision.ops.roi_align(feats, rois, output_size=(8, 8), spatial_scale=1.0 / 32, sampling_ratio=2)
    x = pooled.flatten(1)
    x = F.relu(F.linear(x, sd[rh + ".box_head.fc6.weight"], sd[rh + ".box_head.fc6.bias"]))
    x = F.relu(F.linear(x, sd[rh + ".box_head.fc7.weight"], sd[rh + ".box_head.fc7.bias"]))
    mean = x.mean(0)
    sd[rh + ".box_predictor.cls_score.bias"] = -(sd[rh + ".box_predictor.cls_score.weight"] @ mean)


def make_state_dict(seed: int = 0, cache: bool = True) -> dict:
    """Canonical-alias synthetic checkpoint (fp32 CPU tensors).  ~2.4 GB; cached under
    $RGRG_SYNTH_CACHE (default /tmp/rgrg_b200_synth) because generation takes ~1 min."""
    cache_dir = os.environ.get("RGRG_SYNTH_CACHE", "/tmp/rgrg_b200_synth")
    path = os.path.join(cache_dir, "synth_v1_seed%d.pt" % seed)
    if cache and os.path.exists(path):
        try:
            return torch.load(path, map_location="cpu")
        except Exception:
            pass
    nthreads = torch.get_num_threads()
    sd = _raw_state_dict(seed)
    feats = _calibrate_bn_and_trunk_features(sd)
    _centre_cls_score(sd, feats)
    torch.set_num_threads(nthreads)
    if cache:
        try:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = path + ".tmp%d" % os.getpid()
            torch.save(sd, tmp)
            os.replace(tmp, path)
        except OSError:
            pass
    return sd


def synthetic_images(batch: int, size: int = 512, seed: int = 1000) -> torch.Tensor:
    """N(0,1) pixels ~ the distribution after Normalize(mean .471, std .302)
    (generate_reports_for_images.py:29-30,138)."""
    return torch.randn(batch, 1, size, size, generator=torch.Generator().manual_seed(seed))


def crafted_logits(seed: int, steps: int, rows: int, eos_mask=None, eos_logit: float = 14.0, scale: float = 2.0) -> torch.Tensor:
    """Seeded logits [steps, rows, 50257] for the bookkeeping tests (greedy / beam loops driven by GIVEN logits instead
    of the model forward): N(0, scale^2) base, EOS (50256) raised to `eos_logit` where eos_mask[t, r] is set.  The same
    function builds the golden vectors (oracle/make_golden.py, through the reference's own search loops) and the
    inputs of the GPU tests."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(steps, rows, VOCAB, generator=g) * scale
    if eos_mask is not None:
        m = torch.as_tensor(eos_mask, dtype=torch.bool)
        logits[:, :, VOCAB - 1] = torch.where(m, torch.full_like(logits[:, :, 0], eos_logit), logits[:, :, VOCAB - 1])
    return logits

