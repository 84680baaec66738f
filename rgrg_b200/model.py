"""Host-side mirror of the reference's inference surface, backed by the C-ABI engine.

    model = ReportGenerationModel(pretrain_without_lm_model=True)          # report_generation_model.py:21-33
    model.load_state_dict(checkpoint["model"]); model.to(device); model.eval()  # generate_reports_for_images.py:160-163
    output = model.generate(images, max_length=..., num_beams=..., early_stopping=...)  # :109-114

`generate()` keeps the reference's signature, return tuple, `-1` sentinel and error behaviour
(report_generation_model.py:212-276; language_model.py:428-479).  Everything numeric happens inside
librgrg_b200.so; this file only converts buffers to the tensors the reference's callers expect.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import Engine


def _check_generation_mode(max_length, num_beams, num_beam_groups, do_sample, num_return_sequences):
    """language_model.py:422-479: same mode dispatch, same exception types and messages."""
    is_greedy = (num_beams == 1) and (num_beam_groups == 1) and do_sample is False
    is_sample = (num_beams == 1) and (num_beam_groups == 1) and do_sample is True
    is_beam = (num_beams > 1) and (num_beam_groups == 1) and do_sample is False
    is_beam_sample = (num_beams > 1) and (num_beam_groups == 1) and do_sample is True
    is_group_beam = (num_beams > 1) and (num_beam_groups > 1)
    if num_beam_groups > num_beams:
        raise ValueError("'num_beam_groups' has to be smaller or equal to 'num_beams'")
    if is_group_beam and do_sample is True:
        raise ValueError("Diverse beam search cannot be used in sampling mode. Make sure that 'do_sample' is set to 'False'.")
    if is_greedy:
        if num_return_sequences > 1:
            raise ValueError(f"num_return_sequences has to be 1, but is {num_return_sequences} when doing greedy search.")
        return "greedy"
    if is_sample:
        raise NotImplementedError("Multinomial sampling is not implemented.")
    if is_beam:
        if num_return_sequences > num_beams:
            raise ValueError("'num_return_sequences' has to be smaller or equal to 'num_beams'.")
        if max_length is None:
            raise ValueError("max_length has to be set for beam generation.")
        return "beam"
    if is_beam_sample:
        raise NotImplementedError("Beam-search multinomial sampling is not implemented.")
    if is_group_beam:
        raise NotImplementedError("Diverse beam-search decoding is not implemented.")
    raise ValueError("unsupported generation mode")


OPEN_ENDED_MAX_LENGTH = 300


def _resolve_max_length(max_length):
    """Greedy decoding with max_length=None runs "until every row emitted EOS" in the reference (language_model.py:649),
    growing its cache step by step.  The engine preallocates the KV cache (98 KB per row and token), so an open-ended
    call is capped at OPEN_ENDED_MAX_LENGTH tokens — the limit the reference's own script uses
    (generate_reports_for_images.py:27 MAX_NUM_TOKENS_GENERATE = 300; trained sentences are at most ~60 tokens,
    run_configurations.py:50-52) — instead of the 1024 positions GPT-2 allows (93 GB for 32 images).  max_length = 1 would
    return only the BOS column in the reference; the engine needs at least one decode step."""
    if max_length is None:
        return OPEN_ENDED_MAX_LENGTH
    max_length = int(max_length)
    if max_length < 2:
        raise ValueError("rgrg_b200 needs max_length >= 2 (one decode step); the reference returns the BOS column for 1")
    return max_length


class LanguageModel:
    """Mirror of src/language_model/language_model.py `LanguageModel.generate` (the entry
    evaluate_bbox_variations.py:131-136 calls directly)."""

    def __init__(self, owner: "ReportGenerationModel"):
        self._owner = owner

    @torch.no_grad()
    def generate(self, image_hidden_states, max_length=None, num_beams=1, num_beam_groups=1, do_sample=False,
                 num_return_sequences=1, early_stopping=False) -> torch.LongTensor:
        _check_generation_mode(max_length, num_beams, num_beam_groups, do_sample, num_return_sequences)
        max_length = _resolve_max_length(max_length)
        ids = self._owner._engine().lm_generate(image_hidden_states, int(max_length), int(num_beams), bool(early_stopping))
        return torch.from_numpy(ids.astype(np.int64)).to(self._owner.device)


def get_bbox_features(model: "ReportGenerationModel", images, bbox_coordinates) -> torch.Tensor:
    """Drop-in for evaluate_bbox_variations.py:92-110 `get_bbox_features(model, images, bbox_coordinates)`:
    bbox_coordinates is a list (len = batch) of [29,4] tensors; returns [(batch*29), 1024] region features that
    `model.language_model.generate` turns into one sentence per box (:131-136)."""
    feats = model._engine().bbox_features(images, bbox_coordinates)
    return torch.from_numpy(feats).to(model.device)


# one representative key per weight group the engine reads (SURVEY.md §8(b) canonical alias set); checked by
# load_state_dict(strict=True) so that a wrong checkpoint fails at load time, like nn.Module.load_state_dict
_REQUIRED_KEYS = (
    "object_detector.backbone.0.weight", "object_detector.backbone.7.2.conv3.weight",
    "object_detector.rpn.head.cls_logits.weight", "object_detector.roi_heads.box_head.fc6.weight",
    "object_detector.roi_heads.box_predictor.cls_score.weight", "object_detector.roi_heads.dim_reduction.weight",
    "binary_classifier_region_selection.classifier.0.weight", "language_model.gpt2_blocks.23.3.c_proj.weight",
    "language_model.final_layernorm.weight", "language_model.feature_space_transformation_nn.2.weight",
)


def get_image_tensor(model: "ReportGenerationModel", image) -> torch.Tensor:
    """Drop-in for generate_reports_for_images.py:129-147 `get_image_tensor`, on the GPU: `image` is a path (read with
    cv2.imread(..., IMREAD_UNCHANGED) like the reference) or an 8-bit grayscale array [H, W]; returns the
    [1, 1, 512, 512] fp32 CUDA tensor `model.generate` consumes (INTER_AREA resize to longest side 512, centre zero
    pad, normalise with mean 0.471 / std 0.302 — bit-exact against cv2 + albumentations 1.1.0)."""
    if isinstance(image, str):
        import cv2

        image = cv2.imread(image, cv2.IMREAD_UNCHANGED)
        if image is None:
            raise FileNotFoundError(image)
    return model._engine().preprocess([image])


class ReportGenerationModel:
    def __init__(self, pretrain_without_lm_model: bool = False, device: Optional[torch.device] = None):
        self.pretrain_without_lm_model = pretrain_without_lm_model
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        self.training = False
        self._eng: Optional[Engine] = None
        self._state_dict: Optional[Dict[str, torch.Tensor]] = None
        self.language_model = LanguageModel(self)

    # ---- nn.Module-flavoured plumbing the reference's script uses (generate_reports_for_images.py:160-163)
    def load_state_dict(self, state_dict, strict: bool = True):
        if strict:
            missing = [k for k in _REQUIRED_KEYS if k not in state_dict]
            if missing:
                raise RuntimeError("Error(s) in loading state_dict for ReportGenerationModel:\n\tMissing key(s) in state_dict: "
                                   + ", ".join('"%s"' % k for k in missing[:8]) + (" ..." if len(missing) > 8 else ""))
        self._state_dict = state_dict
        if self._eng is not None:
            self._eng.close()
            self._eng = None
        return self

    def to(self, device, non_blocking: bool = False):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("rgrg_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        # normalise first: torch.device("cuda") != torch.device("cuda", 0), and a spurious mismatch would tear down the
        # engine (and re-upload the whole checkpoint on the next call)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        if self._eng is not None and device != self.device:
            self._eng.close()
            self._eng = None
        self.device = device
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("rgrg_b200 is an inference engine (the training path is out of scope)")
        return self

    def _engine(self) -> Engine:
        if self._eng is None:
            if self._state_dict is None:
                raise RuntimeError("load_state_dict() must be called before generate()")
            eng = Engine(self.device.index or 0)
            eng.load_state_dict(self._state_dict)
            self._eng = eng
        return self._eng

    @torch.no_grad()
    def generate(self, images, max_length: int = None, num_beams: int = 1, num_beam_groups: int = 1,
                 do_sample: bool = False, num_return_sequences: int = 1, early_stopping: bool = False):
        """report_generation_model.py:212-276.  Returns (output_ids int64 [R, T'], selected_regions bool [B,29],
        detections {"top_region_boxes" [B,29,4], "top_scores" [B,29]}, class_detected bool [B,29]) or -1."""
        _check_generation_mode(max_length, num_beams, num_beam_groups, do_sample, num_return_sequences)
        max_length = _resolve_max_length(max_length)
        out = self._engine().generate(images, int(max_length), int(num_beams), bool(early_stopping))
        if out["R"] == 0:
            return -1
        dev = self.device
        output_ids = torch.from_numpy(out["ids"].astype(np.int64)).to(dev)
        selected_regions = torch.from_numpy(out["selected"]).to(dev)
        detections = {"top_region_boxes": torch.from_numpy(out["boxes"]).to(dev),
                      "top_scores": torch.from_numpy(out["scores"]).to(dev)}
        class_detected = torch.from_numpy(out["detected"]).to(dev)
        return output_ids, selected_regions, detections, class_detected
