"""CPU tests of the host-side mirror (no GPU): return contract of ReportGenerationModel.generate with a stub engine,
bench arithmetic, and the rule that the product never imports the oracle."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _StubEngine:
    def __init__(self, R):
        self.R = R
        self.calls = []

    def generate(self, images, max_length, num_beams, early_stopping):
        self.calls.append(("generate", tuple(images.shape), max_length, num_beams, early_stopping))
        B = images.shape[0]
        sel = np.zeros((B, 29), dtype=bool)
        sel.reshape(-1)[: self.R] = True
        return {"R": self.R, "ids": np.full((self.R, 5), 7, dtype=np.int32), "selected": sel, "detected": np.ones((B, 29), dtype=bool),
                "boxes": np.zeros((B, 29, 4), np.float32), "scores": np.zeros((B, 29), np.float32)}

    def lm_generate(self, feats, max_length, num_beams, early_stopping):
        self.calls.append(("lm_generate", tuple(feats.shape), max_length, num_beams, early_stopping))
        return np.full((feats.shape[0], 3), 9, dtype=np.int32)


def _model(stub):
    from rgrg_b200 import ReportGenerationModel

    m = ReportGenerationModel(pretrain_without_lm_model=True, device="cpu")  # device only labels the returned tensors here
    m._eng = stub
    return m


def test_generate_return_contract_matches_reference():
    """report_generation_model.py:276: (output_ids int64, selected_regions bool, detections dict, class_detected bool)."""
    m = _model(_StubEngine(R=3))
    out = m.generate(torch.zeros(2, 1, 512, 512), max_length=5, num_beams=4, early_stopping=True)
    ids, selected, detections, detected = out
    assert ids.dtype == torch.int64 and ids.shape == (3, 5)
    assert selected.dtype == torch.bool and selected.shape == (2, 29) and int(selected.sum()) == 3
    assert set(detections) == {"top_region_boxes", "top_scores"}
    assert detections["top_region_boxes"].shape == (2, 29, 4) and detections["top_scores"].shape == (2, 29)
    assert detected.dtype == torch.bool
    assert m._eng.calls == [("generate", (2, 1, 512, 512), 5, 4, True)]


def test_generate_returns_minus_one_when_nothing_selected():
    """report_generation_model.py:260-261."""
    assert _model(_StubEngine(R=0)).generate(torch.zeros(1, 1, 512, 512), max_length=5) == -1


def test_language_model_generate_mirror():
    m = _model(_StubEngine(R=1))
    ids = m.language_model.generate(torch.zeros(4, 1024), max_length=3, num_beams=1)
    assert ids.dtype == torch.int64 and ids.shape == (4, 3)
    with pytest.raises(ValueError):
        m.language_model.generate(torch.zeros(4, 1024), max_length=None, num_beams=4)


def test_to_normalises_the_device_before_comparing(monkeypatch):
    """.to("cuda") / .to(torch.device("cuda")) must not tear down a live engine on cuda:0."""
    from rgrg_b200 import ReportGenerationModel

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    m = ReportGenerationModel(pretrain_without_lm_model=True)

    class _E:
        closed = False

        def close(self):
            self.closed = True

    e = _E()
    m._eng = e
    m.to("cuda")
    m.to(torch.device("cuda"))
    m.to(torch.device("cuda", 0))
    assert m._eng is e and not e.closed and m.device == torch.device("cuda", 0)
    m.to("cuda:1")
    assert e.closed and m._eng is None
    with pytest.raises(RuntimeError):
        m.to("cpu")


def test_strict_load_reports_missing_keys_and_open_ended_length_is_capped(lm_sd):
    from rgrg_b200 import ReportGenerationModel
    from rgrg_b200 import model as M

    with pytest.raises(RuntimeError, match="Missing key"):
        ReportGenerationModel().load_state_dict(lm_sd)  # decoder-only checkpoint: the detector weights are missing
    ReportGenerationModel().load_state_dict(lm_sd, strict=False)
    stub = _StubEngine(R=1)
    _model(stub).language_model.generate(torch.zeros(2, 1024), max_length=None)
    assert stub.calls[-1][2] == M.OPEN_ENDED_MAX_LENGTH == 300
    with pytest.raises(ValueError):
        _model(stub).generate(torch.zeros(1, 1, 512, 512), max_length=1)


def test_generate_without_weights_fails_loudly():
    from rgrg_b200 import ReportGenerationModel

    with pytest.raises(RuntimeError):
        ReportGenerationModel().generate(torch.zeros(1, 1, 512, 512), max_length=4)


def test_bench_flop_model_matches_survey_example():
    """S=512, P=850, R=29, T=64 -> ~1.59 TFLOP / image."""
    import bench

    assert abs(bench.flops_per_image(512, 850, 29, 64) / 1e12 - 1.59) < 0.01
    assert bench.category_flops("lm_head", 928, 0, 32, 512) == 2.0 * 928 * 50257 * 1024


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rgrg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+(rgrg_oracle|beam_scorer|ref_harness|oracle)\b", text, re.M), f


def test_every_engine_option_is_documented_in_the_header():
    """rgrg_set_option keys accepted by the engine == keys described in include/rgrg_b200.h (the drop-in boundary's documentation)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    eng = open(os.path.join(root, "rgrg_b200", "csrc", "engine.cu")).read()
    hdr = open(os.path.join(root, "include", "rgrg_b200.h")).read()
    keys = set(re.findall(r'k == "([a-z0-9_]+)"', eng))
    documented = set(re.findall(r'"([a-z0-9_]+)"', hdr))
    assert keys, "option parser not found"
    assert not (keys - documented), "undocumented options: %s" % sorted(keys - documented)
