"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/rgrg_b200.h declares,
and the Python mirror keeps the reference's argument checking (no compute calls: there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rgrg_b200.h")).read()
    return sorted(set(re.findall(r"RGRG_API [\w\* ]+?\b(rgrg_\w+)\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__

    __graft_entry__.build()
    import ctypes

    from rgrg_b200 import _cabi

    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(_cabi.exported_symbols()) == declared
    assert b"sm_100a" in _cabi.load().rgrg_version()


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rgrg_b200 import Engine

    with pytest.raises(RuntimeError):
        Engine(0)


def test_generate_argument_errors_match_reference():
    """language_model.py:428-479 raises before any device work."""
    from rgrg_b200 import ReportGenerationModel

    m = ReportGenerationModel(pretrain_without_lm_model=True)
    with pytest.raises(NotImplementedError):
        m.generate(None, max_length=8, do_sample=True)
    with pytest.raises(NotImplementedError):
        m.generate(None, max_length=8, num_beams=4, num_beam_groups=2)
    with pytest.raises(NotImplementedError):
        m.generate(None, max_length=8, num_beams=4, do_sample=True)
    with pytest.raises(ValueError):
        m.generate(None, max_length=None, num_beams=4)
    with pytest.raises(ValueError):
        m.generate(None, max_length=8, num_beams=2, num_beam_groups=4)
    with pytest.raises(ValueError):
        m.generate(None, max_length=8, num_return_sequences=2)
    with pytest.raises(RuntimeError):
        m.to("cpu")
