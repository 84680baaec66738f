"""Import the UNMODIFIED reference (ttanida/rgrg, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  /root/reference exists only in the build container, never on the
GPU box.  This file is used
  * by oracle/make_golden.py to generate the committed fixtures under tests/golden/ (from /root/reference), and
  * by bench.py's CPU legs (`--impl reference`, `cpu_baseline`), which time the reference's own generate() on the host
    cores from oracle/_ref/ — byte-for-byte copies of the eight hot-path files made by oracle/install_ref.py
    (git-ignored, shipped to the GPU box like a built .so); when oracle/_ref/ is absent the CPU legs fall back to
    the oracle port and say so (`kind: "port"`).
The `-m gpu` tests and smoke() never import it.

Two stubs + two patches, plus `torch.cuda.is_available() -> False` while the reference is imported and
built (it picks its device at import time; the CPU legs must stay on the CPU on a GPU box); no reference file is edited:
  1. sys.modules['torchinfo']                         (language_model.py:6  `from torchinfo import summary`)
  2. sys.modules['transformers.generation_beam_search'] -> oracle BeamSearchScorer restatement
                                                      (language_model.py:8; transformers 4.19.2 is not installed)
  3. GPT2LMHeadModel.from_pretrained -> random-init gpt2-medium config (language_model.py:205; no network)
  4. object_detector.resnet50 -> resnet50(weights=None)              (object_detector.py:51; no network)
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_VENDORED = os.path.join(_HERE, "_ref")  # unmodified copies made by oracle/install_ref.py (git-ignored; travels to the GPU box)


def _default_root():
    env = os.environ.get("RGRG_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/src/full_model"):
        return "/root/reference"
    return _VENDORED


REFERENCE_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "full_model"))


class _pin_to_cpu:
    """The reference picks its device at import / construction time (`device = cuda if torch.cuda.is_available() else cpu`:
    language_model.py:199, binary_classifier_region_selection.py:4, ...) while the harness keeps the model on the CPU; on a
    box WITH a GPU that mixes devices inside greedy_search.  The CPU legs therefore import and build the reference with
    torch.cuda.is_available() answering False (restored afterwards) — the reference files stay untouched."""

    def __enter__(self):
        import torch

        self._saved = torch.cuda.is_available
        torch.cuda.is_available = lambda: False

    def __exit__(self, *exc):
        import torch

        torch.cuda.is_available = self._saved
        return False


def import_reference():
    """Returns the reference's ReportGenerationModel class (unmodified code, patched deps)."""
    if not reference_available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    import torch  # noqa: F401
    import transformers
    from transformers import GPT2Config, GPT2LMHeadModel

    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import beam_scorer  # oracle restatement of transformers==4.19.2 BeamSearchScorer

    if "torchinfo" not in sys.modules:
        ti = types.ModuleType("torchinfo")
        ti.summary = lambda *a, **k: None
        sys.modules["torchinfo"] = ti
    if "transformers.generation_beam_search" not in sys.modules:
        gb = types.ModuleType("transformers.generation_beam_search")
        gb.BeamSearchScorer = beam_scorer.BeamSearchScorer
        sys.modules["transformers.generation_beam_search"] = gb
        transformers.generation_beam_search = gb

    def _from_pretrained(cls, *a, **k):
        cfg = GPT2Config(n_embd=1024, n_layer=24, n_head=16, n_positions=1024,
                         vocab_size=50257, activation_function="gelu_new")
        return GPT2LMHeadModel(cfg)

    GPT2LMHeadModel.from_pretrained = classmethod(_from_pretrained)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torchvision
    import src.object_detector.object_detector as od

    od.resnet50 = lambda weights=None: torchvision.models.resnet50(weights=None)
    from src.full_model.report_generation_model import ReportGenerationModel
    return ReportGenerationModel


def build_reference_model(state_dict=None):
    import torch
    with _pin_to_cpu():
        RGM = import_reference()
        torch.manual_seed(0)
        model = RGM(pretrain_without_lm_model=True)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        # the reference aliases the GPT-2 weights 3x; our synthetic state_dict carries every alias
        assert not unexpected, unexpected[:5]
        assert not missing, missing[:5]
    model.eval()
    return model
