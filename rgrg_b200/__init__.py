"""rgrg_b200 — B200-native (sm_100a) engine for the region-guided report-generation inference path of ttanida/rgrg.

The numeric path lives entirely in librgrg_b200.so (hand-written CUDA, C ABI in include/rgrg_b200.h); this package is
the thin host side: `ReportGenerationModel` mirrors the reference's call surface, `Engine` owns one C-ABI handle.
Nothing here falls back to PyTorch or the CPU: without the built library the engine raises.
"""
from . import _cabi  # noqa: F401
from .engine import Engine  # noqa: F401
from . import report_assembly  # noqa: F401
from .model import LanguageModel, ReportGenerationModel, get_bbox_features, get_image_tensor  # noqa: F401

__all__ = ["Engine", "ReportGenerationModel", "LanguageModel", "get_bbox_features", "get_image_tensor", "report_assembly"]
