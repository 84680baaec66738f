"""ctypes binding of include/rgrg_b200.h (the C-ABI shared library built in-tree by __graft_entry__.build())."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librgrg_b200.so")

_lib = None

c_p = C.c_void_p
_i, _f = C.c_int, C.c_float
_PROTOTYPES = {
    "rgrg_create": (_i, [_i, C.POINTER(c_p)]),
    "rgrg_destroy": (None, [c_p]),
    "rgrg_last_error": (C.c_char_p, [c_p]),
    "rgrg_version": (C.c_char_p, []),
    "rgrg_load_weight": (_i, [c_p, C.c_char_p, c_p, C.POINTER(C.c_int64), _i]),
    "rgrg_finalize_weights": (_i, [c_p]),
    "rgrg_generate": (_i, [c_p, c_p, _i, _i, _i, _i, _i, _i, c_p, C.POINTER(_i), c_p, c_p, c_p, c_p, C.POINTER(_i), c_p]),
    "rgrg_lm_generate": (_i, [c_p, c_p, _i, _i, _i, _i, _i, c_p, C.POINTER(_i), c_p]),
    "rgrg_detect": (_i, [c_p, c_p, _i, _i, _i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, C.POINTER(_i), c_p]),
    "rgrg_bbox_features": (_i, [c_p, c_p, _i, _i, _i, c_p, c_p, c_p]),
    "rgrg_lm_forced_logits": (_i, [c_p, c_p, _i, c_p, _i, c_p, c_p]),
    "rgrg_comm_unique_id": (_i, [c_p]),
    "rgrg_comm_init": (_i, [c_p, c_p, _i, _i]),
    "rgrg_allgather_results": (_i, [c_p, _i, _i, c_p, C.c_size_t, c_p]),
    "rgrg_preprocess": (_i, [c_p, c_p, _i, _i, _i, c_p, _i, c_p]),
    "rgrg_greedy_bookkeeping": (_i, [c_p, c_p, _i, _i, _i, c_p, C.POINTER(_i), c_p]),
    "rgrg_beam_bookkeeping": (_i, [c_p, c_p, _i, _i, _i, _i, _i, c_p, C.POINTER(_i), c_p]),
    "rgrg_rpn_filter": (_i, [c_p, c_p, c_p, c_p, _i, _i, _i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "rgrg_roi_align": (_i, [c_p, c_p, c_p, c_p, _i, _i, _i, _i, c_p, c_p]),
    "rgrg_roi_tail": (_i, [c_p, c_p, c_p, c_p, c_p, _i, _i, c_p, c_p, c_p, c_p, c_p]),
    "rgrg_gemm_bf16": (_i, [c_p, c_p, c_p, c_p, _i, _i, _i, _i, _i, c_p, c_p]),
    "rgrg_gemm_bench": (_i, [c_p, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float), c_p, _i]),
    "rgrg_conv3x3_bf16": (_i, [c_p, c_p, c_p, c_p, _i, _i, _i, _i, _i, _i, _i, c_p, c_p]),
    "rgrg_backbone": (_i, [c_p, c_p, _i, _i, c_p, c_p]),
    "rgrg_debug_read": (_i, [c_p, C.c_char_p, c_p, C.c_size_t]),
    "rgrg_set_option": (_i, [c_p, C.c_char_p, _i]),
    "rgrg_profile_read": (_i, [c_p, C.c_char_p, C.c_size_t]),
    "rgrg_kernel_launches": (C.c_int64, [c_p]),
}


def exported_symbols():
    return sorted(_PROTOTYPES)


def load():
    """Loads librgrg_b200.so.  There is NO fallback: a missing library is an error (build it with
    `python __graft_entry__.py`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "rgrg_b200: %s not found — the CUDA extension is required (no CPU / PyTorch fallback). "
                "Build it with `python __graft_entry__.py`." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
