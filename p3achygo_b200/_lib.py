"""ctypes binding of libp3b200.so (include/p3_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``p3achygo_b200/csrc/Makefile``.
There is no Python or CPU fallback: if the shared object is missing the import of this module
raises, and every compute entry point returns ``P3_ERR_NO_DEVICE`` without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# P3_LIB: an experimental build of the same library (csrc/Makefile VARIANT=...), for perf A/B runs only
LIB_PATH = os.environ.get("P3_LIB") or os.path.join(_HERE, "libp3b200.so")

from .layout import (AUX_RESULT_DTYPE, GO_FEATURES_DTYPE, INFER_RESULT_DTYPE, LEAF_RESULT_DTYPE, NUM_LOCS, NUM_MOVES,  # noqa: F401
                     P3_ERR_CUDA, P3_ERR_INVALID_ARG, P3_ERR_IO, P3_ERR_NO_DEVICE, P3_ERR_UNSUPPORTED, P3_OK, PRECISION_BF16,
                     PRECISION_FP16, PRECISION_FP32, RESULT_FULL, RESULT_LEAF, GoFeatures, Loc)


class P3Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libp3b200 error {code}: {msg}")
        self.code = code


# every symbol include/p3_b200.h declares
EXPORTS = [
    "p3_engine_create", "p3_engine_destroy", "p3_engine_load_batch", "p3_engine_load_batch_sym", "p3_engine_run_inference", "p3_engine_get_batch",
    "p3_engine_load_batch_bank", "p3_engine_submit", "p3_engine_wait", "p3_engine_get_batch_bank", "p3_engine_load_game_bank",
    "p3_engine_get_ownership", "p3_engine_path", "p3_engine_batch_size", "p3_engine_get_planes", "p3_engine_get_aux",
    "p3_engine_run_device", "p3_engine_upload", "p3_engine_profile", "p3_engine_stage_ms", "p3_engine_launches_per_run", "p3_engine_flops_per_position",
    "p3_engine_set_cuda_graph", "p3_encode_features", "p3_board_liberties", "p3_legal_mask", "p3_game_derive", "p3_gumbel_topk",
    "p3_conv_test", "p3_broadcast_test", "p3_block_boundary_test", "p3_last_error", "p3_version",
    "p3_engine_set_result_mode", "p3_engine_get_leaf", "p3_engine_get_leaf_bank", "p3_engine_get_aux_bank",
    "p3_engine_get_ownership_bank", "p3_engine_gumbel_topk_bank", "p3_engine_range_check", "p3_engine_first_layer",
]


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). p3achygo_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.p3_last_error.restype = ctypes.c_char_p
    lib.p3_version.restype = ctypes.c_char_p
    lib.p3_engine_create.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, ctypes.POINTER(vp)]
    lib.p3_engine_destroy.argtypes = [vp]
    lib.p3_engine_destroy.restype = None
    lib.p3_engine_load_batch.argtypes = [vp, ci, vp]
    lib.p3_engine_load_batch_sym.argtypes = [vp, ci, vp, ci]
    lib.p3_engine_run_inference.argtypes = [vp]
    lib.p3_engine_get_batch.argtypes = [vp, ci, vp]
    lib.p3_engine_load_batch_bank.argtypes = [vp, ci, ci, vp, ci]
    lib.p3_engine_load_game_bank.argtypes = [vp, ci, ci, vp, ci, ci, cf, vp, ci]
    lib.p3_engine_submit.argtypes = [vp, ci]
    lib.p3_engine_wait.argtypes = [vp, ci]
    lib.p3_engine_get_batch_bank.argtypes = [vp, ci, ci, vp]
    lib.p3_engine_get_ownership.argtypes = [vp, ci, vp]
    lib.p3_engine_path.argtypes = [vp]
    lib.p3_engine_path.restype = ctypes.c_char_p
    lib.p3_engine_batch_size.argtypes = [vp]
    lib.p3_engine_get_planes.argtypes = [vp, ci, vp, vp]
    lib.p3_engine_get_aux.argtypes = [vp, ci, vp]
    lib.p3_engine_run_device.argtypes = [vp, ctypes.POINTER(cf)]
    lib.p3_engine_upload.argtypes = [vp]
    lib.p3_engine_profile.argtypes = [vp, vp, vp, vp, vp]
    lib.p3_engine_stage_ms.argtypes = [vp, ctypes.POINTER(cf * 3)]
    lib.p3_engine_launches_per_run.argtypes = [vp]
    lib.p3_engine_flops_per_position.argtypes = [vp]
    lib.p3_engine_flops_per_position.restype = ctypes.c_double
    lib.p3_engine_set_cuda_graph.argtypes = [vp, ci]
    lib.p3_encode_features.argtypes = [ci, vp, ci, ci, vp, vp]
    lib.p3_board_liberties.argtypes = [ci, vp, ci, vp]
    lib.p3_legal_mask.argtypes = [ci, vp, vp, vp, ci, vp]
    lib.p3_game_derive.argtypes = [ci, vp, vp, ci, vp, vp, ci, vp, vp, vp, vp]
    lib.p3_gumbel_topk.argtypes = [ci, vp, vp, vp, ci, cf, ci, vp, vp, vp]
    lib.p3_conv_test.argtypes = [ci, ci, vp, vp, ci, ci, ci, ci, vp]
    lib.p3_broadcast_test.argtypes = [ci, ci, vp, vp, vp, ci, ci, vp]
    lib.p3_block_boundary_test.argtypes = [ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, vp]
    lib.p3_engine_set_result_mode.argtypes = [vp, ci]
    lib.p3_engine_get_leaf.argtypes = [vp, ci, vp]
    lib.p3_engine_get_leaf_bank.argtypes = [vp, ci, ci, vp]
    lib.p3_engine_get_aux_bank.argtypes = [vp, ci, ci, vp]
    lib.p3_engine_get_ownership_bank.argtypes = [vp, ci, ci, vp]
    lib.p3_engine_range_check.argtypes = [vp, ctypes.POINTER(cf), ctypes.POINTER(ctypes.c_longlong)]
    lib.p3_engine_first_layer.argtypes = [vp, vp, vp, ctypes.c_size_t]
    lib.p3_engine_gumbel_topk_bank.argtypes = [vp, ci, vp, ci, vp, vp, cf, ci, vp, vp, vp]
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != P3_OK:
        raise P3Error(rc, lib.p3_last_error().decode(errors="replace"))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)
