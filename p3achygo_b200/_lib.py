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
LIB_PATH = os.path.join(_HERE, "libp3b200.so")

NUM_LOCS = 361
NUM_MOVES = 362

P3_OK = 0
P3_ERR_INVALID_ARG, P3_ERR_NO_DEVICE, P3_ERR_CUDA, P3_ERR_IO, P3_ERR_UNSUPPORTED = 1, 2, 3, 4, 5
PRECISION_FP32, PRECISION_BF16 = 0, 1


class P3Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libp3b200 error {code}: {msg}")
        self.code = code


class Loc(ctypes.Structure):
    _fields_ = [("i", ctypes.c_int32), ("j", ctypes.c_int32)]


class GoFeatures(ctypes.Structure):
    """nn::GoFeatures, cc/nn/engine/go_features.h:12-22 (1860 bytes)."""
    _fields_ = [
        ("bsize", ctypes.c_int32),
        ("color", ctypes.c_int8),
        ("komi", ctypes.c_float),
        ("board", ctypes.c_int8 * NUM_LOCS),
        ("last_moves", Loc * 5),
        ("stones_atari", ctypes.c_int8 * NUM_LOCS),
        ("stones_two_liberties", ctypes.c_int8 * NUM_LOCS),
        ("stones_three_liberties", ctypes.c_int8 * NUM_LOCS),
        ("stones_laddered", ctypes.c_int8 * NUM_LOCS),
    ]


# numpy view of the same 1860-byte record (for bulk fixtures)
GO_FEATURES_DTYPE = np.dtype({
    "names": ["bsize", "color", "komi", "board", "last_moves", "stones_atari", "stones_two_liberties",
              "stones_three_liberties", "stones_laddered"],
    "formats": ["<i4", "i1", "<f4", ("i1", NUM_LOCS), ("<i4", (5, 2)), ("i1", NUM_LOCS), ("i1", NUM_LOCS),
                ("i1", NUM_LOCS), ("i1", NUM_LOCS)],
    "offsets": [0, 4, 8, 12, 376, 416, 777, 1138, 1499],
    "itemsize": 1860,
})

# nn::NNInferResult, cc/nn/engine/engine.h:12-20 (7568 bytes, opt_move_probs 16-byte aligned)
INFER_RESULT_DTYPE = np.dtype({
    "names": ["move_logits", "move_probs", "value_probs", "score_probs", "opt_move_probs", "err2_outcome"],
    "formats": [("<f4", NUM_MOVES), ("<f4", NUM_MOVES), ("<f4", 2), ("<f4", 800), ("<f4", NUM_MOVES), "<f4"],
    "offsets": [0, 1448, 2896, 2904, 6112, 7560],
    "itemsize": 7568,
})

AUX_RESULT_DTYPE = np.dtype([
    ("pi_logits_aux", "<f4", NUM_MOVES), ("pi_logits_soft", "<f4", NUM_MOVES), ("pi_logits_optimistic", "<f4", NUM_MOVES),
    ("outcome_logits", "<f4", 2), ("score_logits", "<f4", 800), ("gamma", "<f4"), ("q", "<f4", 3), ("q_err", "<f4", 3),
    ("q_score", "<f4", 3), ("q_score_err", "<f4", 3), ("mcts_dist_logits", "<f4", 51), ("mcts_dist_probs", "<f4", 51),
    ("ownership", "<f4", NUM_LOCS), ("value", "<f4"), ("score_mean", "<f4"), ("score_var", "<f4"),
])

assert ctypes.sizeof(GoFeatures) == 1860 and GO_FEATURES_DTYPE.itemsize == 1860

# every symbol include/p3_b200.h declares
EXPORTS = [
    "p3_engine_create", "p3_engine_destroy", "p3_engine_load_batch", "p3_engine_load_batch_sym", "p3_engine_run_inference", "p3_engine_get_batch",
    "p3_engine_load_batch_bank", "p3_engine_submit", "p3_engine_wait", "p3_engine_get_batch_bank", "p3_engine_load_game_bank",
    "p3_engine_get_ownership", "p3_engine_path", "p3_engine_batch_size", "p3_engine_get_planes", "p3_engine_get_aux",
    "p3_engine_run_device", "p3_engine_upload", "p3_engine_profile", "p3_engine_stage_ms", "p3_engine_launches_per_run", "p3_engine_flops_per_position",
    "p3_engine_set_cuda_graph", "p3_encode_features", "p3_board_liberties", "p3_legal_mask", "p3_game_derive", "p3_gumbel_topk",
    "p3_conv_test", "p3_broadcast_test", "p3_block_boundary_test", "p3_last_error", "p3_version",
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
    lib.p3_engine_profile.argtypes = [vp, vp, vp, vp]
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
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != P3_OK:
        raise P3Error(rc, lib.p3_last_error().decode(errors="replace"))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)
