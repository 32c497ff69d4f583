"""Record layouts of the C ABI (include/p3_b200.h) as ctypes / numpy types.

Pure data description: importing this module does NOT load libp3b200.so, so tools that only need the layouts
(``bench.py --impl reference``, fixture generators) never map the CUDA library.
"""
from __future__ import annotations

import ctypes

import numpy as np

NUM_LOCS = 361
NUM_MOVES = 362

P3_OK = 0
P3_ERR_INVALID_ARG, P3_ERR_NO_DEVICE, P3_ERR_CUDA, P3_ERR_IO, P3_ERR_UNSUPPORTED = 1, 2, 3, 4, 5
PRECISION_FP32, PRECISION_BF16, PRECISION_FP16 = 0, 1, 2


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


# p3_leaf_result (include/p3_b200.h): what mcts::LeafEvaluator InitFields keeps, cc/mcts/leaf_evaluator.cc:83-112 (4360 bytes)
LEAF_RESULT_DTYPE = np.dtype([
    ("move_logits", "<f4", NUM_MOVES), ("move_probs", "<f4", NUM_MOVES), ("opt_move_probs", "<f4", NUM_MOVES),
    ("value", "<f4"), ("score_mean", "<f4"), ("score_var", "<f4"), ("err", "<f4"),
])
assert LEAF_RESULT_DTYPE.itemsize == 4360
RESULT_FULL, RESULT_LEAF = 0, 1
