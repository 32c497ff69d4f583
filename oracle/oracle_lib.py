"""TEST INFRASTRUCTURE — ctypes loaders for the two checkers.

``oracle()``  -> libp3oracle.so, the C restatement (oracle/features_oracle.c); always available after
                 ``make -C oracle libp3oracle.so``.
``ref()``     -> oracle/_ref/libp3ref.so, the unmodified reference sources (oracle/ref_driver.cc); built only
                 where /root/reference exists, otherwise the prebuilt file (or None).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = None
_REF = None

vp, ci, cf, u64p = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.POINTER(ctypes.c_uint64)


def build(quiet: bool = True) -> None:
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.STDOUT if quiet else None)


def oracle() -> ctypes.CDLL:
    global _ORACLE
    if _ORACLE is None:
        path = os.path.join(HERE, "libp3oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", HERE, "libp3oracle.so"], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        L.orc_transform_index.argtypes = [ci, ci]
        L.orc_transform_inv.argtypes = [ci, ci]
        L.orc_apply_symmetry_i8.argtypes = [ci, vp, vp]
        L.orc_apply_inverse_f32.argtypes = [ci, vp, vp]
        L.orc_load_go_features.argtypes = [vp, ci, ci, vp, vp]
        L.orc_stones_with_liberties.argtypes = [vp, ci, vp]
        L.orc_legal_mask_nohist.argtypes = [vp, ctypes.c_int8, vp, vp]
        L.orc_prng_seed.argtypes = [ctypes.c_uint64]
        L.orc_prng_seed.restype = ctypes.c_uint64
        L.orc_prng_next.argtypes = [u64p]
        L.orc_prng_next.restype = ctypes.c_uint32
        L.orc_uniform.argtypes = [u64p]
        L.orc_uniform.restype = cf
        L.orc_gumbel.argtypes = [u64p]
        L.orc_gumbel.restype = cf
        L.orc_rand_range.argtypes = [u64p, ci, ci]
        L.orc_gumbel_topk.argtypes = [u64p, vp, vp, cf, ci, vp, vp]
        L.orc_softmax.argtypes = [ci, vp, vp]
        L.orc_init_fields.argtypes = [vp, vp, vp]
        L.orc_game_derive.argtypes = [vp, ci, vp, ctypes.c_int8, vp, vp, vp]
        _ORACLE = L
    return _ORACLE


def ref():
    """The compiled reference, or None when it is not available (GPU box without a prebuilt copy)."""
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "libp3ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/cc/game"):
                subprocess.run(["make", "-C", HERE, "_ref/libp3ref.so"], check=True, stdout=subprocess.DEVNULL)
            else:
                return None
        L = ctypes.CDLL(path)
        L.ref_game_new.restype = vp
        L.ref_game_new.argtypes = [cf, ci]
        L.ref_game_free.argtypes = [vp]
        L.ref_game_play.argtypes = [vp, ci, ci, ci]
        L.ref_game_num_moves.argtypes = [vp]
        L.ref_game_is_over.argtypes = [vp]
        L.ref_game_board.argtypes = [vp, vp]
        L.ref_game_liberties.argtypes = [vp, ci, vp]
        L.ref_game_laddered.argtypes = [vp, vp]
        L.ref_game_legal_mask.argtypes = [vp, ci, vp]
        L.ref_game_features.argtypes = [vp, ci, ci, vp]
        L.ref_game_moves.argtypes = [vp, vp, ci]
        L.ref_game_move_status.argtypes = [vp, ci, vp]
        L.ref_load_go_features.argtypes = [vp, ci, ci, vp, vp]
        L.ref_transform_index.argtypes = [ci, ci]
        L.ref_transform_inv.argtypes = [ci, ci]
        L.ref_prob_new.restype = vp
        L.ref_prob_new.argtypes = [ctypes.c_uint64]
        L.ref_prob_free.argtypes = [vp]
        L.ref_prob_gumbel.restype = cf
        L.ref_prob_gumbel.argtypes = [vp]
        L.ref_prob_uniform.restype = cf
        L.ref_prob_uniform.argtypes = [vp]
        L.ref_prob_next.restype = ctypes.c_uint32
        L.ref_prob_next.argtypes = [vp]
        L.ref_prob_rand_range.argtypes = [vp, ci, ci]
        L.ref_prob_random_symmetry.argtypes = [vp]
        L.ref_softmax362.argtypes = [vp, vp]
        L.ref_gumbel_topk.argtypes = [vp, vp, vp, cf, ci, vp, vp]
        _REF = L
    return _REF


def P(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- numpy-level helpers over the C restatement ---------------------------------------------------
def load_go_features(feats: np.ndarray, version: int = 1):
    n = len(feats)
    npl, ns = (13, 7) if version == 0 else (15, 8)
    planes = np.empty((n, 19, 19, npl), dtype=np.float32)
    scalars = np.empty((n, ns), dtype=np.float32)
    feats = np.ascontiguousarray(feats)
    oracle().orc_load_go_features(P(feats), n, version, P(planes), P(scalars))
    return planes, scalars


def stones_with_liberties(boards: np.ndarray) -> np.ndarray:
    boards = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, 361)
    out = np.zeros((len(boards), 3, 361), dtype=np.int8)
    for b in range(len(boards)):
        for k in range(3):
            oracle().orc_stones_with_liberties(P(boards[b]), k + 1, P(out[b, k]))
    return out


def legal_mask_nohist(boards: np.ndarray, colors: np.ndarray, forbidden=None) -> np.ndarray:
    boards = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, 361)
    out = np.zeros((len(boards), 362), dtype=np.uint8)
    for b in range(len(boards)):
        fb = None if forbidden is None else P(np.ascontiguousarray(forbidden[b], dtype=np.int8))
        oracle().orc_legal_mask_nohist(P(boards[b]), int(colors[b]), fb, P(out[b]))
    return out


def gumbel_topk(state: int, logits: np.ndarray, legal: np.ndarray, noise_scaling: float, k: int):
    st = ctypes.c_uint64(state)
    moves = np.full(k, -1, dtype=np.int32)
    scores = np.zeros(k, dtype=np.float32)
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    legal = np.ascontiguousarray(legal, dtype=np.uint8)
    kv = oracle().orc_gumbel_topk(ctypes.byref(st), P(logits), P(legal), noise_scaling, k, P(moves), P(scores))
    return moves, scores, kv, st.value


def game_derive(moves: np.ndarray, num_moves: np.ndarray, colors: np.ndarray, forbidden=None, want_legal: bool = True):
    """oracle/features_oracle.c::orc_game_derive over a batch: (boards, laddered, legal | None, status)."""
    L = oracle()
    moves = np.ascontiguousarray(moves, dtype=np.int16)
    n = len(moves)
    boards = np.zeros((n, 361), dtype=np.int8)
    lad = np.zeros((n, 361), dtype=np.int8)
    legal = np.zeros((n, 362), dtype=np.uint8) if want_legal else None
    status = np.zeros(n, dtype=np.int32)
    for b in range(n):
        fb = np.ascontiguousarray(forbidden[b], dtype=np.int8) if forbidden is not None else None
        status[b] = L.orc_game_derive(P(moves[b]), int(num_moves[b]), P(fb) if fb is not None else None, int(colors[b]),
                                      P(boards[b]), P(lad[b]), P(legal[b]) if want_legal else None)
    return boards, lad, legal, status
