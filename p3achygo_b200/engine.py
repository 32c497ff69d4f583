"""Host-side mirror of the reference's engine interface for the B200 evaluator.

Same names, argument meaning and error behaviour as the reference's C++ ``nn::Engine``
(``cc/nn/engine/engine.h:22-43``) and factory (``cc/nn/engine/engine_factory.cc:16-73``), so the
parity tests read like the reference's own engine tests (``cc/nn/engine/benchmark_engine.cc:77-109``).
Everything computes in libp3b200.so through the C ABI (``include/p3_b200.h``); PyTorch is not involved.
The C++ adapter a reference maintainer would add (``B200Engine : nn::Engine``) is in INTEGRATION.md and
mirrored by ``p3achygo_b200/host/b200_engine.h``.
"""
from __future__ import annotations

import ctypes
import enum
import os
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (AUX_RESULT_DTYPE, GO_FEATURES_DTYPE, INFER_RESULT_DTYPE, LEAF_RESULT_DTYPE, NUM_LOCS, NUM_MOVES,
                   PRECISION_BF16, PRECISION_FP16, PRECISION_FP32, RESULT_FULL, RESULT_LEAF, P3Error, check, lib, ptr)


class Kind(enum.IntEnum):
    """``nn::Engine::Kind`` (engine.h:24-30) plus the one new kind this build adds."""
    kUnknown = 0
    kTrt = 1
    kTF = 2
    kTFTrt = 3
    kTFXla = 4
    kB200 = 5


def KindToString(kind: Kind) -> str:
    """engine.h:45-57"""
    return {Kind.kTrt: "TensorRT", Kind.kTF: "TF", Kind.kTFTrt: "TF-TRT", Kind.kTFXla: "TF-XLA",
            Kind.kB200: "B200"}.get(kind, "??")


def KindFromEnginePath(path: str) -> Kind:
    """engine_factory.cc:16-35, with ``.p3w`` (flat P3W1 weight file) -> kB200."""
    if os.path.isfile(path):
        ext = os.path.splitext(path)[1]
        if ext == ".trt":
            return Kind.kTrt
        if ext == ".pb":
            return Kind.kTFXla
        if ext == ".p3w":
            return Kind.kB200
        return Kind.kUnknown
    if os.path.basename(os.path.normpath(path)) == "_trt":
        return Kind.kTFTrt
    return Kind.kTF


def GetVersionFromModelPath(path: str) -> int:
    """engine_factory.cc:37-54: a ``VERSION`` file beside the model selects the feature version (default 1)."""
    parent = os.path.dirname(path) if os.path.isfile(path) else path
    vf = os.path.join(parent, "VERSION")
    if os.path.isfile(vf):
        try:
            with open(vf) as f:
                return int(f.read().split()[0])
        except (ValueError, IndexError):
            pass
    return 1


class B200Engine:
    """``nn::Engine`` implemented by libp3b200 (replaces ``TrtEngineImpl``, trt_engine.cc:37-84)."""

    def __init__(self, path: str, batch_size: int, version: int = 1, precision: Optional[int] = None, device: int = 0):
        if precision is None:
            precision = {"bf16": PRECISION_BF16, "fp16": PRECISION_FP16, "fp32": PRECISION_FP32}[os.environ.get("P3_PRECISION", "bf16")]
        self._h = ctypes.c_void_p()
        self._path = path
        self.batch_size = batch_size
        self.version = version
        self.precision = precision
        self.num_planes = 13 if version == 0 else 15
        self.num_scalars = 7 if version == 0 else 8
        check(lib.p3_engine_create(path.encode(), device, batch_size, version, precision, ctypes.byref(self._h)))

    # -- nn::Engine ---------------------------------------------------------------------------
    def kind(self) -> Kind:
        return Kind.kB200

    def path(self) -> str:
        return self._path

    def LoadBatch(self, batch_id: int, features) -> None:
        """engine.h:35. ``features``: a ``GoFeatures`` ctypes struct or a 1-record GO_FEATURES_DTYPE array."""
        if isinstance(features, np.ndarray) or isinstance(features, np.void):
            arr = np.ascontiguousarray(np.asarray(features, dtype=GO_FEATURES_DTYPE).reshape(1))
            check(lib.p3_engine_load_batch(self._h, batch_id, ptr(arr)))
        else:
            check(lib.p3_engine_load_batch(self._h, batch_id, ctypes.cast(ctypes.byref(features), ctypes.c_void_p)))

    def LoadBatchSym(self, batch_id: int, features, sym: int) -> None:
        """NNInterface::LoadBatch with the symmetry applied on the GPU (nn_interface.cc:245-277): ``features`` in the game's
        own orientation, ``sym`` = game::Symmetry 0..7; GetBatch then returns the un-rotated policies (nn_interface.h:263-287)."""
        arr = np.ascontiguousarray(np.asarray(features, dtype=GO_FEATURES_DTYPE).reshape(1))
        check(lib.p3_engine_load_batch_sym(self._h, batch_id, ptr(arr), int(sym)))

    def RunInference(self) -> None:
        """engine.h:36"""
        check(lib.p3_engine_run_inference(self._h))

    def GetBatch(self, batch_id: int) -> np.ndarray:
        """engine.h:37: returns one INFER_RESULT_DTYPE record (the caller-owned NNInferResult)."""
        out = np.zeros(1, dtype=INFER_RESULT_DTYPE)
        check(lib.p3_engine_get_batch(self._h, batch_id, ptr(out)))
        return out[0]

    # -- pipelined form: two slot banks (include/p3_b200.h) ---------------------------
    def LoadBatchBank(self, bank: int, batch_id: int, features, sym: int = 0) -> None:
        arr = np.ascontiguousarray(np.asarray(features, dtype=GO_FEATURES_DTYPE).reshape(1))
        check(lib.p3_engine_load_batch_bank(self._h, bank, batch_id, ptr(arr), int(sym)))

    def LoadGameBank(self, bank: int, batch_id: int, moves: np.ndarray, color: int, komi: float, forbidden=None, sym: int = 0) -> None:
        """NNInterface::LoadBatch from the game record (nn_interface.cc:245-277): the derived grids are computed on the GPU."""
        mv = np.ascontiguousarray(moves, dtype=np.int16)
        fb = np.ascontiguousarray(forbidden, dtype=np.int8) if forbidden is not None else None
        check(lib.p3_engine_load_game_bank(self._h, bank, batch_id, ptr(mv), len(mv), int(color), float(komi),
                                           ptr(fb) if fb is not None else None, int(sym)))

    def Submit(self, bank: int) -> None:
        """Asynchronous half of RunInference for one bank (H2D -> step -> D2H on copy streams); returns at once."""
        check(lib.p3_engine_submit(self._h, bank))

    def Wait(self, bank: int) -> None:
        check(lib.p3_engine_wait(self._h, bank))

    def GetBatchBank(self, bank: int, batch_id: int) -> np.ndarray:
        out = np.zeros(1, dtype=INFER_RESULT_DTYPE)
        check(lib.p3_engine_get_batch_bank(self._h, bank, batch_id, ptr(out)))
        return out[0]

    # -- compact leaf results: what mcts::LeafEvaluator InitFields keeps, 4360 B instead of 7568 B per leaf
    def SetResultMode(self, mode: int) -> None:
        check(lib.p3_engine_set_result_mode(self._h, int(mode)))

    def GetLeaf(self, batch_id: int) -> np.ndarray:
        out = np.zeros(1, dtype=LEAF_RESULT_DTYPE)
        check(lib.p3_engine_get_leaf(self._h, batch_id, ptr(out)))
        return out[0]

    def GetLeafBank(self, bank: int, batch_id: int) -> np.ndarray:
        out = np.zeros(1, dtype=LEAF_RESULT_DTYPE)
        check(lib.p3_engine_get_leaf_bank(self._h, bank, batch_id, ptr(out)))
        return out[0]

    def GetAuxBank(self, bank: int, batch_id: int) -> np.ndarray:
        out = np.zeros(1, dtype=AUX_RESULT_DTYPE)
        check(lib.p3_engine_get_aux_bank(self._h, bank, batch_id, ptr(out)))
        return out[0]

    def GetOwnershipBank(self, bank: int, batch_id: int) -> np.ndarray:
        own = np.zeros(NUM_LOCS, dtype=np.float32)
        check(lib.p3_engine_get_ownership_bank(self._h, bank, batch_id, ptr(own)))
        return own

    def GumbelTopKBank(self, bank: int, slots, legal: np.ndarray, prng_state: np.ndarray, noise_scaling: float, k: int):
        """Gumbel root sampling (cc/mcts/gumbel.cc:283-321) on move_logits still resident in HBM after ``bank``'s last run."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        legal = np.ascontiguousarray(legal, dtype=np.uint8).reshape(-1, NUM_MOVES)
        n = len(slots)
        assert prng_state.dtype == np.uint64 and prng_state.flags.c_contiguous and len(prng_state) == n and len(legal) == n
        moves = np.zeros((n, k), dtype=np.int32)
        scores = np.zeros((n, k), dtype=np.float32)
        kvalid = np.zeros(n, dtype=np.int32)
        check(lib.p3_engine_gumbel_topk_bank(self._h, bank, ptr(slots), n, ptr(legal), ptr(prng_state), noise_scaling, k,
                                             ptr(moves), ptr(scores), ptr(kvalid)))
        return moves, scores, kvalid

    def GetOwnership(self, batch_id: int) -> np.ndarray:
        """engine.h:38-39"""
        own = np.zeros(NUM_LOCS, dtype=np.float32)
        check(lib.p3_engine_get_ownership(self._h, batch_id, ptr(own)))
        return own

    # -- parity / measurement hooks -------------------------------------------------------------
    def LoadBatchAll(self, feats: np.ndarray) -> None:
        for b in range(min(len(feats), self.batch_size)):
            check(lib.p3_engine_load_batch(self._h, b, ctypes.c_void_p(feats[b:b + 1].ctypes.data)))

    def GetPlanes(self, batch_id: int):
        planes = np.zeros((19, 19, self.num_planes), dtype=np.float32)
        scalars = np.zeros(self.num_scalars, dtype=np.float32)
        check(lib.p3_engine_get_planes(self._h, batch_id, ptr(planes), ptr(scalars)))
        return planes, scalars

    def GetAux(self, batch_id: int) -> np.ndarray:
        out = np.zeros(1, dtype=AUX_RESULT_DTYPE)
        check(lib.p3_engine_get_aux(self._h, batch_id, ptr(out)))
        return out[0]

    def RunDevice(self) -> float:
        ms = ctypes.c_float()
        check(lib.p3_engine_run_device(self._h, ctypes.byref(ms)))
        return ms.value

    def Upload(self) -> None:
        check(lib.p3_engine_upload(self._h))

    KERNEL_CLASSES = ["encode", "init_conv", "conv1x1", "conv3x3", "broadcast", "head_conv", "heads", "boundary"]

    def Profile(self):
        """One eager pass with an event around every launch -> {class: (ms, launches, flops, algorithmic HBM bytes)}."""
        n = len(self.KERNEL_CLASSES)
        ms = np.zeros(n, dtype=np.float32)
        launches = np.zeros(n, dtype=np.int32)
        flops = np.zeros(n, dtype=np.float64)
        nbytes = np.zeros(n, dtype=np.float64)
        check(lib.p3_engine_profile(self._h, ptr(ms), ptr(launches), ptr(flops), ptr(nbytes)))
        return {k: (float(ms[i]), int(launches[i]), float(flops[i]), float(nbytes[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def RangeCheck(self):
        """(max |x| of the residual stream over all blocks, number of values at the fp16 clamp) for the resident inputs."""
        mx, sat = ctypes.c_float(), ctypes.c_longlong()
        check(lib.p3_engine_range_check(self._h, ctypes.byref(mx), ctypes.byref(sat)))
        return mx.value, sat.value

    def FirstLayer(self, C: int):
        """Test hook: (residual stream as float16, activated copy as raw uint16) of the first layer, each [batch, 400, C] (C = the net's trunk width)."""
        x = np.zeros((self.batch_size, 400, C), dtype=np.float16)
        a = np.zeros((self.batch_size, 400, C), dtype=np.uint16)
        check(lib.p3_engine_first_layer(self._h, ptr(x), ptr(a), x.nbytes))
        return x, a

    def StageMs(self):
        arr = (ctypes.c_float * 3)()
        check(lib.p3_engine_stage_ms(self._h, ctypes.byref(arr)))
        return list(arr)

    def launches_per_run(self) -> int:
        return lib.p3_engine_launches_per_run(self._h)

    def flops_per_position(self) -> float:
        return lib.p3_engine_flops_per_position(self._h)

    def set_cuda_graph(self, enabled: bool) -> None:
        check(lib.p3_engine_set_cuda_graph(self._h, int(enabled)))

    def close(self) -> None:
        if self._h:
            lib.p3_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def CreateEngine(kind: Kind, path: str, batch_size: int, version: int, **kw) -> B200Engine:
    """engine_factory.cc:56-73. Only the kind this build provides can be constructed here."""
    if kind == Kind.kB200:
        return B200Engine(path, batch_size, version, **kw)
    raise RuntimeError("Unknown Engine Kind.")  # LOG(FATAL) in the reference


# ---- stand-alone kernels ----------------------------------------------------------------------
def encode_features(feats: np.ndarray, version: int = 1, device: int = 0):
    feats = np.ascontiguousarray(feats, dtype=GO_FEATURES_DTYPE)
    n = len(feats)
    npl, ns = (13, 7) if version == 0 else (15, 8)
    planes = np.zeros((n, 19, 19, npl), dtype=np.float32)
    scalars = np.zeros((n, ns), dtype=np.float32)
    check(lib.p3_encode_features(device, ptr(feats), n, version, ptr(planes), ptr(scalars)))
    return planes, scalars


def board_liberties(boards: np.ndarray, device: int = 0) -> np.ndarray:
    boards = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, NUM_LOCS)
    out = np.zeros((len(boards), 3, NUM_LOCS), dtype=np.int8)
    check(lib.p3_board_liberties(device, ptr(boards), len(boards), ptr(out)))
    return out


def legal_mask(boards: np.ndarray, colors: np.ndarray, forbidden: Optional[np.ndarray] = None, device: int = 0):
    boards = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, NUM_LOCS)
    colors = np.ascontiguousarray(colors, dtype=np.int8)
    out = np.zeros((len(boards), NUM_MOVES), dtype=np.uint8)
    fb = None
    if forbidden is not None:
        forbidden = np.ascontiguousarray(forbidden, dtype=np.int8).reshape(-1, NUM_LOCS)
        fb = ptr(forbidden)
    check(lib.p3_legal_mask(device, ptr(boards), ptr(colors), fb, len(boards), ptr(out)))
    return out


MOVE_WHITE = 512


def game_derive(moves: np.ndarray, num_moves: np.ndarray, colors: Optional[np.ndarray] = None, forbidden: Optional[np.ndarray] = None,
                want_ladder: bool = True, device: int = 0):
    """Replay move lists on the GPU -> (boards [n,361] i8, laddered [n,361] i8 | None, legal [n,362] u8 | None, status [n] i32).
    Board::GetLadderedStones (cc/game/board.cc:692-899) and Game::IsValidMove incl. superko (board.cc:595-644)."""
    moves = np.ascontiguousarray(moves, dtype=np.int16)
    n, max_moves = moves.shape
    num_moves = np.ascontiguousarray(num_moves, dtype=np.int32)
    boards = np.zeros((n, NUM_LOCS), dtype=np.int8)
    laddered = np.zeros((n, NUM_LOCS), dtype=np.int8) if want_ladder else None
    legal = np.zeros((n, 362), dtype=np.uint8) if colors is not None else None
    status = np.zeros(n, dtype=np.int32)
    cols = np.ascontiguousarray(colors, dtype=np.int8) if colors is not None else None
    fb = np.ascontiguousarray(forbidden, dtype=np.int8) if forbidden is not None else None
    check(lib.p3_game_derive(device, ptr(moves), ptr(num_moves), max_moves, ptr(fb) if fb is not None else None,
                             ptr(cols) if cols is not None else None, n, ptr(boards), ptr(laddered) if want_ladder else None,
                             ptr(legal) if legal is not None else None, ptr(status)))
    return boards, laddered, legal, status


def gumbel_topk(logits: np.ndarray, legal: np.ndarray, prng_state: np.ndarray, noise_scaling: float, k: int,
                device: int = 0):
    """cc/mcts/gumbel.cc:283-321 for n roots; ``prng_state`` (uint64[n]) is updated in place."""
    logits = np.ascontiguousarray(logits, dtype=np.float32).reshape(-1, NUM_MOVES)
    legal = np.ascontiguousarray(legal, dtype=np.uint8).reshape(-1, NUM_MOVES)
    n = len(logits)
    assert prng_state.dtype == np.uint64 and prng_state.flags.c_contiguous and len(prng_state) == n
    moves = np.zeros((n, k), dtype=np.int32)
    scores = np.zeros((n, k), dtype=np.float32)
    kvalid = np.zeros(n, dtype=np.int32)
    check(lib.p3_gumbel_topk(device, ptr(logits), ptr(legal), ptr(prng_state), n, noise_scaling, k, ptr(moves),
                             ptr(scores), ptr(kvalid)))
    return moves, scores, kvalid


def conv_test(x: np.ndarray, w: np.ndarray, precision: int, device: int = 0) -> np.ndarray:
    """x [n,361,cin], w OIHW [cout,cin,k,k] -> y [n,361,cout] (no BN / activation)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    n, _, cin = x.shape
    cout, _, k, _ = w.shape
    y = np.zeros((n, NUM_LOCS, cout), dtype=np.float32)
    check(lib.p3_conv_test(device, precision, ptr(x), ptr(w), n, cin, cout, k, ptr(y)))
    return y


def broadcast_test(x: np.ndarray, w: np.ndarray, bias: np.ndarray, precision: int, device: int = 0) -> np.ndarray:
    """x [n,361,C], w [361,361] (in, out), bias [361] -> mish(W^T x + bias) [n,361,C]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    bias = np.ascontiguousarray(bias, dtype=np.float32)
    n, _, C = x.shape
    y = np.zeros((n, NUM_LOCS, C), dtype=np.float32)
    check(lib.p3_broadcast_test(device, precision, ptr(x), ptr(w), ptr(bias), n, C, ptr(y)))
    return y


def block_boundary_test(t: np.ndarray, x: np.ndarray, w1: np.ndarray, w2: np.ndarray, scale1, shift1, scale2, shift2,
                        fused: bool, device: int = 0):
    """Expand 1x1 + residual -> BN/mish -> reduce 1x1 -> BN/mish of the bf16 engine (p3_block_boundary_test).
    t [n,361,k1], x [n,361,n1], w1 [n1,k1], w2 [n2,n1]; returns (x' [n,361,n1], out [n,361,n2])."""
    t = np.ascontiguousarray(t, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    w1 = np.ascontiguousarray(w1, dtype=np.float32)
    w2 = np.ascontiguousarray(w2, dtype=np.float32)
    vecs = [np.ascontiguousarray(v, dtype=np.float32) for v in (scale1, shift1, scale2, shift2)]
    n, _, k1 = t.shape
    n1, n2 = w1.shape[0], w2.shape[0]
    xprime = np.empty((n, 361, n1), dtype=np.float32)
    out = np.empty((n, 361, n2), dtype=np.float32)
    check(lib.p3_block_boundary_test(device, 1 if fused else 0, ptr(t), ptr(x), ptr(w1), ptr(w2), ptr(vecs[0]), ptr(vecs[1]),
                                     ptr(vecs[2]), ptr(vecs[3]), n, k1, n1, n2, ptr(xprime), ptr(out)))
    return xprime, out

