"""GPU parity of the whole leaf evaluation (encode -> tower -> heads) through the reference-shaped engine API
(LoadBatch / RunInference / GetBatch, cc/nn/engine/engine.h:22-43), against the PyTorch restatement of
python/model.py (oracle/model_ref.py) on the same seeded positions and weights.

Tolerances (north star): fp32 mode max-abs 1e-3 on logits / probabilities / value; bf16 mode: documented bound below.
"""
import numpy as np
import pytest
import torch

from oracle import oracle_lib
from oracle.model_ref import RefModel

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-3
# bf16 operands (8-bit mantissa) through up to 15 residual blocks with an fp32 residual stream and fp32 accumulation.
# Measured on B200 (round 1): logits <= 2.4e-2, probabilities <= 1e-4, value_probs <= 9e-4, E[score] <= 0.6 points.
BF16_TOL = {"logits": 6e-2, "probs": 1.5e-2, "value": 2e-2, "score_mean": 2.0}


def _same(a, b):
    """Bit-exact equality of two NNInferResult records, field by field (the struct has 8 alignment-padding bytes)."""
    return all(np.array_equal(np.asarray(a[f]), np.asarray(b[f])) for f in a.dtype.names)


def _oracle_outputs(cfg, tensors, feats, dtype=torch.float32):
    planes, scalars = oracle_lib.load_go_features(feats, 1)
    return RefModel(cfg, tensors, dtype=dtype).forward(planes, scalars)


def _run_engine(path, feats, precision, batch=None):
    from p3achygo_b200 import engine as E
    batch = batch or len(feats)
    eng = E.CreateEngine(E.Kind.kB200, path, batch, 1, precision=precision)
    for b in range(len(feats)):
        eng.LoadBatch(b, feats[b])
    eng.RunInference()
    res = [eng.GetBatch(b) for b in range(len(feats))]
    aux = [eng.GetAux(b) for b in range(len(feats))]
    return eng, res, aux


def _compare(res, aux, o, tol_logits, tol_probs, tol_value, tol_smean):
    n = len(res)
    stack = lambda key, src: np.stack([np.asarray(r[key], dtype=np.float64) for r in src])
    checks = [
        ("move_logits", stack("move_logits", res), o["pi_logits"], tol_logits),
        ("move_probs", stack("move_probs", res), o["pi"], tol_probs),
        ("value_probs", stack("value_probs", res), o["outcome"], tol_value),
        ("score_probs", stack("score_probs", res), o["score_probs"], tol_probs),
        ("opt_move_probs", stack("opt_move_probs", res), o["opt_move_probs"], tol_probs),
        ("err2_outcome", stack("err2_outcome", res), o["q_err"][:, 0], tol_logits),
        ("pi_logits_aux", stack("pi_logits_aux", aux), o["pi_logits_aux"], tol_logits),
        ("pi_logits_soft", stack("pi_logits_soft", aux), o["pi_logits_soft"], tol_logits),
        ("pi_logits_optimistic", stack("pi_logits_optimistic", aux), o["pi_logits_optimistic"], tol_logits),
        ("outcome_logits", stack("outcome_logits", aux), o["outcome_logits"], tol_logits),
        ("score_logits", stack("score_logits", aux), o["score_logits"], tol_logits * 4),
        ("gamma", stack("gamma", aux), o["gamma"], tol_logits),
        ("q", stack("q", aux), o["q"], tol_value),
        ("q_score", stack("q_score", aux), o["q_score"], tol_logits),
        ("mcts_dist_probs", stack("mcts_dist_probs", aux), o["mcts_dist_probs"], tol_probs),
        ("ownership", stack("ownership", aux), o["own"], tol_value),
        ("value", stack("value", aux), o["value"], 2 * tol_value),
        ("score_mean", stack("score_mean", aux), o["score_mean"], tol_smean),
        # outputs no round-1 test compared (python/model.py:955-963, :903; cc/mcts/leaf_evaluator.cc:95-107)
        ("q_err", stack("q_err", aux), o["q_err"], tol_logits),
        ("q_score_err", stack("q_score_err", aux), o["q_score_err"], tol_logits),
        ("mcts_dist_logits", stack("mcts_dist_logits", aux), o["mcts_dist_logits"], tol_logits * 4),
    ]
    # Var[score] = E[s^2] - E[s]^2 over 800 bins of +-400 points: compared relative to E[s^2] (the cancellation's scale)
    sv = stack("score_var", aux).reshape(n)
    es2 = np.asarray(o["score_var"], dtype=np.float64).reshape(n) + np.asarray(o["score_mean"], dtype=np.float64).reshape(n) ** 2
    sv_err = float((np.abs(sv - np.asarray(o["score_var"], dtype=np.float64).reshape(n)) / np.maximum(es2, 1.0)).max())
    worst = {}
    for name, got, exp, tol in checks:
        err = float(np.abs(got.reshape(n, -1) - np.asarray(exp).reshape(n, -1)).max())
        worst[name] = (err, tol)
    worst["score_var(rel E[s^2])"] = (sv_err, 50 * tol_probs)
    bad = {k: v for k, v in worst.items() if not (v[0] <= v[1])}
    assert not bad, f"out of tolerance: {bad}\nall: {worst}"
    return worst


@pytest.mark.parametrize("config,n", [("tiny", 32), ("small", 32), ("b10c128btl3", 6), ("b12c256btl3", 4), ("b14c384btl3", 2),
                                      ("b15c192_classic", 3), ("b8c128nbt", 4)])
def test_fp32_engine_matches_oracle(config, n, weight_dir, golden_positions):
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"][10:10 + n]
    o = _oracle_outputs(cfg, tensors, feats, torch.float64)
    eng, res, aux = _run_engine(path, feats, E.PRECISION_FP32)
    worst = _compare(res, aux, o, FP32_TOL, FP32_TOL, FP32_TOL, 5e-2)
    print(config, "fp32 worst errors:", {k: f"{v[0]:.2e}" for k, v in worst.items()})
    # planes produced inside the engine are the reference planes, bit for bit
    oplanes, oscalars = oracle_lib.load_go_features(feats, 1)
    for b in range(n):
        planes, scalars = eng.GetPlanes(b)
        assert np.array_equal(planes, oplanes[b]) and np.array_equal(scalars, oscalars[b])
    own = eng.GetOwnership(1)
    assert np.array_equal(own, np.asarray(aux[1]["ownership"]))
    eng.close()


@pytest.mark.parametrize("config,n", [("small", 32), ("b10c128btl3", 6), ("b12c256btl3", 4), ("b14c384btl3", 2), ("b15c192_classic", 3),
                                      ("b8c128nbt", 4), ("b12c256nbt", 2)])
def test_bf16_engine_within_documented_bound(config, n, weight_dir, golden_positions):
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"][100:100 + n]
    o = _oracle_outputs(cfg, tensors, feats, torch.float64)
    eng, res, aux = _run_engine(path, feats, E.PRECISION_BF16)
    worst = _compare(res, aux, o, BF16_TOL["logits"], BF16_TOL["probs"], BF16_TOL["value"], BF16_TOL["score_mean"])
    print(config, "bf16 worst errors:", {k: f"{v[0]:.2e}" for k, v in worst.items()})
    eng.close()


# IEEE fp16 operands (11-bit mantissa, tcgen05 kind::f16 at the bf16 rate): the reference's production precision is whole-graph fp16
# (python/rl_loop/model_utils.py:181).  Documented bound: 1/5 of the bf16 one (three more mantissa bits in the operands; the fp16
# residual stream is common to both).  Measured on B200 (round 2): logits <= 8.6e-3 (3e-3 .. 6e-3 typical), probabilities <= 6e-5,
# value_probs <= 1.3e-3, E[score] <= 0.52 points.
FP16_TOL = {"logits": 1.2e-2, "probs": 3e-3, "value": 4e-3, "score_mean": 0.75}


@pytest.mark.parametrize("config,n", [("small", 32), ("b10c128btl3", 6), ("b12c256btl3", 4), ("b14c384btl3", 2), ("b15c192_classic", 3),
                                      ("b8c128nbt", 4), ("b12c256nbt", 2)])
def test_fp16_engine_within_documented_bound(config, n, weight_dir, golden_positions):
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"][100:100 + n]
    o = _oracle_outputs(cfg, tensors, feats, torch.float64)
    eng, res, aux = _run_engine(path, feats, E.PRECISION_FP16)
    worst = _compare(res, aux, o, FP16_TOL["logits"], FP16_TOL["probs"], FP16_TOL["value"], FP16_TOL["score_mean"])
    print(config, "fp16 worst errors:", {k: f"{v[0]:.2e}" for k, v in worst.items()})
    eng.close()


def test_bf16_unsupported_for_tiny(weight_dir):
    from p3achygo_b200 import engine as E
    path, _, _ = weight_dir("tiny")
    with pytest.raises(E.P3Error) as ei:
        E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_BF16)
    assert ei.value.code == E._lib.P3_ERR_UNSUPPORTED


@pytest.mark.parametrize("precision_name", ["fp32", "bf16", "fp16"])
def test_engine_contract(precision_name, weight_dir, golden_positions):
    """Slot semantics of nn::Engine: full batch every run, stale slots tolerated, outputs of a slot stay
    intact until the next RunInference, results independent of slot index and of other slots' contents, CUDA-graph
    replay == plain launches, deterministic across runs."""
    from p3achygo_b200 import engine as E
    prec = {"fp32": E.PRECISION_FP32, "bf16": E.PRECISION_BF16, "fp16": E.PRECISION_FP16}[precision_name]
    path, cfg, tensors = weight_dir("b10c128btl3")
    feats = golden_positions["feats"][:8]
    B = 8
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=prec)
    eng.RunInference()                      # nothing loaded yet: must not fail (benchmark_engine.cc:79-82 warm-up)
    for b in range(B):
        eng.LoadBatch(b, feats[b])
    eng.RunInference()
    first = [eng.GetBatch(b).copy() for b in range(B)]
    again = [eng.GetBatch(b).copy() for b in range(B)]
    for a, c in zip(first, again):
        assert _same(a, c)
    eng.RunInference()                      # deterministic
    for b in range(B):
        assert _same(eng.GetBatch(b), first[b])
    eng.set_cuda_graph(False)               # graph replay == eager launches
    eng.RunInference()
    for b in range(B):
        assert _same(eng.GetBatch(b), first[b])
    eng.set_cuda_graph(True)
    # permute slots: position p evaluated in slot (p+3)%B gives the same bits (slots are independent)
    for b in range(B):
        eng.LoadBatch((b + 3) % B, feats[b])
    eng.RunInference()
    for b in range(B):
        assert _same(eng.GetBatch((b + 3) % B), first[b])
    # partial reload: only slot 2 changes, the others keep their (stale) inputs and outputs
    eng.LoadBatch(2, feats[7])
    eng.RunInference()
    assert _same(eng.GetBatch(2), first[7])
    assert _same(eng.GetBatch(5), first[2])
    for r in first:
        assert abs(float(np.sum(r["move_probs"])) - 1.0) < 1e-4 and abs(float(np.sum(r["score_probs"])) - 1.0) < 1e-4
        assert abs(float(np.sum(r["opt_move_probs"])) - 1.0) < 1e-4 and abs(float(np.sum(r["value_probs"])) - 1.0) < 1e-5
        assert 0.0 <= float(r["err2_outcome"]) <= 4.0
    eng.close()


def test_full_size_properties(weight_dir, golden_positions):
    """BASELINE size (b12c256btl3, batch 1024) where the oracle is too slow: size-independent properties.
    (a) a batch of 1024 equals the same positions evaluated in batches of 4 (slot independence, bit-exact);
    (b) the bf16 result stays within the documented bound of the fp32 engine on a sample."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir("b12c256btl3")
    feats = golden_positions["feats"]
    B = 1024
    big = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    big.LoadBatchAll(feats[:B])
    big.RunInference()
    sample = [0, 1, 2, 3, 511, 512, 1020, 1021, 1022, 1023]
    big_res = {b: big.GetBatch(b).copy() for b in sample}
    ms = big.RunDevice()
    assert ms > 0
    big.close()
    small = E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_BF16)
    ref32 = E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_FP32)
    for grp in (sample[:4], sample[4:8], sample[6:10]):
        for s, b in enumerate(grp):
            small.LoadBatch(s, feats[b])
            ref32.LoadBatch(s, feats[b])
        small.RunInference()
        ref32.RunInference()
        for s, b in enumerate(grp):
            assert _same(small.GetBatch(s), big_res[b])
            r32 = ref32.GetBatch(s)
            assert np.abs(np.asarray(r32["move_logits"]) - np.asarray(big_res[b]["move_logits"])).max() < BF16_TOL["logits"]
            assert np.abs(np.asarray(r32["value_probs"]) - np.asarray(big_res[b]["value_probs"])).max() < BF16_TOL["value"]
    small.close()
    ref32.close()


# ---- the engine against the reference's own model code (tests/golden/model_golden.npz, see tests/test_model_golden.py)
_GOLDEN_MAP = [("move_logits", "res", "pi_logits"), ("move_probs", "res", "pi"), ("value_probs", "res", "outcome"),
               ("score_probs", "res", "score_probs"), ("pi_logits_aux", "aux", "pi_logits_aux"),
               ("pi_logits_soft", "aux", "pi_logits_soft"), ("pi_logits_optimistic", "aux", "pi_logits_optimistic"),
               ("outcome_logits", "aux", "outcome_logits"), ("score_logits", "aux", "score_logits"), ("gamma", "aux", "gamma"),
               ("mcts_dist_probs", "aux", "mcts_dist_probs"), ("ownership", "aux", "own"),
               ("mcts_dist_logits", "aux", "mcts_dist_logits")]


@pytest.mark.parametrize("precision_name", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("config", ["tiny", "small", "b10c128btl3", "b12c256btl3", "b14c384btl3", "b15c192_classic", "b8c128nbt"])
def test_engine_matches_reference_model_golden(config, precision_name, weight_dir, golden_positions):
    """The CUDA engine on the positions / weights of the fixture produced by the reference's unmodified python/model.py
    (float64, on oracle/tf_shim).  fp32 engine: max-abs 1e-3 (north star); bf16 engine: the documented bound."""
    import os
    from p3achygo_b200 import engine as E
    if config == "tiny" and precision_name != "fp32":
        pytest.skip("tiny (C=16) is below the tensor-core tile sizes; the tensor engines report UNSUPPORTED (tested above)")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_golden.npz"))
    first, n = (int(v) for v in z[f"{config}/first"])
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"][first:first + n]
    prec = {"fp32": E.PRECISION_FP32, "bf16": E.PRECISION_BF16, "fp16": E.PRECISION_FP16}[precision_name]
    eng, res, aux = _run_engine(path, feats, prec)
    tl, tp, tv = {"fp32": (FP32_TOL, FP32_TOL, FP32_TOL), "bf16": (BF16_TOL["logits"], BF16_TOL["probs"], BF16_TOL["value"]),
                  "fp16": (FP16_TOL["logits"], FP16_TOL["probs"], FP16_TOL["value"])}[precision_name]
    tol_of = {"move_logits": tl, "pi_logits_aux": tl, "pi_logits_soft": tl, "pi_logits_optimistic": tl, "outcome_logits": tl,
              "score_logits": 4 * tl, "gamma": tl, "move_probs": tp, "score_probs": tp, "mcts_dist_probs": tp, "value_probs": tv,
              "ownership": tv, "mcts_dist_logits": 4 * tl}
    worst = {}
    for field, src, gname in _GOLDEN_MAP:
        got = np.stack([np.asarray((res if src == "res" else aux)[b][field], dtype=np.float64) for b in range(n)]).reshape(n, -1)
        ref = z[f"{config}/{gname}"].reshape(n, -1)
        worst[field] = (float(np.abs(got - ref).max()), tol_of[field])
    bad = {k: v for k, v in worst.items() if not v[0] <= v[1]}
    assert not bad, f"out of tolerance vs the reference model code: {bad}\nall: {worst}"
    for field, suffix, tol in (("q", "", tv), ("q_err", "_err", tl), ("q_score", "_score", tl), ("q_score_err", "_score_err", tl)):
        q = np.stack([np.asarray(aux[b][field], dtype=np.float64) for b in range(n)])
        for k, nm in enumerate(("q6", "q16", "q50")):
            err = float(np.abs(q[:, k] - z[f"{config}/{nm}{suffix}"].reshape(n)).max())
            worst[f"{nm}{suffix}"] = (err, tol)
            assert err <= tol, (field, nm, err)
    e2 = np.array([float(res[b]["err2_outcome"]) for b in range(n)])
    assert np.abs(e2 - z[f"{config}/q6_err"].reshape(n)).max() <= tl   # 12:q6_err -> err2_outcome (cc/nn/engine/trt_names.h:19)
    print(config, precision_name, "worst vs reference model code:", {k: f"{v[0]:.2e}" for k, v in worst.items()})
    eng.close()


@pytest.mark.parametrize("precision_name", ["fp32", "bf16"])
def test_symmetry_on_gpu_equals_host_symmetry(precision_name, weight_dir, golden_positions):
    """p3_engine_load_batch_sym (symmetry applied by the encode kernel, un-applied by the heads kernel) against the
    reference's host-side handling restated with the pinned oracle: ApplySymmetry on board / derived grids / last moves
    before LoadBatch (cc/nn/nn_interface.cc:245-277), ApplyInverse on the 361 board entries of move_logits, move_probs and
    opt_move_probs after GetBatch (cc/nn/nn_interface.h:263-287).  Same planes bit for bit, same results bit for bit."""
    import ctypes
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    L = oracle_lib.oracle()
    prec = E.PRECISION_FP32 if precision_name == "fp32" else E.PRECISION_BF16
    path, cfg, tensors = weight_dir("b10c128btl3")
    feats = golden_positions["feats"][40:48]
    B = 8
    host = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=prec)
    dev = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=prec)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for b in range(B):
        sym = b  # all eight symmetries
        f = np.array(feats[b:b + 1], dtype=GO_FEATURES_DTYPE)
        fs = f.copy()
        for grid in ("board", "stones_atari", "stones_two_liberties", "stones_three_liberties", "stones_laddered"):
            src = np.ascontiguousarray(f[grid][0], dtype=np.int8)
            dst = np.zeros(361, dtype=np.int8)
            L.orc_apply_symmetry_i8(sym, vp(src), vp(dst))
            fs[grid][0] = dst
        for k in range(5):
            i, j = (int(v) for v in f["last_moves"][0][k])
            if 0 <= i < 19 and 0 <= j < 19:  # passes {19,0} and no-ops {-1,-1} are not transformed
                t = L.orc_transform_index(sym, i * 19 + j)
                fs["last_moves"][0][k] = (t // 19, t % 19)
        host.LoadBatch(b, fs[0])
        dev.LoadBatchSym(b, f[0], sym)
    host.RunInference()
    dev.RunInference()
    for b in range(B):
        ph, sh = host.GetPlanes(b)
        pd, sd = dev.GetPlanes(b)
        assert np.array_equal(ph, pd) and np.array_equal(sh, sd), f"planes differ for symmetry {b}"
        rh, rd = host.GetBatch(b), dev.GetBatch(b)
        for field in ("move_logits", "move_probs", "opt_move_probs"):
            src = np.ascontiguousarray(np.asarray(rh[field], dtype=np.float32)[:361])
            inv = np.zeros(361, dtype=np.float32)
            L.orc_apply_inverse_f32(b, vp(src), vp(inv))
            got = np.asarray(rd[field], dtype=np.float32)
            assert np.array_equal(got[:361], inv), f"{field} differs for symmetry {b}"
            assert got[361] == np.asarray(rh[field], dtype=np.float32)[361]
        for field in ("value_probs", "score_probs", "err2_outcome"):
            assert np.array_equal(np.asarray(rh[field]), np.asarray(rd[field]))
    host.close()
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision_name", ["fp32", "bf16"])
def test_slot_banks_equal_serial_calls(precision_name, weight_dir, golden_positions):
    """The pipelined form (p3_engine_submit / p3_engine_wait over two slot banks) returns, bit for bit, what
    LoadBatch -> RunInference -> GetBatch returns for the same positions - with both banks in flight at once, over several
    rounds, with the symmetry on the GPU, and a bank cannot be submitted twice without a wait."""
    from p3achygo_b200 import engine as E
    prec = E.PRECISION_FP32 if precision_name == "fp32" else E.PRECISION_BF16
    path, cfg, tensors = weight_dir("b10c128btl3")
    B, rounds = 8, 3
    feats = golden_positions["feats"][:2 * B * rounds]
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=prec)
    want = []
    for lo in range(0, len(feats), B):
        for b in range(B):
            eng.LoadBatchSym(b, feats[lo + b], (lo + b) % 8)
        eng.RunInference()
        want.append([eng.GetBatch(b).copy() for b in range(B)])
    with pytest.raises(E.P3Error):
        eng.Wait(0)                          # nothing submitted
    for r in range(rounds):
        for bank in (0, 1):
            lo = (2 * r + bank) * B
            for b in range(B):
                eng.LoadBatchBank(bank, b, feats[lo + b], (lo + b) % 8)
            eng.Submit(bank)                 # bank 0 and bank 1 are in flight together
        with pytest.raises(E.P3Error):
            eng.Submit(1)                    # not waited yet
        for bank in (1, 0):                  # wait order is free
            eng.Wait(bank)
            for b in range(B):
                assert _same(eng.GetBatchBank(bank, b), want[2 * r + bank][b]), (r, bank, b)
    # the serial calls still work afterwards and share bank 0's slots
    for b in range(B):
        eng.LoadBatchSym(b, feats[b], b % 8)
    eng.RunInference()
    for b in range(B):
        assert _same(eng.GetBatch(b), want[0][b])
    eng.close()


@pytest.mark.gpu
def test_garbage_slots_do_not_fault(weight_dir, golden_positions):
    """LoadBatch may overlap RunInference (cc/nn/nn_interface.cc:276), so the H2D copy can pick up a half-written slot; the
    reference just does not mark such a slot ready (nn_interface.cc:355-369).  Whatever bytes a slot holds, the step must run
    to completion and leave the other slots' results untouched."""
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    path, cfg, tensors = weight_dir("b10c128btl3")
    B = 8
    feats = golden_positions["feats"][:B]
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    for b in range(B):
        eng.LoadBatch(b, feats[b])
    eng.RunInference()
    want = [eng.GetBatch(b).copy() for b in range(B)]
    rng = np.random.default_rng(5)
    junk = rng.integers(0, 256, size=(4, 1860), dtype=np.uint8).view(GO_FEATURES_DTYPE).reshape(-1)
    for k, b in enumerate((1, 3, 4, 6)):
        eng.LoadBatch(b, junk[k])
    eng.RunInference()
    for b in (0, 2, 5, 7):
        assert _same(eng.GetBatch(b), want[b])
    for b in (1, 3, 4, 6):
        eng.LoadBatch(b, feats[b])
    eng.RunInference()
    for b in range(B):
        assert _same(eng.GetBatch(b), want[b])
    eng.close()


# ---- batch independence at the OTHER BASELINE configurations: each takes a different tile / plan path (N = 96 pair tile for
# C = 128, C = 384 off the fused boundary kernel, classic blocks on the residual 3x3 kernel) -----------------------------------
@pytest.mark.parametrize("config,B", [("b10c128btl3", 256), ("b14c384btl3", 2048), ("b15c192_classic", 1024), ("small", 32)])
def test_full_size_batch_independence_other_configs(config, B, weight_dir, golden_positions):
    """BASELINE.json's other configurations at their full batch sizes: a position's result does not depend on the batch it is
    evaluated in, nor on its slot (bit-exact against the same positions in a batch of 4), and stays within the documented
    bf16 bound of the fp32 engine."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"]
    reps = (B + len(feats) - 1) // len(feats)
    # (np.concatenate would re-pack the 1860-byte records of the offset dtype: tile the raw bytes instead)
    raw = np.ascontiguousarray(feats).view(np.uint8).reshape(len(feats), feats.dtype.itemsize)
    allf = np.ascontiguousarray(np.tile(raw, (reps, 1))[:B]).reshape(-1).view(feats.dtype)
    big = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    big.LoadBatchAll(allf)
    big.RunInference()
    sample = sorted({0, 1, 2, 3, B // 2 - 1, B // 2, B - 3, B - 2, B - 1, B // 3, (2 * B) // 3, B // 5})
    big_res = {b: big.GetBatch(b).copy() for b in sample}
    # the golden positions repeat with period len(feats): EVERY slot must equal the slot one period earlier, bit for bit
    # (a race between a CTA's consecutive positions shows up here, not in a 12-slot sample)
    P = len(feats)
    logits = np.stack([np.asarray(big.GetBatch(b)["move_logits"]) for b in range(B)])
    value = np.stack([np.asarray(big.GetBatch(b)["value_probs"]) for b in range(B)])
    for b in range(P, B):
        assert np.array_equal(logits[b], logits[b - P]) and np.array_equal(value[b], value[b - P]), (config, b)
    big.close()
    small = E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_BF16)
    ref32 = E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_FP32)
    for lo in range(0, len(sample), 4):
        grp = sample[lo:lo + 4]
        for s, b in enumerate(grp):
            small.LoadBatch(s, allf[b])
            ref32.LoadBatch(s, allf[b])
        small.RunInference()
        ref32.RunInference()
        for s, b in enumerate(grp):
            assert _same(small.GetBatch(s), big_res[b]), (config, b)
            r32 = ref32.GetBatch(s)
            assert np.abs(np.asarray(r32["move_logits"]) - np.asarray(big_res[b]["move_logits"])).max() < BF16_TOL["logits"]
            assert np.abs(np.asarray(r32["value_probs"]) - np.asarray(big_res[b]["value_probs"])).max() < BF16_TOL["value"]
    small.close()
    ref32.close()


# ---- the first layer alone (init_tc2.cu): against the fp32 restatement, and bit-identical to the explicit-im2col kernel ------------
@pytest.mark.parametrize("config,B", [("b12c256btl3", 600), ("b15c192_classic", 450), ("b14c384btl3", 330), ("b10c128btl3", 700)])
def test_first_layer_against_oracle_and_explicit_im2col_kernel(config, B, weight_dir, golden_positions, monkeypatch):
    """python/model.py:1230-1237 (conv5x5 of the planes + dense of the game state) through p3_engine_first_layer, several
    positions per persistent CTA: the fp16 residual stream is within operand rounding of the fp32 oracle, the activated copy's
    padding rows / columns are zero (the 3x3 kernels read them as the conv's zero padding), and the shifted-view kernel
    (init_tc2.cu) equals the explicit-im2col one (init_tc.cu, P3_INIT_TC=1) bit for bit on every board point."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"]
    reps = (B + len(feats) - 1) // len(feats)
    raw = np.ascontiguousarray(feats).view(np.uint8).reshape(len(feats), feats.dtype.itemsize)
    allf = np.ascontiguousarray(np.tile(raw, (reps, 1))[:B]).reshape(-1).view(feats.dtype)
    C = cfg.channels
    live = np.array([20 + r * 20 + c for r in range(19) for c in range(19)])
    pad = np.setdiff1d(np.arange(400), live)
    out = {}
    for form in ("2", "1"):
        monkeypatch.setenv("P3_INIT_TC", form)
        eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
        eng.LoadBatchAll(allf)
        eng.RunInference()  # leaves the tower's own traffic in both buffers first
        out[form] = eng.FirstLayer(C)
        eng.close()
    x2, a2 = out["2"]
    x1, a1 = out["1"]
    assert np.array_equal(x2[:, live], x1[:, live]) and np.array_equal(a2[:, live], a1[:, live])
    assert not a2[:, pad].any() and not a1[:, pad].any()
    assert not x2[:, pad].any()
    # fp32 restatement on the first len(feats) positions (the rest repeat them)
    planes, scalars = oracle_lib.load_go_features(feats, 1)
    m = RefModel(cfg, tensors)
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(planes)).float().permute(0, 3, 1, 2)
        ref = m.conv("model/init_conv", x) + m.dense("model/init_game_state", torch.from_numpy(np.ascontiguousarray(scalars)).float())[:, :, None, None]
        ref = ref.permute(0, 2, 3, 1).reshape(len(feats), 361, C).numpy()
    got = x2[:, live].astype(np.float32)
    for b in range(B):
        r = ref[b % len(feats)]
        assert np.abs(got[b] - r).max() <= 2e-2 * max(1.0, np.abs(r).max()), (config, b)


def test_stream_padding_rows_stay_zero_over_repeated_runs(weight_dir, golden_positions):
    """The first layer writes board points only (init_tc2.cu): nothing may leave values in the residual stream's padding rows /
    columns, or they would pile up run after run (the 3x3 residual kernel's outputs there are not zero by themselves) and trip
    the range check of a healthy net.  Classic blocks = the 3x3 kernel with a residual."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir("b15c192_classic")
    feats = golden_positions["feats"]
    B = 160
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    eng.LoadBatchAll(feats[np.arange(B) % len(feats)])
    eng.RunInference()
    mx0, sat0 = eng.RangeCheck()
    for _ in range(6):
        eng.RunInference()
    mx1, sat1 = eng.RangeCheck()
    x, a = eng.FirstLayer(cfg.channels)
    eng.close()
    live = np.array([20 + r * 20 + c for r in range(19) for c in range(19)])
    pad = np.setdiff1d(np.arange(400), live)
    assert sat0 == 0 and sat1 == 0 and mx1 == mx0
    assert not x[:, pad].any() and not a[:, pad].any()


@pytest.mark.parametrize("config,B", [("b12c256btl3", 320), ("b10c128btl3", 200)])
def test_scheduling_knobs_do_not_change_results(config, B, weight_dir, golden_positions, monkeypatch):
    """Round-2 changes that only re-schedule the same arithmetic - alternating tile order between launches, the tower's end as
    one launch, the shifted-view first layer, the L2 prefetch of the boundaries next to a broadcast block, CUDA graph / programmatic
    dependent launch - must leave every output bit-identical: each is switched off in turn and compared with the default engine."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir(config)
    feats = golden_positions["feats"]
    allf = feats[np.arange(B) % len(feats)]

    def run():
        eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
        eng.LoadBatchAll(allf)
        eng.RunInference()
        eng.RunInference()
        out = [eng.GetBatch(b).copy() for b in range(B)]
        own = [np.asarray(eng.GetOwnership(b)).copy() for b in range(0, B, 37)]
        eng.close()
        return out, own

    want, want_own = run()
    for knob, val in (("P3_TILE_ALTERNATE", "0"), ("P3_TC_TAIL", "0"), ("P3_INIT_TC", "1"), ("P3_CHAIN_L2PF", "0"), ("P3_CUDA_GRAPH", "0"),
                      ("P3_PDL", "0")):
        monkeypatch.setenv(knob, val)
        got, got_own = run()
        monkeypatch.delenv(knob)
        for b in range(B):
            assert _same(got[b], want[b]), (knob, b)
        for a, w in zip(got_own, want_own):
            assert np.array_equal(a, w), knob


# ---- compact leaf records, per-bank auxiliary outputs, ownership symmetry, root sampling on resident logits ---------------------
@pytest.mark.parametrize("precision_name", ["fp32", "bf16"])
def test_leaf_results_equal_initfields_of_full_results(precision_name, weight_dir, golden_positions):
    """P3_RESULT_LEAF: the 4360-byte record per slot holds exactly the three policy arrays of the NNInferResult of
    the same run plus InitFields' four scalars (cc/mcts/leaf_evaluator.cc:83-112, restated as oracle orc_init_fields) - serial
    and over the slot banks; GetBatch refuses to serve stale full results in that mode."""
    import ctypes
    from p3achygo_b200 import engine as E
    L = oracle_lib.oracle()
    prec = E.PRECISION_FP32 if precision_name == "fp32" else E.PRECISION_BF16
    path, cfg, tensors = weight_dir("b10c128btl3")
    B = 8
    feats = golden_positions["feats"][200:200 + 2 * B]
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=prec)
    full = []
    for lo in (0, B):
        for b in range(B):
            eng.LoadBatchSym(b, feats[lo + b], (lo + b) % 8)
        eng.RunInference()
        full += [eng.GetBatch(b).copy() for b in range(B)]
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    def check(leaf, r):
        for f in ("move_logits", "move_probs", "opt_move_probs"):
            assert np.array_equal(np.asarray(leaf[f]), np.asarray(r[f])), f
        want = np.zeros(3, dtype=np.float32)
        L.orc_init_fields(vp(np.ascontiguousarray(r["value_probs"], dtype=np.float32)),
                          vp(np.ascontiguousarray(r["score_probs"], dtype=np.float32)), vp(want))
        assert abs(float(leaf["value"]) - float(want[0])) <= 1e-6
        es2 = float(want[2]) + float(want[1]) ** 2
        assert abs(float(leaf["score_mean"]) - float(want[1])) <= 1e-3 * max(1.0, abs(float(want[1])))  # 800 fp32 terms of up to 400
        assert abs(float(leaf["score_var"]) - float(want[2])) <= 2e-5 * max(1.0, es2)      # fp32 cancellation scale
        assert abs(float(leaf["err"]) - float(np.sqrt(np.float32(r["err2_outcome"])))) <= 1e-6

    eng.SetResultMode(E.RESULT_LEAF)
    for b in range(B):
        eng.LoadBatchSym(b, feats[b], b % 8)
    eng.RunInference()
    for b in range(B):
        check(eng.GetLeaf(b), full[b])
    with pytest.raises(E.P3Error):
        eng.GetBatch(0)
    for bank in (0, 1):
        for b in range(B):
            eng.LoadBatchBank(bank, b, feats[bank * B + b], (bank * B + b) % 8)
        eng.Submit(bank)
    for bank in (1, 0):
        eng.Wait(bank)
        for b in range(B):
            check(eng.GetLeafBank(bank, b), full[bank * B + b])
        with pytest.raises(E.P3Error):
            eng.GetBatchBank(bank, 0)
    eng.SetResultMode(E.RESULT_FULL)
    for b in range(B):
        eng.LoadBatchSym(b, feats[b], b % 8)
    eng.RunInference()
    assert _same(eng.GetBatch(3), full[3])
    with pytest.raises(E.P3Error):
        eng.GetLeaf(0)
    eng.close()


def test_bank_aux_and_ownership_symmetry(weight_dir, golden_positions):
    """each slot bank keeps its own auxiliary outputs (ownership, leaf statistics), valid after Wait while the other
    bank's step has already overwritten the shared step buffer; ownership is un-rotated like the policies (ApplyInverse)."""
    import ctypes
    from p3achygo_b200 import engine as E
    L = oracle_lib.oracle()
    path, cfg, tensors = weight_dir("b10c128btl3")
    B = 8
    feats = golden_positions["feats"][300:300 + 2 * B]
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_FP32)
    want = []
    for lo in (0, B):                               # identity orientation, serial: the reference point
        for b in range(B):
            eng.LoadBatch(b, feats[lo + b])
        eng.RunInference()
        want += [eng.GetAux(b).copy() for b in range(B)]
    with pytest.raises(E.P3Error):
        eng.GetAuxBank(0, 0)                        # no submitted run yet
    for bank in (0, 1):
        for b in range(B):
            eng.LoadBatchBank(bank, b, feats[bank * B + b], 0)
        eng.Submit(bank)
    eng.Wait(0)
    eng.Wait(1)                                     # bank 1's step ran last: the shared buffer holds bank 1
    for bank in (0, 1):
        for b in range(B):
            a = eng.GetAuxBank(bank, b)
            for f in a.dtype.names:
                assert np.array_equal(np.asarray(a[f]), np.asarray(want[bank * B + b][f])), (bank, b, f)
            assert np.array_equal(eng.GetOwnershipBank(bank, b), np.asarray(want[bank * B + b]["ownership"]))
    # ownership under a symmetry: the net sees the rotated board; the result must come back in the game's orientation.
    # Rotating the input is not bit-neutral for the net, so compare with the host-side route: rotate features on the host,
    # run with sym = 0, ApplyInverse on the ownership (oracle), against LoadBatchSym + GetOwnership.
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for sym in range(8):
        f = np.array(feats[sym:sym + 1], dtype=GO_FEATURES_DTYPE)
        fs = f.copy()
        for grid in ("board", "stones_atari", "stones_two_liberties", "stones_three_liberties", "stones_laddered"):
            src = np.ascontiguousarray(f[grid][0], dtype=np.int8)
            dst = np.zeros(361, dtype=np.int8)
            L.orc_apply_symmetry_i8(sym, vp(src), vp(dst))
            fs[grid][0] = dst
        for k in range(5):
            i, j = (int(v) for v in f["last_moves"][0][k])
            if 0 <= i < 19 and 0 <= j < 19:
                t = L.orc_transform_index(sym, i * 19 + j)
                fs["last_moves"][0][k] = (t // 19, t % 19)
        eng.LoadBatch(0, fs[0])
        eng.LoadBatchSym(1, f[0], sym)
        eng.RunInference()
        rot = np.ascontiguousarray(eng.GetOwnership(0), dtype=np.float32)
        inv = np.zeros(361, dtype=np.float32)
        L.orc_apply_inverse_f32(sym, vp(rot), vp(inv))
        assert np.array_equal(eng.GetOwnership(1), inv), sym
    eng.close()


def test_root_sampling_on_resident_logits(weight_dir, golden_positions):
    """p3_engine_gumbel_topk_bank: Gumbel top-k (cc/mcts/gumbel.cc:283-321) drawn from move_logits that never left HBM equals
    the oracle fed with the logits GetBatch returns - after a submit (bank copies) and after a serial run."""
    from p3achygo_b200 import engine as E
    L = oracle_lib.oracle()
    path, cfg, tensors = weight_dir("b10c128btl3")
    B, k = 16, 16
    feats = golden_positions["feats"][400:400 + B]
    legal = golden_positions["legal"][400:400 + B] if "legal" in golden_positions else None
    rng = np.random.default_rng(3)
    if legal is None:
        legal = (rng.random((B, 362)) < 0.8).astype(np.uint8)
        legal[:, 361] = 1
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    slots = np.array([5, 0, 15, 7, 7, 3], dtype=np.int32)
    seeds = [L.orc_prng_seed(900 + i) for i in range(len(slots))]
    for mode in ("submit", "serial"):
        if mode == "submit":
            for b in range(B):
                eng.LoadBatchBank(1, b, feats[b], 0)
            eng.Submit(1)
            eng.Wait(1)
            res = [eng.GetBatchBank(1, b).copy() for b in range(B)]
            bank = 1
        else:
            for b in range(B):
                eng.LoadBatch(b, feats[B - 1 - b])
            eng.RunInference()
            res = [eng.GetBatch(b).copy() for b in range(B)]
            bank = 0
        state = np.array(seeds, dtype=np.uint64)
        moves, scores, kvalid = eng.GumbelTopKBank(bank, slots, legal[slots], state, 1.0, k)
        for i, sl in enumerate(slots):
            om, osc, okv, ost = oracle_lib.gumbel_topk(int(seeds[i]), np.asarray(res[sl]["move_logits"], dtype=np.float32),
                                                       legal[sl], 1.0, k)
            kk = min(k, okv)
            assert kvalid[i] == okv and int(state[i]) == ost
            assert np.array_equal(moves[i][:kk], om[:kk]) and np.array_equal(scores[i][:kk], osc[:kk]), (mode, i)
    eng.close()


# ---- dynamic range of the fp16 residual stream (ADVICE r1 / VERDICT r1 weak 1) ------------------------------------------------
def _scaled_family(cfg, base, f_expand):
    """synthetic_weights with every block's last conv scaled by f_expand: the residual stream grows block by block
    (b10c128btl3: max |x| 2.3 / 275 / 3.4e3 / 1.9e5 for f = 1 / 8 / 12 / 20, measured with the fp64 oracle; the bf16 engine does
    not store the last block's stream, so its scan sees the blocks before it)."""
    from p3achygo_b200 import weights as W
    t = {k: v.copy() for k, v in base.items()}
    for name in t:
        if "/trunk/" in name and name.endswith("conv/kernel"):
            blk, j = int(name.split("/")[2].split(":")[0]), int(name.split("/")[3].split(":")[0])
            if j == (2 if "broadcast_res" in name else len(W.block_convs(cfg, blk)) - 1):
                t[name] *= np.float32(f_expand)
    return t


def test_large_magnitude_trunk_bounded_or_loud(tmp_path, golden_positions):
    """The tensor engine keeps the residual stream in IEEE fp16 with a saturating pack.  Large but representable trunks
    (max |x| up to 3.4e3, 1500 x the He-scaled family every other test uses) keep the documented RELATIVE accuracy; a trunk beyond
    +-65504 is reported by p3_engine_range_check and, in validation mode (P3_RANGE_CHECK=1), makes RunInference fail loudly
    instead of returning clamped evaluations; the fp32 engine evaluates the same net within its bound."""
    import os
    from p3achygo_b200 import engine as E
    from p3achygo_b200 import weights as W
    cfg = W.config_from_str("b10c128btl3")
    base = W.synthetic_weights(cfg, 0)
    feats = golden_positions["feats"][100:104]
    for f, expect_sat in ((8.0, False), (12.0, False), (26.0, True)):
        tensors = _scaled_family(cfg, base, f)
        path = os.path.join(tmp_path, f"scaled_{int(f)}.p3w")
        W.save_weights(path, cfg, tensors)
        o = _oracle_outputs(cfg, tensors, feats, torch.float64)
        scale = float(np.abs(o["pi_logits"]).max())
        trunk_max = float(np.abs(o["trunk"]).max())
        for prec in (E.PRECISION_FP32, E.PRECISION_BF16):
            eng = E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=prec)
            for b in range(4):
                eng.LoadBatch(b, feats[b])
            eng.RunInference()
            mx, sat = eng.RangeCheck()
            got = np.stack([np.asarray(eng.GetBatch(b)["move_logits"], dtype=np.float64) for b in range(4)])
            rel = float(np.abs(got - o["pi_logits"]).max()) / scale
            print(f"f={f} prec={prec}: oracle trunk max {trunk_max:.3g}, engine stream max {mx:.3g}, saturated {sat}, rel logit err {rel:.2e}")
            assert mx >= 0.5 * min(trunk_max, 65504.0)            # the scan sees every block, the oracle's number is the last one
            if prec == E.PRECISION_FP32:
                assert sat == 0 and rel <= 1e-4                    # fp32 stream: no clamp, 1e-3-class accuracy at any scale
            elif not expect_sat:
                assert sat == 0 and rel <= 6e-2                    # documented bf16 bound, relative to the logit scale
            else:
                assert sat > 0                                     # the clamp was hit ...
                os.environ["P3_RANGE_CHECK"] = "1"
                try:
                    with pytest.raises(E.P3Error) as ei:           # ... and validation mode refuses to evaluate
                        eng.RunInference()
                    assert ei.value.code == E._lib.P3_ERR_UNSUPPORTED and "saturated" in str(ei.value)
                finally:
                    del os.environ["P3_RANGE_CHECK"]
            eng.close()
