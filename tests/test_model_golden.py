"""The net-forward oracle (oracle/model_ref.py) against golden vectors produced by the reference's OWN model code.

tests/golden/model_golden.npz holds the 25 outputs of `P3achyGoModel.call` (python/model.py:1222-1295) — the unmodified
reference sources executed in float64 on oracle/tf_shim (TensorFlow / Keras are absent from this image; the shim restates the
calls model.py makes, see oracle/tf_shim/README.md) — on committed golden positions with this repo's seeded synthetic
weights.  Generator: tests/golden/make_model_golden.py (runs only where /root/reference exists).

This pins the block structure, BN / activation placement, head wiring, output order and weight-tensor layout of the
restatement; the GPU parity tests then compare the CUDA engine with the restatement and (test_gpu_engine.py) with this
fixture directly.
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_lib
from oracle.model_ref import RefModel
from p3achygo_b200 import weights as W

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "model_golden.npz")

# fixture name -> (RefModel.forward key, column selector)
DIRECT = ["pi_logits", "pi", "outcome_logits", "outcome", "score_logits", "score_probs", "gamma", "pi_logits_aux",
          "pi_logits_soft", "pi_logits_optimistic", "mcts_dist_logits", "mcts_dist_probs"]
Q_GROUPS = {"q": ["q6", "q16", "q50"], "q_err": ["q6_err", "q16_err", "q50_err"], "q_score": ["q6_score", "q16_score", "q50_score"],
            "q_score_err": ["q6_score_err", "q16_score_err", "q50_score_err"]}


def golden_cases():
    z = np.load(GOLDEN)
    return sorted({k.split("/")[0] for k in z.files})


@pytest.mark.parametrize("config", golden_cases())
def test_restatement_matches_reference_model_code(config, golden_positions):
    z = np.load(GOLDEN)
    first, n = (int(v) for v in z[f"{config}/first"])
    cfg = W.config_from_str(config)
    tensors = W.synthetic_weights(cfg, 0)
    feats = golden_positions["feats"][first:first + n]
    planes, scalars = oracle_lib.load_go_features(feats, 1)
    o = RefModel(cfg, tensors, dtype=torch.float64).forward(planes, scalars)
    tol = 1e-9  # both sides are float64; only the summation order differs
    for name in DIRECT:
        ref = z[f"{config}/{name}"].reshape(n, -1)
        got = np.asarray(o[name]).reshape(n, -1)
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), (name, np.abs(got - ref).max())
    own = z[f"{config}/own"].reshape(n, 361)
    assert np.abs(o["own"] - own).max() <= tol
    for key, names in Q_GROUPS.items():
        for k, nm in enumerate(names):
            assert np.abs(o[key][:, k] - z[f"{config}/{nm}"].reshape(n)).max() <= tol * max(1.0, np.abs(z[f'{config}/{nm}']).max()), nm


def test_golden_covers_every_baseline_config():
    """The five BASELINE.json configs' architectures (tiny for config/test.json-class nets) are all pinned."""
    have = set(golden_cases())
    assert {"tiny", "b10c128btl3", "b12c256btl3", "b14c384btl3", "b15c192_classic"} <= have
