#!/usr/bin/env python
"""Generates tests/golden/model_golden.npz: outputs of the reference's OWN model code on seeded positions and weights.

Run in the build container only (needs /root/reference):
    python tests/golden/make_model_golden.py

What runs: the unmodified reference sources python/{model,model_transformer,model_config,constants}.py —
`P3achyGoModel.create(ModelConfig.<cfg>(), ...)` and its `call` (python/model.py:1222-1295) — on top of
oracle/tf_shim, a restatement of the TensorFlow 2.16 / Keras 3 calls that code makes (those libraries are not in this
image; see oracle/tf_shim/README.md), in float64.  The weights are this repo's seeded synthetic tensors
(p3achygo_b200.weights.synthetic_weights), assigned to the reference model's layers by attribute, the inputs are the
feature planes of committed golden positions (tests/golden/positions.npz, made by the compiled reference).

The fixture pins oracle/model_ref.py (the PyTorch restatement the parity tests use) and, through it, the CUDA engine.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PY = "/root/reference/python"
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))  # `tensorflow`, `keras`
sys.path.insert(0, REF_PY)
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import model as ref_model  # noqa: E402  (the reference's python/model.py)
from model_config import ModelConfig as RefConfig  # noqa: E402

from oracle import oracle_lib  # noqa: E402
from p3achygo_b200 import weights as W  # noqa: E402
from p3achygo_b200._lib import GO_FEATURES_DTYPE  # noqa: E402

# (config, first position index, number of positions): small batches, the net is evaluated in float64 on the CPU
CASES = [("tiny", 10, 8), ("b10c128btl3", 100, 3), ("b12c256btl3", 100, 2), ("b14c384btl3", 100, 1), ("b15c192_classic", 100, 1),
         ("b8c128nbt", 100, 2), ("small", 100, 2)]

OUTPUT_NAMES = ["pi_logits", "pi", "outcome_logits", "outcome", "own", "score_logits", "score_probs", "gamma", "pi_logits_aux",
                "q6", "q16", "q50", "q6_err", "q16_err", "q50_err", "q6_score", "q16_score", "q50_score", "q6_score_err",
                "q16_score_err", "q50_score_err", "pi_logits_soft", "pi_logits_optimistic", "mcts_dist_logits", "mcts_dist_probs"]


def set_conv(layer, w_oihw):
    layer.kernel.assign(np.transpose(w_oihw, (2, 3, 1, 0)))  # OIHW -> Keras HWIO


def set_dense(layer, tensors, tag):
    layer.kernel.assign(tensors[f"{tag}/dense/kernel"])
    layer.bias.assign(tensors[f"{tag}/dense/bias"])


def set_bn(layer, tensors, tag):
    assert abs(layer.epsilon - float(tensors[f"{tag}/batch_norm/epsilon"][0])) < 1e-9  # fp32(1e-3) in the weight file
    layer.gamma.assign(tensors[f"{tag}/batch_norm/gamma"])
    layer.beta.assign(tensors[f"{tag}/batch_norm/beta"])
    layer.moving_mean.assign(tensors[f"{tag}/batch_norm/moving_mean"])
    layer.moving_variance.assign(tensors[f"{tag}/batch_norm/moving_variance"])


def set_conv_block(block, tensors, tag):
    """ConvPreActivation (python/model.py:266-281): norm_layer = BatchNormalization, conv = Conv2D without bias."""
    assert isinstance(block, ref_model.ConvPreActivation)
    set_bn(block.norm_layer, tensors, tag)
    set_conv(block.conv, tensors[f"{tag}/conv/kernel"])


def assign_weights(m, cfg, tensors):
    set_conv(m.init_board_conv, tensors["model/init_conv/conv/kernel"])
    set_dense(m.init_game_layer, tensors, "model/init_game_state")
    for i, block in enumerate(m.blocks):
        tag = W.block_tag(cfg, i)
        if isinstance(block, ref_model.BroadcastResidualBlock):
            set_conv_block(block.blocks[0], tensors, f"{tag}/00:conv_block")
            set_dense(block.blocks[1].dense, tensors, f"{tag}/01:broadcast")
            set_conv_block(block.blocks[2], tensors, f"{tag}/02:conv_block")
        elif isinstance(block, ref_model.NbtResidualBlock):
            # reduce conv, two classic blocks of two convs each, expand conv (python/model.py:431-470): flattened 0..5
            flat = [block.blocks[0], *block.blocks[1].blocks, *block.blocks[2].blocks, block.blocks[3]]
            for j, cb in enumerate(flat):
                set_conv_block(cb, tensors, f"{tag}/{j:02d}:conv_block")
        else:
            for j, cb in enumerate(block.blocks):
                set_conv_block(cb, tensors, f"{tag}/{j:02d}:conv_block")
    ph, p = m.policy_head, "model/policy_head"
    set_conv(ph.conv_p, tensors[f"{p}/conv_policy/conv/kernel"])
    set_conv(ph.conv_g, tensors[f"{p}/conv_global/conv/kernel"])
    set_bn(ph.gpool.g_norm_layer, tensors, f"{p}/global_pool_bias")
    set_dense(ph.gpool.dense, tensors, f"{p}/global_pool_bias")
    set_conv(ph.output_moves, tensors[f"{p}/conv_moves/conv/kernel"])
    set_dense(ph.output_pass, tensors, f"{p}/dense_pass")
    set_conv(ph.soft_policy_moves, tensors[f"{p}/conv_soft_moves/conv/kernel"])
    set_dense(ph.soft_policy_pass, tensors, f"{p}/dense_soft_pass")
    set_conv(ph.optimistic_policy_moves, tensors[f"{p}/conv_optimistic_moves/conv/kernel"])
    set_dense(ph.optimistic_policy_pass, tensors, f"{p}/dense_optimistic_pass")
    vh, v = m.value_head, "model/value_head"
    set_conv(vh.conv, tensors[f"{v}/conv_value/conv/kernel"])
    set_dense(vh.outcome_q_embed, tensors, f"{v}/dense_outcome_pre")
    set_dense(vh.outcome_q_output, tensors, f"{v}/dense_outcome")
    set_dense(vh.outcome_mcts_dist, tensors, f"{v}/dense_mcts_dist")
    set_conv(vh.conv_ownership, tensors[f"{v}/ownership/conv/kernel"])
    set_dense(vh.gamma_pre, tensors, f"{v}/dense_gamma_pre")
    set_dense(vh.gamma_output, tensors, f"{v}/dense_gamma")
    set_dense(vh.score_pre, tensors, f"{v}/dense_scores_pre")
    set_dense(vh.score_output, tensors, f"{v}/dense_scores")


def main():
    z = np.load(os.path.join(HERE, "positions.npz"))
    feats_all = np.ascontiguousarray(z["feats"]).view(GO_FEATURES_DTYPE).reshape(-1)
    out = {}
    for name, first, n in CASES:
        if name not in W.CONFIGS:
            print(f"skip {name}: not a config of this repo yet")
            continue
        cfg = W.config_from_str(name)
        tensors = W.synthetic_weights(cfg, 0)
        rcfg = RefConfig.from_str(name)
        m = ref_model.P3achyGoModel.create(rcfg, 19, cfg.num_input_planes, cfg.num_input_features, name)
        feats = feats_all[first:first + n]
        planes, scalars = oracle_lib.load_go_features(feats, 1)
        x = torch.from_numpy(planes.astype(np.float64))
        g = torch.from_numpy(scalars.astype(np.float64))
        with torch.no_grad():
            m(x[:1], g[:1])  # builds every layer
            assign_weights(m, cfg, tensors)
            ys = m(x, g)
        assert len(ys) == len(OUTPUT_NAMES)
        out[f"{name}/first"] = np.array([first, n], dtype=np.int64)
        for nm, y in zip(OUTPUT_NAMES, ys):
            out[f"{name}/{nm}"] = y.numpy().astype(np.float64)
        print(name, "pi_logits[0,:4] =", out[f"{name}/pi_logits"][0, :4], "outcome[0] =", out[f"{name}/outcome"][0])
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "model_golden.npz"))


if __name__ == "__main__":
    main()
