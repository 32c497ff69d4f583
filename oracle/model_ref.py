"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement (PyTorch, fp32 or fp64) of the forward pass of the reference's Keras model
``P3achyGoModel.call`` (``python/model.py:1222-1295``).  The arithmetic of the reference net lives
in third-party modules that are absent here — TensorFlow 2.16.2 / Keras 3 (``requirements.txt``;
Conv2D, BatchNormalization, Dense, mish, softmax, softplus, tanh, sigmoid) and TensorRT 10 — so
their published definitions are restated and each function cites the reference call site.

PARITY PINNED against the reference's own model code: tests/golden/model_golden.npz holds the outputs of the UNMODIFIED
``python/model.py`` (``P3achyGoModel.create(...)`` + ``call``) run in this container in float64 on ``oracle/tf_shim`` (a
restatement of the TensorFlow / Keras calls model.py makes, since those libraries cannot be installed here), with this repo's
seeded weights on committed golden positions (generator: tests/golden/make_model_golden.py).  tests/test_model_golden.py
checks this restatement against that fixture to 1e-9 on all 25 outputs for every BASELINE config.  What stays unpinned is
only TensorFlow's own kernel arithmetic (conv / matmul rounding order), which no restatement can pin without TensorFlow; the
reference's own tests hold no numeric vectors for it (``python/test/model_v1_test.py:74-279`` checks shapes / ranges).
(Integer features / masks / PRNG are pinned separately: oracle/features_oracle.c.)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

NUM_LOCS = 361


def mish(x: torch.Tensor) -> torch.Tensor:
    """keras.activations.mish: x * tanh(softplus(x)) (model.py:269-281)."""
    return x * torch.tanh(F.softplus(x))


class RefModel:
    """Forward-only restatement. ``cfg`` is p3achygo_b200.weights.ModelConfig, ``w`` the tensor dict
    in the P3W1 naming (conv kernels OIHW, dense kernels (in, out))."""

    def __init__(self, cfg, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}

    # -- layers ------------------------------------------------------------------------------
    def conv(self, tag: str, x: torch.Tensor) -> torch.Tensor:
        """make_conv: Conv2D padding='same', no bias (model.py:101-117). x is NCHW here."""
        k = self.w[f"{tag}/conv/kernel"]
        return F.conv2d(x, k, padding=k.shape[-1] // 2)

    def dense(self, tag: str, x: torch.Tensor) -> torch.Tensor:
        """make_dense: Dense with bias (model.py:120-126)."""
        return x @ self.w[f"{tag}/dense/kernel"] + self.w[f"{tag}/dense/bias"]

    def bn(self, tag: str, x: torch.Tensor) -> torch.Tensor:
        """BatchNormalization(momentum .99, eps 1e-3) in inference form (model.py:231):
        gamma * (x - mean) / sqrt(var + eps) + beta, per channel (dim 1 of NCHW)."""
        g = self.w[f"{tag}/batch_norm/gamma"]
        b = self.w[f"{tag}/batch_norm/beta"]
        m = self.w[f"{tag}/batch_norm/moving_mean"]
        v = self.w[f"{tag}/batch_norm/moving_variance"]
        eps = self.w[f"{tag}/batch_norm/epsilon"][0]
        scale = g / torch.sqrt(v + eps)
        shift = b - m * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)

    def conv_pre_act(self, tag: str, x: torch.Tensor) -> torch.Tensor:
        """ConvPreActivation.call: conv(mish(BN(x))) (model.py:276-292)."""
        return self.conv(tag, mish(self.bn(tag, x)))

    def broadcast(self, tag: str, x: torch.Tensor) -> torch.Tensor:
        """BroadcastPreAct.call (model.py:570-581): per channel, mish then Dense(361 -> 361) over the
        flattened board; one shared kernel + bias for all channels."""
        b, c, h, w = x.shape
        t = mish(x).reshape(b, c, h * w)
        t = t @ self.w[f"{tag}/dense/kernel"] + self.w[f"{tag}/dense/bias"]
        return t.reshape(b, c, h, w)

    def block(self, i: int, x: torch.Tensor) -> torch.Tensor:
        """ResidualBlock.call: res + blocks(x), linear activation (model.py:316-327)."""
        from p3achygo_b200.weights import block_convs, block_tag

        convs = block_convs(self.cfg, i)
        y = x
        if self.cfg.is_broadcast(i):  # BroadcastResidualBlock, model.py:583-607
            y = self.conv_pre_act(convs[0][0], y)
            y = self.broadcast(f"{block_tag(self.cfg, i)}/01:broadcast", y)
            y = self.conv_pre_act(convs[1][0], y)
        elif self.cfg.trunk_block_type == "nbt":  # NbtResidualBlock (model.py:431-470): two nested classic blocks at Cb
            t = self.conv_pre_act(convs[0][0], y)
            t = t + self.conv_pre_act(convs[2][0], self.conv_pre_act(convs[1][0], t))
            t = t + self.conv_pre_act(convs[4][0], self.conv_pre_act(convs[3][0], t))
            y = self.conv_pre_act(convs[5][0], t)
        else:  # Bottleneck (model.py:372-412) / Classic (model.py:330-354)
            for tag, _, _, _ in convs:
                y = self.conv_pre_act(tag, y)
        return x + y

    @staticmethod
    def gpool(x: torch.Tensor) -> torch.Tensor:
        """GlobalPool.call: concat(mean_HW, max_HW) (model.py:643-647)."""
        return torch.cat([x.mean(dim=(2, 3)), x.amax(dim=(2, 3))], dim=1)

    # -- forward -----------------------------------------------------------------------------
    def trunk(self, planes_nhwc: np.ndarray, feats: np.ndarray) -> torch.Tensor:
        x = torch.from_numpy(np.ascontiguousarray(planes_nhwc)).to(self.dtype).permute(0, 3, 1, 2)
        g = torch.from_numpy(np.ascontiguousarray(feats)).to(self.dtype)
        x = self.conv("model/init_conv", x)                                  # model.py:1230
        x = x + self.dense("model/init_game_state", g)[:, :, None, None]     # model.py:1231-1237
        for i in range(self.cfg.blocks):                                     # model.py:1239-1240
            x = self.block(i, x)
        return x

    def forward(self, planes_nhwc: np.ndarray, feats: np.ndarray) -> Dict[str, np.ndarray]:
        with torch.no_grad():
            x = self.trunk(planes_nhwc, feats)
            out = {}
            out.update(self.policy_head(x))
            out.update(self.value_head(x))
            out["pi"] = torch.softmax(out["pi_logits"], dim=1)                   # model.py:1265
            out["outcome"] = torch.softmax(out["outcome_logits"], dim=1)         # model.py:1266
            out["score_probs"] = torch.softmax(out["score_logits"], dim=1)       # model.py:1267
            # what the C++ consumer derives (trt_engine.cc:347; leaf_evaluator.cc:83-112)
            out["opt_move_probs"] = torch.softmax(out["pi_logits_optimistic"], dim=1)
            s = torch.arange(800, dtype=self.dtype) - 400 + 0.5
            mean = (out["score_probs"] * s).sum(1)
            out["score_mean"] = mean
            out["score_var"] = (out["score_probs"] * s * s).sum(1) - mean * mean
            out["value"] = out["outcome"][:, 1] - out["outcome"][:, 0]
            out["trunk"] = x.permute(0, 2, 3, 1).reshape(x.shape[0], NUM_LOCS, -1)
            return {k: v.to(torch.float64).numpy() for k, v in out.items()}

    def policy_head(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """PolicyHead.call (model.py:783-812) + GlobalPoolBias.call (model.py:691-702)."""
        ph = "model/policy_head"
        b = x.shape[0]
        p = self.conv(f"{ph}/conv_policy", x)
        g = self.conv(f"{ph}/conv_global", x)
        g = mish(self.bn(f"{ph}/global_pool_bias", g))
        g_pooled = self.gpool(g)
        p = p + self.dense(f"{ph}/global_pool_bias", g_pooled)[:, :, None, None]
        p = mish(p)
        pi = self.conv(f"{ph}/conv_moves", p)                       # [B,2,19,19]
        pass_logits = self.dense(f"{ph}/dense_pass", g_pooled) - 3  # model.py:795
        pi_soft = self.conv(f"{ph}/conv_soft_moves", p).reshape(b, NUM_LOCS)
        pass_soft = self.dense(f"{ph}/dense_soft_pass", g_pooled) - 3
        pi_opt = self.conv(f"{ph}/conv_optimistic_moves", p).reshape(b, NUM_LOCS)
        pass_opt = self.dense(f"{ph}/dense_optimistic_pass", g_pooled) - 3
        return {
            "pi_logits": torch.cat([pi[:, 0].reshape(b, NUM_LOCS), pass_logits[:, 0:1]], dim=1),
            "pi_logits_aux": torch.cat([pi[:, 1].reshape(b, NUM_LOCS), pass_logits[:, 1:2]], dim=1),
            "pi_logits_soft": torch.cat([pi_soft, pass_soft], dim=1),
            "pi_logits_optimistic": torch.cat([pi_opt, pass_opt], dim=1),
        }

    def value_head(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """ValueHead.call (model.py:887-979)."""
        vh = "model/value_head"
        b = x.shape[0]
        v = self.conv(f"{vh}/conv_value", x)
        v_pooled = self.gpool(v)
        e = mish(self.dense(f"{vh}/dense_outcome_pre", v_pooled))
        o = self.dense(f"{vh}/dense_outcome", e)
        mcts_logits = self.dense(f"{vh}/dense_mcts_dist", e)
        own = torch.tanh(self.conv(f"{vh}/ownership", v)).reshape(b, NUM_LOCS)
        gamma = self.dense(f"{vh}/dense_gamma", mish(self.dense(f"{vh}/dense_gamma_pre", v_pooled)))
        scores = self.w[f"{vh}/scores"]                                         # model.py:1225-1228
        v_scores = torch.cat([v_pooled[:, None, :].expand(b, 800, v_pooled.shape[1]),
                              scores[None, :, None].expand(b, 800, 1)], dim=-1)  # model.py:925-944
        s = self.dense(f"{vh}/dense_scores", mish(self.dense(f"{vh}/dense_scores_pre", v_scores)))[..., 0]
        score_logits = torch.clamp(F.softplus(gamma), max=10.0) * s             # model.py:949-951
        return {
            "outcome_logits": o[:, 0:2],
            "own": own,
            "score_logits": score_logits,
            "gamma": gamma[:, 0],
            "q": torch.tanh(o[:, 2:5]),
            "q_err": 4 * torch.sigmoid(o[:, 5:8]),
            "q_score": o[:, 8:11],
            "q_score_err": torch.abs(o[:, 11:14]),
            "mcts_dist_logits": mcts_logits,
            "mcts_dist_probs": torch.softmax(mcts_logits, dim=1),
        }
