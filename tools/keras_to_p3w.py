#!/usr/bin/env python
"""Real-checkpoint importer: p3achygo `.keras` checkpoint -> flat `.p3w` weight file of the B200 engine.

    python tools/keras_to_p3w.py model_0123.keras model_0123.p3w

What a `.keras` file is (python/rl_loop/model_utils.py:197-204 -> `model.save`): a zip with `config.json` (the
`P3achyGoModel.get_config()` dict, python/scripts/migrate_checkpoint.py:49-66 reads it the same way) and `model.weights.h5`.
Keras 3 stores one dataset per variable at `<object path>/vars/<i>`, where the object path follows the Python ATTRIBUTES of the
model ("Keras 3 keys h5 weights by attribute name", migrate_checkpoint.py:3-9: `layers/value_head/outcome_q_embed`), items of a
list attribute are named by the snake-cased class name with `_1`, `_2`, ... for repeats, and `<i>` is the position in
`layer.weights` (Conv2D: kernel; Dense: kernel, bias; BatchNormalization: gamma, beta, moving_mean, moving_variance).

The tensor mapping is the one tests/golden/make_model_golden.py::assign_weights applies to the reference's own model object
(attribute by attribute), expressed as attribute paths; conv kernels go from Keras HWIO to the OIHW of the `.p3w` format
(python/export_weights.py:16-90 tag tree, p3achygo_b200/weights.py).  A path is resolved against the file by trying the
attribute-path form at the root and under `layers/`, and the class-name form `layers/<snake class>[_k]` for the model's direct
children (which of them Keras used depends on its visit order; every candidate is shape-checked and exactly one must match).

No TensorFlow / Keras / h5py needed: tools/minih5.py reads the HDF5 subset Keras writes.  `write_keras_checkpoint` produces a
checkpoint with the same layout from a tensor dict; the round-trip test (tests/test_keras_import.py) uses it, since this image
has no Keras to write a real one - so the naming rules above are pinned to the reference's scripts, not to a Keras run.
"""
from __future__ import annotations

import io
import json
import os
import re
import sys
import zipfile
from typing import Dict, List, Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import minih5  # noqa: E402
from p3achygo_b200 import weights as W  # noqa: E402

# A path element is an attribute name, or (list attribute, index, class name of the item, class names of the items before it).
Elem = object


def snake(name: str) -> str:
    """keras.src.utils.naming.to_snake_case for the class names that occur here."""
    s = re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", s).lower()


def list_item_name(classes: List[str], i: int) -> str:
    """Name of item i of a list attribute: snake class name, `_k` for the k-th repeat (saving_lib._save_container_state)."""
    k = sum(1 for c in classes[:i] if c == classes[i])
    return snake(classes[i]) + (f"_{k}" if k else "")


def trunk_block_classes(cfg: W.ModelConfig) -> List[str]:
    kind = {"btl": "BottleneckResidualConvBlock", "classic": "ClassicResidualBlock", "nbt": "NbtResidualBlock"}[cfg.trunk_block_type]
    return ["BroadcastResidualBlock" if cfg.is_broadcast(i) else kind for i in range(cfg.blocks)]


def inner_paths(cfg: W.ModelConfig, i: int) -> List[Tuple[str, List[str]]]:
    """(p3w conv_block tag, attribute path below model.blocks[i]) of every ConvPreActivation of trunk block i
    (python/model.py:330-486, 583-607; same flattening as make_model_golden.assign_weights)."""
    convs = W.block_convs(cfg, i)
    if cfg.is_broadcast(i):
        classes = ["ConvPreActivation", "BroadcastPreAct", "ConvPreActivation"]
        return [(convs[0][0], ["blocks", list_item_name(classes, 0)]), (convs[1][0], ["blocks", list_item_name(classes, 2)])]
    if cfg.trunk_block_type == "nbt":
        outer = ["ConvPreActivation", "ClassicResidualBlock", "ClassicResidualBlock", "ConvPreActivation"]
        inner = ["ConvPreActivation", "ConvPreActivation"]
        paths = [["blocks", list_item_name(outer, 0)]]
        for r in (1, 2):
            for j in (0, 1):
                paths.append(["blocks", list_item_name(outer, r), "blocks", list_item_name(inner, j)])
        paths.append(["blocks", list_item_name(outer, 3)])
        return [(convs[k][0], paths[k]) for k in range(6)]
    classes = ["ConvPreActivation"] * len(convs)
    return [(convs[k][0], ["blocks", list_item_name(classes, k)]) for k in range(len(convs))]


def tensor_plan(cfg: W.ModelConfig):
    """[(p3w tensor name, top-level attribute, class of that attribute's value | None for list items, path below it, var index,
    kind)] for every tensor the engine reads.  kind: 'conv' (HWIO -> OIHW) or 'raw'."""
    plan = []

    def conv(tag, top, sub, cls=None):
        plan.append((f"{tag}/conv/kernel", top, cls, sub, 0, "conv"))

    def dense(tag, top, sub, cls=None):
        plan.append((f"{tag}/dense/kernel", top, cls, sub, 0, "raw"))
        plan.append((f"{tag}/dense/bias", top, cls, sub, 1, "raw"))

    def bn(tag, top, sub, cls=None):
        for k, d in enumerate(("gamma", "beta", "moving_mean", "moving_variance")):
            plan.append((f"{tag}/batch_norm/{d}", top, cls, sub, k, "raw"))

    conv("model/init_conv", "init_board_conv", [], "Conv2D")
    dense("model/init_game_state", "init_game_layer", [], "Dense")
    classes = trunk_block_classes(cfg)
    for i in range(cfg.blocks):
        top = ("blocks", i, classes)
        for tag, sub in inner_paths(cfg, i):
            bn(tag, top, sub + ["norm_layer"])
            conv(tag, top, sub + ["conv"])
        if cfg.is_broadcast(i):
            bclasses = ["ConvPreActivation", "BroadcastPreAct", "ConvPreActivation"]
            dense(f"{W.block_tag(cfg, i)}/01:broadcast", top, ["blocks", list_item_name(bclasses, 1), "dense"])
    ph, vh = "model/policy_head", "model/value_head"
    conv(f"{ph}/conv_policy", "policy_head", ["conv_p"], "PolicyHead")
    conv(f"{ph}/conv_global", "policy_head", ["conv_g"], "PolicyHead")
    bn(f"{ph}/global_pool_bias", "policy_head", ["gpool", "g_norm_layer"], "PolicyHead")
    dense(f"{ph}/global_pool_bias", "policy_head", ["gpool", "dense"], "PolicyHead")
    conv(f"{ph}/conv_moves", "policy_head", ["output_moves"], "PolicyHead")
    dense(f"{ph}/dense_pass", "policy_head", ["output_pass"], "PolicyHead")
    conv(f"{ph}/conv_soft_moves", "policy_head", ["soft_policy_moves"], "PolicyHead")
    dense(f"{ph}/dense_soft_pass", "policy_head", ["soft_policy_pass"], "PolicyHead")
    conv(f"{ph}/conv_optimistic_moves", "policy_head", ["optimistic_policy_moves"], "PolicyHead")
    dense(f"{ph}/dense_optimistic_pass", "policy_head", ["optimistic_policy_pass"], "PolicyHead")
    conv(f"{vh}/conv_value", "value_head", ["conv"], "ValueHead")
    dense(f"{vh}/dense_outcome_pre", "value_head", ["outcome_q_embed"], "ValueHead")
    dense(f"{vh}/dense_outcome", "value_head", ["outcome_q_output"], "ValueHead")
    dense(f"{vh}/dense_mcts_dist", "value_head", ["outcome_mcts_dist"], "ValueHead")
    conv(f"{vh}/ownership", "value_head", ["conv_ownership"], "ValueHead")
    dense(f"{vh}/dense_gamma_pre", "value_head", ["gamma_pre"], "ValueHead")
    dense(f"{vh}/dense_gamma", "value_head", ["gamma_output"], "ValueHead")
    dense(f"{vh}/dense_scores_pre", "value_head", ["score_pre"], "ValueHead")
    dense(f"{vh}/dense_scores", "value_head", ["score_output"], "ValueHead")
    return plan


def model_layer_classes(cfg: W.ModelConfig) -> List[Tuple[str, str]]:
    """(top-level attribute or ('blocks', i), class) of the model's direct sublayers in the order `model.layers` lists them
    (= assignment order in P3achyGoModel.__init__, python/model.py:1152-1190)."""
    out = [("init_board_conv", "Conv2D"), ("init_game_layer", "Dense")]
    out += [(("blocks", i), c) for i, c in enumerate(trunk_block_classes(cfg))]
    out += [("policy_head", "PolicyHead"), ("value_head", "ValueHead"), ("identity", "Activation")]
    return out


def candidate_groups(cfg: W.ModelConfig, top, cls: Optional[str], sub: List[str]) -> List[str]:
    """h5 group paths the object could be stored under (see the module docstring)."""
    tail = "/".join(sub)
    join = lambda *parts: "/".join(p for p in parts if p)
    cands = []
    if isinstance(top, tuple):
        _, i, classes = top
        item = list_item_name(classes, i)
        cands += [join("blocks", item, tail), join("layers", "blocks", item, tail)]
        key = ("blocks", i)
    else:
        cands += [join(top, tail), join("layers", top, tail)]
        key = top
    order = model_layer_classes(cfg)
    names = [c for _, c in order]
    idx = [k for k, (a, _) in enumerate(order) if a == key][0]
    cands.append(join("layers", list_item_name(names, idx), tail))
    return list(dict.fromkeys(cands))


def config_from_keras(cfg_json: dict) -> W.ModelConfig:
    def find(obj):
        if isinstance(obj, dict):
            if obj.get("class_name") == "P3achyGoModel":
                return obj.get("config", {})
            for v in obj.values():
                r = find(v)
                if r:
                    return r
        if isinstance(obj, list):
            for v in obj:
                r = find(v)
                if r:
                    return r
        return None

    c = find(cfg_json)
    if c is None:
        raise ValueError("config.json holds no P3achyGoModel config")
    if c.get("generic_arch") or c.get("is_transformer"):
        raise ValueError("generic / transformer trunks are not supported by the B200 engine")
    if int(c.get("board_len", 19)) != 19:
        raise ValueError("only 19x19 nets are supported")
    return W.ModelConfig(
        c.get("name") or "from_keras", blocks=int(c["num_blocks"]), conv_size=int(c["conv_size"]),
        broadcast_interval=int(c["broadcast_interval"]), inner_bottleneck_layers=int(c["bottleneck_length"]) - 2,  # model.py:1661
        channels=int(c["num_channels"]), bottleneck_channels=int(c["num_bottleneck_channels"]),
        head_channels=int(c["num_head_channels"]), c_val=int(c["c_val"]), trunk_block_type=c.get("trunk_block_type", "btl"),
        num_input_planes=int(c["num_input_planes"]), num_input_features=int(c["num_input_features"]))


def convert_tensors(cfg: W.ModelConfig, h5: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    shapes = W.tensor_shapes(cfg)
    out: Dict[str, np.ndarray] = {}
    for name, top, cls, sub, var, kind in tensor_plan(cfg):
        want = shapes[name]
        h5_shape = (want[2], want[3], want[1], want[0]) if kind == "conv" else want  # Keras HWIO
        hits = []
        for g in candidate_groups(cfg, top, cls, sub):
            key = f"{g}/vars/{var}"
            if key in h5 and tuple(h5[key].shape) == tuple(h5_shape):
                hits.append(key)
        if len(hits) != 1:
            tried = [f"{g}/vars/{var}" for g in candidate_groups(cfg, top, cls, sub)]
            near = [k for k in h5 if sub and sub[-1] in k][:6]
            raise KeyError(f"{name}: expected exactly one of {tried} with shape {h5_shape}; found {hits or 'none'}"
                           f" (similar keys in the file: {near})")
        t = np.asarray(h5[hits[0]], dtype=np.float32)
        out[name] = np.ascontiguousarray(np.transpose(t, (3, 2, 0, 1)) if kind == "conv" else t)
    for name in shapes:  # constants that are not Keras variables
        if name.endswith("/batch_norm/epsilon"):
            out[name] = np.array([1e-3], dtype=np.float32)  # keras.layers.BatchNormalization(epsilon=1e-3), python/model.py:231
        elif name.endswith("/scores"):
            out[name] = (0.05 * np.arange(-400, 400, dtype=np.float32) + 0.025).astype(np.float32)  # python/model.py:1225-1228
    missing = set(shapes) - set(out)
    if missing:
        raise KeyError(f"tensors without a source: {sorted(missing)[:5]}")
    return out


def read_keras_checkpoint(path: str) -> Tuple[W.ModelConfig, Dict[str, np.ndarray]]:
    with zipfile.ZipFile(path, "r") as zf:
        cfg = config_from_keras(json.loads(zf.read("config.json")))
        h5 = minih5.read_h5(zf.read("model.weights.h5"))
    return cfg, convert_tensors(cfg, h5)


def keras_to_p3w(keras_path: str, p3w_path: str) -> W.ModelConfig:
    cfg, tensors = read_keras_checkpoint(keras_path)
    W.save_weights(p3w_path, cfg, tensors)
    return cfg


# ---- fixture writer (tests): a checkpoint with the layout described above, from a `.p3w`-style tensor dict ---------------------
def write_keras_checkpoint(path: str, cfg: W.ModelConfig, tensors: Dict[str, np.ndarray], layout: str = "attributes") -> None:
    """layout 'attributes': objects stored where Keras' sorted attribute walk finds them first (`blocks/...`, `init_board_conv`,
    `init_game_layer` at the root; `policy_head`, `value_head` under `layers/` by class name, as the migrate script's keys show);
    'layers': every direct child under `layers/<snake class>[_k]`.  Extra variables the engine does not read (the heads' unused
    `norm_layer`) are written too, as Keras would."""
    tree: dict = {}

    def put(group: str, var: int, arr: np.ndarray):
        node = tree
        for part in group.split("/") + ["vars"]:
            node = node.setdefault(part, {})
        node[str(var)] = np.ascontiguousarray(arr, dtype=np.float32)

    order = model_layer_classes(cfg)
    names = [c for _, c in order]
    for name, top, cls, sub, var, kind in tensor_plan(cfg):
        cands = candidate_groups(cfg, top, cls, sub)
        if layout == "layers":
            group = cands[-1]
        elif isinstance(top, tuple) or top in ("init_board_conv", "init_game_layer"):
            group = cands[0]
        else:
            group = cands[-1]  # policy_head / value_head: reached through model.layers first
        t = tensors[name]
        put(group, var, np.transpose(t, (2, 3, 1, 0)) if kind == "conv" else t)
    Ch = cfg.head_channels
    for head in ("PolicyHead", "ValueHead"):  # BatchNormalization attributes that exist but are not used by call()
        g = "layers/" + list_item_name(names, names.index(head)) + "/norm_layer"
        for k, v in enumerate((np.ones(Ch), np.zeros(Ch), np.zeros(Ch), np.ones(Ch))):
            put(g, k, v.astype(np.float32))
    config = {"module": "model", "class_name": "P3achyGoModel", "registered_name": "p3achygo>P3achyGoModel",
              "config": {"board_len": 19, "num_input_planes": cfg.num_input_planes, "num_input_features": cfg.num_input_features,
                         "num_blocks": cfg.blocks, "num_channels": cfg.channels, "num_bottleneck_channels": cfg.bottleneck_channels,
                         "num_head_channels": cfg.head_channels, "c_val": cfg.c_val, "bottleneck_length": cfg.inner_bottleneck_layers + 2,
                         "conv_size": cfg.conv_size, "broadcast_interval": cfg.broadcast_interval,
                         "trunk_block_type": cfg.trunk_block_type, "generic_arch": None, "c_l2": 1e-4, "name": cfg.name}}
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as zf:
        zf.writestr("metadata.json", json.dumps({"keras_version": "3.3.3", "date_saved": "fixture"}))
        zf.writestr("config.json", json.dumps(config))
        zf.writestr("model.weights.h5", minih5.write_h5(tree))


if __name__ == "__main__":
    if len(sys.argv) != 3:
        print(__doc__)
        sys.exit(2)
    c = keras_to_p3w(sys.argv[1], sys.argv[2])
    print(f"wrote {sys.argv[2]}: {c}")
