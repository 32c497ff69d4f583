"""Real-checkpoint importer: `.keras` (zip of config.json + model.weights.h5, Keras-3 attribute-path keys,
python/scripts/migrate_checkpoint.py:1-48, python/rl_loop/model_utils.py:197-204) -> `.p3w`.

No Keras / h5py in this image: the checkpoint fixtures are written by the importer module's own writer with the naming rules of
Keras' saving_lib (documented in tools/keras_to_p3w.py); the HDF5 layer (tools/minih5.py) is checked separately."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import keras_to_p3w as K  # noqa: E402
import minih5  # noqa: E402
from p3achygo_b200 import weights as W  # noqa: E402


def test_minih5_round_trip_and_structures():
    rng = np.random.default_rng(0)
    tree = {"layers": {"value_head": {"outcome_q_embed": {"vars": {"0": rng.standard_normal((64, 80)).astype(np.float32),
                                                                   "1": rng.standard_normal(80).astype(np.float32)}}}},
            "blocks": {f"bottleneck_residual_conv_block{'' if i == 0 else '_' + str(i)}": {"vars": {"0": np.arange(i + 1, dtype=np.float32)}}
                       for i in range(40)},          # > 8 links: several group nodes under one B-tree node
            "vars": {}, "scalar": np.array(3.5, dtype=np.float64), "ints": np.arange(10, dtype=np.int32).reshape(2, 5),
            "half": np.array([1.5, -2.25], dtype=np.float16), "empty": np.zeros((0, 3), dtype=np.float32)}
    img = minih5.write_h5(tree)
    assert img[:8] == b"\x89HDF\r\n\x1a\n" and img[8] == 0
    flat = minih5.read_h5(img)

    def walk(n, p=""):
        for k, v in n.items():
            q = f"{p}/{k}" if p else k
            if isinstance(v, dict):
                yield from walk(v, q)
            else:
                yield q, np.asarray(v)

    want = dict(walk(tree))
    assert set(want) == set(flat)
    for k in want:
        assert want[k].dtype == flat[k].dtype and want[k].shape == flat[k].shape and np.array_equal(want[k], flat[k]), k
    with pytest.raises(minih5.H5FormatError):
        minih5.read_h5(b"not an hdf5 file at all")
    with pytest.raises(minih5.H5FormatError):
        minih5.read_h5(img[:8] + b"\x02" + img[9:])       # superblock version 2: refused loudly, not misread


@pytest.mark.parametrize("config", ["tiny", "b10c128btl3", "b15c192_classic", "b8c128nbt"])
@pytest.mark.parametrize("layout", ["attributes", "layers"])
def test_keras_checkpoint_round_trip(config, layout, tmp_path):
    """tensors -> .keras (Keras naming, HWIO kernels) -> importer -> .p3w -> load: every tensor bit-identical, config recovered."""
    cfg = W.config_from_str(config)
    tensors = W.synthetic_weights(cfg, 3)
    kpath, ppath = os.path.join(tmp_path, "model_0001.keras"), os.path.join(tmp_path, "model_0001.p3w")
    K.write_keras_checkpoint(kpath, cfg, tensors, layout=layout)
    got_cfg = K.keras_to_p3w(kpath, ppath)
    for f in ("blocks", "channels", "bottleneck_channels", "head_channels", "c_val", "inner_bottleneck_layers", "broadcast_interval",
              "trunk_block_type", "conv_size", "num_input_planes", "num_input_features"):
        assert getattr(got_cfg, f) == getattr(cfg, f), f
    lcfg, loaded = W.load_weights(ppath)
    assert set(loaded) == set(tensors)
    for name, t in tensors.items():
        assert np.array_equal(loaded[name], t), name


def test_keras_key_layout_follows_the_reference_scripts(tmp_path):
    """The one h5 key the reference itself documents (migrate_checkpoint.py:5-9): layers/value_head/outcome_q_embed/vars/{0,1}."""
    import zipfile
    cfg = W.config_from_str("tiny")
    K.write_keras_checkpoint(os.path.join(tmp_path, "m.keras"), cfg, W.synthetic_weights(cfg, 0))
    with zipfile.ZipFile(os.path.join(tmp_path, "m.keras")) as zf:
        assert {"config.json", "model.weights.h5", "metadata.json"} <= set(zf.namelist())
        keys = minih5.read_h5(zf.read("model.weights.h5"))
    assert "layers/value_head/outcome_q_embed/vars/0" in keys and "layers/value_head/outcome_q_embed/vars/1" in keys
    assert keys["blocks/bottleneck_residual_conv_block_1/blocks/conv_pre_activation_2/conv/vars/0"].shape[:2] == (1, 1)   # HWIO
    assert "blocks/broadcast_residual_block/blocks/broadcast_pre_act/dense/vars/0" in keys


def test_importer_fails_loudly_on_a_missing_or_misshapen_tensor(tmp_path):
    import zipfile
    cfg = W.config_from_str("tiny")
    tensors = W.synthetic_weights(cfg, 0)
    path = os.path.join(tmp_path, "m.keras")
    K.write_keras_checkpoint(path, cfg, tensors)
    with zipfile.ZipFile(path) as zf:
        h5 = minih5.read_h5(zf.read("model.weights.h5"))
    del h5["layers/value_head/outcome_mcts_dist/vars/0"]      # a pre-mcts_dist checkpoint (what migrate_checkpoint.py exists for)
    with pytest.raises(KeyError) as ei:
        K.convert_tensors(cfg, h5)
    assert "dense_mcts_dist" in str(ei.value)
