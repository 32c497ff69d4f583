"""Flat binary weight files ("P3W1") and net shape configs for the B200 leaf evaluator.

The reference's "exported weight format" is the tag tree of ``python/export_weights.py:16-90``
(an HDF5 layout with no reader anywhere in ``cc/``).  HDF5 is not available
here and that exporter is stale against the current ``python/model.py`` (it never writes BN
gamma and misses the v1 head layers), so this module defines a flat little-endian mirror of
the same tree:

    bytes 0..7   magic  b"P3ACHYW1"
    u32          n_meta, then n_meta x { u16 len, utf-8 key, i32 value }
    u32          n_tensors, then n_tensors x
                 { u16 len, utf-8 name, u8 ndim, u32 dims[ndim], pad to 4, f32 data[prod(dims)] }

Tensor names are ``/``-joined tags following ``export_weights.py`` (``ModelTags``,
``BlockTags``, ``LayerTags``, ``DatasetTags``): e.g.
``model/trunk/03:bottleneck_res/01:conv_block/conv/kernel``.  Layout conventions follow that
exporter: conv kernels are OIHW ``(C', C, k, k)`` (``export_weights.py:134-137``), dense kernels
are Keras ``(in, out)``, batch-norm carries ``gamma, beta, moving_mean, moving_variance`` and a
1-element ``epsilon``.  Metadata keys are ``MetadataTags`` (``export_weights.py:27-39``) plus
``broadcast_interval`` and ``trunk_block_type`` (0 = btl, 1 = classic), which the stale exporter
cannot express.

Net shapes mirror ``python/model_config.py:62-128``.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict, Tuple

import numpy as np

MAGIC = b"P3ACHYW1"
BOARD_LEN = 19
NUM_LOCS = 361
TRUNK_BTL, TRUNK_CLASSIC, TRUNK_NBT = 0, 1, 2


@dataclasses.dataclass(frozen=True)
class ModelConfig:
    """Mirror of ``ModelConfig`` in ``python/model_config.py:21-60`` (conv/FC nets only)."""

    name: str
    blocks: int = 16
    conv_size: int = 3
    broadcast_interval: int = 8
    inner_bottleneck_layers: int = 2
    channels: int = 128
    bottleneck_channels: int = 64
    head_channels: int = 32
    c_val: int = 64
    trunk_block_type: str = "btl"
    num_input_planes: int = 15
    num_input_features: int = 8

    def is_broadcast(self, i: int) -> bool:
        # python/model.py:1003
        return i % self.broadcast_interval == self.broadcast_interval - 1

    def flops_per_position(self) -> float:
        """2*MAC of every conv / dense of the forward pass."""
        C, Cb, Ch, Cv, P = self.channels, self.bottleneck_channels, self.head_channels, self.c_val, NUM_LOCS
        k = self.conv_size
        mac = (k + 2) ** 2 * self.num_input_planes * C * P + self.num_input_features * C
        for i in range(self.blocks):
            if self.is_broadcast(i):
                mac += 2 * C * C * P + C * P * P
            elif self.trunk_block_type == "btl":
                mac += (2 * C * Cb + self.inner_bottleneck_layers * k * k * Cb * Cb) * P
            elif self.trunk_block_type == "nbt":  # NbtResidualBlock: 1x1, two classic blocks of two convs at Cb, 1x1
                mac += (2 * C * Cb + 4 * k * k * Cb * Cb) * P
            else:
                mac += 2 * k * k * C * C * P
        mac += 2 * C * Ch * P + 4 * Ch * P + 2 * Ch * (Ch + 4)
        mac += C * Ch * P + Ch * P + 2 * 2 * Ch * Cv + Cv * (14 + 51 + 1) + 2 * Ch * Cv + 2 * 800 * Cv
        return 2.0 * mac

    def tower_flops_per_position(self) -> float:
        C, Cb, P = self.channels, self.bottleneck_channels, NUM_LOCS
        k = self.conv_size
        mac = (k + 2) ** 2 * self.num_input_planes * C * P + self.num_input_features * C
        for i in range(self.blocks):
            if self.is_broadcast(i):
                mac += 2 * C * C * P + C * P * P
            elif self.trunk_block_type == "btl":
                mac += (2 * C * Cb + self.inner_bottleneck_layers * k * k * Cb * Cb) * P
            elif self.trunk_block_type == "nbt":
                mac += (2 * C * Cb + 4 * k * k * Cb * Cb) * P
            else:
                mac += 2 * k * k * C * C * P
        return 2.0 * mac


# python/model_config.py:62-128
CONFIGS: Dict[str, ModelConfig] = {
    "tiny": ModelConfig("tiny", blocks=6, broadcast_interval=4, inner_bottleneck_layers=1, channels=16,
                        bottleneck_channels=8, head_channels=8, c_val=16),
    "small": ModelConfig("small"),
    "b10c128btl3": ModelConfig("b10c128btl3", blocks=10, broadcast_interval=4, inner_bottleneck_layers=3,
                               channels=128, bottleneck_channels=64),
    "b5c256btl3": ModelConfig("b5c256btl3", blocks=5, broadcast_interval=2, inner_bottleneck_layers=3,
                              channels=256, bottleneck_channels=128),
    "b12c256btl3": ModelConfig("b12c256btl3", blocks=12, broadcast_interval=5, inner_bottleneck_layers=3,
                               channels=256, bottleneck_channels=128),
    "b14c384btl3": ModelConfig("b14c384btl3", blocks=14, broadcast_interval=6, inner_bottleneck_layers=3,
                               channels=384, bottleneck_channels=192, head_channels=32, c_val=80),
    "b15c192_classic": ModelConfig("b15c192_classic", blocks=15, broadcast_interval=6, channels=192,
                                   head_channels=32, c_val=80, trunk_block_type="classic"),
    # nested-bottleneck nets (python/model_config.py:131-163, NbtResidualBlock python/model.py:431-470)
    "b8c128nbt": ModelConfig("b8c128nbt", blocks=8, broadcast_interval=3, channels=128, bottleneck_channels=64,
                             head_channels=32, trunk_block_type="nbt"),
    "b12c256nbt": ModelConfig("b12c256nbt", blocks=12, broadcast_interval=3, channels=256, bottleneck_channels=128,
                              head_channels=32, c_val=80, trunk_block_type="nbt"),
    "b10c384nbt": ModelConfig("b10c384nbt", blocks=10, broadcast_interval=4, channels=384, bottleneck_channels=192,
                              head_channels=32, c_val=80, trunk_block_type="nbt"),
}


def config_from_str(name: str) -> ModelConfig:
    """``ModelConfig.from_str``, python/model_config.py:174-196."""
    if name not in CONFIGS:
        raise Exception("Unknown Model Config")
    return CONFIGS[name]


# ---------------------------------------------------------------------------------------------
# tensor naming (tag tree of python/export_weights.py)
# ---------------------------------------------------------------------------------------------
def block_tag(cfg: ModelConfig, i: int) -> str:
    if cfg.is_broadcast(i):
        return f"model/trunk/{i:02d}:broadcast_res"
    if cfg.trunk_block_type == "btl":
        return f"model/trunk/{i:02d}:bottleneck_res"
    if cfg.trunk_block_type == "nbt":
        return f"model/trunk/{i:02d}:nbt_res"
    return f"model/trunk/{i:02d}:classic_res"


def block_convs(cfg: ModelConfig, i: int):
    """[(tag, cin, cout, ksize)] of the ConvPreActivation layers of trunk block i
    (python/model.py:330-427, 490-630). The broadcast block's middle layer is listed separately."""
    C, Cb, k = cfg.channels, cfg.bottleneck_channels, cfg.conv_size
    t = block_tag(cfg, i)
    if cfg.is_broadcast(i):
        return [(f"{t}/00:conv_block", C, C, 1), (f"{t}/02:conv_block", C, C, 1)]
    if cfg.trunk_block_type == "btl":
        out = [(f"{t}/00:conv_block", C, Cb, 1)]
        for j in range(cfg.inner_bottleneck_layers):
            out.append((f"{t}/{j + 1:02d}:conv_block", Cb, Cb, k))
        out.append((f"{t}/{cfg.inner_bottleneck_layers + 1:02d}:conv_block", Cb, C, 1))
        return out
    if cfg.trunk_block_type == "nbt":
        # reduce 1x1, nbt_res0 = two convs, nbt_res1 = two convs (each with its own inner residual), expand 1x1; flattened 0..5
        return ([(f"{t}/00:conv_block", C, Cb, 1)] + [(f"{t}/{j:02d}:conv_block", Cb, Cb, k) for j in range(1, 5)] +
                [(f"{t}/05:conv_block", Cb, C, 1)])
    return [(f"{t}/00:conv_block", C, C, k), (f"{t}/01:conv_block", C, C, k)]


def tensor_shapes(cfg: ModelConfig) -> Dict[str, Tuple[int, ...]]:
    """name -> shape for every tensor of a net of this config."""
    C, Ch, Cv = cfg.channels, cfg.head_channels, cfg.c_val
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(tag, cin, cout, k):
        s[f"{tag}/conv/kernel"] = (cout, cin, k, k)

    def dense(tag, cin, cout):
        s[f"{tag}/dense/kernel"] = (cin, cout)
        s[f"{tag}/dense/bias"] = (cout,)

    def bn(tag, c):
        for d in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"{tag}/batch_norm/{d}"] = (c,)
        s[f"{tag}/batch_norm/epsilon"] = (1,)

    conv("model/init_conv", cfg.num_input_planes, C, cfg.conv_size + 2)  # model.py:1152-1160
    dense("model/init_game_state", cfg.num_input_features, C)            # model.py:1161
    for i in range(cfg.blocks):
        for tag, cin, cout, k in block_convs(cfg, i):
            bn(tag, cin)
            conv(tag, cin, cout, k)
        if cfg.is_broadcast(i):
            dense(f"{block_tag(cfg, i)}/01:broadcast", NUM_LOCS, NUM_LOCS)  # model.py:546
    ph = "model/policy_head"                                              # model.py:741-779
    conv(f"{ph}/conv_policy", C, Ch, 1)
    conv(f"{ph}/conv_global", C, Ch, 1)
    bn(f"{ph}/global_pool_bias", Ch)
    dense(f"{ph}/global_pool_bias", 2 * Ch, Ch)
    conv(f"{ph}/conv_moves", Ch, 2, 1)
    dense(f"{ph}/dense_pass", 2 * Ch, 2)
    conv(f"{ph}/conv_soft_moves", Ch, 1, 1)
    dense(f"{ph}/dense_soft_pass", 2 * Ch, 1)
    conv(f"{ph}/conv_optimistic_moves", Ch, 1, 1)
    dense(f"{ph}/dense_optimistic_pass", 2 * Ch, 1)
    vh = "model/value_head"                                               # model.py:846-885
    conv(f"{vh}/conv_value", C, Ch, 1)
    dense(f"{vh}/dense_outcome_pre", 2 * Ch, Cv)
    dense(f"{vh}/dense_outcome", Cv, 14)
    dense(f"{vh}/dense_mcts_dist", Cv, 51)
    conv(f"{vh}/ownership", Ch, 1, 1)
    dense(f"{vh}/dense_gamma_pre", 2 * Ch, Cv)
    dense(f"{vh}/dense_gamma", Cv, 1)
    dense(f"{vh}/dense_scores_pre", 2 * Ch + 1, Cv)
    dense(f"{vh}/dense_scores", Cv, 1)
    s[f"{vh}/scores"] = (800,)
    return s


def config_meta(cfg: ModelConfig) -> Dict[str, int]:
    """MetadataTags of python/export_weights.py:27-39,104-116 (+ two keys it cannot express)."""
    return {
        "ninput_planes": cfg.num_input_planes,
        "ninput_features": cfg.num_input_features,
        "nlayers": cfg.blocks,
        "nchannels": cfg.channels,
        "nbtl_channels": cfg.bottleneck_channels,
        "nhead_channels": cfg.head_channels,
        "nval_channels": cfg.c_val,
        "nbtl": cfg.inner_bottleneck_layers,
        "board_len": BOARD_LEN,
        "conv_size": cfg.conv_size,
        "broadcast_interval": cfg.broadcast_interval,
        "trunk_block_type": {"btl": TRUNK_BTL, "classic": TRUNK_CLASSIC, "nbt": TRUNK_NBT}[cfg.trunk_block_type],
    }


def config_from_meta(meta: Dict[str, int], name: str = "from_file") -> ModelConfig:
    return ModelConfig(
        name, blocks=meta["nlayers"], conv_size=meta.get("conv_size", 3),
        broadcast_interval=meta["broadcast_interval"], inner_bottleneck_layers=meta["nbtl"],
        channels=meta["nchannels"], bottleneck_channels=meta["nbtl_channels"],
        head_channels=meta["nhead_channels"], c_val=meta["nval_channels"],
        trunk_block_type={TRUNK_BTL: "btl", TRUNK_CLASSIC: "classic", TRUNK_NBT: "nbt"}[meta["trunk_block_type"]],
        num_input_planes=meta["ninput_planes"], num_input_features=meta["ninput_features"])


def synthetic_weights(cfg: ModelConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded random-init weights of this architecture (there are no trained checkpoints in the
    reference repo and no network): kernels ~ N(0, gain/fan_in), BN statistics randomised so BN is
    not the identity, and a non-zero ``dense_gamma`` (Keras zero-inits it, model.py:872) so the
    score head is exercised."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in tensor_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if name.endswith("/scores"):
            # model.py:1225-1228
            t = (0.05 * np.arange(-400, 400, dtype=np.float32) + 0.025).astype(np.float32)
        elif leaf == "kernel":
            # gains chosen so a random net stays in a trained net's numeric range (trunk O(1),
            # logits O(1..10)); otherwise every softmax saturates and parity would be vacuous.
            fan_in = shape[1] * shape[2] * shape[3] if len(shape) == 4 else shape[0]
            gain = 1.0
            if "/trunk/" in name and "conv_block" in name:
                last = int(name.split("/")[3].split(":")[0])
                n_convs = len(block_convs(cfg, int(name.split("/")[2].split(":")[0])))
                is_last = last == (2 if "broadcast_res" in name else n_convs - 1) or ("nbt_res" in name and last in (2, 4))
                gain = 0.1 if is_last else 1.5
            elif "broadcast/dense" in name:
                gain = 0.5
            elif "dense_scores_pre" in name:
                gain = 8.0
            elif any(t in name for t in ("conv_moves", "conv_soft_moves", "conv_optimistic_moves",
                                         "dense_outcome/", "dense_mcts_dist/", "dense_scores/")):
                gain = 16.0
            elif "policy_head" in name or "value_head" in name:
                gain = 1.0
            t = rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(gain / fan_in))
        elif leaf == "bias":
            t = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
        elif leaf == "gamma":
            t = rng.uniform(0.7, 1.3, shape).astype(np.float32)
        elif leaf == "beta":
            t = rng.uniform(-0.3, 0.3, shape).astype(np.float32)
        elif leaf == "moving_mean":
            t = rng.uniform(-0.3, 0.3, shape).astype(np.float32)
        elif leaf == "moving_variance":
            t = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif leaf == "epsilon":
            t = np.array([1e-3], dtype=np.float32)  # model.py:231
        else:
            raise AssertionError(name)
        out[name] = np.ascontiguousarray(t, dtype=np.float32)
    return out


def save_weights(path: str, cfg: ModelConfig, tensors: Dict[str, np.ndarray]) -> None:
    expected = tensor_shapes(cfg)
    missing = set(expected) - set(tensors)
    if missing:
        raise ValueError(f"missing tensors: {sorted(missing)[:4]} ...")
    meta = config_meta(cfg)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(meta)))
        for k, v in meta.items():
            kb = k.encode()
            f.write(struct.pack("<H", len(kb)) + kb + struct.pack("<i", int(v)))
        f.write(struct.pack("<I", len(expected)))
        for name, shape in expected.items():
            t = np.ascontiguousarray(tensors[name], dtype="<f4")
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: shape {t.shape} != {shape}")
            nb = name.encode()
            f.write(struct.pack("<H", len(nb)) + nb + struct.pack("<B", t.ndim))
            f.write(struct.pack(f"<{t.ndim}I", *t.shape))
            f.write(b"\0" * ((-f.tell()) % 4))
            f.write(t.tobytes())


def load_weights(path: str):
    """-> (ModelConfig, {name: float32 ndarray})"""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a P3W1 weight file")
    off = 8
    (n_meta,) = struct.unpack_from("<I", buf, off); off += 4
    meta: Dict[str, int] = {}
    for _ in range(n_meta):
        (ln,) = struct.unpack_from("<H", buf, off); off += 2
        key = buf[off:off + ln].decode(); off += ln
        (val,) = struct.unpack_from("<i", buf, off); off += 4
        meta[key] = val
    (n_t,) = struct.unpack_from("<I", buf, off); off += 4
    tensors: Dict[str, np.ndarray] = {}
    for _ in range(n_t):
        (ln,) = struct.unpack_from("<H", buf, off); off += 2
        name = buf[off:off + ln].decode(); off += ln
        (nd,) = struct.unpack_from("<B", buf, off); off += 1
        dims = struct.unpack_from(f"<{nd}I", buf, off); off += 4 * nd
        off += (-off) % 4
        cnt = int(np.prod(dims)) if nd else 1
        tensors[name] = np.frombuffer(buf, dtype="<f4", count=cnt, offset=off).reshape(dims).copy()
        off += 4 * cnt
    return config_from_meta(meta), tensors


def make_synthetic_weight_file(path: str, config_name: str, seed: int = 0) -> ModelConfig:
    cfg = config_from_str(config_name)
    save_weights(path, cfg, synthetic_weights(cfg, seed))
    return cfg
