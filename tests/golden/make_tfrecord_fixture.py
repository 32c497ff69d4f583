"""Writes tests/golden/positions_8.tfrecord (plain, 8 records) and positions_64.tfrecord.zz (zlib, 64): golden positions as
tf.train.Example records with the fields cc/recorder/make_tf_example.h:39-49 writes and nn::GoDataset reads
(cc/nn/engine/go_dataset.cc:60-120).  No TensorFlow: the Example wire format is spelled out here (map<string, Feature>,
BytesList / FloatList), the framing comes from libp3host.so (p3_host_tfrecord_frame).  Labels are seeded: a one-hot policy on a
legal move and a score margin.

    python tests/golden/make_tfrecord_fixture.py
"""
import ctypes
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def ld(field: int, payload: bytes) -> bytes:  # length-delimited field
    return varint((field << 3) | 2) + varint(len(payload)) + payload


def bytes_feature(data: bytes) -> bytes:
    return ld(1, ld(1, data))  # Feature.bytes_list { value }


def float_feature(v: float) -> bytes:
    return ld(2, ld(1, np.float32(v).tobytes()))  # Feature.float_list { value (packed) }


def example(fields: dict) -> bytes:
    feats = b"".join(ld(1, ld(1, k.encode()) + ld(2, v)) for k, v in fields.items())  # map entries
    return ld(1, feats)  # Example.features


def labels(n: int, legal: np.ndarray, seed: int = 5):
    rng = np.random.default_rng(seed)
    policy = np.zeros((n, 362), dtype=np.float32)
    for i in range(n):
        policy[i, int(rng.choice(np.flatnonzero(legal[i])))] = 1.0
    margin = (rng.integers(-40, 41, size=n) + 0.5).astype(np.float32)
    return policy, margin


def main():
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    z = np.load(os.path.join(HERE, "positions.npz"))
    feats = np.ascontiguousarray(z["feats"]).view(GO_FEATURES_DTYPE).reshape(-1)[:64]
    policy, margin = labels(64, z["legal"][:64])
    host = ctypes.CDLL(os.path.join(ROOT, "p3achygo_b200", "libp3host.so"))
    host.p3_host_tfrecord_frame.restype = ctypes.c_longlong
    host.p3_host_tfrecord_frame.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong]
    buf = np.zeros(1 << 20, dtype=np.uint8)
    at = 0
    at8 = 0
    for i in range(64):
        f = feats[i]
        lm = np.array([(-1 if m[0] < 0 else 361 if m[0] == 19 else m[0] * 19 + m[1]) for m in f["last_moves"]], dtype=np.int16)
        rec = example({
            "bsize": bytes_feature(np.uint8(19).tobytes()), "board": bytes_feature(f["board"].tobytes()),
            "last_moves": bytes_feature(lm.tobytes()), "stones_atari": bytes_feature(f["stones_atari"].tobytes()),
            "stones_two_liberties": bytes_feature(f["stones_two_liberties"].tobytes()),
            "stones_three_liberties": bytes_feature(f["stones_three_liberties"].tobytes()),
            "stones_in_ladder": bytes_feature(f["stones_laddered"].tobytes()), "color": bytes_feature(np.int8(f["color"]).tobytes()),
            "komi": float_feature(float(f["komi"])), "own": bytes_feature(np.zeros(361, dtype=np.int8).tobytes()),
            "pi": bytes_feature(policy[i].tobytes()), "pi_aux": bytes_feature(np.int16(361).tobytes()),
            "score_margin": float_feature(float(margin[i])), "q6": float_feature(0.0), "q16": float_feature(0.0), "q50": float_feature(0.0),
        })
        at = host.p3_host_tfrecord_frame(rec, len(rec), buf.ctypes.data_as(ctypes.c_void_p), at, len(buf))
        assert at > 0
        if i == 7:
            at8 = at
    raw = buf[:at].tobytes()
    open(os.path.join(HERE, "positions_8.tfrecord"), "wb").write(raw[:at8])
    open(os.path.join(HERE, "positions_64.tfrecord.zz"), "wb").write(zlib.compress(raw, 6))
    print(len(raw), "bytes,", len(zlib.compress(raw, 6)), "compressed")


if __name__ == "__main__":
    main()
