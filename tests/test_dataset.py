"""The TFRecord / tf.Example position reader (p3achygo_b200/host/go_dataset.{h,cc}, mirror of nn::GoDataset,
cc/nn/engine/go_dataset.cc:32-123) and the nn::Benchmark loop over it with the reference's accuracy statistics
(cc/nn/engine/benchmark_engine.cc:24-109).

CPU: the committed fixtures (tests/golden/positions_8.tfrecord plain, positions_64.tfrecord.zz zlib; written by
tests/golden/make_tfrecord_fixture.py from the golden positions) read back field for field; where /root/reference exists, the
reference's own python/test_data/mixed_schema.tfrecord against the values its generator (python/test_data/generate.py) writes.
GPU: the dataset benchmark's statistics equal the ones computed here from the engine's results.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)
vp = ctypes.c_void_p


def _host():
    h = ctypes.CDLL(os.path.join(ROOT, "p3achygo_b200", "libp3host.so"))
    h.p3_host_dataset_read.restype = ctypes.c_longlong
    h.p3_host_dataset_read.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, ctypes.POINTER(ctypes.c_longlong)]
    h.p3_host_benchmark_dataset.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                            ctypes.c_int, vp]
    return h


def _read(path, batch, cap=128):
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    f = np.zeros(cap, dtype=GO_FEATURES_DTYPE)
    pol = np.zeros((cap, 362), dtype=np.float32)
    sm = np.zeros(cap, dtype=np.float32)
    dw = np.zeros(cap, dtype=np.uint8)
    nb = ctypes.c_longlong(0)
    n = _host().p3_host_dataset_read(path.encode(), batch, cap, f.ctypes.data_as(vp), pol.ctypes.data_as(vp), sm.ctypes.data_as(vp),
                                     dw.ctypes.data_as(vp), ctypes.byref(nb))
    return int(n), int(nb.value), f[:n], pol[:n], sm[:n], dw[:n]


@pytest.mark.parametrize("name,count", [("positions_8.tfrecord", 8), ("positions_64.tfrecord.zz", 64)])
def test_fixture_round_trip(name, count, golden_positions):
    import make_tfrecord_fixture as M
    n, nb, f, pol, sm, dw = _read(os.path.join(GOLDEN, name), 16)
    assert n == count and nb == (count + 15) // 16          # trailing partial batch kept (go_dataset.cc:125-126)
    src = golden_positions["feats"][:count]
    policy, margin = M.labels(64, golden_positions["legal"][:64])
    for field in ("bsize", "color", "komi", "board", "stones_atari", "stones_two_liberties", "stones_three_liberties", "stones_laddered"):
        assert np.array_equal(f[field], src[field]), field
    # last moves travel as int16 encodings and come back through game::AsLoc (loc.h:29-31): pass 361 -> {19, 0}; the no-op -1 -> {0, -1}
    want = src["last_moves"].copy()
    noop = want[:, :, 0] < 0
    want[noop] = (0, -1)
    assert np.array_equal(f["last_moves"], want)
    assert np.array_equal(pol, policy[:count]) and np.array_equal(sm, margin[:count])
    assert np.array_equal(dw, (margin[:count] >= 0).astype(np.uint8))


@pytest.mark.skipif(not os.path.exists("/root/reference/python/test_data/mixed_schema.tfrecord"), reason="reference tree not present")
def test_reads_the_reference_test_file():
    """python/test_data/generate.py: 3 old-schema + 3 new-schema records, empty boards, komi 6.5, one-hot policy at move 0."""
    n, nb, f, pol, sm, dw = _read("/root/reference/python/test_data/mixed_schema.tfrecord", 4)
    assert (n, nb) == (6, 2)
    assert f["color"].tolist() == [1, -1, 1, 1, -1, 1] and np.all(f["komi"] == 6.5) and np.all(f["bsize"] == 19)
    assert not f["board"].any() and not f["stones_laddered"].any()
    assert sm.tolist() == [5.5, -5.5, 0.5, 3.5, -3.5, 0.0] and dw.tolist() == [1, 0, 1, 1, 0, 1]
    assert np.all(pol.argmax(axis=1) == 0) and np.all(pol.sum(axis=1) == 1.0)
    assert np.all(f["last_moves"] == np.array([0, -1]))      # AsLoc(-1) under C++ division


@pytest.mark.gpu
def test_dataset_benchmark_statistics(weight_dir, golden_positions):
    import make_tfrecord_fixture as M
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir("b10c128btl3")
    out = np.zeros(7, dtype=np.float64)
    ds = os.path.join(GOLDEN, "positions_64.tfrecord.zz")
    _host().p3_host_benchmark_dataset(path.encode(), 0, 16, 1, E.PRECISION_BF16, ds.encode(), 3, out.ctypes.data_as(vp))
    n, nb, f, pol, sm, dw = _read(ds, 16)
    eng = E.CreateEngine(E.Kind.kB200, path, 16, 1, precision=E.PRECISION_BF16)
    pl, ol, pp, op, sd = [], [], [], [], []
    for lo in range(0, 64, 16):
        for b in range(16):
            eng.LoadBatch(b, f[lo + b])
        eng.RunInference()
        for b in range(16):
            r = eng.GetBatch(b)
            mv = int(pol[lo + b].argmax())
            won = int(dw[lo + b])
            pl.append(-np.log(np.float32(r["move_probs"][mv])))
            ol.append(-np.log(np.float32(r["value_probs"][won])))
            pp.append(float(int(np.argmax(r["move_probs"])) == mv))
            op.append(float(int(np.argmax(r["value_probs"])) == won))
            sd.append(abs(float(sm[lo + b]) - int(int(np.argmax(r["score_probs"])) + 0.5 - 400)))
    eng.close()
    assert out[0] == 64 and out[1] > 0
    assert np.allclose(out[2:7], [np.mean(pl), np.mean(ol), np.mean(pp), np.mean(op), np.mean(sd)], rtol=1e-5, atol=1e-6)
