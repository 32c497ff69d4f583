"""CPU: the C-ABI library loads and exports every symbol include/p3_b200.h declares, struct layouts mirror the
reference structs, compute entry points fail loudly without a GPU (no CPU fallback), and the host-side mirror of the
reference factory behaves like cc/nn/engine/engine_factory.cc:16-73."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "p3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p3_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from p3achygo_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in include/p3_b200.h but not exported by libp3b200.so"
    assert set(declared) == set(_lib.EXPORTS), set(declared) ^ set(_lib.EXPORTS)
    assert b"sm_100a" in _lib.lib.p3_version()


def test_struct_layouts_mirror_reference():
    """sizes / offsets of nn::GoFeatures (1860 B) and nn::NNInferResult (7568 B, opt_move_probs 16-byte aligned),
    SURVEY.md 8a / 8b; the reference driver static_asserts the same size against the real struct."""
    from p3achygo_b200._lib import AUX_RESULT_DTYPE, GO_FEATURES_DTYPE, INFER_RESULT_DTYPE, GoFeatures
    assert ctypes.sizeof(GoFeatures) == 1860 and GO_FEATURES_DTYPE.itemsize == 1860
    assert GoFeatures.board.offset == 12 and GoFeatures.last_moves.offset == 376
    assert GoFeatures.stones_atari.offset == 416 and GoFeatures.stones_two_liberties.offset == 777
    assert GoFeatures.stones_three_liberties.offset == 1138 and GoFeatures.stones_laddered.offset == 1499
    assert INFER_RESULT_DTYPE.itemsize == 7568
    assert INFER_RESULT_DTYPE.fields["opt_move_probs"][1] == 6112 and INFER_RESULT_DTYPE.fields["opt_move_probs"][1] % 16 == 0
    assert INFER_RESULT_DTYPE.fields["err2_outcome"][1] == 7560
    assert AUX_RESULT_DTYPE.itemsize == 4 * (362 * 3 + 2 + 800 + 1 + 12 + 51 * 2 + 361 + 3)


def test_host_library_loads():
    host = ctypes.CDLL(os.path.join(ROOT, "p3achygo_b200", "libp3host.so"))
    for sym in ("p3_host_benchmark", "p3_host_benchmark_pipelined", "p3_host_benchmark_games", "p3_host_iface_sync_test", "p3_host_iface_run",
                "p3_host_iface_run_banks", "p3_host_iface_run_games", "p3_host_dataset_read", "p3_host_benchmark_dataset",
                "p3_host_tfrecord_frame"):
        assert hasattr(host, sym), sym


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(tmp_path):
    """Without a CUDA device every compute entry point must fail with P3_ERR_NO_DEVICE — never silently compute."""
    from p3achygo_b200 import engine as E
    from p3achygo_b200 import weights as W
    from p3achygo_b200._lib import GO_FEATURES_DTYPE, P3_ERR_NO_DEVICE
    with pytest.raises(E.P3Error) as ei:
        E.encode_features(np.zeros(1, dtype=GO_FEATURES_DTYPE))
    assert ei.value.code == P3_ERR_NO_DEVICE
    with pytest.raises(E.P3Error) as ei:
        E.board_liberties(np.zeros((1, 361), np.int8))
    assert ei.value.code == P3_ERR_NO_DEVICE
    path = str(tmp_path / "tiny.p3w")
    W.make_synthetic_weight_file(path, "tiny")
    with pytest.raises(E.P3Error) as ei:
        E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_FP32)
    assert ei.value.code == P3_ERR_NO_DEVICE


def test_engine_create_argument_errors(tmp_path):
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import P3_ERR_INVALID_ARG, P3_ERR_IO
    with pytest.raises(E.P3Error) as ei:
        E.B200Engine(str(tmp_path / "missing.p3w"), 4)
    assert ei.value.code == P3_ERR_IO
    bad = tmp_path / "bad.p3w"
    bad.write_bytes(b"not a weight file")
    with pytest.raises(E.P3Error) as ei:
        E.B200Engine(str(bad), 4)
    assert ei.value.code == P3_ERR_IO
    with pytest.raises(E.P3Error) as ei:
        E.B200Engine(str(bad), 0)
    assert ei.value.code == P3_ERR_INVALID_ARG


def test_factory_mirrors_reference(tmp_path):
    """KindFromEnginePath / GetVersionFromModelPath, cc/nn/engine/engine_factory.cc:16-54 (+ the new .p3w kind)."""
    from p3achygo_b200 import engine as E
    d = tmp_path / "model"
    d.mkdir()
    for name in ("m.trt", "m.pb", "m.p3w", "m.bin"):
        (d / name).write_bytes(b"x")
    assert E.KindFromEnginePath(str(d / "m.trt")) == E.Kind.kTrt
    assert E.KindFromEnginePath(str(d / "m.pb")) == E.Kind.kTFXla
    assert E.KindFromEnginePath(str(d / "m.p3w")) == E.Kind.kB200
    assert E.KindFromEnginePath(str(d / "m.bin")) == E.Kind.kUnknown
    trt_dir = tmp_path / "_trt"
    trt_dir.mkdir()
    assert E.KindFromEnginePath(str(trt_dir)) == E.Kind.kTFTrt
    assert E.KindFromEnginePath(str(d)) == E.Kind.kTF
    assert E.GetVersionFromModelPath(str(d / "m.p3w")) == 1            # no VERSION file -> 1
    (d / "VERSION").write_text("0\n")
    assert E.GetVersionFromModelPath(str(d / "m.p3w")) == 0
    assert E.GetVersionFromModelPath(str(d)) == 0
    (d / "VERSION").write_text("garbage")
    assert E.GetVersionFromModelPath(str(d / "m.p3w")) == 1            # unparsable -> default 1 (with a warning)
    assert E.KindToString(E.Kind.kTrt) == "TensorRT" and E.KindToString(E.Kind.kUnknown) == "??"
    with pytest.raises(RuntimeError):
        E.CreateEngine(E.Kind.kTrt, str(d / "m.trt"), 4, 1)
