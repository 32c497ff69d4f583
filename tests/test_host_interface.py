"""The host-side NNInterface mirror (p3achygo_b200/host/nn_interface.{h,cc}): the reference's slot-synchronisation contract
(cc/nn/nn_interface.cc:286-371) re-stated over nn::Engine without the 256-thread cap.

CPU: the port of cc/nn/__tests__/nn_interface_sync_test.cc (a counting engine, jittered and deliberately slow workers,
partial timed-out batches): no torn read, no stale result, no slot mix-up.
GPU: 256 worker threads drive a 256-slot B200 engine through the interface; every result equals the engine's own
result for that position, bit for bit, with the symmetry handled on the host side or on the GPU.
"""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host():
    h = ctypes.CDLL(os.path.join(ROOT, "p3achygo_b200", "libp3host.so"))
    h.p3_host_iface_sync_test.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    h.p3_host_iface_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    h.p3_host_iface_run_banks.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double)]
    return h


@pytest.mark.parametrize("threads,timeout_us", [(128, 200), (16, 400), (2, 0)])
def test_sync_invariants(threads, timeout_us):
    out = (ctypes.c_longlong * 6)()
    _host().p3_host_iface_sync_test(threads, 2500, timeout_us, out)
    race, stale, wrong_slot, wrong_value, cycles, served = list(out)
    assert (race, stale, wrong_slot, wrong_value) == (0, 0, 0, 0)
    assert cycles > 0 and served >= cycles  # partial batches happen, every served result came from a later cycle


@pytest.mark.gpu
@pytest.mark.parametrize("use_sym", [0, 1])
def test_interface_over_b200_engine(use_sym, weight_dir, golden_positions):
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import INFER_RESULT_DTYPE
    path, cfg, tensors = weight_dir("b10c128btl3")
    threads, n = 256, 512
    feats = np.ascontiguousarray(golden_positions["feats"][:n])
    results = np.zeros(n, dtype=INFER_RESULT_DTYPE)
    ninf = ctypes.c_longlong(0)
    _host().p3_host_iface_run(path.encode(), 0, threads, 1, E.PRECISION_BF16, feats.ctypes.data_as(ctypes.c_void_p), n, use_sym, 400,
                              results.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ninf))
    assert ninf.value >= n // threads
    # the engine's own answer for the same positions (slot-independent, deterministic)
    eng = E.CreateEngine(E.Kind.kB200, path, 64, 1, precision=E.PRECISION_BF16)
    for lo in range(0, n, 64):
        for s in range(64):
            if use_sym:
                eng.LoadBatchSym(s, feats[lo + s], (lo + s) % 8)
            else:
                eng.LoadBatch(s, feats[lo + s])
        eng.RunInference()
        for s in range(64):
            r = eng.GetBatch(s)
            for f in r.dtype.names:
                assert np.array_equal(np.asarray(r[f]), np.asarray(results[lo + s][f])), (lo + s, f)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("use_sym", [0, 1])
def test_double_buffered_interface_equals_single(use_sym, weight_dir, golden_positions):
    """NNInterfaceB200 with two slot banks (256 workers over a 128-slot engine, each bank with its own infer thread calling
    Submit + Wait) serves every worker the same bits as the single-bank interface."""
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import INFER_RESULT_DTYPE
    path, cfg, tensors = weight_dir("b10c128btl3")
    threads, n = 256, 1024
    feats = np.ascontiguousarray(golden_positions["feats"][:n])
    assert len(feats) == n and feats.dtype.itemsize == 1860
    out = {}
    for banks in (1, 2):
        results = np.zeros(n, dtype=INFER_RESULT_DTYPE)
        ninf, secs = ctypes.c_longlong(0), ctypes.c_double(0)
        _host().p3_host_iface_run_banks(path.encode(), 0, threads, 1, E.PRECISION_BF16, feats.ctypes.data_as(ctypes.c_void_p), n,
                                        use_sym, 400, banks, results.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ninf),
                                        ctypes.byref(secs))
        assert ninf.value >= n // threads * banks
        out[banks] = results
    for f in INFER_RESULT_DTYPE.names:
        assert np.array_equal(out[1][f], out[2][f]), f
