"""The reference's callers on the B200 engine, UNCHANGED (north star; VERDICT r1 missing #1).

oracle/Makefile compiles, where they lie under /root/reference and without modification,
  * cc/nn/nn_interface.{h,cc} + cc/nn/__tests__/nn_interface_sync_test.cc   -> oracle/_ref/nn_interface_sync_test
  * cc/nn/nn_interface.cc + cc/mcts/{gumbel,tree,leaf_evaluator,search_policy,node_table}.cc + cc/game + cc/core, linked with the
    product's adapter nn::B200Engine built against the REAL nn::Engine (-DP3_REFERENCE_TREE)   -> oracle/_ref/libp3refnn.so
against header-only abseil / doctest / boost stand-ins (oracle/absl_shim); the only reference-side change is edit 1 of
INTEGRATION.md (Engine::Kind::kB200), applied to a build-time copy by oracle/ref_patches/0001-engine-kind-b200.patch.
The built files travel to the GPU box; /root/reference itself is not needed at run time.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
SYNC_TEST = os.path.join(REFDIR, "nn_interface_sync_test")
REFNN = os.path.join(REFDIR, "libp3refnn.so")


def test_reference_sync_test_unmodified_on_the_shim():
    """cc/nn/__tests__/nn_interface_sync_test.cc as the reference wrote it (its CountingEngine, its four TEST_CASEs: kGenCounter /
    kMutex x single / dual interface), 3 s per case instead of 120: green on the abseil stand-in the B200 run below uses."""
    if not os.path.exists(SYNC_TEST):
        pytest.skip("oracle/_ref/nn_interface_sync_test not built (needs /root/reference at build time)")
    env = dict(os.environ, TEST_SECONDS="3")
    r = subprocess.run([SYNC_TEST], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "4 test case(s), 0 failed check(s)" in r.stderr


def _refnn():
    if not os.path.exists(REFNN):
        pytest.skip("oracle/_ref/libp3refnn.so not built (needs /root/reference at build time)")
    L = ctypes.CDLL(REFNN)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.ref_nn_b200_sync.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, ci, ctypes.POINTER(ctypes.c_longlong)]
    L.ref_selfplay_gumbel.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, ci, ctypes.c_double, ci, ci, ctypes.POINTER(ctypes.c_longlong),
                                      ctypes.POINTER(ctypes.c_double)]
    return L


@pytest.fixture(scope="module")
def records():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ladder_games.npz"))
    idx = np.arange(200, 328)                      # random playouts
    return (np.ascontiguousarray(z["moves"][idx]), np.ascontiguousarray(z["num_moves"][idx].astype(np.int32)),
            np.ascontiguousarray(z["colors"][idx].astype(np.int8)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("wake,dual", [(1, 0), (0, 0), (1, 1), (0, 1)])
def test_unmodified_nninterface_on_b200_engine(wake, dual, records, weight_dir):
    """The four scenarios of the reference's sync test with its UNMODIFIED NNInterface driving the real engine and real games
    (NNInterface::LoadBatch builds GoFeatures with the reference's own Board code, random symmetry per call, timeout 200 us, every
    8th worker slow enough to force partial batches): no RunInference / GetBatch overlap, no stale result, and every result
    bit-identical to the same (game, symmetry) evaluated alone."""
    L = _refnn()
    path, cfg, tensors = weight_dir("b10c128btl3")
    mv, nm, col = records
    vp = ctypes.c_void_p
    out = (ctypes.c_longlong * 6)()
    rc = L.ref_nn_b200_sync(path.encode(), 0, 128, 12, 200, wake, dual, 0, mv.ctypes.data_as(vp), nm.ctypes.data_as(vp),
                            col.ctypes.data_as(vp), mv.shape[1], len(mv), out)
    race, stale, differ, runs, served, compared = list(out)
    print(f"wake={wake} dual={dual}: runs {runs}, served {served}, compared {compared}")
    assert rc == 0
    assert (race, stale, differ) == (0, 0, 0)
    assert served == 128 * 12 and compared >= 256 and runs >= 12   # cache off: every call reached the engine; partial batches => runs > iters


REFNN_GR = os.path.join(ROOT, "oracle", "_ref", "libp3refnn_gr.so")


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_patched_nn_interface_loads_slots_as_game_records(weight_dir, records):
    """INTEGRATION.md's optional edit 5 (oracle/ref_patches/0002-load-game-records.patch): NNInterface::LoadBatch hands the move
    list to nn::Engine::LoadGameRecord and the engine derives board, liberty grids, laddered stones and last moves on the GPU; the
    slot's result comes back un-rotated.  The reference's sync scenario over that build, with the serial re-evaluation forced onto
    the GoFeatures path (FeaturesOnlyEngine): every worker slot was a record, and every NNInferResult equals the host-features one
    bit for bit - through the reference's own NNInterface, random symmetries included.  Then the Gumbel self-play harness on it."""
    if not os.path.exists(REFNN_GR):
        pytest.skip("oracle/_ref/libp3refnn_gr.so not built")
    L = ctypes.CDLL(REFNN_GR)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.ref_nn_b200_sync.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, ci, ctypes.POINTER(ctypes.c_longlong)]
    L.ref_selfplay_gumbel.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, ci, ctypes.c_double, ci, ci, ctypes.POINTER(ctypes.c_longlong),
                                      ctypes.POINTER(ctypes.c_double)]
    L.ref_record_loads.restype = ctypes.c_longlong
    L.ref_selfplay_record_loads.restype = ctypes.c_longlong
    path, cfg, tensors = weight_dir("b10c128btl3")
    mv, nm, col = records
    out = (ctypes.c_longlong * 6)()
    before = L.ref_record_loads()
    rc = L.ref_nn_b200_sync(path.encode(), 0, 64, 10, 200, 1, 0, 0, mv.ctypes.data_as(vp), nm.ctypes.data_as(vp),
                            col.ctypes.data_as(vp), mv.shape[1], len(mv), out)
    race, stale, differ, runs, served, compared = list(out)
    loaded = L.ref_record_loads() - before
    print(f"record slots: runs {runs}, served {served}, compared {compared}, loaded as records {loaded}")
    assert rc == 0 and (race, stale, differ) == (0, 0, 0)
    assert served == 64 * 10 and loaded == served and compared >= 256
    out4 = (ctypes.c_longlong * 4)()
    secs = ctypes.c_double(0)
    rc = L.ref_selfplay_gumbel(path.encode(), 0, 2, 32, 16, 4, 4.0, 1 << 16, 40, out4, ctypes.byref(secs))
    moves, served, runs, games = list(out4)
    print(f"self-play on record slots: {moves} moves, {served} leaf evals, {L.ref_selfplay_record_loads()} record loads")
    assert rc == 0 and moves > 64 and L.ref_selfplay_record_loads() >= served > moves


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_gumbel_search_root_runs_on_b200_engine(weight_dir):
    """GumbelEvaluator::SearchRoot (cc/mcts/gumbel.cc:260), unmodified, self-playing on the engine: games progress, every move
    costs leaf evaluations, and the engines are driven in batches."""
    L = _refnn()
    path, cfg, tensors = weight_dir("b10c128btl3")
    out = (ctypes.c_longlong * 4)()
    secs = ctypes.c_double(0)
    rc = L.ref_selfplay_gumbel(path.encode(), 0, 2, 32, 16, 4, 4.0, 1 << 16, 40, out, ctypes.byref(secs))
    moves, served, runs, games = list(out)
    print(f"self-play smoke: {moves} moves, {served} leaf evals, {runs} engine runs, {games} games in {secs.value:.2f} s")
    assert rc == 0 and moves > 64 and served > moves and runs > 0
    assert served / runs > 4        # batches, not single positions
