"""CPU (only where the compiled reference is available: oracle/_ref/libp3ref.so, built from /root/reference by
oracle/Makefile): pins the C restatement against the reference's own code on FRESH seeded positions, beyond the committed
golden vectors — including the 10-move position of cc/nn/__tests__/nn_board_utils_test.cc."""
import ctypes

import numpy as np
import pytest

from oracle import oracle_lib
from oracle.oracle_lib import P

R = oracle_lib.ref()
pytestmark = pytest.mark.skipif(R is None, reason="compiled reference not available on this machine")

FEAT = np.dtype({"names": ["bsize", "color", "komi", "board", "last_moves", "a", "b", "c", "d"],
                 "formats": ["<i4", "i1", "<f4", ("i1", 361), ("<i4", (5, 2)), ("i1", 361), ("i1", 361), ("i1", 361), ("i1", 361)],
                 "offsets": [0, 4, 8, 12, 376, 416, 777, 1138, 1499], "itemsize": 1860})


def _playout(rng, n_moves, komi=7.5):
    g = R.ref_game_new(komi, 1)
    mask = np.zeros(362, dtype=np.uint8)
    color = 1
    for _ in range(n_moves):
        R.ref_game_legal_mask(g, color, P(mask))
        cand = np.flatnonzero(mask[:361])
        if len(cand) == 0 or rng.random() < 0.03:
            R.ref_game_play(g, 19, 0, color)
        else:
            m = int(rng.choice(cand))
            assert R.ref_game_play(g, m // 19, m % 19, color) == 1
        color = -color
        if R.ref_game_is_over(g):
            break
    return g, color


def test_features_planes_liberties_legal_on_fresh_positions():
    rng = np.random.default_rng(4242)
    L = oracle_lib.oracle()
    for trial in range(40):
        g, color = _playout(rng, int(rng.integers(1, 280)), komi=float(rng.choice([0.5, 6.5, 7.5])))
        sym = int(rng.integers(0, 8))
        f = np.zeros(1, dtype=FEAT)
        R.ref_game_features(g, color, sym, P(f))
        for version, (npl, ns) in ((1, (15, 8)), (0, (13, 7))):
            rp, rs = np.empty((1, 19, 19, npl), np.float32), np.empty((1, ns), np.float32)
            R.ref_load_go_features(P(f), 1, version, P(rp), P(rs))
            op, os_ = oracle_lib.load_go_features(f, version)
            assert np.array_equal(rp, op) and np.array_equal(rs.view(np.uint32), os_.view(np.uint32))
        board = np.zeros(361, np.int8)
        R.ref_game_board(g, P(board))
        for n in (1, 2, 3):
            ref_grid, got = np.zeros(361, np.int8), np.zeros(361, np.int8)
            R.ref_game_liberties(g, n, P(ref_grid))
            L.orc_stones_with_liberties(P(board), n, P(got))
            assert np.array_equal(ref_grid, got)
        legal = np.zeros(362, np.uint8)
        R.ref_game_legal_mask(g, color, P(legal))
        got = oracle_lib.legal_mask_nohist(board[None], np.array([color], np.int8))[0]
        assert np.all(got >= legal)
        forb = (got != legal)[:361].astype(np.int8)
        assert np.array_equal(oracle_lib.legal_mask_nohist(board[None], np.array([color], np.int8), forb[None])[0], legal)
        R.ref_game_free(g)


def test_nn_board_utils_position():
    """cc/nn/__tests__/nn_board_utils_test.cc:55-125: B(0,0) W(1,0) B(0,1) W(1,1) ... 10 alternating moves; identity symmetry."""
    g = R.ref_game_new(7.5, 1)
    moves = [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2), (0, 3), (1, 3), (0, 4), (1, 4)]
    color = 1
    for (i, j) in moves:
        assert R.ref_game_play(g, i, j, color) == 1
        color = -color
    f = np.zeros(1, dtype=FEAT)
    R.ref_game_features(g, 1, 0, P(f))
    planes, scalars = oracle_lib.load_go_features(f, 1)
    rp, rs = np.empty((1, 19, 19, 15), np.float32), np.empty((1, 8), np.float32)
    R.ref_load_go_features(P(f), 1, 1, P(rp), P(rs))
    assert np.array_equal(planes, rp) and np.array_equal(scalars, rs)
    assert planes[0, :, :, 0].sum() == 5 and planes[0, :, :, 1].sum() == 5       # own / opponent stones
    for k, (i, j) in enumerate(moves[-5:]):                                       # history oldest -> newest on planes 2..6
        assert planes[0, i, j, 2 + k] == 1.0
    assert list(scalars[0]) == [1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -0.5]          # black to move, komi 7.5 -> -7.5/15
    R.ref_game_free(g)


def test_symmetry_prng_softmax_gumbel_direct():
    L = oracle_lib.oracle()
    for s in range(8):
        for i in range(0, 361, 7):
            assert L.orc_transform_index(s, i) == R.ref_transform_index(s, i)
            assert L.orc_transform_inv(s, i) == R.ref_transform_inv(s, i)
    rng = np.random.default_rng(5)
    for seed in (3, 99, 2 ** 50 + 1):
        p = R.ref_prob_new(seed)
        st = ctypes.c_uint64(L.orc_prng_seed(seed))
        for _ in range(200):
            assert np.float32(R.ref_prob_gumbel(p)) == np.float32(L.orc_gumbel(ctypes.byref(st)))
        R.ref_prob_free(p)
    logits = (rng.standard_normal(362) * 3).astype(np.float32)
    a, b = np.zeros(362, np.float32), np.zeros(362, np.float32)
    R.ref_softmax362(P(logits), P(a))
    L.orc_softmax(362, P(logits), P(b))
    assert np.array_equal(a, b)
    legal = (rng.random(362) < 0.7).astype(np.uint8)
    for k in (1, 4, 16, 64):
        p = R.ref_prob_new(17)
        rm, rsc = np.full(k, -1, np.int32), np.zeros(k, np.float32)
        kv = R.ref_gumbel_topk(p, P(logits), P(legal), 1.0, k, P(rm), P(rsc))
        R.ref_prob_free(p)
        om, osc, okv, _ = oracle_lib.gumbel_topk(L.orc_prng_seed(17), logits, legal, 1.0, k)
        assert kv == okv and np.array_equal(rm[:min(k, kv)], om[:min(k, kv)]) and np.array_equal(rsc[:min(k, kv)], osc[:min(k, kv)])


def test_logf_restatement_is_the_c_library_logf():
    """probability.cc:12-15 calls libm's logf.  orc_logf (the glibc 2.39 algorithm restated; the CUDA kernel mirrors it) equals
    this machine's logf - the one libp3ref.so links - on 2^21 uniforms of Probability::Uniform's grid, on their -logf, on
    random floats of every binade, and Probability::GumbelSample of the compiled reference equals the restated composition."""
    L = oracle_lib.oracle()
    L.orc_logf.restype = L.orc_libm_logf.restype = L.orc_gumbel_from_uniform.restype = ctypes.c_float
    L.orc_logf.argtypes = L.orc_libm_logf.argtypes = L.orc_gumbel_from_uniform.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(2024)
    bits = rng.integers(0, 2 ** 23, 1 << 21, dtype=np.uint32)
    u = ((np.uint32(127 << 23) | bits).view(np.float32) - np.float32(1.0)).astype(np.float32)
    xs = np.concatenate([u, rng.integers(1, 0x7f800000, 1 << 19, dtype=np.uint32).view(np.float32),
                         np.array([0.0, 1.0, np.inf, 1e-45, 1.17549435e-38, np.float32(1) - np.float32(2 ** -24)], np.float32)])
    f_orc = np.vectorize(L.orc_logf, otypes=[np.float32])
    f_lib = np.vectorize(L.orc_libm_logf, otypes=[np.float32])
    # numpy's float32 log is not glibc's; go through ctypes in chunks (vectorize is a python loop: keep it to ~2.6 M calls)
    a, b = f_orc(xs), f_lib(xs)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    inner = (-a[: 1 << 18]).astype(np.float32)
    assert np.array_equal(f_orc(inner).view(np.uint32), f_lib(inner).view(np.uint32))
    for seed in (1, 123456789, 2 ** 61 + 5):
        p = R.ref_prob_new(seed)
        st = ctypes.c_uint64(L.orc_prng_seed(seed))
        for _ in range(20000):
            assert np.float32(R.ref_prob_gumbel(p)) == np.float32(L.orc_gumbel(ctypes.byref(st)))
        R.ref_prob_free(p)


def test_game_rules_on_fresh_games():
    """orc_game_derive against Board::GetLadderedStones / Game::IsValidMove of the compiled reference on fresh games."""
    rng = np.random.default_rng(31337)
    recs = []
    for _ in range(40):
        g, color = _playout(rng, int(rng.integers(30, 320)))
        moves = np.full(448, -1, dtype=np.int16)
        n = R.ref_game_moves(g, P(moves), 448)
        lad = np.zeros(361, dtype=np.int8)
        legal = np.zeros(362, dtype=np.uint8)
        st = np.zeros(361, dtype=np.uint8)
        board = np.zeros(361, dtype=np.int8)
        R.ref_game_board(g, P(board))
        R.ref_game_laddered(g, P(lad))
        R.ref_game_legal_mask(g, color, P(legal))
        R.ref_game_move_status(g, color, P(st))
        recs.append((moves, n, color, board, lad, legal, (st == 4).astype(np.int8)))
        R.ref_game_free(g)
    boards, lad, legal, status = oracle_lib.game_derive(np.stack([r[0] for r in recs]), np.array([r[1] for r in recs]),
                                                        np.array([r[2] for r in recs], dtype=np.int8))
    assert not status.any()
    assert np.array_equal(boards, np.stack([r[3] for r in recs]))
    assert np.array_equal(lad, np.stack([r[4] for r in recs]))
    assert np.array_equal(legal, np.stack([r[5] for r in recs]))
