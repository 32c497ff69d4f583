"""CPU: pins the C restatement (oracle/features_oracle.c) against golden vectors produced by the
reference's own sources (tests/golden/make_golden.py) — the files the GPU parity tests then trust."""
import ctypes
import hashlib
import os

import numpy as np

from oracle import oracle_lib
from oracle.oracle_lib import P


def _digest(planes, scalars):
    return np.frombuffer(hashlib.sha256(planes.tobytes() + scalars.tobytes()).digest()[:8], dtype="<u8")[0]


def test_planes_match_reference_digests(golden_positions):
    feats = golden_positions["feats"]
    planes, scalars = oracle_lib.load_go_features(feats, 1)
    got = np.array([_digest(planes[i], scalars[i]) for i in range(len(feats))], dtype=np.uint64)
    assert np.array_equal(got, golden_positions["planes_digest"])
    assert np.array_equal(planes[:64], golden_positions["planes_first"].astype(np.float32))
    assert np.array_equal(scalars, golden_positions["scalars"])
    # planes are exactly {0, 1}; scalar 7 is +-komi/15 (go_features.cc:54-59)
    assert set(np.unique(planes)) <= {0.0, 1.0}


def test_planes_v0(golden_positions):
    feats = golden_positions["feats"][:64]
    planes, scalars = oracle_lib.load_go_features(feats, 0)
    assert planes.shape[-1] == 13 and scalars.shape[-1] == 7
    assert np.array_equal(planes, golden_positions["planes_v0_first"].astype(np.float32))
    assert np.array_equal(scalars, golden_positions["scalars_v0_first"])


def test_plane_semantics_nn_board_utils(golden_positions):
    """Re-expression of cc/nn/__tests__/nn_board_utils_test.cc:55-200 on LoadGoFeatures: plane indices,
    move-history order, pass flags, colour perspective."""
    f = golden_positions["feats"]
    planes, scalars = oracle_lib.load_go_features(f, 1)
    for i in range(0, len(f), 37):
        color = int(f["color"][i])
        board = f["board"][i].reshape(19, 19)
        assert np.array_equal(planes[i, :, :, 0], (board == color).astype(np.float32))
        assert np.array_equal(planes[i, :, :, 1], (board == -color).astype(np.float32))
        assert scalars[i, 0] == (1.0 if color == 1 else 0.0) and scalars[i, 1] == (1.0 if color == -1 else 0.0)
        for k in range(5):
            mi, mj = f["last_moves"][i][k]
            is_pass = (mi, mj) == (19, 0)
            assert scalars[i, 2 + k] == (1.0 if is_pass else 0.0)
            if (mi, mj) != (-1, -1) and not is_pass:
                assert planes[i, mi, mj, 2 + k] == 1.0 and planes[i, :, :, 2 + k].sum() == 1.0
            else:
                assert planes[i, :, :, 2 + k].sum() == 0.0
        assert scalars[i, 7] == np.float32(np.float32(-1.0 if color == 1 else 1.0) * f["komi"][i]) / np.float32(15.0)


def test_symmetry_tables(known_answers):
    L = oracle_lib.oracle()
    fwd = np.array([[L.orc_transform_index(s, i) for i in range(361)] for s in range(8)])
    inv = np.array([[L.orc_transform_inv(s, i) for i in range(361)] for s in range(8)])
    assert np.array_equal(fwd, known_answers["sym_fwd"]) and np.array_equal(inv, known_answers["sym_inv"])
    for s in range(8):  # inverse round trip (cc/game/__tests__/symmetry_test.cc:90-105)
        assert np.array_equal(inv[s][fwd[s]], np.arange(361))


def test_liberties_and_symmetry_consistency(golden_positions):
    boards, libs, feats, syms = (golden_positions[k] for k in ("boards", "libs", "feats", "syms"))
    got = oracle_lib.stones_with_liberties(boards[:256])
    assert np.array_equal(got, libs[:256])
    # GoFeatures grids are the symmetry-applied raw grids (nn_interface.cc:264-273)
    L = oracle_lib.oracle()
    out = np.zeros(361, dtype=np.int8)
    for i in range(0, 256, 5):
        L.orc_apply_symmetry_i8(int(syms[i]), P(np.ascontiguousarray(boards[i])), P(out))
        assert np.array_equal(out, feats["board"][i])
        L.orc_apply_symmetry_i8(int(syms[i]), P(np.ascontiguousarray(libs[i, 0])), P(out))
        assert np.array_equal(out, feats["stones_atari"][i])
        L.orc_apply_symmetry_i8(int(syms[i]), P(np.ascontiguousarray(golden_positions["ladder"][i])), P(out))
        assert np.array_equal(out, feats["stones_laddered"][i])


def test_legal_mask_without_history(golden_positions):
    boards, colors, legal = (golden_positions[k] for k in ("boards", "colors", "legal"))
    got = oracle_lib.legal_mask_nohist(boards[:512], colors[:512])
    # history-free legality is a superset; the difference is exactly superko / pass-alive prohibition
    assert np.all(got >= legal[:512])
    extra = (got != legal[:512])
    assert extra.sum() < 0.01 * legal[:512].sum()
    got2 = oracle_lib.legal_mask_nohist(boards[:512], colors[:512], forbidden=extra[:, :361].astype(np.int8))
    assert np.array_equal(got2, legal[:512])


def test_prng_and_probability(known_answers):
    L = oracle_lib.oracle()
    for si, seed in enumerate(known_answers["seeds"]):
        for name, fn in (("nexts", L.orc_prng_next), ("unis", L.orc_uniform), ("gums", L.orc_gumbel)):
            st = ctypes.c_uint64(L.orc_prng_seed(int(seed)))
            got = np.array([fn(ctypes.byref(st)) for _ in range(32)], dtype=known_answers[name].dtype)
            assert np.array_equal(got, known_answers[name][si]), name
        st = ctypes.c_uint64(L.orc_prng_seed(int(seed)))
        got = [L.orc_rand_range(ctypes.byref(st), 3, 3 + 1 + (j * 37) % 361) for j in range(32)]
        assert got == list(known_answers["rand_range"][si])
        st = ctypes.c_uint64(L.orc_prng_seed(int(seed)))
        got = [L.orc_rand_range(ctypes.byref(st), 0, 8) for _ in range(32)]  # GetRandomSymmetry, symmetry.h:33
        assert got == list(known_answers["rand_sym"][si])


def test_softmax(known_answers):
    L = oracle_lib.oracle()
    out = np.zeros(362, dtype=np.float32)
    for i in range(len(known_answers["softmax_in"])):
        L.orc_softmax(362, P(np.ascontiguousarray(known_answers["softmax_in"][i])), P(out))
        assert np.array_equal(out, known_answers["softmax_out"][i])
    # known answers of cc/core/__tests__/vmath_test.cc style: uniform logits -> uniform probs
    L.orc_softmax(362, P(np.zeros(362, dtype=np.float32)), P(out))
    assert np.allclose(out, 1.0 / 362, rtol=1e-6)


def test_gumbel_topk(known_answers):
    ka = known_answers
    k = int(ka["g_k"])
    L = oracle_lib.oracle()
    for i in range(len(ka["g_seed"])):
        moves, scores, kv, st = oracle_lib.gumbel_topk(L.orc_prng_seed(int(ka["g_seed"][i])), ka["g_logits"][i],
                                                       ka["g_legal"][i], 1.0, k)
        assert kv == ka["g_kvalid"][i]
        kk = min(k, kv)
        assert np.array_equal(moves[:kk], ka["g_moves"][i][:kk])
        assert np.array_equal(scores[:kk], ka["g_scores"][i][:kk])
        stc = ctypes.c_uint64(st)
        assert L.orc_prng_next(ctypes.byref(stc)) == ka["g_next"][i]  # exactly k_valid draws were consumed


def test_init_fields_matches_formula():
    L = oracle_lib.oracle()
    rng = np.random.default_rng(3)
    p = rng.random(800).astype(np.float32)
    p /= p.sum()
    v = np.array([0.3, 0.7], dtype=np.float32)
    out = np.zeros(3, dtype=np.float32)
    L.orc_init_fields(P(v), P(p), P(out))
    s = np.arange(800) - 400 + 0.5
    assert abs(out[0] - 0.4) < 1e-6
    assert abs(out[1] - float((p * s).sum())) < 1e-2
    assert abs(out[2] - float((p * s * s).sum() - (p * s).sum() ** 2)) < 1.0


def test_oracle_game_rules_match_reference_fixture():
    """oracle/features_oracle.c::orc_game_derive (replay, ladder reader, exact legal mask) against the compiled reference's
    outputs for 1297 game records (tests/golden/ladder_games.npz), among them the reference's own 17 ladder test positions with
    the values cc/game/__tests__/board_test.cc asserts."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ladder_games.npz"))
    boards, lad, legal, status = oracle_lib.game_derive(z["moves"], z["num_moves"], z["colors"])   # pass-alive points restated too
    assert not status.any() and int(z["forbidden"].sum()) > 300
    assert np.array_equal(boards, z["boards"])
    assert np.array_equal(lad, z["ladder"])
    assert np.array_equal(legal, z["legal"])
    for t in range(int(z["n_reftest"])):
        for i, j, c in z["reftest_expect"][t]:
            if c != -9:
                assert lad[t, i * 19 + j] == c, (str(z["reftest_names"][t]), i, j)
