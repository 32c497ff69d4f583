"""GPU parity, integer / byte work (bit-exact bar): encode, liberties, legal mask, Gumbel root sampling.
All calls go through the C ABI (p3achygo_b200.engine -> libp3b200.so); the checker is the pinned C oracle
and the golden vectors produced by the reference sources."""
import hashlib

import numpy as np
import pytest

from oracle import oracle_lib

pytestmark = pytest.mark.gpu


def _digest(planes, scalars):
    return np.frombuffer(hashlib.sha256(planes.tobytes() + scalars.tobytes()).digest()[:8], dtype="<u8")[0]


def test_encode_bit_exact_vs_oracle_and_golden(golden_positions):
    from p3achygo_b200 import engine as E
    feats = golden_positions["feats"]
    planes, scalars = E.encode_features(feats, 1)
    oplanes, oscalars = oracle_lib.load_go_features(feats, 1)
    assert np.array_equal(planes, oplanes)
    assert np.array_equal(scalars.view(np.uint32), oscalars.view(np.uint32))  # bit-exact incl. komi/15
    got = np.array([_digest(planes[i], scalars[i]) for i in range(len(feats))], dtype=np.uint64)
    assert np.array_equal(got, golden_positions["planes_digest"])


def test_encode_v0_and_edge_cases(golden_positions):
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    feats = golden_positions["feats"][:64]
    planes, scalars = E.encode_features(feats, 0)
    assert np.array_equal(planes, golden_positions["planes_v0_first"].astype(np.float32))
    assert np.array_equal(scalars, golden_positions["scalars_v0_first"])
    # empty batch, empty board, all-pass / all-noop history, both colours, odd komi
    p0, s0 = E.encode_features(np.zeros(0, dtype=GO_FEATURES_DTYPE), 1)
    assert p0.shape[0] == 0 and s0.shape[0] == 0
    edge = np.zeros(4, dtype=GO_FEATURES_DTYPE)
    edge["bsize"] = 19
    edge["color"] = [1, -1, 1, -1]
    edge["komi"] = [7.5, 0.5, -3.25, 1e-3]
    edge["last_moves"][0] = [[-1, -1]] * 5
    edge["last_moves"][1] = [[19, 0]] * 5
    edge["last_moves"][2] = [[-1, -1], [19, 0], [0, 0], [18, 18], [9, 9]]
    edge["last_moves"][3] = [[3, 3]] * 5
    edge["board"][3][:] = 1
    edge["stones_atari"][3][:] = -1
    planes, scalars = E.encode_features(edge, 1)
    oplanes, oscalars = oracle_lib.load_go_features(edge, 1)
    assert np.array_equal(planes, oplanes) and np.array_equal(scalars.view(np.uint32), oscalars.view(np.uint32))


def test_encode_idempotent_and_full_size(golden_positions):
    """Size-independent properties at the bench batch size: plane sums equal stone counts; re-encoding is idempotent."""
    from p3achygo_b200 import engine as E
    feats = np.tile(golden_positions["feats"], 4)[:4096]
    planes, scalars = E.encode_features(feats, 1)
    planes2, _ = E.encode_features(feats, 1)
    assert np.array_equal(planes, planes2)
    stones = (feats["board"] != 0).sum(axis=1)
    assert np.array_equal(planes[..., 0].sum(axis=(1, 2)) + planes[..., 1].sum(axis=(1, 2)), stones.astype(np.float32))
    assert np.all(scalars[:, 0] + scalars[:, 1] == 1.0)


def test_liberties_bit_exact(golden_positions):
    from p3achygo_b200 import engine as E
    boards, libs = golden_positions["boards"], golden_positions["libs"]
    got = E.board_liberties(boards)
    assert np.array_equal(got, libs)                                         # the reference's own grids
    assert np.array_equal(got[:128], oracle_lib.stones_with_liberties(boards[:128]))
    assert np.array_equal(E.board_liberties(np.zeros((1, 361), np.int8)), np.zeros((1, 3, 361), np.int8))  # empty board
    full = np.ones((1, 361), np.int8)                                         # one 361-stone group, 0 liberties
    assert not E.board_liberties(full).any()


def test_legal_mask(golden_positions):
    from p3achygo_b200 import engine as E
    boards, colors, legal = (golden_positions[k] for k in ("boards", "colors", "legal"))
    got = E.legal_mask(boards, colors)
    exp = oracle_lib.legal_mask_nohist(boards, colors)
    assert np.array_equal(got, exp)
    # with the history-dependent prohibitions supplied by the host, it equals Game::IsValidMove exactly
    forbidden = (exp != legal)[:, :361].astype(np.int8)
    assert np.array_equal(E.legal_mask(boards, colors, forbidden), legal)


def test_gumbel_topk_matches_reference(known_answers):
    from p3achygo_b200 import engine as E
    ka = known_answers
    k = int(ka["g_k"])
    L = oracle_lib.oracle()
    state = np.array([L.orc_prng_seed(int(s)) for s in ka["g_seed"]], dtype=np.uint64)
    moves, scores, kvalid = E.gumbel_topk(ka["g_logits"], ka["g_legal"], state, 1.0, k)
    assert np.array_equal(kvalid, ka["g_kvalid"])
    import ctypes
    n_set_equal = 0
    for i in range(len(kvalid)):
        kk = min(k, int(kvalid[i]))
        # scores: uniform bits are exact; -log(-log(u)) within 2 ulp of glibc logf (documented in gumbel.cu)
        exp_sorted = ka["g_scores"][i][:kk]
        np.testing.assert_allclose(scores[i][:kk], exp_sorted, rtol=0, atol=4e-6 * max(1.0, float(np.abs(exp_sorted).max())))
        assert np.all(moves[i][kk:] == -1)
        n_set_equal += int(np.array_equal(moves[i][:kk], ka["g_moves"][i][:kk]))
        st = ctypes.c_uint64(int(state[i]))
        assert L.orc_prng_next(ctypes.byref(st)) == ka["g_next"][i]          # PRNG advanced by exactly k_valid draws
    assert n_set_equal >= len(kvalid) - 1                                      # order identical barring a 1-ulp tie
