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
    for i in range(len(kvalid)):
        kk = min(k, int(kvalid[i]))
        # bit-exact: PCG32 jump-ahead, the uniform, glibc's logf (restated on the device) and the two roundings of the score
        assert np.array_equal(scores[i][:kk], ka["g_scores"][i][:kk])
        assert np.array_equal(moves[i][:kk], ka["g_moves"][i][:kk])
        assert np.all(moves[i][kk:] == -1)
        st = ctypes.c_uint64(int(state[i]))
        assert L.orc_prng_next(ctypes.byref(st)) == ka["g_next"][i]          # PRNG advanced by exactly k_valid draws


@pytest.mark.gpu
def test_gumbel_topk_noise_scaling_and_many_roots():
    """4096 roots, noise_scaling != 1 (the product must be rounded before the sum, gumbel.cc:296-300), random legality
    incl. all-illegal-but-pass and fully legal roots: moves, scores, k_valid and the PRNG state equal the oracle's."""
    from p3achygo_b200 import engine as E
    L = oracle_lib.oracle()
    rng = np.random.default_rng(77)
    n, k = 4096, 16
    logits = (rng.standard_normal((n, 362)) * 3.0).astype(np.float32)
    legal = (rng.random((n, 362)) < rng.random((n, 1))).astype(np.uint8)
    legal[:, 361] = 1
    legal[0, :361] = 0
    legal[1, :] = 1
    seeds = [L.orc_prng_seed(int(s)) for s in rng.integers(0, 2 ** 62, n)]
    for scale in (0.3, 1.7):
        state = np.array(seeds, dtype=np.uint64)
        moves, scores, kvalid = E.gumbel_topk(logits, legal, state, scale, k)
        for i in range(0, n, 7):
            om, osc, okv, ost = oracle_lib.gumbel_topk(int(seeds[i]), logits[i], legal[i], scale, k)
            kk = min(k, okv)
            assert kvalid[i] == okv and int(state[i]) == ost
            assert np.array_equal(moves[i][:kk], om[:kk]) and np.array_equal(scores[i][:kk], osc[:kk])
