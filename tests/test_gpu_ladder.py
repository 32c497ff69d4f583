"""Rules on the GPU from game records (p3_game_derive, csrc/ladder.cu): position replay, Board::GetLadderedStones
(cc/game/board.cc:692-899) and the exact legal-move mask incl. positional superko and pass-alive points
(Game::IsValidMove, cc/game/board.cc:595-644) - bit-exact against
  * tests/golden/ladder_games.npz: outputs of the compiled, unmodified reference for 1297 game records, among them the 17
    positions of the reference's own ladder tests (cc/game/__tests__/board_test.cc "LadderTest") with the values those tests
    assert, random and fighting playouts, and ko positions;
  * the compiled reference run live on fresh random games (oracle/_ref/libp3ref.so travels to the GPU box).
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ladder_games.npz")


@pytest.fixture(scope="module")
def games():
    z = np.load(GOLDEN)
    return {k: z[k] for k in z.files}


def test_fixture_holds_reference_asserted_values(games):
    """CPU: the values the reference's ladder tests CHECK are what the reference computed for the fixture."""
    n = int(games["n_reftest"])
    assert n == 17
    for t in range(n):
        seen = 0
        for i, j, c in games["reftest_expect"][t]:
            if c != -9:
                assert games["ladder"][t, i * 19 + j] == c, (games["reftest_names"][t], i, j)
                seen += 1
        assert seen >= 1
    assert int((games["ladder"] != 0).any(axis=1).sum()) > 500


@pytest.mark.gpu
def test_reference_ladder_tests_on_gpu(games):
    from p3achygo_b200 import engine as E
    n = int(games["n_reftest"])
    boards, lad, legal, status = E.game_derive(games["moves"][:n], games["num_moves"][:n], colors=games["colors"][:n],
                                               forbidden=games["forbidden"][:n])
    assert not status.any()
    for t in range(n):
        name = str(games["reftest_names"][t])
        for i, j, c in games["reftest_expect"][t]:
            if c != -9:
                assert lad[t, i * 19 + j] == c, (name, i, j, c)          # what board_test.cc asserts
        assert np.array_equal(lad[t], games["ladder"][t]), name           # the reference's whole grid
        assert np.array_equal(boards[t], games["boards"][t]), name        # "the board is unchanged" / replay is exact


@pytest.mark.gpu
def test_game_derive_matches_reference_fixture(games):
    from p3achygo_b200 import engine as E
    boards, lad, legal, status = E.game_derive(games["moves"], games["num_moves"], colors=games["colors"],
                                               forbidden=games["forbidden"])
    assert not status.any()
    assert np.array_equal(boards, games["boards"])
    bad = np.flatnonzero((lad != games["ladder"]).any(axis=1))
    assert len(bad) == 0, (bad[:10], len(bad))
    bad = np.flatnonzero((legal != games["legal"]).any(axis=1))
    assert len(bad) == 0, (bad[:10], len(bad))
    # without the pass-alive grid only those points differ
    _, _, legal2, _ = E.game_derive(games["moves"], games["num_moves"], colors=games["colors"], want_ladder=False)
    diff = legal2[:, :361] != games["legal"][:, :361]
    assert np.array_equal(diff, (games["forbidden"] != 0) & (games["boards"] == 0) & diff)
    # history-free legality (p3_legal_mask) differs from the exact one exactly at the superko points
    free = E.legal_mask(games["boards"], games["colors"], games["forbidden"])
    assert int((free != legal).sum()) >= 80 and not (legal & ~free).any()


@pytest.mark.gpu
def test_game_derive_edge_cases():
    from p3achygo_b200 import engine as E
    W = E.MOVE_WHITE
    moves = np.full((4, 8), -1, dtype=np.int16)
    nm = np.array([0, 3, 2, 8], dtype=np.int32)
    moves[1, :3] = [361, 361 + W, 0]              # passes, then a stone
    moves[2, :2] = [5, 5 + W]                     # occupied point: not a legal record
    moves[3, :8] = [1, 0 + W, 19, 361 + W, 361, 361 + W, 361, 361 + W]   # black captures the corner stone
    boards, lad, legal, status = E.game_derive(moves, nm, colors=np.array([1, -1, 1, -1], dtype=np.int8))
    assert not boards[0].any() and legal[0].all() and not lad[0].any()
    assert boards[1, 0] == 1 and boards[1].sum() == 1 and legal[1, 0] == 0 and legal[1, 361] == 1
    assert status[2] == 1 and status[[0, 1, 3]].tolist() == [0, 0, 0]
    assert boards[3, 0] == 0 and boards[3, 1] == 1 and boards[3, 19] == 1   # captured
    assert legal[3, 0] == 0                       # white may not play into the corner: self-capture
    with pytest.raises(E.P3Error):
        E.game_derive(np.zeros((1, 0), dtype=np.int16), np.zeros(1, dtype=np.int32))


@pytest.mark.gpu
def test_game_derive_matches_live_reference():
    """Fresh seeded games played through the compiled reference on this box, compared move list by move list."""
    from oracle import oracle_lib
    from oracle.oracle_lib import P
    from p3achygo_b200 import engine as E
    R = oracle_lib.ref()
    if R is None:
        pytest.skip("oracle/_ref/libp3ref.so not present")
    rng = np.random.default_rng(2026)
    mask = np.zeros(362, dtype=np.uint8)
    recs = []
    for game in range(24):
        g = R.ref_game_new(7.5, 1)
        color = 1
        for mv in range(int(rng.integers(40, 330))):
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
            if mv % 23 == 22:
                moves = np.full(448, -1, dtype=np.int16)
                n = R.ref_game_moves(g, P(moves), 448)
                lad = np.zeros(361, dtype=np.int8)
                legal = np.zeros(362, dtype=np.uint8)
                st = np.zeros(361, dtype=np.uint8)
                R.ref_game_laddered(g, P(lad))
                R.ref_game_legal_mask(g, color, P(legal))
                R.ref_game_move_status(g, color, P(st))
                recs.append((moves, n, color, lad, legal, (st == 4).astype(np.int8)))
        R.ref_game_free(g)
    moves = np.stack([r[0] for r in recs])
    nm = np.array([r[1] for r in recs], dtype=np.int32)
    colors = np.array([r[2] for r in recs], dtype=np.int8)
    boards, lad, legal, status = E.game_derive(moves, nm, colors=colors, forbidden=np.stack([r[5] for r in recs]))
    assert not status.any() and len(recs) > 100
    assert np.array_equal(lad, np.stack([r[3] for r in recs]))
    assert np.array_equal(legal, np.stack([r[4] for r in recs]))
