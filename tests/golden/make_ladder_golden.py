"""Generates tests/golden/ladder_games.npz from the UNMODIFIED reference sources (oracle/_ref/libp3ref.so): game records
(move lists) with what the reference's own Board computes for them - position, Board::GetLadderedStones
(cc/game/board.cc:692-899), Game::IsValidMove over all encodings (superko and pass-alive included), the pass-alive
prohibited points (PlayMoveDry status kPassAliveRegion).  Run in the container that has /root/reference:

    python tests/golden/make_ladder_golden.py

Two families:
  * "playout": seeded random legal playouts (as make_golden.py; 2 % passes, so pass-alive regions and superko occur),
    plus "fight" playouts that prefer moves next to stones (many ataris, captures, kos -> long ladder searches);
  * "reftest": the 19x19 positions of the reference's own ladder tests (cc/game/__tests__/board_test.cc, TEST_CASE
    "LadderTest"), parsed from the test source at generation time and built the way game::ParseBoardDSL builds them
    (all black stones in scan order, then all white stones; cc/game/board_dsl.cc:91-118), with the values the test
    CHECKs (`laddered_stones[AsIndex(Loc{i, j}, BOARD_LEN)] == COLOR`) recorded next to the reference's full grid.
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_lib  # noqa: E402
from oracle.oracle_lib import P  # noqa: E402

MAX_MOVES = 448
WHITE_BIT = 512


def snapshot(R, g, color, rec):
    mv = np.full(MAX_MOVES, -1, dtype=np.int16)
    n = R.ref_game_moves(g, P(mv), MAX_MOVES)
    assert n <= MAX_MOVES
    board = np.zeros(361, dtype=np.int8)
    lad = np.zeros(361, dtype=np.int8)
    legal = np.zeros(362, dtype=np.uint8)
    status = np.zeros(361, dtype=np.uint8)
    R.ref_game_board(g, P(board))
    R.ref_game_laddered(g, P(lad))
    R.ref_game_legal_mask(g, color, P(legal))
    R.ref_game_move_status(g, color, P(status))
    rec["moves"].append(mv)
    rec["num_moves"].append(n)
    rec["boards"].append(board)
    rec["ladder"].append(lad)
    rec["legal"].append(legal)
    rec["colors"].append(color)
    rec["forbidden"].append((status == 4).astype(np.int8))
    rec["status"].append(status)


def playouts(R, rec, n, seed, fight):
    rng = np.random.default_rng(seed)
    mask = np.zeros(362, dtype=np.uint8)
    board = np.zeros(361, dtype=np.int8)
    count = 0
    while count < n:
        g = R.ref_game_new(7.5, 1)
        length = int(rng.integers(20, 381))
        snaps = set(int(x) for x in rng.integers(10, length + 1, size=6)) | {length}
        color = 1
        for mv in range(1, length + 1):
            R.ref_game_legal_mask(g, color, P(mask))
            cand = np.flatnonzero(mask[:361])
            if len(cand) == 0 or rng.random() < 0.02:
                R.ref_game_play(g, 19, 0, color)
            else:
                if fight and mv > 4:
                    R.ref_game_board(g, P(board))
                    b2 = board.reshape(19, 19) != 0
                    near = np.zeros((19, 19), dtype=bool)
                    near[1:] |= b2[:-1]; near[:-1] |= b2[1:]; near[:, 1:] |= b2[:, :-1]; near[:, :-1] |= b2[:, 1:]
                    c2 = cand[near.reshape(-1)[cand]]
                    if len(c2) and rng.random() < 0.85:
                        cand = c2
                m = int(rng.choice(cand))
                assert R.ref_game_play(g, m // 19, m % 19, color) == 1
            color = -color
            if R.ref_game_is_over(g):
                break
            if mv in snaps and count < n:
                snapshot(R, g, color, rec)
                count += 1
        R.ref_game_free(g)


def ko_games(R, rec, n, seed):
    """A ko just taken by BLACK, WHITE to move: the retake is a positional-superko repeat (kRepeatedPosition), the taking
    stone is a group in atari whose ladder search has to respect it; random stones elsewhere."""
    rng = np.random.default_rng(seed)
    mask = np.zeros(362, dtype=np.uint8)
    for _ in range(n):
        r, c = int(rng.integers(0, 17)), int(rng.integers(0, 16))
        black = [(r, c + 1), (r + 1, c), (r + 2, c + 1)]
        white = [(r + 1, c + 1), (r, c + 2), (r + 2, c + 2), (r + 1, c + 3)]
        box = {(i, j) for i in range(r - 1, r + 4) for j in range(c - 1, c + 5)}
        g = R.ref_game_new(7.5, 1)
        for k in range(4):
            if k < 3:
                assert R.ref_game_play(g, black[k][0], black[k][1], 1) == 1
            else:
                R.ref_game_play(g, 19, 0, 1)
            assert R.ref_game_play(g, white[k][0], white[k][1], -1) == 1
        color = 1
        for _ in range(int(rng.integers(0, 120)) * 2):
            R.ref_game_legal_mask(g, color, P(mask))
            cand = [m for m in np.flatnonzero(mask[:361]) if (m // 19, m % 19) not in box]
            if not cand:
                R.ref_game_play(g, 19, 0, color)
            else:
                m = int(rng.choice(cand))
                assert R.ref_game_play(g, m // 19, m % 19, color) == 1
            color = -color
        if R.ref_game_play(g, r + 1, c + 2, 1) != 1:   # random stones may have changed the shape's liberties
            R.ref_game_free(g)
            continue
        snapshot(R, g, -1, rec)
        R.ref_game_free(g)


def reference_ladder_tests(R, rec):
    src = open("/root/reference/cc/game/__tests__/board_test.cc").read()
    start = src.index('TEST_CASE("LadderTest")')
    body = src[start:]
    names, expects = [], []
    for m in re.finditer(r'SUBCASE\("([^"]+)"\)\s*\{(.*?)(?=SUBCASE\(|\Z)', body, re.S):
        name, sub = m.group(1), m.group(2)
        dsl = re.search(r'ParseBoardDSL\(R"\((.*?)\)"\)', sub, re.S)
        if not dsl:
            continue
        rows = [r.split() for r in dsl.group(1).strip().splitlines()]
        if len(rows) != 19 or any(len(r) != 19 for r in rows):
            continue
        checks = re.findall(r'laddered_stones\[AsIndex\(Loc\{(\d+),\s*(\d+)\},\s*BOARD_LEN\)\]\s*==\s*(\w+)', sub)
        g = R.ref_game_new(7.5, 1)
        for sym, col in (("x", 1), ("o", -1)):
            for i in range(19):
                for j in range(19):
                    if rows[i][j].lower() == sym:
                        assert R.ref_game_play(g, i, j, col) == 1, (name, i, j)
        snapshot(R, g, 1, rec)
        R.ref_game_free(g)
        ex = np.full((16, 3), -9, dtype=np.int16)
        for k, (i, j, c) in enumerate(checks[:16]):
            ex[k] = (int(i), int(j), {"WHITE": -1, "BLACK": 1, "EMPTY": 0}[c])
        names.append(name)
        expects.append(ex)
    return names, np.array(expects)


def main():
    R = oracle_lib.ref()
    assert R is not None, "reference library unavailable (needs /root/reference)"
    rec = {k: [] for k in ("moves", "num_moves", "boards", "ladder", "legal", "colors", "forbidden", "status")}
    names, expects = reference_ladder_tests(R, rec)
    n_ref = len(names)
    playouts(R, rec, 600, seed=4242, fight=False)
    playouts(R, rec, 600, seed=777, fight=True)
    ko_games(R, rec, 80, seed=99)
    out = {k: np.array(v) for k, v in rec.items()}
    out["moves"] = out["moves"].astype(np.int16)
    out["num_moves"] = out["num_moves"].astype(np.int32)
    out["colors"] = out["colors"].astype(np.int8)
    st = out.pop("status")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ladder_games.npz")
    np.savez_compressed(path, n_reftest=np.int32(n_ref), reftest_names=np.array(names), reftest_expect=expects, **out)
    lad = out["ladder"]
    print("reference ladder tests:", n_ref, names)
    print("positions", len(lad), "with laddered stones", int((lad != 0).any(1).sum()), "laddered stones", int((lad != 0).sum()),
          "pass-alive forbidden points", int(out["forbidden"].sum()), "superko-illegal points", int((st == 6).sum()),
          "self-capture points", int((st == 5).sum()), "bytes", os.path.getsize(path))
    # the reference's own asserted values agree with the grid it computes here
    for t in range(n_ref):
        for i, j, c in expects[t]:
            if c != -9:
                assert lad[t, i * 19 + j] == c, (names[t], i, j, c)


if __name__ == "__main__":
    main()
