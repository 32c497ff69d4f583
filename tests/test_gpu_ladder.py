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
    # the pass-alive points themselves (Benson at the game's last qualifying pass, board.cc:223-462, 582-593) are computed on the
    # GPU as well: nothing changes when the host grid is left out
    b2, lad2, legal2, st2 = E.game_derive(games["moves"], games["num_moves"], colors=games["colors"])
    assert not st2.any() and int(games["forbidden"].sum()) > 300
    bad = np.flatnonzero((legal2 != games["legal"]).any(axis=1))
    assert len(bad) == 0, (bad[:10], len(bad))
    assert np.array_equal(lad2, games["ladder"])
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
    boards, lad, legal, status = E.game_derive(moves, nm, colors=colors)   # pass-alive points (r[5]) are derived on the GPU too
    assert int(np.stack([r[5] for r in recs]).sum()) >= 0
    assert not status.any() and len(recs) > 100
    assert np.array_equal(lad, np.stack([r[3] for r in recs]))
    assert np.array_equal(legal, np.stack([r[4] for r in recs]))


def _features_from_fixture(games, idx, komi=7.5):
    """The GoFeatures NNInterface::LoadBatch would build (identity orientation) from the reference's own outputs in the fixture."""
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import GO_FEATURES_DTYPE
    boards = games["boards"][idx]
    libs = E.board_liberties(boards)            # bit-exact vs Board::GetStonesWithLiberties (tests/test_gpu_features.py)
    feats = np.zeros(len(idx), dtype=GO_FEATURES_DTYPE)
    feats["bsize"] = 19
    feats["color"] = games["colors"][idx]
    feats["komi"] = komi
    feats["board"] = boards
    feats["stones_atari"] = libs[:, 0]
    feats["stones_two_liberties"] = libs[:, 1]
    feats["stones_three_liberties"] = libs[:, 2]
    feats["stones_laddered"] = games["ladder"][idx]
    for k, g in enumerate(idx):
        nm = int(games["num_moves"][g])
        for j in range(5):
            off = nm - 5 + j
            if off < 0:
                feats["last_moves"][k, j] = (-1, -1)
            else:
                p = int(games["moves"][g, off]) & 511
                feats["last_moves"][k, j] = (19, 0) if p == 361 else (p // 19, p % 19)
    return feats


@pytest.mark.gpu
@pytest.mark.parametrize("pipelined", [False, True])
def test_engine_game_record_slots_equal_feature_slots(pipelined, games, weight_dir):
    """p3_engine_load_game_bank: a slot loaded as a game record gives, bit for bit, the planes and the NNInferResult of the
    same position loaded as the GoFeatures the reference builds - all symmetries, mixed with feature slots in one batch."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir("b10c128btl3")
    B = 64
    idx = np.concatenate([np.arange(0, 17), np.arange(100, 100 + 2 * B - 17)])   # the reference's ladder tests + playouts
    feats = _features_from_fixture(games, idx)
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    for half in range(2):
        sel = idx[half * B:(half + 1) * B]
        for b in range(B):
            eng.LoadBatchSym(b, feats[half * B + b], b % 8)
        eng.RunInference()
        want = [eng.GetBatch(b).copy() for b in range(B)]
        want_planes = [eng.GetPlanes(b) for b in range(B)]
        for b in range(B):
            g = sel[b]
            if b % 5 == 4:   # some slots stay GoFeatures slots
                if pipelined:
                    eng.LoadBatchBank(1, b, feats[half * B + b], b % 8)
                else:
                    eng.LoadBatchSym(b, feats[half * B + b], b % 8)
                continue
            mv = games["moves"][g][: games["num_moves"][g]]
            eng.LoadGameBank(1 if pipelined else 0, b, mv, int(games["colors"][g]), 7.5, games["forbidden"][g], b % 8)
        if pipelined:
            eng.Submit(1)
            eng.Wait(1)
        else:
            eng.RunInference()
        for b in range(B):
            got = eng.GetBatchBank(1, b) if pipelined else eng.GetBatch(b)
            for f in got.dtype.names:
                assert np.array_equal(got[f], want[b][f]), (half, b, f)
            pl, sc = eng.GetPlanes(b)
            assert np.array_equal(pl, want_planes[b][0]) and np.array_equal(sc, want_planes[b][1]), (half, b)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path_kind", ["derive", "run_inference", "banks"])
def test_reader_watchdog_retries_unsplit(path_kind, games, weight_dir, monkeypatch, capfd):
    """The ladder reader's watchdog (status bit 3: a warp polled ~10 s for a split search's item) is not the end of the batch: the
    entry points run it again with splitting off - every search stays on the warp that claimed it, nothing waits for another
    warp - and the results are the reference's.  P3_LADDER_FORCE_WATCHDOG raises the bit after every SPLIT run, so whatever
    comes back here came from the retry."""
    from p3achygo_b200 import engine as E
    monkeypatch.setenv("P3_LADDER_FORCE_WATCHDOG", "1")
    if path_kind == "derive":
        sel = slice(0, 400)
        boards, lad, legal, status = E.game_derive(games["moves"][sel], games["num_moves"][sel], colors=games["colors"][sel],
                                                   forbidden=games["forbidden"][sel])
        assert not status.any()
        assert np.array_equal(lad, games["ladder"][sel]) and np.array_equal(legal, games["legal"][sel])
        assert "unsplit" in capfd.readouterr().err
        return
    path, cfg, tensors = weight_dir("b10c128btl3")
    B = 48
    idx = np.concatenate([np.arange(0, 17), np.arange(100, 100 + B - 17)])
    feats = _features_from_fixture(games, idx)
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    for b in range(B):
        eng.LoadBatchSym(b, feats[b], 0)
    eng.RunInference()
    want = [eng.GetBatch(b).copy() for b in range(B)]
    bank = 1 if path_kind == "banks" else 0
    for b in range(B):
        g = idx[b]
        eng.LoadGameBank(bank, b, games["moves"][g][: games["num_moves"][g]], int(games["colors"][g]), 7.5, games["forbidden"][g], 0)
    if path_kind == "banks":
        eng.Submit(1)
        eng.Wait(1)
    else:
        eng.RunInference()
    for b in range(B):
        got = eng.GetBatchBank(1, b) if path_kind == "banks" else eng.GetBatch(b)
        for f in got.dtype.names:
            assert np.array_equal(got[f], want[b][f]), (b, f)
    eng.close()
    assert "unsplit" in capfd.readouterr().err


@pytest.mark.gpu
def test_game_derive_matches_oracle_on_fresh_games():
    """GPU against the CPU restatement (oracle/features_oracle.c::orc_game_derive) on fresh seeded games that neither the fixture
    nor the reference has seen: the games are played with the oracle's own exact legality (no /root/reference needed)."""
    from oracle import oracle_lib
    from p3achygo_b200 import engine as E
    rng = np.random.default_rng(777001)
    recs = []
    for game in range(10):
        moves = np.full(448, -1, dtype=np.int16)
        n, color = 0, 1
        length = int(rng.integers(60, 300))
        while n < length:
            _, _, legal, st = oracle_lib.game_derive(moves[None], np.array([n]), np.array([color], dtype=np.int8))
            cand = np.flatnonzero(legal[0, :361])
            if len(cand) == 0 or rng.random() < 0.02:
                moves[n] = 361 + (E.MOVE_WHITE if color < 0 else 0)
            else:
                near = cand if n < 6 or rng.random() < 0.3 else cand[np.argsort(rng.random(len(cand)))[: max(8, len(cand) // 6)]]
                moves[n] = int(rng.choice(near)) + (E.MOVE_WHITE if color < 0 else 0)
            n += 1
            color = -color
            if n % 37 == 0 or n == length:
                recs.append((moves.copy(), n, color))
    mv = np.stack([r[0] for r in recs])
    nm = np.array([r[1] for r in recs], dtype=np.int32)
    col = np.array([r[2] for r in recs], dtype=np.int8)
    want = oracle_lib.game_derive(mv, nm, col)
    got = E.game_derive(mv, nm, colors=col)
    assert not got[3].any() and not want[3].any() and len(recs) > 40
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    assert int((want[1] != 0).sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("banks", [1, 2])
def test_interface_from_game_records(banks, games, weight_dir):
    """NNInterfaceB200::LoadAndGetInferenceGame: 128 workers hand game records to the interface (one or two slot banks); every
    result equals the engine's answer for the GoFeatures the reference builds for that position."""
    import ctypes
    from p3achygo_b200 import engine as E
    from p3achygo_b200._lib import INFER_RESULT_DTYPE
    path, cfg, tensors = weight_dir("b10c128btl3")
    host = ctypes.CDLL(os.path.join(ROOT, "p3achygo_b200", "libp3host.so"))
    vp, ci = ctypes.c_void_p, ctypes.c_int
    host.p3_host_iface_run_games.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, vp, vp, vp, vp, ci, ci, ci, ci, vp,
                                             ctypes.POINTER(ctypes.c_longlong)]
    threads, n = 128, 256
    idx = np.arange(0, n)
    mv = np.ascontiguousarray(games["moves"][idx])
    nm = np.ascontiguousarray(games["num_moves"][idx])
    col = np.ascontiguousarray(games["colors"][idx])
    fb = np.ascontiguousarray(games["forbidden"][idx])
    results = np.zeros(n, dtype=INFER_RESULT_DTYPE)
    ninf = ctypes.c_longlong(0)
    host.p3_host_iface_run_games(path.encode(), 0, threads, 1, E.PRECISION_BF16, mv.ctypes.data_as(vp), nm.ctypes.data_as(vp),
                                 col.ctypes.data_as(vp), fb.ctypes.data_as(vp), mv.shape[1], n, 400, banks,
                                 results.ctypes.data_as(vp), ctypes.byref(ninf))
    assert ninf.value >= n // threads
    feats = _features_from_fixture(games, idx)
    eng = E.CreateEngine(E.Kind.kB200, path, 64, 1, precision=E.PRECISION_BF16)
    for lo in range(0, n, 64):
        for s in range(64):
            eng.LoadBatchSym(s, feats[lo + s], (lo + s) % 8)
        eng.RunInference()
        for s in range(64):
            r = eng.GetBatch(s)
            for f in r.dtype.names:
                assert np.array_equal(np.asarray(r[f]), np.asarray(results[lo + s][f])), (lo + s, f)
    eng.close()


@pytest.mark.gpu
def test_engine_rejects_impossible_game_record(games, weight_dir):
    """A move list that is not a legal game (a stone onto an occupied point) makes the run fail loudly, as the reference's CHECKs
    would; valid slots of the same batch still get their results, and the engine keeps working afterwards."""
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir("b10c128btl3")
    eng = E.CreateEngine(E.Kind.kB200, path, 4, 1, precision=E.PRECISION_BF16)
    good = games["moves"][20][: games["num_moves"][20]]
    for b in range(4):
        eng.LoadGameBank(0, b, good, int(games["colors"][20]), 7.5)
    eng.RunInference()
    want = eng.GetBatch(0).copy()
    eng.LoadGameBank(0, 2, np.array([5, 5 + E.MOVE_WHITE], dtype=np.int16), 1, 7.5)
    with pytest.raises(E.P3Error) as ei:
        eng.RunInference()
    assert "slot 2" in str(ei.value)
    got = eng.GetBatch(0)
    assert all(np.array_equal(got[f], want[f]) for f in got.dtype.names)
    eng.LoadGameBank(0, 2, good, int(games["colors"][20]), 7.5)
    eng.RunInference()
    got = eng.GetBatch(2)
    assert all(np.array_equal(got[f], want[f]) for f in got.dtype.names)
    with pytest.raises(E.P3Error):
        eng.LoadGameBank(0, 0, np.zeros(2000, dtype=np.int16), 1, 7.5)     # longer than P3_MAX_GAME_MOVES
    eng.close()


@pytest.mark.gpu
def test_game_derive_is_symmetry_equivariant_at_full_size(games, known_answers):
    """Size-independent property at 8 x 1297 = 10 376 records (more than the bench's 8192 positions): playing the same game on a
    rotated / reflected board gives the rotated / reflected board, laddered stones and legal mask - the reference's rules have no
    preferred direction, while the kernels' row / lane layout does."""
    from p3achygo_b200 import engine as E
    sym_fwd = known_answers["sym_fwd"]        # [8, 361] TransformIndex tables of the compiled reference
    mv = games["moves"].astype(np.int32)
    point = mv & 511
    is_stone = (mv >= 0) & (point < 361)
    all_moves = []
    for s in range(8):
        t = mv.copy()
        t[is_stone] = sym_fwd[s][point[is_stone]] + (mv[is_stone] & 512)
        all_moves.append(t.astype(np.int16))
    n = len(mv)
    boards, lad, legal, status = E.game_derive(np.concatenate(all_moves), np.tile(games["num_moves"], 8), colors=np.tile(games["colors"], 8))
    assert not status.any()
    for s in range(8):
        sl = slice(s * n, (s + 1) * n)
        for got, ref in ((boards[sl], boards[:n]), (lad[sl], lad[:n]), (legal[sl][:, :361], legal[:n][:, :361])):
            want = np.zeros_like(ref)
            want[:, sym_fwd[s]] = ref          # sym_grid[T(i)] = grid[i]  (cc/game/symmetry.h:42-51)
            assert np.array_equal(got, want), s
    assert np.array_equal(lad[:n], games["ladder"]) and np.array_equal(legal[:n], games["legal"])


@pytest.mark.gpu
def test_root_sampling_from_game_records(games):
    """The root's candidate moves end to end on the GPU: game record -> exact legal mask (p3_game_derive) -> Gumbel top-k
    (p3_gumbel_topk, cc/mcts/gumbel.cc:283-321), against the oracle fed with the reference's own Game::IsValidMove mask
    (superko-illegal and pass-alive points are not drawn for, which shifts every later PRNG draw if a mask bit is wrong)."""
    from oracle import oracle_lib
    from p3achygo_b200 import engine as E
    idx = np.concatenate([np.arange(1217, 1297), np.arange(200, 248)])      # the ko positions + playouts
    rng = np.random.default_rng(12)
    logits = (rng.standard_normal((len(idx), 362)) * 2.0).astype(np.float32)
    _, _, legal, status = E.game_derive(games["moves"][idx], games["num_moves"][idx], colors=games["colors"][idx], want_ladder=False)
    assert not status.any() and np.array_equal(legal, games["legal"][idx])
    L = oracle_lib.oracle()
    seeds = [L.orc_prng_seed(1000 + 17 * i) for i in range(len(idx))]
    state = np.array(seeds, dtype=np.uint64)
    k = 16
    moves, scores, kvalid = E.gumbel_topk(logits, legal, state, 1.0, k)
    for i in range(len(idx)):
        om, osc, okv, ost = oracle_lib.gumbel_topk(int(seeds[i]), logits[i], games["legal"][idx[i]], 1.0, k)
        kk = min(k, okv)
        assert kvalid[i] == okv and int(state[i]) == ost                      # same number of draws
        assert np.array_equal(scores[i][:kk], osc[:kk]) and np.array_equal(moves[i][:kk], om[:kk])   # bit-exact


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_slot_reloaded_during_the_copy_is_not_an_error(games, weight_dir):
    """LoadBatch may overlap RunInference (cc/nn/nn_interface.cc:276).  A game-record slot that is being
    re-loaded while its bank is copied can reach the GPU torn (move list of one game, length of another) and be rejected by the
    replay; that slot's result is unread by contract, so the run must NOT fail - while a genuinely impossible record in a
    stable slot still does (test_engine_rejects_impossible_game_record).  One thread hammers slot 0 with two different games
    while the main thread runs 400 steps over both submit/wait and the serial call; the other slots' results never change."""
    import threading
    from p3achygo_b200 import engine as E
    path, cfg, tensors = weight_dir("b10c128btl3")
    B = 8
    eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=E.PRECISION_BF16)
    order = np.argsort(games["num_moves"])
    long_i, short_i = int(order[-1]), int(order[len(order) // 8])
    recs = [(games["moves"][i][: games["num_moves"][i]].copy(), int(games["colors"][i])) for i in (long_i, short_i)]
    assert len(recs[0][0]) > 200 and 5 < len(recs[1][0]) < 120
    stable = [int(i) for i in order[-B:-1]]
    for bank in (0, 1):
        eng.LoadGameBank(bank, 0, recs[0][0], recs[0][1], 7.5)
        for b in range(1, B):
            i = stable[b - 1]
            eng.LoadGameBank(bank, b, games["moves"][i][: games["num_moves"][i]], int(games["colors"][i]), 7.5)
    eng.RunInference()
    want = [eng.GetBatch(b).copy() for b in range(B)]
    stop = threading.Event()

    def hammer():
        k = 0
        while not stop.is_set():
            for bank in (0, 1):
                eng.LoadGameBank(bank, 0, recs[k & 1][0], recs[k & 1][1], 7.5)
            k += 1

    t = threading.Thread(target=hammer)
    t.start()
    try:
        for it in range(400):
            if it % 4 == 3:
                eng.RunInference()
                got = [eng.GetBatch(b) for b in range(1, B)]
            else:
                bank = it & 1
                eng.Submit(bank)
                eng.Wait(bank)
                got = [eng.GetBatchBank(bank, b) for b in range(1, B)]
            for b in range(1, B):
                assert all(np.array_equal(got[b - 1][f], want[b][f]) for f in want[b].dtype.names), (it, b)
    finally:
        stop.set()
        t.join()
    # a stable impossible record is still an error
    eng.LoadGameBank(0, 0, np.array([5, 5 + E.MOVE_WHITE], dtype=np.int16), 1, 7.5)
    with pytest.raises(E.P3Error):
        eng.RunInference()
    eng.close()
