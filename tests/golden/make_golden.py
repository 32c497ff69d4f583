"""Generates tests/golden/*.npz from the UNMODIFIED reference sources (oracle/_ref/libp3ref.so, built by
oracle/Makefile from /root/reference).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py            # test fixtures (1024 positions)
    python tests/golden/make_golden.py --bench    # + bench_positions.npz (8192 positions)

Positions are seeded random legal playouts from the empty board: length uniform in
[5, 300], uniformly random legal non-pass moves with 2 % passes, komi from {0.5, 6.5, 7.5}, a uniformly
random symmetry, colour to move alternating.  Everything stored is an output of the reference itself:
GoFeatures records (NNInterface::LoadBatch semantics), legal masks (Game::IsValidMove), liberty grids,
feature planes (nn::LoadGoFeatures), symmetry tables, PCG32 / Gumbel draws, root top-k samples, softmax.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_lib  # noqa: E402
from oracle.oracle_lib import P  # noqa: E402

FEAT_DTYPE = np.dtype({
    "names": ["bsize", "color", "komi", "board", "last_moves", "stones_atari", "stones_two_liberties",
              "stones_three_liberties", "stones_laddered"],
    "formats": ["<i4", "i1", "<f4", ("i1", 361), ("<i4", (5, 2)), ("i1", 361), ("i1", 361), ("i1", 361), ("i1", 361)],
    "offsets": [0, 4, 8, 12, 376, 416, 777, 1138, 1499],
    "itemsize": 1860,
})


def gen_positions(n: int, seed: int, per_game: int = 4):
    R = oracle_lib.ref()
    assert R is not None, "reference library unavailable (needs /root/reference)"
    rng = np.random.default_rng(seed)
    feats = np.zeros(n, dtype=FEAT_DTYPE)
    legal = np.zeros((n, 362), dtype=np.uint8)
    boards = np.zeros((n, 361), dtype=np.int8)       # un-rotated raw positions
    colors = np.zeros(n, dtype=np.int8)
    libs = np.zeros((n, 3, 361), dtype=np.int8)
    ladder = np.zeros((n, 361), dtype=np.int8)
    syms = np.zeros(n, dtype=np.int8)
    count = 0
    mask = np.zeros(362, dtype=np.uint8)
    while count < n:
        komi = float(rng.choice([0.5, 6.5, 7.5]))
        g = R.ref_game_new(komi, 1)
        length = int(rng.integers(5, 301))
        snaps = set(int(x) for x in rng.integers(5, length + 1, size=per_game)) | {length}
        color = 1
        for mv in range(1, length + 1):
            R.ref_game_legal_mask(g, color, P(mask))
            cand = np.flatnonzero(mask[:361])
            if len(cand) == 0 or rng.random() < 0.02:
                R.ref_game_play(g, 19, 0, color)
            else:
                m = int(rng.choice(cand))
                ok = R.ref_game_play(g, m // 19, m % 19, color)
                assert ok == 1
            color = -color
            if R.ref_game_is_over(g):
                break
            if mv in snaps and count < n:
                sym = int(rng.integers(0, 8))
                R.ref_game_features(g, color, sym, ctypes.c_void_p(feats[count:count + 1].ctypes.data))
                R.ref_game_legal_mask(g, color, P(legal[count]))
                R.ref_game_board(g, P(boards[count]))
                for k in range(3):
                    R.ref_game_liberties(g, k + 1, P(libs[count, k]))
                R.ref_game_laddered(g, P(ladder[count]))
                colors[count] = color
                syms[count] = sym
                count += 1
        R.ref_game_free(g)
    return dict(feats=feats, legal=legal, boards=boards, colors=colors, libs=libs, ladder=ladder, syms=syms)


def ref_planes(feats: np.ndarray, version: int):
    R = oracle_lib.ref()
    n = len(feats)
    npl, ns = (13, 7) if version == 0 else (15, 8)
    planes = np.empty((n, 19, 19, npl), dtype=np.float32)
    scalars = np.empty((n, ns), dtype=np.float32)
    R.ref_load_go_features(P(np.ascontiguousarray(feats)), n, version, P(planes), P(scalars))
    return planes, scalars


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--n", type=int, default=1024)
    args = ap.parse_args()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    R = oracle_lib.ref()

    pos = gen_positions(args.n, seed=20261018)
    planes, scalars = ref_planes(pos["feats"], 1)
    planes0, scalars0 = ref_planes(pos["feats"][:64], 0)
    digests = np.array([np.frombuffer(hashlib.sha256(planes[i].tobytes() + scalars[i].tobytes()).digest()[:8], dtype="<u8")[0]
                        for i in range(len(planes))], dtype=np.uint64)
    np.savez_compressed(
        os.path.join(out_dir, "positions.npz"), feats=pos["feats"].view(np.uint8).reshape(-1, 1860), legal=pos["legal"],
        boards=pos["boards"], colors=pos["colors"], libs=pos["libs"], ladder=pos["ladder"], syms=pos["syms"],
        planes_digest=digests, planes_first=planes[:64].astype(np.uint8), scalars=scalars,
        planes_v0_first=planes0.astype(np.uint8), scalars_v0_first=scalars0)

    # ---- symmetry tables, PRNG / probability known answers, softmax, Gumbel top-k
    sym_fwd = np.array([[R.ref_transform_index(s, i) for i in range(361)] for s in range(8)], dtype=np.int32)
    sym_inv = np.array([[R.ref_transform_inv(s, i) for i in range(361)] for s in range(8)], dtype=np.int32)
    seeds = np.array([0, 1, 7, 42, 12345, 2**40 + 17, 2**63 + 5], dtype=np.uint64)
    nexts = np.zeros((len(seeds), 32), dtype=np.uint32)
    unis = np.zeros((len(seeds), 32), dtype=np.float32)
    gums = np.zeros((len(seeds), 32), dtype=np.float32)
    rr = np.zeros((len(seeds), 32), dtype=np.int32)
    rsym = np.zeros((len(seeds), 32), dtype=np.int32)
    for si, s in enumerate(seeds):
        p = R.ref_prob_new(int(s))
        nexts[si] = [R.ref_prob_next(p) for _ in range(32)]
        R.ref_prob_free(p)
        p = R.ref_prob_new(int(s))
        unis[si] = [R.ref_prob_uniform(p) for _ in range(32)]
        R.ref_prob_free(p)
        p = R.ref_prob_new(int(s))
        gums[si] = [R.ref_prob_gumbel(p) for _ in range(32)]
        R.ref_prob_free(p)
        p = R.ref_prob_new(int(s))
        rr[si] = [R.ref_prob_rand_range(p, 3, 3 + 1 + (j * 37) % 361) for j in range(32)]
        R.ref_prob_free(p)
        p = R.ref_prob_new(int(s))
        rsym[si] = [R.ref_prob_random_symmetry(p) for _ in range(32)]
        R.ref_prob_free(p)

    rng = np.random.default_rng(99)
    sm_in = (rng.standard_normal((16, 362)) * rng.uniform(0.5, 6.0, (16, 1))).astype(np.float32)
    sm_out = np.zeros_like(sm_in)
    for i in range(16):
        R.ref_softmax362(P(sm_in[i]), P(sm_out[i]))

    n_roots, k = 128, 16
    g_logits = (rng.standard_normal((n_roots, 362)) * 2.0).astype(np.float32)
    g_legal = pos["legal"][:n_roots].copy()
    g_legal[5, :] = 0
    g_legal[5, [3, 77, 361]] = 1       # fewer legal moves than k
    g_legal[6, 361] = 0                # pass disabled
    g_seed = np.arange(n_roots, dtype=np.uint64) * np.uint64(7919) + np.uint64(11)
    g_moves = np.full((n_roots, k), -1, dtype=np.int32)
    g_scores = np.zeros((n_roots, k), dtype=np.float32)
    g_kvalid = np.zeros(n_roots, dtype=np.int32)
    g_next = np.zeros(n_roots, dtype=np.uint32)    # next() of the stream after sampling (pins the draw count)
    for i in range(n_roots):
        p = R.ref_prob_new(int(g_seed[i]))
        g_kvalid[i] = R.ref_gumbel_topk(p, P(g_logits[i]), P(g_legal[i]), 1.0, k, P(g_moves[i]), P(g_scores[i]))
        g_next[i] = R.ref_prob_next(p)
        R.ref_prob_free(p)
    np.savez_compressed(
        os.path.join(out_dir, "known_answers.npz"), sym_fwd=sym_fwd, sym_inv=sym_inv, seeds=seeds, nexts=nexts, unis=unis,
        gums=gums, rand_range=rr, rand_sym=rsym, softmax_in=sm_in, softmax_out=sm_out, g_logits=g_logits, g_legal=g_legal,
        g_seed=g_seed, g_moves=g_moves, g_scores=g_scores, g_kvalid=g_kvalid, g_next=g_next, g_k=np.int32(k))

    if args.bench:
        bp = gen_positions(8192, seed=777, per_game=8)
        np.savez_compressed(os.path.join(out_dir, "bench_positions.npz"), feats=bp["feats"].view(np.uint8).reshape(-1, 1860))
    for f in sorted(os.listdir(out_dir)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
