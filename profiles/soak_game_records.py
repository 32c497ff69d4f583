import sys, os, ctypes, time, numpy as np
sys.path.insert(0, os.getcwd())
from p3achygo_b200 import engine as E, weights as W, _lib
cfg = W.config_from_str("b12c256btl3")
W.save_weights("/tmp/w12.p3w", cfg, W.synthetic_weights(cfg, 0))
gz = np.load("tests/golden/ladder_games.npz")
g_moves = np.ascontiguousarray(gz["moves"][17:], dtype=np.int16)
g_num = np.ascontiguousarray(gz["num_moves"][17:], dtype=np.int32)
g_col = np.ascontiguousarray(gz["colors"][17:], dtype=np.int8)
host = ctypes.CDLL("p3achygo_b200/libp3host.so")
host.p3_host_benchmark_games.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p]
out = np.zeros(5)
t0 = time.time()
host.p3_host_benchmark_games(b"/tmp/w12.p3w", 0, 1024, 1, E.PRECISION_BF16, _lib.ptr(g_moves), _lib.ptr(g_num), _lib.ptr(g_col),
                             g_moves.shape[1], len(g_moves), 5, 400, 8, _lib.ptr(out))
print("soak games: 400 steps", time.time() - t0, "s; cycle us", out[0], "checksum", out[4], flush=True)
sym = np.load("tests/golden/known_answers.npz")["sym_fwd"]
mv = gz["moves"].astype(np.int32); pt = mv & 511; stone = (mv >= 0) & (pt < 361)
allm = []
for s in range(8):
    t = mv.copy(); t[stone] = sym[s][pt[stone]] + (mv[stone] & 512); allm.append(t.astype(np.int16))
allm = np.concatenate(allm); nm = np.tile(gz["num_moves"], 8); col = np.tile(gz["colors"], 8)
ref = None
t0 = time.time()
for rep in range(10):
    r = E.game_derive(allm, nm, colors=col)
    assert not r[3].any()
    if ref is None: ref = r
    else: assert all(np.array_equal(a, b) for a, b in zip(r[:3], ref[:3]))
print("derive stress ok: 10 x", len(allm), "records in", time.time() - t0, "s", flush=True)
