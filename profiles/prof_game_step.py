import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from p3achygo_b200 import engine as E, weights as W
z = np.load("tests/golden/ladder_games.npz")
cfg = W.config_from_str("b12c256btl3")
W.save_weights("/tmp/w12.p3w", cfg, W.synthetic_weights(cfg, 0))
B = 1024
eng = E.CreateEngine(E.Kind.kB200, "/tmp/w12.p3w", B, 1, precision=E.PRECISION_BF16)
eng.set_cuda_graph(False)   # plain launches so that ncu lists every kernel of the step
for rep in range(3):
    for b in range(B):
        g = 17 + (rep * B + b) % 1280
        eng.LoadGameBank(0, b, z["moves"][g][: z["num_moves"][g]], int(z["colors"][g]), 7.5, None, b % 8)
    eng.RunInference()
print("ok", float(eng.GetBatch(0)["value_probs"][1]))
eng.close()
