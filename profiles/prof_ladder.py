import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from p3achygo_b200 import engine as E
z = np.load("tests/golden/ladder_games.npz")
sel = slice(17, 17 + 1024)
out = E.game_derive(z["moves"][sel], z["num_moves"][sel], colors=z["colors"][sel], forbidden=z["forbidden"][sel])
assert np.array_equal(out[1], z["ladder"][sel])
print("ok")
