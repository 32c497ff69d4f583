"""Minimal driver for profiling: b12c256btl3 @ 1024 (or argv), N device-resident steps. Usage: run_step.py [config] [batch] [steps]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from p3achygo_b200 import engine as E, weights as W
from p3achygo_b200.layout import GO_FEATURES_DTYPE

config = sys.argv[1] if len(sys.argv) > 1 else "b12c256btl3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_positions.npz"))
key = "feats" if "feats" in z.files else z.files[0]
raw = np.ascontiguousarray(z[key]).view(np.uint8).reshape(-1, 1860)
feats = np.ascontiguousarray(raw[:B]).reshape(-1).view(GO_FEATURES_DTYPE)
cfg = W.config_from_str(config)
d = tempfile.mkdtemp()
path = os.path.join(d, "w.p3w")
W.save_weights(path, cfg, W.synthetic_weights(cfg, 0))
prec = {"bf16": E.PRECISION_BF16, "fp16": E.PRECISION_FP16, "fp32": E.PRECISION_FP32}[os.environ.get("P3_PRECISION", "bf16")]
eng = E.CreateEngine(E.Kind.kB200, path, B, 1, precision=prec)
eng.LoadBatchAll(feats)
eng.Upload()
ms = [eng.RunDevice() for _ in range(steps)]
r0 = eng.GetAux(0)
print("ms per step: min %.4f median %.4f" % (min(ms), float(np.median(ms))), "lib", os.environ.get("P3_LIB", "base").split("/")[-1],
      "checksum", float(np.sum(np.asarray(r0["pi_logits_aux"], dtype=np.float64))))
if os.environ.get("P3_PROFILE_CLASSES"):
    print(eng.Profile())
eng.close()
