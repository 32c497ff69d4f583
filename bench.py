#!/usr/bin/env python
"""Leaf-evaluation throughput of the B200 evaluator (BASELINE.json metric: leaf evals/sec, b12c256btl3, batch 1024).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference path restated on the host cores

A "step" is one pass of the hot path over one batch of 1024 synthetic positions (encode -> tower -> heads).
  value : whole-job leaf evals/s with the batch's game state already resident in HBM (device time, CUDA events on the
          engine's stream, max over ranks).
  e2e   : the same through the engine calls with HOST buffers, timed by the C++ harness that mirrors nn::Benchmark
          (p3achygo_b200/host/benchmark_engine.cc; reference: cc/nn/engine/benchmark_engine.cc:77-109): every position is
          loaded from host memory, copied H2D, evaluated, copied D2H and read into a caller-owned NNInferResult.
          e2e.value uses the engine's two slot banks (LoadBatchBank x B -> Submit | Wait -> GetBatchBank x B, one bank's host
          phases and copies overlapping the other's kernels); e2e.serial is the reference's un-overlapped cycle
          LoadBatch x B -> RunInference -> GetBatch x B.
  e2e_from_game_records : the slot-bank cycle with the slots loaded as move lists (rules derived on the GPU inside the step).
  sustained : 150 back-to-back device steps with the NVML clocks of that interval (the board power limit shows here).
  game_records (rank 0) : p3_game_derive stand-alone on 1024 fixture records, checked against the reference fixture, with the
          compiled reference timed on the same records when oracle/_ref is present.
Prints ONE JSON line on rank 0 (stdout carries nothing else).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG = "b12c256btl3"
BATCH = 1024
H2D_PER_POS = 1860   # sizeof(GoFeatures)
D2H_PER_POS = 7568   # sizeof(NNInferResult)


def workload_config(config: str, batch: int, world: int) -> dict:
    """`config` of the JSON line: the SAME dict in both arms (the driver compares them), nothing arm-specific in it."""
    return {"workload": f"{config} leaf evaluation (encode -> tower -> heads), batch {batch} per GPU, seeded random-playout "
                        "positions, seeded synthetic weights",
            "batch_per_gpu": batch, "parallelism": f"replicas x{world} (independent games, no collective on the data path)",
            "l2": "inputs larger than L2: the activation working set per step (~1 GB at batch 1024) exceeds the 126 MB L2 and the "
                  "inputs cycle over 8192 positions"}


def kernel_source_hash() -> str:
    """Identifies the kernels a committed ncu capture was taken from: sha256 over the tensor-path sources."""
    import hashlib
    h = hashlib.sha256()
    for name in ("conv_tc.cu", "chain_tc.cu", "pw_tc.cu", "broadcast_tc.cu", "init_tc.cu", "init_tc2.cu", "heads.cu", "tc_util.cuh", "ptx.cuh", "common.cuh"):
        with open(os.path.join(ROOT, "p3achygo_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def selfplay_child(wpath: str, device: int, seconds: float):
    """Body of the self-play leg; runs in a child process (see selfplay_leg) and prints its result as one SELFPLAY_JSON line."""
    ci = ctypes.c_int

    def load(name):
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(path):
            return None
        L = ctypes.CDLL(path)
        L.ref_selfplay_gumbel.argtypes = [ctypes.c_char_p, ci, ci, ci, ci, ci, ctypes.c_double, ci, ci, ctypes.POINTER(ctypes.c_longlong),
                                          ctypes.POINTER(ctypes.c_double)]
        L.ref_selfplay_record_loads.restype = ctypes.c_longlong
        return L

    def run(L, n, k, interfaces=8, slots=128):
        out = (ctypes.c_longlong * 4)()
        secs = ctypes.c_double(0)
        rc = L.ref_selfplay_gumbel(wpath.encode(), device, interfaces, slots, n, k, seconds, 1 << 20, 300, out, ctypes.byref(secs))
        if rc != 0 or secs.value <= 0:
            return None
        return {"n": n, "k": k, "interfaces": interfaces, "slots_per_interface": slots, "moves": int(out[0]), "seconds": secs.value,
                "moves_per_s": out[0] / secs.value,
                "leaf_evals_per_s": out[1] / secs.value, "avg_engine_batch": out[1] / max(out[2], 1), "games_finished": int(out[3]),
                "slots_loaded_as_game_records": int(L.ref_selfplay_record_loads())}

    L = load("libp3refnn.so")
    runs = [run(L, n, k) for n, k in ((96, 8), (600, 16))]   # default and selected n / k of config/v3-b12c256btl3-2000k-inf.json
    if any(r is None for r in runs):
        runs = None
    res = {"runs": runs}
    # the same search over the reference with INTEGRATION.md's optional edit 5 (oracle/ref_patches/0002): NNInterface::LoadBatch hands
    # the game record to the engine, which derives board / liberties / laddered stones / last moves on the GPU (p3_engine_load_game_bank)
    G = load("libp3refnn_gr.so")
    if G is not None and runs is not None:
        res["game_record_slots"] = [run(G, 96, 8), run(G, 96, 8, 8, 256)]
    if runs is not None:  # kMaxNumThreads slots per interface (cc/constants/constants.h:78) instead of 128: fuller engine batches
        res["runs_256_slots"] = [run(L, 96, 8, 6, 256)]
    print("SELFPLAY_JSON " + json.dumps(res), flush=True)


def selfplay_leg(wpath: str, device: int, seconds: float):
    """Self-play moves/s through the reference's OWN, unmodified search: oracle/_ref/libp3refnn.so = cc/nn/nn_interface.cc +
    cc/mcts/* + cc/game/* compiled from /root/reference (oracle/Makefile) and linked with the product's nn::B200Engine adapter;
    game threads call GumbelEvaluator::SearchRoot per move (cc/mcts/gumbel.cc:260).  The engine under test is libp3b200; the
    callers are the reference's (that is the point: they drop onto it unchanged).  None when the harness was not built.
    Runs in a child process: 1024 game threads of reference code are test infrastructure, and whatever happens to them must not
    take the bench line with it ({"error": ...} instead)."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libp3refnn.so")):
        return None
    cmd = [sys.executable, os.path.abspath(__file__), "--selfplay-child", wpath, str(device), str(seconds)]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=10 * seconds + 300)
    except subprocess.TimeoutExpired:
        return {"error": "self-play child timed out"}
    for ln in p.stdout.splitlines():
        if ln.startswith("SELFPLAY_JSON "):
            res = json.loads(ln[len("SELFPLAY_JSON "):])
            if not res.get("runs"):
                return {"error": "self-play harness returned an error"}
            selfplay_leg.game_record_slots = res.get("game_record_slots")
            selfplay_leg.runs_256 = res.get("runs_256_slots")
            return res["runs"]
    return {"error": f"self-play child exited {p.returncode}: {(p.stderr or '').strip().splitlines()[-1:]}"}


def load_positions():
    from p3achygo_b200.layout import GO_FEATURES_DTYPE
    z = np.load(os.path.join(ROOT, "tests", "golden", "bench_positions.npz"))
    return np.ascontiguousarray(z["feats"]).view(GO_FEATURES_DTYPE).reshape(-1)


def game_record_leg(device: int):
    """Rules from game records on the GPU (p3_game_derive: replay -> ladder reader -> exact legal mask), 1024 of the fixture
    games of tests/golden/ladder_games.npz, host buffers in and out; the compiled reference timed beside it when present."""
    from p3achygo_b200 import engine as E
    z = np.load(os.path.join(ROOT, "tests", "golden", "ladder_games.npz"))
    sel = slice(17, 17 + 1024)
    mv, nm, col, fb = z["moves"][sel], z["num_moves"][sel], z["colors"][sel], z["forbidden"][sel]
    E.game_derive(mv, nm, colors=col, forbidden=fb, device=device)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        boards, lad, legal, status = E.game_derive(mv, nm, colors=col, forbidden=fb, device=device)
    dt = (time.perf_counter() - t0) / reps
    ok = bool(np.array_equal(lad, z["ladder"][sel]) and np.array_equal(legal, z["legal"][sel]) and not status.any())
    out = {"positions": 1024, "ms_per_batch": dt * 1e3, "positions_per_s": 1024 / dt, "bit_exact_vs_reference_fixture": ok,
           "what": "p3_game_derive: move lists -> boards, Board::GetLadderedStones, Game::IsValidMove (superko) for 1024 "
                   "random-playout game records, wall time incl. H2D / D2H and per-call scratch allocation"}
    try:
        from oracle import oracle_lib
        from oracle.oracle_lib import P
        R = oracle_lib.ref()
    except Exception:
        R = None
    if R is not None and hasattr(R, "ref_game_move_status"):
        lad1 = np.zeros(361, dtype=np.int8)
        mask = np.zeros(362, dtype=np.uint8)
        tl = tm = 0.0
        n_cpu = 128
        for k in range(17, 17 + n_cpu):
            g = R.ref_game_new(7.5, 1)
            for code in z["moves"][k][: z["num_moves"][k]]:
                pnt, c = int(code) & 511, (-1 if int(code) & 512 else 1)
                R.ref_game_play(g, 19 if pnt == 361 else pnt // 19, 0 if pnt == 361 else pnt % 19, c)
            t0 = time.perf_counter(); R.ref_game_laddered(g, P(lad1)); tl += time.perf_counter() - t0
            t0 = time.perf_counter(); R.ref_game_legal_mask(g, int(z["colors"][k]), P(mask)); tm += time.perf_counter() - t0
            R.ref_game_free(g)
        out["reference_cpu"] = {"kind": "reference", "cores": 1, "sample": f"{n_cpu} of the same game records",
                                "laddered_us_per_position": tl / n_cpu * 1e6, "legal_mask_us_per_position": tm / n_cpu * 1e6}
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU WHILE the timed region runs: NVML polled every 5 ms from a thread
    (the timed region of a default run is ~0.1 s, too short for `nvidia-smi -lms`), nvidia-smi as a fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self._nvml = None
        self._thread = threading.Thread(target=self._poll if self._nvml else self._poll_smi, daemon=True)
        self._thread.start()

    def _poll(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self._visible_index())],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                return

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": "nvml" if self._nvml else "nvidia-smi"}


def dist_setup(n_gpus: int):
    """One process per GPU under torchrun; returns (rank, world, local_rank, reduce_max, barrier)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return rank, world, local, (lambda x, op="max": x), (lambda: None)
    import torch
    import torch.distributed as dist
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local) if backend == "nccl" else torch.device("cpu")
    if backend == "nccl":
        dist.init_process_group(backend=backend, device_id=dev)
    else:
        dist.init_process_group(backend=backend)

    def reduce_max(x: float, op: str = "max") -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    def barrier():
        dist.barrier()
        if backend == "nccl":
            torch.cuda.synchronize()

    return rank, world, local, reduce_max, barrier


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's path restated on the host cores (features: C oracle; net: PyTorch restatement)
# --------------------------------------------------------------------------------------------------
def cpu_leaf_evals(feats, cfg, tensors, n_sample: int, threads: int):
    import torch
    from oracle import oracle_lib
    from oracle.model_ref import RefModel
    torch.set_num_threads(threads)
    model = RefModel(cfg, tensors, dtype=torch.float32)
    sample = np.resize(feats, n_sample) if n_sample > len(feats) else feats[:n_sample]
    t0 = time.perf_counter()
    for lo in range(0, n_sample, 128):  # slices bound the activation memory of the CPU forward
        planes, scalars = oracle_lib.load_go_features(sample[lo:lo + 128], 1)
        out = model.forward(planes, scalars)
        assert np.isfinite(out["pi_logits"]).all()
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from p3achygo_b200 import weights as W
    cfg = W.config_from_str(CONFIG)
    tensors = W.synthetic_weights(cfg, 0)
    feats = load_positions()
    cores = os.cpu_count() or 1
    # one step = one full batch of the workload (1024 positions, ~3 s on 24 cores), evaluated in slices of 128
    n_sample = int(os.environ.get("P3_CPU_SAMPLE", str(BATCH)))
    for _ in range(max(min(args.warmup, 2), 1)):
        cpu_leaf_evals(feats, cfg, tensors, 64, cores)
    total_t, total_n = 0.0, 0
    for s in range(args.steps):
        _, dt = cpu_leaf_evals(np.roll(feats, -s * n_sample), cfg, tensors, n_sample, cores)
        total_t += dt
        total_n += n_sample
    value = total_n / total_t
    line = {
        "impl": "reference", "metric": "leaf evals/sec", "value": value, "unit": "positions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(CONFIG, BATCH, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "positions/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} positions/step x {args.steps} steps; features: C restatement of "
                                   "go_features.cc, net: PyTorch-CPU restatement of python/model.py (TensorFlow absent)"},
        "e2e": {"value": value, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    if len(sys.argv) == 5 and sys.argv[1] == "--selfplay-child":
        return selfplay_child(sys.argv[2], int(sys.argv[3]), float(sys.argv[4]))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=CONFIG)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, ...) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world, local, reduce_max, barrier = dist_setup(args.gpus)
    from p3achygo_b200 import engine as E
    from p3achygo_b200 import weights as W
    from p3achygo_b200 import _lib

    cfg = W.config_from_str(args.config)
    B = args.batch
    feats = load_positions()
    # games are independent: rank r evaluates its own shard of the synthetic positions (no collective on the data path)
    shard = np.roll(feats, -rank * (len(feats) // max(world, 1)))
    tmpdir = tempfile.mkdtemp(prefix="p3bench_")
    wpath = os.path.join(tmpdir, f"{args.config}.p3w")
    tensors = W.synthetic_weights(cfg, 0)
    W.save_weights(wpath, cfg, tensors)
    precision = {"bf16": E.PRECISION_BF16, "fp16": E.PRECISION_FP16, "fp32": E.PRECISION_FP32}[args.precision]
    eng = E.CreateEngine(E.KindFromEnginePath(wpath), wpath, B, 1, precision=precision, device=local)

    def stage_batch(i: int):
        base = (i * B) % len(shard)
        idx = (np.arange(B) + base) % len(shard)
        eng.LoadBatchAll(shard[idx])
        eng.Upload()  # H2D outside the timed region: `value` is measured with inputs resident in HBM

    # ---- device-resident throughput (value)
    for i in range(args.warmup):
        stage_batch(i)
        eng.RunDevice()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ms_steps = []
    for i in range(args.steps):
        stage_batch(args.warmup + i)
        ms_steps.append(eng.RunDevice())
    barrier()
    clocks = sampler.stop()
    total_ms = reduce_max(float(np.sum(ms_steps)))
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---- per-kernel-class device times (eager pass, an event around every launch), averaged over a few passes
    prof = {}
    n_prof = 3
    for _ in range(n_prof):
        p = eng.Profile()
        for k, (ms, n, fl, nb) in p.items():
            a = prof.setdefault(k, [0.0, n, fl, nb])
            a[0] += ms / n_prof
    peaks = measured_peaks()
    dom = prof["conv3x3"] if prof["conv3x3"][1] > 0 else prof["conv1x1"]
    dom_name = "tc_conv3x3_pair_kernel (3x3 layers)" if prof["conv3x3"][1] > 0 else "1x1 layers"
    # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
    traffic, traffic_src, ncu_share, ncu_tensor = None, None, None, None
    # (tagged with the hash of the kernel sources it was captured from, and dropped when they have changed since)
    tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    chain_traffic = None
    if os.path.exists(tpath) and args.config == CONFIG and B == BATCH and prof["conv3x3"][1] > 0:
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("kernel_source_hash") == kernel_source_hash():
            k = tj["kernels"].get("tc_conv3x3_pair_kernel")
            if k:
                traffic, traffic_src = k["dram_bytes"], tj["source"]
                ncu_share, ncu_tensor = k.get("share_of_step_in_launch_list"), k.get("tensor_pipe_active_pct")
            kc = tj["kernels"].get("tc_chain_pair_kernel")
            chain_traffic = kc["dram_bytes"] if kc else None
        else:
            traffic_src = "profiles/r2_ncu_traffic.json is stale (kernel sources changed since the capture): traffic not reported"
    achieved = dom[2] / (dom[0] * 1e-3) / 1e12 if dom[0] > 0 else 0.0
    all_conv_ms = prof["conv1x1"][0] + prof["conv3x3"][0] + prof["head_conv"][0] + prof["boundary"][0]
    all_conv_fl = prof["conv1x1"][2] + prof["conv3x3"][2] + prof["head_conv"][2] + prof["boundary"][2]
    step_ms = sum(v[0] for v in prof.values())
    roofline = {
        # the per-class pass is a ~15 ms eager burst at un-capped clocks (see `clocks`), so the like-for-like denominator is the
        # BURST peak; the sustained pairing is reported from the back-to-back leg below
        "bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_burst"], "frac_burst": achieved / peaks["bf16_burst"],
        "peak_source": peaks["source"] + ", burst figure (the per-kernel pass is a short burst at un-capped clocks)",
        "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)", "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": 2.0 * 400 * cfg.bottleneck_channels * 2 * B if hasattr(cfg, "bottleneck_channels") else None,
        "ncu_share_of_step": ncu_share, "ncu_tensor_pipe_active_pct": ncu_tensor,
        "launches_per_step": dom[1], "avg_launch_ms": dom[0] / max(dom[1], 1),
        "share_of_step": dom[0] / step_ms if step_ms else None,
        "all_conv_tflops": all_conv_fl / (all_conv_ms * 1e-3) / 1e12 if all_conv_ms else None,
        "whole_step_tflops": cfg.flops_per_position() * B / (np.mean(ms_steps) * 1e-3) / 1e12,
        "whole_step_frac_of_burst_peak": cfg.flops_per_position() * B / (np.mean(ms_steps) * 1e-3) / 1e12 / peaks["bf16_burst"],
        # the fused block boundaries (35 % of the step): HBM-bound, algorithmic bytes = every operand read / written once
        "boundary": {"kernel": "tc_chain_pair_kernel", "bound": "hbm", "launches_per_step": prof["boundary"][1],
                     "avg_launch_ms": prof["boundary"][0] / max(prof["boundary"][1], 1),
                     "algorithmic_bytes_per_step": prof["boundary"][3],
                     "achieved_GBps": prof["boundary"][3] / (prof["boundary"][0] * 1e-3) / 1e9 if prof["boundary"][0] else None,
                     "peak_GBps": peaks["hbm"],
                     "hbm_frac": prof["boundary"][3] / (prof["boundary"][0] * 1e-3) / 1e9 / peaks["hbm"] if prof["boundary"][0] else None,
                     # avg_launch_ms is an eager pass with an event around every launch (no graph, nothing of a launch hidden under
                     # its predecessor's tail); the graph-replayed step is shorter than the sum of that pass by this factor, and
                     # the same factor applied to this class gives the in-step estimate
                     "graph_over_eager": float(np.mean(ms_steps)) / step_ms if step_ms else None,
                     "hbm_frac_in_graph_estimate": (prof["boundary"][3] / (prof["boundary"][0] * 1e-3 * float(np.mean(ms_steps)) / step_ms)
                                                    / 1e9 / peaks["hbm"]) if prof["boundary"][0] and step_ms else None,
                     "traffic": chain_traffic},
        "class_hbm_frac": {k: (v[3] / (v[0] * 1e-3) / 1e9 / peaks["hbm"] if v[0] else None) for k, v in prof.items()},
        "kernel_ms": {k: round(v[0], 4) for k, v in prof.items()},
        "kernel_launches": {k: v[1] for k, v in prof.items()},
        "hbm": {"encode_GBps": (H2D_PER_POS + 21692 + 722) * B / (prof["encode"][0] * 1e-3) / 1e9 if prof["encode"][0] else None,
                "heads_GBps": (361 * 3 * cfg.head_channels * 4 + 9000) * B / (prof["heads"][0] * 1e-3) / 1e9 if prof["heads"][0] else None,
                "peak_GBps": peaks["hbm"]},
    }

    # ---- end-to-end through the reference-shaped engine calls with host buffers (C++ harness)
    host = ctypes.CDLL(os.path.join(ROOT, "p3achygo_b200", "libp3host.so"))
    host.p3_host_benchmark.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    out = np.zeros(5, dtype=np.float64)
    threads = min(8, os.cpu_count() or 1)
    barrier()
    host.p3_host_benchmark(wpath.encode(), local, B, 1, precision, _lib.ptr(shard), len(shard), args.warmup, args.steps, threads,
                           _lib.ptr(out))
    barrier()
    cycle_us = reduce_max(float(out[3]))
    run_us = reduce_max(float(out[0]))
    serial_value = world * B / (cycle_us * 1e-6)
    # the same cycle over the engine's two slot banks (p3_engine_submit / p3_engine_wait): host phases and both copies of one
    # bank overlap the other bank's kernels; every position still goes host -> GPU -> host inside the timed region
    host.p3_host_benchmark_pipelined.argtypes = host.p3_host_benchmark.argtypes
    outp = np.zeros(5, dtype=np.float64)
    barrier()
    host.p3_host_benchmark_pipelined(wpath.encode(), local, B, 1, precision, _lib.ptr(shard), len(shard), args.warmup, args.steps,
                                     threads, _lib.ptr(outp))
    barrier()
    pipe_us = reduce_max(float(outp[0]))
    e2e_value = world * B / (pipe_us * 1e-6)
    # ... and with compact leaf records (P3_RESULT_LEAF): only the 4360 B per slot that InitFields keeps cross PCIe
    host.p3_host_benchmark_pipelined_leaf.argtypes = host.p3_host_benchmark.argtypes
    outl = np.zeros(5, dtype=np.float64)
    barrier()
    host.p3_host_benchmark_pipelined_leaf(wpath.encode(), local, B, 1, precision, _lib.ptr(shard), len(shard), args.warmup, args.steps,
                                          threads, _lib.ptr(outl))
    barrier()
    leaf_us = reduce_max(float(outl[0]))
    e2e_leaf = {"value": world * B / (leaf_us * 1e-6), "unit": "positions/s", "cycle_us": leaf_us, "h2d_bytes_per_step": H2D_PER_POS * B,
                "d2h_bytes_per_step": 4360 * B, "get_leaf_us": float(outl[2]), "wait_us": float(outl[3]),
                "path": "as e2e, with p3_engine_set_result_mode(P3_RESULT_LEAF): GetLeafBank x B returns the three policy arrays + value / "
                        "E[score] / Var[score] / err of mcts::LeafEvaluator InitFields (cc/mcts/leaf_evaluator.cc:83-112)"}

    # ---- the same pipelined cycle with the slots loaded as GAME RECORDS (move lists): board, liberty grids, laddered stones and last
    # moves are derived on the GPU inside the step (p3_engine_load_game_bank) - the work NNInterface::LoadBatch does on the host
    gz = np.load(os.path.join(ROOT, "tests", "golden", "ladder_games.npz"))
    g_moves = np.ascontiguousarray(gz["moves"][17:], dtype=np.int16)
    g_num = np.ascontiguousarray(gz["num_moves"][17:], dtype=np.int32)
    g_col = np.ascontiguousarray(gz["colors"][17:], dtype=np.int8)
    host.p3_host_benchmark_games.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_void_p]
    outg = np.zeros(5, dtype=np.float64)
    barrier()
    host.p3_host_benchmark_games(wpath.encode(), local, B, 1, precision, _lib.ptr(g_moves), _lib.ptr(g_num), _lib.ptr(g_col),
                                 g_moves.shape[1], len(g_moves), args.warmup, args.steps, threads, _lib.ptr(outg))
    barrier()
    games_us = reduce_max(float(outg[0]))
    e2e_games = {"value": world * B / (games_us * 1e-6), "unit": "positions/s", "cycle_us": games_us,
                 "h2d_bytes_per_step": int((1860 + 2 * 1024 + 4 + 361 + 1) * B), "d2h_bytes_per_step": D2H_PER_POS * B,
                 "path": "LoadGameBank x B (move lists of 1280 random-playout game records, cycled) -> Submit | Wait -> GetBatchBank x B; "
                         "replay + ladder reader + liberties + feature assembly run on the GPU in front of the encode kernel"}

    # ---- the same step replayed back to back (no host staging between steps): what a saturated evaluator sustains.
    # The board power limit, not the kernels, sets this number: NVML reports sw_power_cap and lower SM clocks here.
    n_sus = int(os.environ.get("P3_SUSTAINED_STEPS", "150"))
    sampler2 = ClockSampler(local)
    for _ in range(10):
        eng.RunDevice()
    barrier()
    sampler2.start()
    sus_ms = [eng.RunDevice() for _ in range(n_sus)]
    barrier()
    clocks2 = sampler2.stop()
    sus_tail = float(np.mean(sus_ms[n_sus // 2:]))
    sus_tail = reduce_max(sus_tail)
    sustained = {"value": world * B / (sus_tail * 1e-3), "unit": "positions/s", "ms_per_step": sus_tail, "steps": n_sus,
                 "note": "back-to-back steps, mean of the second half; the device-resident `value` above has the host's "
                         "staging of the next batch between steps", "clocks": clocks2}

    sus_tflops = cfg.flops_per_position() * B / (sus_tail * 1e-3) / 1e12
    sustained["whole_step_tflops"] = sus_tflops
    sustained["whole_step_frac_of_sustained_peak"] = sus_tflops / peaks["bf16_sustained"]
    if dom[0] > 0:
        # dominant kernel like for like under the cap: its share of the eager pass applied to the back-to-back step time
        roofline["frac_sustained"] = (dom[2] / (sus_tail * 1e-3 * dom[0] / step_ms) / 1e12) / peaks["bf16_sustained"]

    eng.close()

    # ---- self-play through the reference's unmodified NNInterface + Gumbel search (the other half of BASELINE.json's metric)
    selfplay = None
    if os.environ.get("P3_BENCH_SELFPLAY", "1") != "0" and args.config == CONFIG:
        barrier()
        sp = selfplay_leg(wpath, local, float(os.environ.get("P3_SELFPLAY_SECONDS", "8")))
        barrier()
        # every rank takes part in the same three reductions whatever happened to its child
        ok = isinstance(sp, list) and len(sp) == 2
        any_failed = reduce_max(0.0 if ok else 1.0) > 0
        totals = [reduce_max(sp[i]["moves_per_s"] if ok else 0.0, op="sum") for i in range(2)]
        if ok and not any_failed:
            for r, t in zip(sp, totals):
                r["moves_per_s_all_gpus"] = t
            selfplay = {"unit": "moves/s", "host_cores": os.cpu_count(), "interfaces_per_gpu": 8, "slots_per_interface": 128,
                        "runs": sp,
                        "what": "reference cc/nn/nn_interface.cc + cc/mcts (unmodified, compiled from /root/reference into "
                                "oracle/_ref/libp3refnn.so) over nn::B200Engine: one game thread per slot, GumbelEvaluator::SearchRoot per "
                                "move from the empty board, NN cache 2^20 keyed on the last move (cc/selfplay/main.cc:177), timeout 400 us; "
                                "Game -> GoFeatures (ladders, liberties) on the host cores as the reference does it; run in a child "
                                "process"}
            r256 = getattr(selfplay_leg, "runs_256", None)
            if r256 and r256[0]:
                selfplay["runs_256_slots_per_interface"] = r256  # unmodified reference, 6 interfaces x 256 slots; this rank only
            gr = getattr(selfplay_leg, "game_record_slots", None)
            if gr and gr[0]:
                selfplay["with_game_record_slots"] = {
                    "runs": gr,
                    "what": "the same harness over the reference with INTEGRATION.md's optional edit 5 (oracle/ref_patches/0002-load-game-"
                            "records.patch, 3 hunks): NNInterface::LoadBatch hands the move list to nn::Engine::LoadGameRecord and the engine "
                            "derives board, liberty grids, laddered stones and last moves on the GPU; this rank only"}
        elif isinstance(sp, dict):
            selfplay = sp  # {"error": ...}: the child failed on this rank
        elif sp is not None:
            selfplay = {"error": "the self-play child failed on another rank"}

    line = {
        "metric": "leaf evals/sec", "value": value, "unit": "positions/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args.config, B, world),
        "engine": {"cuda_graph": True, "library": "libp3b200.so (hand-written sm_100a kernels; links no cuBLAS / cuDNN / NCCL)",
                   "collectives": "none on the data path; torch.distributed (NCCL) is initialised only for the harness barrier and the "
                                  "max-over-ranks time"},
        "value_is": "device-resident burst throughput (inputs in HBM, host staging between steps lets the clocks recover); `e2e` is the "
                    "host-buffer number, `sustained` the back-to-back power-capped one",
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "positions/s", "h2d_bytes_per_step": H2D_PER_POS * B, "d2h_bytes_per_step": D2H_PER_POS * B,
                "cycle_us": pipe_us, "load_batch_us": float(outp[1]), "get_batch_us": float(outp[2]), "wait_us": float(outp[3]),
                "host_threads": threads,
                "path": "C++ nn::B200Engine over two slot banks: LoadBatchBank x B -> Submit | Wait -> GetBatchBank x B, bank k's "
                        "host phases and copies overlapping bank 1-k's kernels (host/benchmark_engine.cc, p3_host_benchmark_pipelined)",
                "serial": {"value": serial_value, "cycle_us": cycle_us, "run_inference_us": run_us, "load_batch_us": float(out[1]),
                           "get_batch_us": float(out[2]),
                           "path": "the reference's own cycle, nothing overlapped: LoadBatch x B -> RunInference -> GetBatch x B "
                                   "(cc/nn/engine/benchmark_engine.cc:77-109 shape)"}},
        "gpu_launches": None,
        "e2e_leaf": e2e_leaf,
        "e2e_from_game_records": e2e_games,
        "sustained": sustained,
        "selfplay": selfplay,
        "roofline": roofline,
    }
    if rank == 0:
        line["game_records"] = game_record_leg(local)
    line["gpu_launches"] = int(sum(v[1] for v in prof.values())) * args.steps

    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        n_sample = int(os.environ.get("P3_CPU_SAMPLE", "4096"))  # ~10-15 s of CPU work
        cpu_leaf_evals(feats, cfg, tensors, 64, cores)  # warm-up
        v, dt = cpu_leaf_evals(feats, cfg, tensors, n_sample, cores)
        line["cpu_baseline"] = {"value": v, "unit": "positions/s", "cores": cores, "kind": "port",
                                "sample": f"{n_sample} positions of the same workload in {dt:.1f} s; features: C restatement of "
                                          "go_features.cc, net: PyTorch-CPU restatement of python/model.py (TensorFlow absent)"}
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
