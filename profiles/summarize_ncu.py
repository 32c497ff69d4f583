#!/usr/bin/env python
"""Turns the ncu artefacts of one gpurun call into the tracked summaries under profiles/.

    [ROUND=r2] python profiles/summarize_ncu.py <tag> <launch_list.csv> <full_capture.ncu-rep | raw_page.csv> [more ...]

Writes profiles/<round>_launches_<tag>.csv (copy), profiles/<round>_ncu_full_<tag>_summary.csv (one row per captured launch)
and the per-kernel means + share of the step in the launch list: profiles/r1_ncu_traffic_<tag>.json for round 1, and from
round 2 on profiles/<round>_ncu_traffic.json, which carries `kernel_source_hash` (bench.kernel_source_hash() of the sources the
capture was taken from); bench.py reads roofline.traffic from it and drops it when the sources have changed since.  Needs the
`ncu` CLI (no GPU) to read .ncu-rep files.
"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROUND = os.environ.get("ROUND", "r2")
KERNELS = r"(tc_conv3x3_pair_kernel|tc_chain_pair_kernel|tc_pw_pair_kernel|tc_broadcast_kernel|init_tc_kernel|tc_conv3x3_res_kernel|tc_conv_kernel|heads_kernel|encode_kernel|[a-z_0-9]+_kernel)"
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size"]


def kname(full):
    m = re.search(KERNELS, full)
    return m.group(1) if m else full[:40]


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else v * 1000.0 if r[ui] == "ms" else v
        a = agg.setdefault(kname(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    return {k: {"launches": n, "sum_us": v, "share": v / tot} for k, (n, v) in agg.items()}, tot, len(data)


TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def raw_rows(rep):
    """rep: a .ncu-rep, or the `ncu -i rep --page raw --csv` text saved on the GPU box (a full capture of a whole step with
    sources is larger than gpurun brings back).  Times are normalised to us, DRAM bytes to MB (the raw page picks its own units)."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for k in KEYS:
            if k not in hdr:
                continue
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            if k == "gpu__time_duration.sum":
                v = str(float(v.replace(",", "")) * TO_US.get(u, 1.0))
            elif k.startswith("dram__bytes"):
                v = str(float(v.replace(",", "")) * TO_MB.get(u, 1.0))
            d[k] = v
        res.append(d)
    return res


def main():
    tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    shutil.copy(launches, os.path.join(HERE, f"{ROUND}_launches_{tag}.csv"))
    shares, tot, n = launch_shares(launches)
    rows = [d for rep in reps for d in raw_rows(rep)]
    with open(os.path.join(HERE, f"{ROUND}_ncu_full_{tag}_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["# ncu --set full --clock-control none --import-source on; bench.py --steps 2 --warmup 3 --no-cpu-baseline "
                    "(b12c256btl3 @ 1024); one captured launch per row; time in us, dram bytes in MB"])
        w.writerow(KEYS)
        for d in rows:
            w.writerow([d.get(k, "") for k in KEYS])
    acc = collections.defaultdict(list)
    for d in rows:
        acc[kname(d["Kernel Name"])].append(d)
    out = {"source": f"profiles/{ROUND}_ncu_full_{tag}_summary.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                     f"b12c256btl3 batch 1024); shares from profiles/{ROUND}_launches_{tag}.csv ({n} launches, {tot:.0f} us)", "kernels": {}}
    if ROUND != "r1":
        sys.path.insert(0, os.path.dirname(HERE))
        import bench
        out["kernel_source_hash"] = bench.kernel_source_hash()
        out["tag"] = tag
    for k, v in acc.items():
        f = lambda key: sum(float(d.get(key) or 0) for d in v) / len(v)  # noqa: E731
        out["kernels"][k] = {"captured_launches": len(v), "dram_read_bytes": f("dram__bytes_read.sum") * 1e6,
                             "dram_write_bytes": f("dram__bytes_write.sum") * 1e6,
                             "dram_bytes": (f("dram__bytes_read.sum") + f("dram__bytes_write.sum")) * 1e6,
                             "time_us_under_ncu": f("gpu__time_duration.sum"),
                             "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                             "share_of_step_in_launch_list": shares.get(k, {}).get("share")}
    json.dump(out, open(os.path.join(HERE, f"r1_ncu_traffic_{tag}.json" if ROUND == "r1" else f"{ROUND}_ncu_traffic.json"), "w"), indent=1)
    for k, s in shares.items():
        print(f"{k:28s} n={s['launches']:3d} avg={s['sum_us'] / s['launches']:7.1f} us share={s['share']:.3f}")
    for k, v in out["kernels"].items():
        print(f"{k:28s} dram {v['dram_bytes'] / 1e6:7.1f} MB  {v['time_us_under_ncu']:7.1f} us  tensor {v['tensor_pipe_active_pct']:.1f}%")


if __name__ == "__main__":
    main()
