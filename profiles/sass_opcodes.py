#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell-native paths (B200_PROFILING.md): writes profiles/sass_opcodes.txt.
    python profiles/sass_opcodes.py            (needs cuobjdump and the built p3achygo_b200/libp3b200.so)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "p3achygo_b200", "libp3b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU", "HMMA", "FFMA", "REDUX", "SHFL", "VOTE"]
cur, counts, cta2 = None, collections.OrderedDict(), collections.Counter()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "").replace("p3::", "").split("(")[0]
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        if op in WANT:
            counts[cur][op] += 1
        if op == "UTCHMMA" and ".2CTA" in m.group(2):
            cta2[cur] += 1
with open(os.path.join(ROOT, "profiles", "sass_opcodes.txt"), "w") as f:
    f.write("# cuobjdump -sass p3achygo_b200/libp3b200.so (sm_100a), instruction counts per kernel; UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,\n"
            "# UTMALDG / UTMASTG = TMA load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops. No HMMA (legacy mma.sync) anywhere.\n")
    f.write("%-58s %s  UTCHMMA.2CTA\n" % ("kernel", " ".join("%8s" % w for w in WANT)))
    for k, c in counts.items():
        f.write("%-58s %s  %8d\n" % (k[:58], " ".join("%8d" % c[w] for w in WANT), cta2[k]))
print(open(os.path.join(ROOT, "profiles", "sass_opcodes.txt")).read())
