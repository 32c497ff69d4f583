#!/bin/bash
# scratch/ab2.sh <rounds> <variant> ...  : prints compact per-class times
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for v in "$@"; do
    if [ "$v" = "base" ]; then lib=""; else lib="$PWD/p3achygo_b200/libp3b200_$v.so"; fi
    P3_LIB=$lib P3_PROFILE_CLASSES=1 python profiles/run_step.py b12c256btl3 1024 9 2>&1 | tail -2 | tr '\n' ' ' | python -c "
import sys,re,ast
line=sys.stdin.read()
m=re.search(r'ms per step: min ([\d.]+) median ([\d.]+)',line)
if not m: print('$v', line[:300])
else:
    d=ast.literal_eval(line[line.index('{'):line.rindex('}')+1])
    print('round $r %-6s' % '$v','min',m.group(1),'med',m.group(2),' '.join(f'{k}={v[0]:.3f}' for k,v in d.items()), re.search(r'checksum ([\d.]+)',line).group(1))
"
  done
done
