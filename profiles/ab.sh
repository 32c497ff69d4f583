#!/bin/bash
# A/B of experimental library builds on one box: scratch/ab.sh <rounds> <variant> [<variant> ...]   ("base" = the product library)
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for v in "$@"; do
    if [ "$v" = "base" ]; then lib=""; else lib="$PWD/p3achygo_b200/libp3b200_$v.so"; fi
    out=$(P3_LIB=$lib P3_PROFILE_CLASSES=1 python scratch/run_step.py b12c256btl3 1024 7 2>&1 | tail -2 | tr '\n' ' ')
    echo "round $r variant $v: $out"
  done
done
