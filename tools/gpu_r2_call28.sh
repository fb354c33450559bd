#!/bin/bash
# A/B of the predicate-free phase wrap at few voices per GPU (rank 0's shard of N, 8,192-frame launches) and on the full load
mkdir -p gpurun_out
for w in 8 4 2 1; do
  for lib in "" "$PWD/skred_b200/variants/wrap_select/libskred_b200.so"; do
    echo "== world $w  lib ${lib:-default (umin)}"
    SKB_ENGINE_LIB=$lib timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -E "^launch +(8|10|11)" | cut -c1-60
  done
done
