#!/bin/bash
# A/B of engine tuning builds (skred_b200/variants/*): lone-row and full-load class speed
mkdir -p gpurun_out; rm -f gpurun_out/ab.txt
for d in skred_b200/variants/*/; do
  n=$(basename $d); export SKB_ENGINE_LIB=$PWD/$d/libskred_b200.so
  echo "== $n" >> gpurun_out/ab.txt
  python tools/class_bench.py 4096 512 "plain_sine,korg(config3)" 2>&1 | grep -E "kernel ms" | sed 's/^/V=4096  /' >> gpurun_out/ab.txt
  python tools/class_bench.py 65536 512 "plain_sine,korg(config3)" 2>&1 | grep -E "kernel ms" | sed 's/^/V=65536 /' >> gpurun_out/ab.txt
done
cat gpurun_out/ab.txt
