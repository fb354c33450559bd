#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bins_bench.py 1024 4096 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
for t in 1 0; do
  echo "== 512-frame launches of the bench load, SKB_TBL_AFFINE=$t"
  SKB_TBL_AFFINE=$t timeout 300 python tools/bench_probe.py 65536 40 1 512 2>&1 | grep -E "^launch +(16|24|32|39)|phase us"
done
