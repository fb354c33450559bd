#!/bin/bash
# round 2, call 6: per-stage clocks of k_render_rows; warp bins: parity suite + speed
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_event_fuzz.py -m gpu -q -x ) > gpurun_out/pytest_bins.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_bins.log
grep -v "^#" gpurun_out/pytest_bins.log | tail -30 | cut -c1-300
timeout 300 python tools/bins_bench.py 1024 512 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
timeout 300 python tools/gpu_fuzz_sweep.py 1 120 8 2>&1 | grep -v "^#" | tail -8 > gpurun_out/fuzz_dense_bins.txt; cat gpurun_out/fuzz_dense_bins.txt
rm -f gpurun_out/rows_speed.txt
for w in 8; do
  for m in 0 1; do
    echo "== world $w SKB_ROWS=$m" >> gpurun_out/rows_speed.txt
    SKB_EARLY_FLUSH=0 SKB_ROWS=$m timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -E "^launch +(8|11)|k_render_rows us" >> gpurun_out/rows_speed.txt
  done
done
cat gpurun_out/rows_speed.txt
