#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rows_speed.txt
( SKB_ROWS=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_event_fuzz.py -m gpu -q -x ) > gpurun_out/pytest_rows.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_rows.log
grep -v "^#" gpurun_out/pytest_rows.log | tail -4 | cut -c1-300
export SKB_EARLY_FLUSH=0
for m in 0 1; do
  echo "== class-pure, 8192 voices on the GPU (rank 0 of 8 x 65536), 8192-frame calls, SKB_ROWS=$m   [A work, G work, C work, A wait, G wait, C wait] us per CTA" >> gpurun_out/rows_speed.txt
  SKB_CB_WORLD=8 SKB_ROWS=$m timeout 300 python tools/class_bench.py 65536 8192 "plain_sine,lut(config2),korg(config3)" 2>&1 | grep -E "kernel ms|phase us" >> gpurun_out/rows_speed.txt
done
for w in 8 4; do
  for m in 0 1; do
    echo "== world $w SKB_ROWS=$m" >> gpurun_out/rows_speed.txt
    SKB_ROWS=$m timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -E "^launch +(8|11)|k_render_rows us" >> gpurun_out/rows_speed.txt
  done
done
cat gpurun_out/rows_speed.txt
