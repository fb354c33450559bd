#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/rows_speed.txt
for w in 8; do
  for m in 1; do
    echo "== world $w SKB_ROWS=$m" >> gpurun_out/rows_speed.txt
    SKB_EARLY_FLUSH=0 SKB_ROWS=$m timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -E "^launch +(8|11)|k_render_rows us" >> gpurun_out/rows_speed.txt
  done
done
echo "== class-pure, 8192 voices on the GPU (world 8 of 65536), SKB_ROWS=0/1" >> gpurun_out/rows_speed.txt
cat gpurun_out/rows_speed.txt
