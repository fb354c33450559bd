#!/bin/bash
# round 2, call 8: where k_render_rows spends its time (class-pure loads, stage widths, one ncu source capture)
mkdir -p gpurun_out; rm -f gpurun_out/rows_ab.txt
export SKB_EARLY_FLUSH=0
for m in 0 1; do
  echo "== class-pure, 8192 voices on the GPU (rank 0 of 8 x 65536), 512-frame calls, SKB_ROWS=$m" >> gpurun_out/rows_ab.txt
  SKB_CB_WORLD=8 SKB_ROWS=$m timeout 300 python tools/class_bench.py 65536 512 "plain_sine,lut(config2),korg(config3),pcm(config4) alive" 2>&1 | grep -E "kernel ms" >> gpurun_out/rows_ab.txt
done
for v in rows_g8 rows_fb64 rows_fb64_g8; do
  echo "== $v world 8 SKB_ROWS=1" >> gpurun_out/rows_ab.txt
  SKB_ENGINE_LIB=$PWD/skred_b200/variants/$v/libskred_b200.so SKB_ROWS=1 timeout 300 python tools/bench_probe.py 65536 12 1 8192 8 2>&1 | grep -E "^launch +(8|11)|k_render_rows us" >> gpurun_out/rows_ab.txt
done
cat gpurun_out/rows_ab.txt
SKB_ROWS=1 timeout 300 python tools/bench_probe.py 65536 10 1 8192 8 > gpurun_out/plain_rows.log 2>&1 &&
SKB_ROWS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_rows -s 6 -c 1 -f -o gpurun_out/prof_rows python tools/bench_probe.py 65536 10 1 8192 8 > gpurun_out/ncu_rows.log 2>&1
tail -3 gpurun_out/ncu_rows.log; ls -la gpurun_out/prof_rows.ncu-rep
