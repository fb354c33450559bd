#!/bin/bash
# round 2, one GPU: ONE `ncu --set full` capture of the dominant kernel of the bench command (after the same command ran clean
# without ncu), and the counters bench.py quotes (DRAM bytes, warp instructions of one launch) stamped ON THIS BOX with the hash of
# the kernel sources the capture ran.
mkdir -p gpurun_out
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency --no-fast --min-timed-s 0"
timeout 300 python bench.py $NARGS > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-300
python tools/ncu_counters.py gpurun_out/prof.ncu-rep 65536 8192 gpurun_out/r02_ncu_counters.json | cut -c1-1200
ncu -i gpurun_out/prof.ncu-rep --page details > gpurun_out/k_render_free_ncu_full.txt 2>&1
