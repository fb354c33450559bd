#!/bin/bash
# round 2, first GPU call at round-1 HEAD: dense random streams on the CUDA engine
# and the per-CTA / per-warp picture of the bench load at 65,536 and 8,192 voices per GPU.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
( time timeout 1200 python tools/gpu_fuzz_sweep.py 1 200 8 ) > gpurun_out/fuzz_dense.txt 2>&1
tail -3 gpurun_out/fuzz_dense.txt
timeout 300 python tools/bench_probe.py 65536 24 1 8192 > gpurun_out/probe_65536.txt 2>&1
timeout 300 python tools/bench_probe.py 8192 24 1 8192 > gpurun_out/probe_8192.txt 2>&1
timeout 300 python tools/class_bench.py 65536 512 > gpurun_out/class_bench.txt 2>&1
tail -30 gpurun_out/probe_65536.txt
