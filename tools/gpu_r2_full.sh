#!/bin/bash
# round 2, one GPU: the whole GPU suite, smoke, the modulated-voice probe, both bench arms, the launch list of the bench command.
# (The one `ncu --set full` capture of the dominant kernel is tools/gpu_r2_capture.sh, a call of its own.)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -12 | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^#" | tail -3
timeout 300 python tools/bins_bench.py 1024 512 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; tail -c 400 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -3 gpurun_out/bench.err | cut -c1-300; cat gpurun_out/bench.json | cut -c1-1500
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency --no-fast --min-timed-s 0"
timeout 300 python bench.py $NARGS > gpurun_out/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-300
