#!/bin/bash
# Round-end confirmation in one gpurun call: parity tests (full-size jobs excluded: tests/test_gpu_full_size.py
# takes ~5 min of reference CPU time), smoke, both bench arms, launch list + one full ncu capture.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
BARGS="--steps 5 --warmup 3"
timeout 600 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_full_size.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^#" | tail -3
timeout 300 python bench.py --impl reference $BARGS 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; tail -c 300 gpurun_out/bench_ref.json
timeout 300 python bench.py $BARGS 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; cat gpurun_out/bench.json
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency"
timeout 200 python bench.py $NARGS > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
