#!/bin/bash
# One gpurun call: parity tests, smoke, a short bench, launch list, one ncu capture.
# usage: tools/gpu_check.sh [tests] [bench] [ncu]
mkdir -p gpurun_out
WHAT="${@:-tests bench ncu}"
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
BARGS="--steps 5 --warmup 3"
if [[ "$WHAT" == *tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^#" | tail -5
fi
if [[ "$WHAT" == *bench* ]]; then
  timeout 900 python bench.py --impl reference $BARGS 2>gpurun_out/bench_ref.err | grep -v "^#" > gpurun_out/bench_ref.json; tail -c 600 gpurun_out/bench_ref.json
  timeout 900 python bench.py $BARGS 2>gpurun_out/bench.err | grep -v "^#" > gpurun_out/bench.json; echo "bench exit $?"; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
if [[ "$WHAT" == *ncu* ]]; then
  NARGS="--steps 2 --warmup 3 --no-cpu --no-latency"
  timeout 600 python bench.py $NARGS > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
  timeout 600 python bench.py $NARGS > gpurun_out/plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
  tail -3 gpurun_out/ncu2.log
fi
