#!/bin/bash
# round 2, call 2: seed-127 diagnosis, the whole GPU suite (new: 5 s full-width parity), bench at N = 1 in the new format
mkdir -p gpurun_out
for env in "" "SKB_FORCE_GENERIC=1" "SKB_NO_BATCH=1"; do
  echo "== seed 127 dense 8 $env" >> gpurun_out/fuzz127.txt
  env $env timeout 300 python tools/gpu_fuzz_diag.py 127 8 2>&1 | grep -v "^#" >> gpurun_out/fuzz127.txt
done
cat gpurun_out/fuzz127.txt | cut -c1-400
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-3000
