#!/bin/bash
# per-warp envelope slices: parity suite, class bench (attack / decay / sustain launches), bench
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -12 | cut -c1-300
timeout 600 python tools/class_bench.py 65536 512 2>&1 | grep -v "^#" | tee gpurun_out/class_bench.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-fast 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -3 gpurun_out/bench.err | cut -c1-300; cut -c1-400 gpurun_out/bench.json
for ef in 1024 2048 4096; do
  SKB_EARLY_FLUSH=$ef timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-fast --no-latency 2>/dev/null > gpurun_out/bench_ef$ef.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_ef$ef.json").read().strip().splitlines()[-1])
print("SKB_EARLY_FLUSH=$ef  value %.4g  e2e %.4g  e2e ms %.4f  device ms %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["ms_per_step"]), d["e2e"].get("host_ms_per_step"))
PY
done
