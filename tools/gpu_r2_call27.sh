#!/bin/bash
# predicate-free phase wrap (wrap_pick): parity suite, bench, class bench, rows kernel probe at 8,192 voices per GPU
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-fast --no-latency 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -2 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g  device ms %.4f  kernel_ms %.4f frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
PY
timeout 600 python tools/class_bench.py 65536 512 2>&1 | grep "kernel ms" | tee gpurun_out/class_bench.txt
for r in 0 1; do
  echo "== rank 0's shard of 8 (8,192 voices), 8,192-frame launches, SKB_ROWS=$r"
  SKB_ROWS=$r timeout 300 python tools/bench_probe.py 65536 12 1 8192 8 2>&1 | grep -E "^launch +(6|8|10|11)" | cut -c1-120
done
