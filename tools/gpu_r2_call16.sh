#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -x ) > gpurun_out/pytest_x.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_x.log
grep -v "^#" gpurun_out/pytest_x.log | tail -4 | cut -c1-300
timeout 300 python tools/bins_bench.py 1024 4096 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
for b in 0 1; do
  SKB_BALANCE=$b python bench.py --steps 20 --warmup 5 --no-cpu --no-fast --min-timed-s 0.3 --latency-blocks 600 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('SKB_BALANCE=$b: value %.4g  ms/step %.4f (reps min %.4f max %.4f)  kernel_ms %.4f  e2e %.4g (%.4f ms)  p50 block latency %.4f ms  64-voice %.4f' % (d['value'], d['ms_per_step'], d['config']['ms_per_step_repetitions']['min'], d['config']['ms_per_step_repetitions']['max'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['block_latency_ms_p50'], d['block_latency_ms_p50_64_voices']))"
done
