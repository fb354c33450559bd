#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bins_bench.py 1024 4096 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
for v in default fma_only; do
  if [ $v = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$v/libskred_b200.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu --no-fast --min-timed-s 0.1 --latency-blocks 600 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v: value %.4g  ms/step %.4f  kernel_ms %.4f  e2e %.4g (%.4f ms)  p50 block latency %.4f ms  64-voice %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['block_latency_ms_p50'], d['block_latency_ms_p50_64_voices'])); print(json.dumps(d.get('modulation_groups')))" 
done
