#!/bin/bash
# A/B of engine tuning builds (skred_b200/variants/*) on the bench workload: device value + kernel ms
mkdir -p gpurun_out; rm -f gpurun_out/ab_bench.txt
for d in skred_b200/variants/*/; do
  n=$(basename $d); export SKB_ENGINE_LIB=$PWD/$d/libskred_b200.so
  python bench.py --steps 10 --warmup 3 --no-cpu --no-latency 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-12s value %.4g  ms/step %.4f  kernel_ms %.4f  e2e %.4g' % ('$n', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))" >> gpurun_out/ab_bench.txt
done
cat gpurun_out/ab_bench.txt
