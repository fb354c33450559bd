#!/bin/bash
# round 2, session 3, call 4: A/B/C of the packed fp32 ops: default (scalar) / mix only (variants/mixx2) / everywhere (variants/x2)
mkdir -p gpurun_out; rm -f gpurun_out/ab_s3b.txt
for n in default mixx2 x2 default mixx2 x2 default mixx2; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('%-8s value %.4g (ms/step %.4f; unflushed %.4g, %.4f)  kernel_ms %.4f  frac %.4f  e2e %.4g (%.4f ms)' % ('$n', d['value'], d['ms_per_step'], c.get('value_l2_unflushed') or 0, c.get('ms_per_step_l2_unflushed') or 0, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))" >> gpurun_out/ab_s3b.txt
done
cat gpurun_out/ab_s3b.txt
