#!/bin/bash
# round 2, session 3, call 6: the planner's row costs (rows -> CTAs) against the measured per-CTA times
mkdir -p gpurun_out; rm -f gpurun_out/ab_s3d.txt
timeout 200 python tools/bench_probe.py 65536 24 1 8192 2>&1 | grep -v "^#" > gpurun_out/probe_default.txt; grep -A12 "per-CTA body" gpurun_out/probe_default.txt | cut -c1-260
for n in default r6_56 r6_66 r6_56_r4_40 os3 os6 default r6_56 r6_66 r6_56_r4_40 os3 os6; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast --min-timed-s 0.25 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('%-13s value %.4g (ms/step %.4f; unflushed %.4g)  kernel_ms %.4f  frac %.4f  e2e %.4g (%.4f ms)' % ('$n', d['value'], d['ms_per_step'], c.get('value_l2_unflushed') or 0, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))" >> gpurun_out/ab_s3d.txt
done
cat gpurun_out/ab_s3d.txt
