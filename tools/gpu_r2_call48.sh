#!/bin/bash
# round 2, session 3, call 9: predicate-free phase wrap (SKB_WRAP_UMIN=1) on this session's tree
mkdir -p gpurun_out; rm -f gpurun_out/ab_s3f.txt
for n in default umin default umin default umin; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('%-13s value %.4g (ms/step %.4f; unflushed %.4g)  kernel_ms %.4f  frac %.4f  e2e %.4g (%.4f ms)' % ('$n', d['value'], d['ms_per_step'], c.get('value_l2_unflushed') or 0, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))" >> gpurun_out/ab_s3f.txt
done
cat gpurun_out/ab_s3f.txt
export SKB_ENGINE_LIB=$PWD/skred_b200/variants/umin/libskred_b200.so
timeout 300 python tools/class_bench.py 65536 512 2>&1 | grep -v "^#" | grep -v "phase us\|rows per class" > gpurun_out/class_umin.txt; cat gpurun_out/class_umin.txt
