#!/bin/bash
# round 2, session 3, call 5: which pairs to pack — default (mix only) against mix + biquad, mix + gain, mix + both, none
mkdir -p gpurun_out; rm -f gpurun_out/ab_s3c.txt
for n in default mix_biq mix_gain mix_biq_gain mixoff default mix_biq mix_gain mix_biq_gain mixoff; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('%-13s value %.4g (ms/step %.4f; unflushed %.4g)  kernel_ms %.4f  frac %.4f  e2e %.4g (%.4f ms)' % ('$n', d['value'], d['ms_per_step'], c.get('value_l2_unflushed') or 0, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))" >> gpurun_out/ab_s3c.txt
done
unset SKB_ENGINE_LIB
cat gpurun_out/ab_s3c.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -6 | cut -c1-300
