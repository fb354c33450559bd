#!/bin/bash
# round 2, session 3, call 7: the four-wide envelope pre-pass (div_tame) — exactness check of the division, A/B against the
# previous commit's engine (variants/prev), per-CTA probe, class bench (attack / decay columns), GPU suite.
mkdir -p gpurun_out; rm -f gpurun_out/ab_s3e.txt
./build/check_div_tame | tee gpurun_out/check_div_tame.txt
for n in default prev default prev; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['config']
print('%-13s value %.4g (ms/step %.4f; unflushed %.4g)  kernel_ms %.4f  frac %.4f  e2e %.4g (%.4f ms)' % ('$n', d['value'], d['ms_per_step'], c.get('value_l2_unflushed') or 0, d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))" >> gpurun_out/ab_s3e.txt
done
unset SKB_ENGINE_LIB
cat gpurun_out/ab_s3e.txt
timeout 200 python tools/bench_probe.py 65536 24 1 8192 2>&1 | grep -v "^#" > gpurun_out/probe_default.txt; grep -A9 "per-CTA body" gpurun_out/probe_default.txt | cut -c1-200
timeout 300 python tools/class_bench.py 65536 512 2>&1 | grep -v "^#" | grep -v "phase us\|rows per class" > gpurun_out/class_s3.txt; cat gpurun_out/class_s3.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -6 | cut -c1-300
