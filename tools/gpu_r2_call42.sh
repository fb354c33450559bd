#!/bin/bash
# round 2, session 3, call 3: envelope pre-pass dealt per (voice, 32-frame chunk) to warps + two-frame CTA row sum; A/B of the
# packed ops in the mix only (variants/mixx2); GPU suite; class bench; capture + stamped counters of the default build.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -6 | cut -c1-300
rm -f gpurun_out/ab_s3.txt
for n in default mixx2 default mixx2; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-8s value %.4g  ms/step %.4f  kernel_ms %.4f  frac %.4f  e2e %.4g' % ('$n', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/ab_s3.txt
done
unset SKB_ENGINE_LIB
cat gpurun_out/ab_s3.txt
timeout 300 python tools/class_bench.py 65536 512 2>&1 | grep -v "^#" | grep -v "phase us\|rows per class" > gpurun_out/class_s3.txt; cat gpurun_out/class_s3.txt
[ $rc -ne 0 ] && { echo "suite failed: no capture"; exit 1; }
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency --no-fast --min-timed-s 0"
timeout 300 python bench.py $NARGS > gpurun_out/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200
python tools/ncu_counters.py gpurun_out/prof.ncu-rep 65536 8192 gpurun_out/r02_ncu_counters.json | cut -c1-800
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json
