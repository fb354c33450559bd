#!/bin/bash
# round 2, session 3, call 16 (last): after the finished-by-op voice_sample fix — the failing seeds + 150 more, GPU suite, then the
# records of the final tree (launch list, capture + stamped counters, bench)
mkdir -p gpurun_out
( echo "# python tools/gpu_event_fuzz_sweep.py 2000 200 on B200, round-2 FINAL tree (after the finished-by-op voice_sample fix)"
  timeout 500 python tools/gpu_event_fuzz_sweep.py 2000 200 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_event_fuzz_2000.txt
tail -2 gpurun_out/r02_s3_gpu_event_fuzz_2000.txt | cut -c1-300
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -5 | cut -c1-300
[ $rc -ne 0 ] && { echo "suite failed: no records"; exit 1; }
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency --no-fast --min-timed-s 0"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200
python tools/ncu_counters.py gpurun_out/prof.ncu-rep 65536 8192 gpurun_out/r02_ncu_counters.json | cut -c1-300
cp gpurun_out/r02_ncu_counters.json profiles/r02_ncu_counters.json
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','block_latency_ms_p50','block_latency_ms_p50_64_voices')}); print(d['roofline']['frac'], d['roofline']['traffic'], (d['roofline_issue'] or {}).get('frac')); print(d['e2e']['value'])"
