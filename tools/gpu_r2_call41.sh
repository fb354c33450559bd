#!/bin/bash
# round 2, session 3, call 2: direct finish (k_finish_host -> mapped host memory, flag polling) against the staged finish, then the
# final-tree records: GPU suite, smoke, a sparse fuzz sweep, bench both arms, launch list, the ncu capture + stamped counters.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^#" | tail -3
rm -f gpurun_out/ab_finish.txt
for n in direct copy direct copy; do
  if [ $n = copy ]; then export SKB_FINISH_COPY=1; else unset SKB_FINISH_COPY; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d.get('block_latency_parts_ms') or {}
print('%-6s value %.4g ms/step %.4f | e2e %.4g ms/step %.4f | p50 block latency %.4f ms (64 voices %.4f) stream_wait %.4f device %.4f' % ('$n', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['block_latency_ms_p50'], d.get('block_latency_ms_p50_64_voices') or 0, p.get('stream_wait', 0), p.get('device_ms_last_block', 0)))" >> gpurun_out/ab_finish.txt
done
unset SKB_FINISH_COPY
cat gpurun_out/ab_finish.txt
[ $rc -ne 0 ] && { echo "suite failed: no records"; exit 1; }
( echo "# python tools/gpu_fuzz_sweep.py 100 80: sparse random skode streams (voices among all 64), session-3 tree"
  timeout 600 python tools/gpu_fuzz_sweep.py 100 80 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_fuzz_sparse.txt
tail -2 gpurun_out/r02_s3_gpu_fuzz_sparse.txt | cut -c1-300
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; tail -c 300 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -3 gpurun_out/bench.err | cut -c1-300; cut -c1-600 gpurun_out/bench.json
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency --no-fast --min-timed-s 0"
timeout 300 python bench.py $NARGS > gpurun_out/plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200
python tools/ncu_counters.py gpurun_out/prof.ncu-rep 65536 8192 gpurun_out/r02_ncu_counters.json | cut -c1-800
ncu -i gpurun_out/prof.ncu-rep --page details > gpurun_out/k_render_free_ncu_details.txt 2>&1
