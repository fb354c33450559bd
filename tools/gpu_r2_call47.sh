#!/bin/bash
# round 2, session 3, call 8: final one-GPU records of the session's tree: both bench arms, launch list, ncu capture + stamped counters,
# modulated-voice probe, full bench line LAST (so that it finds the counters of this very tree).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^#" | tail -2
NARGS="--steps 2 --warmup 3 --no-cpu --no-latency --no-fast --min-timed-s 0"
timeout 300 python bench.py $NARGS > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $NARGS > gpurun_out/ncu1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 6 -c 1 -o gpurun_out/prof -f python bench.py $NARGS > gpurun_out/ncu2.log 2>&1
tail -1 gpurun_out/ncu2.log | cut -c1-200
python tools/ncu_counters.py gpurun_out/prof.ncu-rep 65536 8192 gpurun_out/r02_ncu_counters.json | cut -c1-600
cp gpurun_out/r02_ncu_counters.json profiles/r02_ncu_counters.json
timeout 300 python tools/bins_bench.py 1024 512 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; tail -6 gpurun_out/bins_bench.txt | cut -c1-200
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; tail -c 250 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -2 gpurun_out/bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','block_latency_ms_p50','block_latency_ms_p50_64_voices')}); print(d['roofline']); print(d['roofline_issue']); print(d['e2e']); print(d['modulation_groups']); print(d['fast_mode']['speedup_vs_parity_build'] if d.get('fast_mode') else None)"
