#!/bin/bash
# round 2, session 3, 2 GPUs: the direct finish on rank 0 of a sharded render — sharded parity (both sum orders), the new
# direct-vs-staged finish test, state migration at re-plan, bench at N = 2.
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(sharded_render_n_gpus and 2) or direct_finish" ) > gpurun_out/pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n2.log
grep -v "^#" gpurun_out/pytest_n2.log | tail -6 | cut -c1-400
cat gpurun_out/sharded_parity_n2.txt | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/gpu_migration_check.py gpurun_out/migration_n2.txt 2>&1 | grep -v "^#" | tail -3 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-fast --no-cpu 2>gpurun_out/bench_n2.err > gpurun_out/scale_n2.json; echo "bench exit $?"; tail -3 gpurun_out/bench_n2.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/scale_n2.json')); print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','block_latency_ms_p50')}); print(d['e2e']['value'], d['roofline']['frac']); print((d.get('weak_scaling') or {}).get('value'))"
