#!/bin/bash
# round 2, 2 GPUs: sharded parity through the C-ABI exchange step (both orders, determinism), state migration at re-plan, bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded_render_n_gpus and 2" ) > gpurun_out/pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n2.log
grep -v "^#" gpurun_out/pytest_n2.log | tail -6 | cut -c1-400
cat gpurun_out/sharded_parity_n2.txt | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/gpu_migration_check.py gpurun_out/migration_n2.txt 2>&1 | grep -v "^#" | tail -4 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-fast 2>gpurun_out/bench_n2.err > gpurun_out/scale_n2.json; echo "bench exit $?"; tail -3 gpurun_out/bench_n2.err | cut -c1-300; cut -c1-600 gpurun_out/scale_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 2>/dev/null > gpurun_out/scale_ref_n2.json; tail -c 300 gpurun_out/scale_ref_n2.json
