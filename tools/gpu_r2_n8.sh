#!/bin/bash
# round 2, 8 GPUs of one box (charged 8x: kept lean): sharded parity through the C-ABI exchange step at N = 8, the plain-C multi-GPU
# host, state migration at re-plan, then the bench at N = 4 and 8 the way the driver launches it (N = 1, 2: tools/gpu_r2_full.sh, _n2.sh)
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded_render_n_gpus and 8" ) > gpurun_out/pytest_n8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n8.log
grep -v "^#" gpurun_out/pytest_n8.log | tail -6 | cut -c1-400
cat gpurun_out/sharded_parity_n8.txt | cut -c1-400
gcc -O2 -Iinclude tools/c/multi_gpu_host.c -Lskred_b200 -lskred_b200 -Wl,-rpath,$PWD/skred_b200 -lm -o /tmp/multi_gpu_host &&
  ( for n in 2 8; do for o in 0 1; do timeout 200 /tmp/multi_gpu_host $n 8192 16 $o; echo "exit $?"; done; done ) 2>&1 | grep -v "^#" | tee gpurun_out/c_host_multi_gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/gpu_migration_check.py gpurun_out/migration_n8.txt 2>&1 | grep -v "^#" | tail -4 | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu --no-fast 2>gpurun_out/bench_n4.err > gpurun_out/scale_n4.json; echo "bench n4 exit $?"; tail -2 gpurun_out/bench_n4.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29568 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/bench_n8.err > gpurun_out/scale_n8.json; echo "bench n8 exit $?"; tail -2 gpurun_out/bench_n8.err | cut -c1-300
for n in 4 8; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print($n, "value %.4g e2e %.4g ms/step %.4f lat %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("block_latency_ms_p50")), (d.get("weak_scaling") or {}).get("value"))
except Exception as e:
    print($n, "unreadable", e)
PY
done
