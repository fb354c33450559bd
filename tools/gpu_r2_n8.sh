#!/bin/bash
# round 2, 8 GPUs of one box: sharded parity through the C-ABI exchange step at N = 2 and 8, the plain-C multi-GPU host, state
# migration at re-plan, then the bench at N = 1, 2, 4, 8 the way the driver launches it (reference arm once).
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sharded_render_n_gpus" ) > gpurun_out/pytest_n8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_n8.log
grep -v "^#" gpurun_out/pytest_n8.log | tail -8 | cut -c1-400
ls gpurun_out/sharded_parity_n*.txt; tail -6 gpurun_out/sharded_parity_n8.txt | cut -c1-400
gcc -O2 -Iinclude tools/c/multi_gpu_host.c -Lskred_b200 -lskred_b200 -Wl,-rpath,$PWD/skred_b200 -lm -o /tmp/multi_gpu_host &&
  ( for n in 2 8; do for o in 0 1; do timeout 300 /tmp/multi_gpu_host $n 8192 16 $o; echo "exit $?"; done; done ) 2>&1 | grep -v "^#" | tee gpurun_out/c_host_multi_gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/gpu_migration_check.py gpurun_out/migration_n8.txt 2>&1 | grep -v "^#" | tail -6 | cut -c1-400
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench_n1.err > gpurun_out/scale_n1.json; echo "bench n1 exit $?"
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29560 + n)) bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/bench_n$n.err > gpurun_out/scale_n$n.json; echo "bench n$n exit $?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-300
done
for n in 1 2 4 8; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print($n, "value %.4g e2e %.4g ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: d[k] for k in d if k.startswith("weak") or k.startswith("block_lat")})
except Exception as e:
    print($n, "unreadable", e)
PY
done
