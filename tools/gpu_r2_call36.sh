#!/bin/bash
# few-voices build of k_render_free (free_lo.cu): parity suite with it (default: used whenever no CTA holds more than 8 rows —
# every 64 ... 4,096-voice test) and without it (SKB_LO=0), bench, shards of the bench job
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -6 | cut -c1-300
( time SKB_LO=0 timeout 1700 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu_lo0.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_lo0.log
grep -v "^#" gpurun_out/pytest_gpu_lo0.log | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-fast 2>gpurun_out/bench.err > gpurun_out/bench.json; echo "bench exit $?"; tail -2 gpurun_out/bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g  device ms %.4f  lat %.4f lat64 %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["block_latency_ms_p50"], d["block_latency_ms_p50_64_voices"]))
PY
for w in 8 4 2; do for lo in 1 0; do
  echo "== world $w SKB_LO=$lo"
  SKB_LO=$lo SKB_EARLY_FLUSH=0 timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -E "^launch +(8|11)" | cut -c1-60
done; done
