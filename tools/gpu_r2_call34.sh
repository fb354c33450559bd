#!/bin/bash
# per-CTA / per-warp picture of one shard of the bench job cut 2 and 8 ways (one launch per call)
mkdir -p gpurun_out
for w in 2 8; do
  echo "== world $w"
  SKB_EARLY_FLUSH=0 timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -v "^#" | grep -vE "^launch +[0-7] " | cut -c1-700
done > gpurun_out/probe_shards.txt
cat gpurun_out/probe_shards.txt
