#!/bin/bash
# lone-warp speed: engine variants (frames per pipeline stage, warps per CTA = register budget) on one shard of 8 and on the full load
mkdir -p gpurun_out
for v in "" sub4_w8 sub8_w8 sub8_w14; do
  lib=""; [ -n "$v" ] && lib="$PWD/skred_b200/variants/$v/libskred_b200.so"
  for w in 8 1; do
    echo "== variant ${v:-default}  world $w"
    SKB_ENGINE_LIB=$lib SKB_EARLY_FLUSH=0 timeout 300 python tools/bench_probe.py 65536 12 1 8192 $w 2>&1 | grep -E "^launch +(8|11)|^   CTA +[0-9]+ [0-9.]+ us:" | head -4 | cut -c1-420
  done
done
