#!/bin/bash
# ncu --set full of k_render_free for single feature classes (sustain launch): tools/gpu_class_prof.sh "classes" [voices]
mkdir -p gpurun_out
CL="${1:-plain_sine}"
V="${2:-65536}"
IFS=',' read -ra ARR <<< "$CL"
for c in "${ARR[@]}"; do
  tag=$(echo "$c" | tr -c 'a-zA-Z0-9' '_')_$V
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 10 -c 1 -o gpurun_out/class_$tag -f \
     python tools/class_bench.py $V 512 "$c" > gpurun_out/class_ncu_$tag.log 2>&1
done
