#!/bin/bash
# ncu --set full of k_render_free for single feature classes (sustain launch): tools/gpu_class_prof.sh "plain_sine,korg(config3)"
mkdir -p gpurun_out
CL="${1:-plain_sine}"
python tools/class_bench.py 65536 512 "$CL" > gpurun_out/class_plain.txt 2>&1 || exit 1
IFS=',' read -ra ARR <<< "$CL"
for c in "${ARR[@]}"; do
  tag=$(echo "$c" | tr -c 'a-zA-Z0-9' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_render_free -s 10 -c 1 -o gpurun_out/class_$tag -f \
     python tools/class_bench.py 65536 512 "$c" > gpurun_out/class_ncu_$tag.log 2>&1
done
grep -v "^#" gpurun_out/class_plain.txt
