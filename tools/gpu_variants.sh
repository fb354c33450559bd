#!/bin/bash
# class_bench for every engine tuning build under skred_b200/variants/: tools/gpu_variants.sh "classes"
mkdir -p gpurun_out
CL="${1:-plain_sine,plain+filter,plain+cz1,lut(config2),korg(config3)}"
for d in skred_b200/variants/*/; do
  n=$(basename $d)
  echo "== $n" >> gpurun_out/variants.txt
  SKB_ENGINE_LIB=$PWD/$d/libskred_b200.so timeout 600 python tools/class_bench.py 65536 512 "$CL" 2>&1 | grep -v "^#" >> gpurun_out/variants.txt
done
cat gpurun_out/variants.txt
