#!/bin/bash
# round 2, session 3, call 1: packed fp32 pairs (FADD2 / FMUL2, SKB_F32X2) — microbenchmark of the instructions, the whole GPU
# parity suite on the new engine, and an A/B against the scalar build of the same tree (skred_b200/variants/nox2).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
./build/ubench_f32x2 > gpurun_out/ubench_f32x2.txt 2>&1; cat gpurun_out/ubench_f32x2.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^#" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
rm -f gpurun_out/ab_x2.txt
for n in x2 nox2 x2 nox2; do
  if [ $n = nox2 ]; then export SKB_ENGINE_LIB=$PWD/skred_b200/variants/nox2/libskred_b200.so; else unset SKB_ENGINE_LIB; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-latency --no-fast 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-6s value %.4g  ms/step %.4f  kernel_ms %.4f  frac %.4f  e2e %.4g' % ('$n', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/ab_x2.txt
done
cat gpurun_out/ab_x2.txt
for n in x2 nox2; do
  if [ $n = nox2 ]; then export SKB_ENGINE_LIB=$PWD/skred_b200/variants/nox2/libskred_b200.so; else unset SKB_ENGINE_LIB; fi
  echo "# class_bench $n" >> gpurun_out/class_x2.txt
  timeout 300 python tools/class_bench.py 65536 512 2>&1 | grep -v "^#" >> gpurun_out/class_x2.txt
done
cat gpurun_out/class_x2.txt
