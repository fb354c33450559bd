#!/bin/bash
# round 2, session 3, call 15: wide random-stream sweeps on the round's final kernels
mkdir -p gpurun_out
( echo "# python tools/gpu_event_fuzz_sweep.py 2000 200 on B200, round-2 final tree"
  timeout 700 python tools/gpu_event_fuzz_sweep.py 2000 200 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_event_fuzz_2000.txt
tail -2 gpurun_out/r02_s3_gpu_event_fuzz_2000.txt | cut -c1-300
( echo "# python tools/gpu_event_fuzz_sweep.py 700 40 1 (negative amplitudes, 37,164 frames per seed) on B200, round-2 final tree"
  timeout 400 python tools/gpu_event_fuzz_sweep.py 700 40 1 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_event_fuzz_neg2.txt
tail -2 gpurun_out/r02_s3_gpu_event_fuzz_neg2.txt | cut -c1-300
( echo "# python tools/gpu_fuzz_sweep.py 2000 100 16: dense random skode streams (voices among the first 16), round-2 final tree"
  timeout 500 python tools/gpu_fuzz_sweep.py 2000 100 16 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_fuzz_dense16.txt
tail -3 gpurun_out/r02_s3_gpu_fuzz_dense16.txt | cut -c1-300
( echo "# python tools/gpu_fuzz_sweep.py 3000 100: sparse random skode streams (voices among all 64), round-2 final tree"
  timeout 500 python tools/gpu_fuzz_sweep.py 3000 100 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_fuzz_sparse2.txt
tail -3 gpurun_out/r02_s3_gpu_fuzz_sparse2.txt | cut -c1-300
