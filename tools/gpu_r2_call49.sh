#!/bin/bash
# round 2, session 3, call 10: random-stream sweeps on the session's final kernels (free + levels + bins + direct finish)
mkdir -p gpurun_out
( echo "# python tools/gpu_event_fuzz_sweep.py 300 60 on B200, session-3 final tree: random TIMESTAMPED event streams over a mixed 1,024-voice load,"
  echo "# 512 ... 8,192-frame synth() calls against the reference fed callback by callback (in-kernel boundary ops, envelope pre-pass, direct finish)"
  timeout 900 python tools/gpu_event_fuzz_sweep.py 300 60 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_event_fuzz.txt
tail -2 gpurun_out/r02_s3_gpu_event_fuzz.txt | cut -c1-300
( echo "# python tools/gpu_fuzz_sweep.py 1 80 8: DENSE random skode streams (voices among the first 8), session-3 final tree"
  timeout 900 python tools/gpu_fuzz_sweep.py 1 80 8 2>&1 | grep -v "^#" ) > gpurun_out/r02_s3_gpu_fuzz_dense.txt
tail -3 gpurun_out/r02_s3_gpu_fuzz_dense.txt | cut -c1-300
