#!/bin/bash
# round 2, one GPU: the random-skode parity sweeps against the compiled reference on the final kernels (bins, levels, free).
mkdir -p gpurun_out
( echo "# python tools/gpu_fuzz_sweep.py 1 200 8 on B200, round-2 final tree: 200 DENSE random skode streams (voices among the first 8: modulation"
  echo "# chains -> k_render_levels for the acyclic ones, k_render_bins_warp for loops; re-plans every few lines), every synth.def array word for word"
  timeout 1500 python tools/gpu_fuzz_sweep.py 1 200 8 2>&1 | grep -v "^#" ) > gpurun_out/r02_gpu_fuzz_dense.txt
tail -4 gpurun_out/r02_gpu_fuzz_dense.txt | cut -c1-400
( echo "# python tools/gpu_fuzz_sweep.py 1000 150 16: voices among the first 16"
  timeout 1200 python tools/gpu_fuzz_sweep.py 1000 150 16 2>&1 | grep -v "^#" ) > gpurun_out/r02_gpu_fuzz_dense16.txt
tail -3 gpurun_out/r02_gpu_fuzz_dense16.txt | cut -c1-400
( echo "# python tools/gpu_fuzz_sweep.py 100 200: sparse streams (voices among all 64)"
  timeout 1200 python tools/gpu_fuzz_sweep.py 100 200 2>&1 | grep -v "^#" ) > gpurun_out/r02_gpu_fuzz_sparse.txt
tail -3 gpurun_out/r02_gpu_fuzz_sparse.txt | cut -c1-400
