#!/bin/bash
# round 2, call 12: levelled modulation (k_render_levels): parity suite, dense fuzz, speed of 0.sk and the C-mod pair load
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_event_fuzz.py -m gpu -q -x ) > gpurun_out/pytest_lev.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_lev.log
grep -v "^#" gpurun_out/pytest_lev.log | tail -30 | cut -c1-400
timeout 300 python tools/bins_bench.py 1024 512 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
SKB_LEVELS=0 timeout 300 python tools/bins_bench.py 1024 512 2>&1 | grep -v "^#" | sed 's/^/[SKB_LEVELS=0] /' >> gpurun_out/bins_bench.txt; tail -4 gpurun_out/bins_bench.txt
timeout 600 python tools/gpu_fuzz_sweep.py 1 200 8 2>&1 | grep -v "^#" | tail -8 > gpurun_out/fuzz_dense_lev.txt; cat gpurun_out/fuzz_dense_lev.txt
