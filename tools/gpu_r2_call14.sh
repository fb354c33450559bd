#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_event_fuzz.py -m gpu -q -x ) > gpurun_out/pytest_lev.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_lev.log
grep -v "^#" gpurun_out/pytest_lev.log | tail -30 | cut -c1-400
timeout 300 python tools/bins_bench.py 1024 512 2>&1 | grep -v "^#" > gpurun_out/bins_bench.txt; cat gpurun_out/bins_bench.txt
