#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "exchange_canary or oversized" 2>&1 | grep -v "^#" | tail -15 | cut -c1-400
cat gpurun_out/exchange_canary.txt | cut -c1-400
