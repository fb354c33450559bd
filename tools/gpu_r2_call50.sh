#!/bin/bash
# round 2, session 3, call 11: which build fails the event-fuzz sweep — this tree, this tree without packed ops, the round's previous final
mkdir -p gpurun_out; rm -f gpurun_out/evfuzz_which.txt
for n in default scalar r2final; do
  if [ $n = default ]; then unset SKB_ENGINE_LIB; else export SKB_ENGINE_LIB=$PWD/skred_b200/variants/$n/libskred_b200.so; fi
  echo "== $n" >> gpurun_out/evfuzz_which.txt
  timeout 300 python tools/gpu_event_fuzz_sweep.py 300 20 2>&1 | grep -v "^#" | tail -1 | cut -c1-300 >> gpurun_out/evfuzz_which.txt
done
unset SKB_ENGINE_LIB
cat gpurun_out/evfuzz_which.txt
timeout 300 python tools/gpu_event_fuzz_diag.py 303 1536 2>&1 | grep -v "^#" | cut -c1-600 > gpurun_out/evfuzz_diag_303.txt; head -40 gpurun_out/evfuzz_diag_303.txt
