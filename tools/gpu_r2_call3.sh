#!/bin/bash
# round 2, call 3: parity of the TMA-staged build, A/B of the table staging variants, the 5 s full-width test again
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_event_fuzz.py -m gpu -q -x ) > gpurun_out/pytest_tma.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tma.log
tail -5 gpurun_out/pytest_tma.log
( time timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x -k first_5s ) > gpurun_out/pytest_5s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_5s.log
tail -4 gpurun_out/pytest_5s.log
rm -f gpurun_out/ab_tma.txt
run_ab() {  # name env...
  n=$1; shift
  echo "== $n $*" >> gpurun_out/ab_tma.txt
  env "$@" python tools/class_bench.py 65536 512 "plain_sine,lut(config2),korg(config3)" 2>&1 | grep -E "kernel ms" | sed 's/^/V=65536 /' >> gpurun_out/ab_tma.txt
  env "$@" python tools/class_bench.py 8192 512 "korg(config3)" 2>&1 | grep -E "kernel ms" | sed 's/^/V=8192  /' >> gpurun_out/ab_tma.txt
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-latency --min-timed-s 0.1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench value %.4g  ms/step %.4f  kernel_ms %.4f  fp32 frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/ab_tma.txt
}
for v in tma0 tma1 tma1_24k tma1_40k_env8; do
  run_ab $v SKB_ENGINE_LIB=$PWD/skred_b200/variants/$v/libskred_b200.so
done
run_ab tma0_noTblAffine SKB_ENGINE_LIB=$PWD/skred_b200/variants/tma0/libskred_b200.so SKB_TBL_AFFINE=0
cat gpurun_out/ab_tma.txt
