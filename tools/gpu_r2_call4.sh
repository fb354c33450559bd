#!/bin/bash
# round 2, call 4: A/B of the mono tile; ncu L1TEX counters of the table-staging A/B (class-pure Korg load)
mkdir -p gpurun_out
rm -f gpurun_out/ab_mono.txt
run_ab() {
  n=$1; shift
  echo "== $n $*" >> gpurun_out/ab_mono.txt
  env "$@" python tools/class_bench.py 65536 512 "plain_sine,lut(config2),korg(config3)" 2>&1 | grep -E "kernel ms" | sed 's/^/V=65536 /' >> gpurun_out/ab_mono.txt
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-latency --min-timed-s 0.1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench value %.4g  ms/step %.4f  kernel_ms %.4f  fp32 frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))" >> gpurun_out/ab_mono.txt
}
for v in tma0 mono1_tma0 mono1_tma1 mono0_tma1; do
  run_ab $v SKB_ENGINE_LIB=$PWD/skred_b200/variants/$v/libskred_b200.so
done
cat gpurun_out/ab_mono.txt
M=gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
for v in tma0 tma1; do
  SKB_ENGINE_LIB=$PWD/skred_b200/variants/$v/libskred_b200.so timeout 300 ncu --metrics $M --clock-control none -k regex:k_render_free -s 9 -c 2 --csv --log-file gpurun_out/ncu_tbl_$v.csv python tools/class_bench.py 65536 512 "korg(config3)" > gpurun_out/ncu_tbl_$v.log 2>&1
done
tail -3 gpurun_out/ncu_tbl_tma1.log
