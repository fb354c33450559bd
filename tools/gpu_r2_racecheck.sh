#!/bin/bash
# round 2, one GPU: compute-sanitizer racecheck (shared-memory hazards) over the kernels that exchange data through shared
# memory between threads: k_render_bins (double-buffered voice_sample[]), k_render_bins_warp, k_render_levels / k_render_rows
# (stage buffers), k_render_free (mix tile, envelope rows).  The same selection runs clean without the tool first.
mkdir -p gpurun_out
SEL='synthetic_vs_port_all_state and (mods or lut_adsr or korg) or batched_launch_applies_events_in_kernel and config4 or oversized_modulation'
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/race_plain.log 2>&1 || { echo "plain selection failed"; tail -5 gpurun_out/race_plain.log; exit 1; }
grep -v "^#" gpurun_out/race_plain.log | tail -2
( echo "# compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -m gpu -k '$SEL'";
  timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" 2>&1 | grep -v "^# " ) > gpurun_out/r02_racecheck.txt
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/r02_racecheck.txt | sort | uniq -c | sort -rn | head -20
