#!/usr/bin/env python3
"""Replay one seed of the random-skode parity stream (tools/gpu_fuzz_sweep.py) on the CUDA drop-in and, at the first
callback after which a voice differs from the reference, print that voice's words on both sides and what the engine
planned for it.   python tools/gpu_fuzz_diag.py <seed> [dense K]   (env SKB_FORCE_GENERIC=1 / SKB_NO_BATCH=1 to bisect)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import test_setter_equivalence as T      # noqa: E402
from oracle import oracle as O           # noqa: E402

seed = int(sys.argv[1])
dense = int(sys.argv[2]) if len(sys.argv) > 2 else 0
plain = T.rand_line


def line(rng):
    if not dense:
        return plain(rng)
    T.V = dense
    try:
        return plain(rng)
    finally:
        T.V = 64


rng = np.random.RandomState(seed)
ref, dut = O.RefSkred(64, run_seq=False), O.DropinCuda(64, run_seq=False)
KEYS = ["voice_phase", "voice_phase_inc", "voice_sample", "voice_amp", "voice_finished", "voice_cz_mode", "voice_cz_distortion",
        "voice_cz_mod_osc", "voice_cz_mod_depth", "voice_wave_table_index", "voice_table_size", "voice_one_shot",
        "voice_freq_mod_osc", "voice_amp_mod_osc", "voice_pan_mod_osc", "voice_use_amp_envelope", "voice_smoother_gain",
        "voice_filter_mode", "voice_disconnect", "voice_sample_hold_max", "voice_quantize", "voice_direction", "voice_loop_enabled"]
for step in range(60):
    ls = [line(rng) for _ in range(rng.randint(1, 12))]
    for ln in ls:
        ref.wire(ln)
        dut.wire(ln)
    n = int(rng.choice([512, 512, 512, 64, 300]))
    pa, pb = T.snapshot(ref), T.snapshot(dut)
    oa, ob = ref.render(n, block=n), dut.render(n, block=n)
    a, b = T.snapshot(ref), T.snapshot(dut)
    bad = set()
    for k in a:
        same = T.bits(a[k]) == T.bits(b[k])
        if a[k].dtype == np.float32:
            same |= np.isnan(a[k]) & np.isnan(b[k])
        if not np.all(same) and a[k].ndim >= 1 and a[k].shape[0] == 64:
            bad |= set(int(i) for i in np.argwhere(~same)[:, 0])
    if bad:
        print("step %d (%d frames): voices %s differ; lines of this step: %s" % (step, n, sorted(bad), ls))
        st = dut.engine_stats()
        print("engine: free %d group voices %d groups %d launches %d replans %d" %
              (st.n_free_voices, st.n_group_voices, st.n_groups, st.kernel_launches, st.replans))
        for v in sorted(bad)[:4]:
            print(" voice", v)
            for k in KEYS:
                print("   %-26s before ref %r dut %r | after ref %r dut %r" % (k, pa[k][v], pb[k][v], a[k][v], b[k][v]))
            print("   filter after ref %s dut %s" % (a["filter"][v][:4], b["filter"][v][:4]))
            for m in ("voice_cz_mod_osc", "voice_freq_mod_osc", "voice_amp_mod_osc", "voice_pan_mod_osc"):
                mv = int(a[m][v])
                if 0 <= mv < 64:
                    print("   modulator %s = %d: its voice_sample before ref %r dut %r, after ref %r dut %r, amp %r finished %r" %
                          (m, mv, pa["voice_sample"][mv], pb["voice_sample"][mv], a["voice_sample"][mv], b["voice_sample"][mv],
                           a["voice_amp"][mv], a["voice_finished"][mv]))
        print(" mix max|diff| %g" % float(np.nanmax(np.abs(oa.astype(np.float64) - ob))))
        break
else:
    print("seed %d: no difference in 60 steps" % seed)
