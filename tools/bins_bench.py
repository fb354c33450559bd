#!/usr/bin/env python3
"""Modulation-group kernels (k_render_bins_warp / k_render_bins) on the GPU: device time per callback of BASELINE
configs[0] (0.sk: a two-voice FM pair) and voice-samples/s of the config-3 sub-variant in which every voice of a pair
CZ-modulates its neighbour, next to the same load without the pairs (free-voice kernel).
  python tools/bins_bench.py [voices] [frames per call]"""
import os
import sys

import numpy as np

os.environ["SKB_EARLY_FLUSH"] = "0"      # one launch per synth() call: last_render_ms then covers the whole call
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skred_b200 import Skred, workloads as W  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F = int(sys.argv[2]) if len(sys.argv) > 2 else 512


def run(name, sk, frames, calls=40):
    out = np.zeros((frames, 2), dtype=np.float32)
    ms = []
    a0 = None
    for k in range(calls):
        if k == calls // 2:
            a0 = sk.stats().active_voice_frames
        sk.lib.synth(out.ctypes.data, None, frames, 2, None)
        ms.append(sk.stats().last_render_ms)
    st = sk.stats()
    act = (st.active_voice_frames - a0) / (calls - calls // 2)
    med = float(np.median(ms[calls // 2:]))
    print("%-44s %d frames per call: device %.4f ms per call (%.4f ms per 512 frames), %.3g rendered voice-samples/s, "
          "%d group voices in %d bins, %d free" % (name, frames, med, med * 512 / frames, act / (med * 1e-3),
                                                    st.n_group_voices, st.n_groups, st.n_free_voices), flush=True)
    print("      segments [staged, generic] = %s; stage clocks us [A, G, C work | A, G, C wait] summed over CTAs and launches: %s" %
          (list(st.class_rows[6:8]), ["%.0f" % (x / 1965.0) for x in st.phase_cycles[:6]]), flush=True)
    sk.lib.synth_free()


# BASELINE configs[0]: 0.sk = "S100; v0 w0 f440 a4 F1,10; v1 w0 f1 a50 m1" (the setter calls wire() makes of it)
for frames in (512, 8192):
    sk = Skred(64, private=True, max_frames=8192)
    sk.apply([("wave_reset", 0, 100), ("wave_set", 0, 0), ("freq_set", 0, 440.0), ("amp_set", 0, 4.0), ("freq_mod_set", 0, 1, 10.0),
              ("wave_set", 1, 0), ("freq_set", 1, 1.0), ("amp_set", 1, 50.0), ("wave_mute", 1, 1)])
    run("0.sk (two-voice FM pair)", sk, frames)

for pairs, label in ((V // 2, "config3 + C-modulation pairs (all voices in bins)"), (0, "config3 without pairs (free-voice kernel)")):
    sk = Skred(V, private=True, max_frames=8192)
    W.install(sk, W.config3(V, seconds=60.0, cmod_pairs=pairs))
    run("%s, %d voices" % (label, V), sk, F)
