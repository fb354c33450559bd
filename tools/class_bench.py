#!/usr/bin/env python3
"""Per-feature-class kernel time of k_render_free (GPU): python tools/class_bench.py [voices] [frames]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skred_b200 import Skred, workloads as W  # noqa: E402
from skred_b200.host import load_engine_lib  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
F = int(sys.argv[2]) if len(sys.argv) > 2 else 512
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))


def plain(v):
    return [("wave_set", v, 0), ("freq_set", v, W._freq(v)), ("amp_set", v, 40.0 / V), ("pan_set", v, (v % 21 - 10) / 10.0)]


CLASSES = {
    "plain_sine": plain,
    "plain+env": lambda v: plain(v) + [("envelope_set", v, 0.01, 0.1, 0.5, 0.2), ("envelope_velocity", v, 1.0)],
    "plain+filter": lambda v: plain(v) + [("filter_mode", v, 1), ("mmf_set_freq", v, 900.0)],
    "plain+cz1": lambda v: plain(v) + [("cz_set", v, 1, 0.4)],
    "plain+cz(1..7)": lambda v: plain(v) + [("cz_set", v, 1 + v % 7, 0.4)],
    "lut(config2)": lambda v: W.lut_voice(v, V, True),
    "korg(config3)": lambda v: W.korg_voice(v, V),
    "pcm(config4) alive": lambda v: W.pcm_voice(v, V) + [("wave_loop", v, 1), ("voice_trigger", v)],
    "generic(s&h)": lambda v: plain(v) + [("hold", v, 3)],
}

eng = load_engine_lib()
only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
for name, fn in CLASSES.items():
    if only and name not in only:
        continue
    sk = Skred(V, private=True, max_frames=max(F, 512), rank=0, world=int(os.environ.get("SKB_CB_WORLD", "1")))
    from skred_b200.host import install_table
    for i, k in enumerate(("sine_lutable_0", "triangle_lutable_0", "impulse_lutable_0")):
        install_table(sk, 200 + i, luts[k])
    calls = []
    for v in range(V):
        calls += fn(v)
    sk.apply(calls)
    out = np.zeros((F, 2), dtype=np.float32)
    ms = []
    for it in range(12):          # 12 blocks: past attack (441 frames) after ~1, decay (4410) after ~10
        sk.lib.synth(out.ctypes.data, None, F, 2, None)
        ms.append(sk.stats().last_render_ms)
    vs = V // int(os.environ.get("SKB_CB_WORLD", "1")) * F
    st_ = sk.stats()
    print("   phase us/CTA-pass [compact setup tables prepass render wait rowsum store]:",
          ["%.1f" % (x / max(st_.cta_batches, 1) / 1965.0) for x in st_.phase_cycles])
    print("   rows per class (none, none+f, pw, pw+f, pow, pow+f, mixed, generic):", list(sk.stats().class_rows))
    print("%-22s kernel ms: first %.3f  decay %.3f  sustain %.3f   -> %.3g voice-samples/s (sustain)" %
          (name, ms[0], ms[5], ms[-1], vs / (ms[-1] * 1e-3)), flush=True)
    sk.lib.synth_free()
