#!/usr/bin/env python3
"""0.sk (BASELINE configs[0]: a two-voice FM pair) in calls of [frames] frames, one launch per call (SKB_EARLY_FLUSH=0):
what the launch list of `ncu --metrics gpu__time_duration.sum` is taken from.   python tools/osk_probe.py [frames] [calls]"""
import os
import sys

import numpy as np

os.environ["SKB_EARLY_FLUSH"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skred_b200 import Skred  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sk = Skred(64, private=True, max_frames=8192)
sk.apply([("wave_reset", 0, 100), ("wave_set", 0, 0), ("freq_set", 0, 440.0), ("amp_set", 0, 4.0), ("freq_mod_set", 0, 1, 10.0),
          ("wave_set", 1, 0), ("freq_set", 1, 1.0), ("amp_set", 1, 50.0), ("wave_mute", 1, 1)])
out = np.zeros((F, 2), dtype=np.float32)
ms = []
for _ in range(N):
    sk.lib.synth(out.ctypes.data, None, F, 2, None)
    ms.append(sk.stats().last_render_ms)
print("0.sk, %d-frame calls, one launch each: device ms per call %s -> %.4f ms per 512-frame callback (median)" %
      (F, ["%.3f" % m for m in ms], float(np.median(ms[2:])) * 512 / F))
