#!/usr/bin/env python3
"""Which voice, which frame: the per-voice taps of the reference and the CUDA drop-in for a seed of
tests/test_event_fuzz.py.  python tools/gpu_event_fuzz_tap.py SEED [CALL]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases                                  # noqa: E402
import full_size as FS                        # noqa: E402
import test_event_fuzz as T                   # noqa: E402
from oracle import oracle as O                # noqa: E402
from skred_b200 import workloads as W         # noqa: E402

seed = int(sys.argv[1])
call = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
luts = cases.load_luts()
V = T.V
frames = 5 * 4096 + 700
rng = np.random.RandomState(seed)
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
timed = sorted(wl["timed"] + T.random_events(rng, frames, 2500), key=lambda x: x[0])
ref, dut = O.RefSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
ref.enable_tap(512)
dut.enable_tap(call)
W.install(ref, wl)
W.install(dut, wl)
FS.queue_events(dut, timed)
oa, ta = ref.render_with_tap(frames, block=512, events=W.bucket(timed))
ob, tb = dut.render_with_tap(frames, block=call)
print("mix err", float(np.abs(oa.astype(np.float64) - ob).max()))
d = ta.view(np.uint32) != tb.view(np.uint32)              # [frame][voice][2]
dv = d.any(axis=2)
bad = np.nonzero(dv.any(axis=0))[0]
print("voices whose tap differs:", bad[:40], "count", len(bad))
for v in bad[:8]:
    f0 = int(np.argmax(dv[:, v]))
    n = int(dv[:, v].sum())
    print("voice %d (v%%3=%d): first differing frame %d (callback %d, offset %d), %d frames differ" % (v, v % 3, f0, f0 // 512, f0 % 512, n))
    print("   events:", [(t, W.callback_for_time(t), c) for t, c in timed if c[1] == v])
    print("   ref tap:", ta[f0 - 1:f0 + 3, v, :].tolist())
    print("   gpu tap:", tb[f0 - 1:f0 + 3, v, :].tolist())
