#!/usr/bin/env python3
"""A failing seed of tests/test_event_fuzz.py, looked at AFTER the whole render (where the test compares): which words differ,
their bits, the voice's parameters and events.  python tools/gpu_event_fuzz_final_diff.py SEED [CALL]"""
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
W.install(ref, wl)
W.install(dut, wl)
FS.queue_events(dut, timed)
want = ref.render(frames, block=512, events=W.bucket(timed))
got = dut.render(frames, block=call)
print("seed %d call %d: max mix err %.3g" % (seed, call, float(np.max(np.abs(want.astype(np.float64) - got)))))
a, b = ref.state(), dut.state()
for key in T.EXACT:
    x, y = FS.bits(a[key]), FS.bits(b[key])
    d = x != y
    if d.ndim > 1:
        d = d.any(axis=1)
    for v in np.nonzero(d)[0]:
        v = int(v)
        print("voice %d (v%%3 = %d) %s: ref %r (bits %s)  gpu %r (bits %s)" % (v, v % 3, key, a[key][v], np.atleast_1d(x[v]), b[key][v], np.atleast_1d(y[v])))
        print("   events:", [(t, c) for t, c in timed if c[1] == v])
        print("   ref: " + ", ".join("%s=%r" % (k2, a[k2][v]) for k2 in T.EXACT))
        print("   gpu: " + ", ".join("%s=%r" % (k2, b[k2][v]) for k2 in T.EXACT))
