#!/usr/bin/env python3
"""Delta-debug a failing seed of tests/test_event_fuzz.py on the GPU: the smallest subset of the random events whose
mix still differs from the CPU restatement's.  python tools/gpu_event_fuzz_ddmin.py SEED [SECONDS]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases                                  # noqa: E402
import full_size as FS                        # noqa: E402
import test_event_fuzz as T                   # noqa: E402
from oracle import oracle as O                # noqa: E402
from skred_b200 import workloads as W         # noqa: E402

seed = int(sys.argv[1])
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
what = sys.argv[3] if len(sys.argv) > 3 else "mix"          # "mix": the mix differs; "sample": voice_sample differs at the end
call = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
luts = cases.load_luts()
V = T.V
frames = 4 * 4096 if what == "mix" else 5 * 4096 + 700
rng = np.random.RandomState(seed)
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=(5 * 4096 + 700) / 44100.0 + 1.0, stationary=True)
extra = T.random_events(rng, 5 * 4096 + 700, 2500)
extra = [e for e in extra if e[0] <= frames + 512]
ntrial = 0


def fails(ev, base=True):
    global ntrial
    ntrial += 1
    timed = sorted((wl["timed"] if base else []) + ev, key=lambda x: x[0])
    a, b = O.PortSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    for s in (a, b):
        W.install(s, wl)
        FS.queue_events(s, timed)
    oa = a.render(frames, block=512)
    ob = b.render(frames, block=call)
    if what == "mix":
        return float(np.abs(oa.astype(np.float64) - ob).max()) > 1e-5
    sa, sb = a.state()["sample"], b.state()["sample"]
    return bool(np.any(sa.view(np.uint32) != sb.view(np.uint32)))


t0 = time.time()
print("full set fails:", fails(extra), len(extra), flush=True)
base = fails(extra, base=False)
print("without the load's own retriggers fails:", base, flush=True)
use_base = not base
ev, n = list(extra), 2
while len(ev) >= 2 and time.time() - t0 < budget:
    chunk = max(1, len(ev) // n)
    reduced = False
    for i in range(n):
        lo, hi = i * chunk, (len(ev) if i == n - 1 else (i + 1) * chunk)
        comp = ev[:lo] + ev[hi:]
        if comp and fails(comp, use_base):
            ev, n, reduced = comp, max(n - 1, 2), True
            break
        if time.time() - t0 > budget:
            break
    if not reduced:
        if n >= len(ev):
            break
        n = min(len(ev), n * 2)
print("trials", ntrial, "seconds %.0f" % (time.time() - t0), "events left", len(ev), "(with the load's retriggers)" if use_base else "")
for t, c in ev:
    print("  when %6d  callback %2d  v%%3=%d  %s" % (t, W.callback_for_time(t), c[1] % 3, c))
