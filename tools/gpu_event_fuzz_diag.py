#!/usr/bin/env python3
"""Diagnose a failing seed of tests/test_event_fuzz.py on the GPU: python tools/gpu_event_fuzz_diag.py SEED [CALL]"""
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


def go(tag):
    rng = np.random.RandomState(seed)
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
    timed = sorted(wl["timed"] + T.random_events(rng, frames, 2500), key=lambda x: x[0])
    ref, dut = O.RefSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    W.install(ref, wl)
    W.install(dut, wl)
    FS.queue_events(dut, timed)
    ev = W.bucket(timed)
    done, k = 0, 0
    first = True
    while done < frames:
        n = min(call, frames - done)
        want = np.zeros((n, 2), dtype=np.float32)
        d2 = 0
        while d2 < n:                      # reference: 512-frame callbacks with their events
            m = min(512, n - d2)
            if k in ev:
                ref.apply(ev[k])
            ref._synth(want[d2:d2 + m], m)
            d2 += m
            k += 1
        got = np.zeros((n, 2), dtype=np.float32)
        dut._synth(got, n)
        err = np.abs(want.astype(np.float64) - got).max(axis=1)
        a, b = ref.state(), dut.state()
        badv = set()
        for key in T.EXACT:
            x, y = FS.bits(a[key]), FS.bits(b[key])
            d = x != y
            if d.ndim > 1:
                d = d.any(axis=1)
            for v in np.nonzero(d)[0]:
                badv.add((int(v), key))
        print("[%s] frames %6d..%6d  max mix err %.3g at frame %d   state diffs: %s" % (
            tag, done, done + n, err.max(), done + int(np.argmax(err)), sorted(badv)[:12]))
        if (badv or err.max() > 1e-5) and first:
            first = False
            ff = done + int(np.argmax(err > 1e-6)) if err.max() > 1e-6 else done
            print("   first frame with err > 1e-6:", ff, "callback", ff // 512)
            vs = sorted({v for v, _ in badv})[:6]
            for v in vs:
                print("   voice %d (v%%3 = %d) events:" % (v, v % 3),
                      [(t, c) for t, c in timed if c[1] == v and t <= done + n + 512])
                print("      ref: " + ", ".join("%s=%r" % (key, a[key][v]) for key in T.EXACT))
                print("      gpu: " + ", ".join("%s=%r" % (key, b[key][v]) for key in T.EXACT))
            cbs = range(max(1, ff // 512 - 1), ff // 512 + 1)
            for cb in cbs:
                print("   events fired before callback %d:" % cb, ev.get(cb, [])[:40])
        done += n


go("batched")
os.environ["SKB_NO_BATCH"] = "1"
go("SKB_NO_BATCH")
