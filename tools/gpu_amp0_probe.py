#!/usr/bin/env python3
"""voice_sample of a voice silenced by amp_set(v, 0) (synth.c:537-542 zeroes it every frame): CUDA drop-in vs port."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases                                  # noqa: E402
import full_size as FS                        # noqa: E402
from oracle import oracle as O                # noqa: E402
from skred_b200 import workloads as W         # noqa: E402

luts = cases.load_luts()
V = 1024
frames = 5 * 4096 + 700


def case(name, voice, cb, call=4096, V=V, make=None):
    wl = make() if make else W.config5(V, seconds=600.0, luts=luts, event_seconds=0.001, stationary=True)
    timed = [(cb * 512 + 100, ("amp_set", voice, 0.0))]
    a, b = O.PortSkred(wl["voices"], run_seq=False), O.DropinCuda(wl["voices"], run_seq=False)
    for s in (a, b):
        W.install(s, wl)
        FS.queue_events(s, timed)
    out = []
    done = 0
    while done < frames:
        n = min(call, frames - done)
        a.render(n, block=512)
        b.render(n, block=n)
        sa, sb = a.state(), b.state()
        out.append("%d:%g/%g" % (done // 512, sa["sample"][voice], sb["sample"][voice]))
        done += n
    st = b.engine_stats()
    bad = np.nonzero(sa["sample"].view(np.uint32) != sb["sample"].view(np.uint32))[0]
    print("%-40s port/gpu sample after each call: %s   replans %d  differing voices %s" % (name, " ".join(out), st.replans, bad[:8]), flush=True)


case("v150 cb35 call4096", 150, 35)
case("v150 cb33 call4096", 150, 33)
case("v150 cb32 call4096", 150, 32)
case("v150 cb35 call512", 150, 35, call=512)
case("v153 cb35 call4096", 153, 35)
case("v0 cb35 call4096", 0, 35)
case("v151 (korg) cb35 call4096", 151, 35)
case("config2 V=64 v2 cb35", 2, 35, make=lambda: W.config2(64, seconds=0.5, luts=luts))
