#!/usr/bin/env python3
"""Minimal cases around tests/test_event_fuzz.py seed 12: one voice with an envelope, an in-kernel op while it is in
its release segment.  Prints the first frame at which the CUDA drop-in's tap differs from the CPU restatement's."""
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
V = 64
frames = 5 * 4096


def case(name, recipe, vel, op_cb, op, release_cb=15, others=0, call=4096):
    setup = []
    for v in range(1 + others):
        setup += [(x[0], v) + tuple(x[2:]) for x in recipe(v + 508, 1024)]
    timed = [(12 * 512 + 100, ("envelope_velocity", 0, vel))]
    if release_cb:
        timed.append((release_cb * 512 + 100, ("envelope_velocity", 0, 0.0)))
    if op:
        timed.append((op_cb * 512 + 100, op))
    wl = {"voices": V, "tables": {200: (luts["sine_lutable_0"], {}), 201: (luts["triangle_lutable_0"], {}),
                                  202: (luts["impulse_lutable_0"], {})}, "setup": setup, "timed": sorted(timed, key=lambda x: x[0])}
    a, b = O.PortSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    a.enable_tap(512)
    b.enable_tap(call)
    for s in (a, b):
        W.install(s, wl)
        FS.queue_events(s, wl["timed"])
    oa, ta = a.render_with_tap(frames, block=512)
    ob, tb = b.render_with_tap(frames, block=call)
    d = (ta.view(np.uint32) != tb.view(np.uint32)).any(axis=(1, 2))
    f0 = int(np.argmax(d)) if d.any() else -1
    print("%-46s first differing frame %6d (callback %s)   mix err %.3g" % (
        name, f0, f0 // 512 if f0 >= 0 else "-", float(np.abs(oa.astype(np.float64) - ob).max())), flush=True)


pan = ("pan_set", 0, 0.75)
case("korg vel .5 release, pan_set cb29", W.korg_voice, 0.5, 29, pan)
case("korg vel 1  release, pan_set cb29", W.korg_voice, 1.0, 29, pan)
case("korg vel .5 release, trigger cb29", W.korg_voice, 0.5, 29, ("voice_trigger", 0))
case("korg vel .5 sustain, pan_set cb29", W.korg_voice, 0.5, 29, pan, release_cb=0)
case("korg vel .5 release, no op", W.korg_voice, 0.5, 29, None)
case("lut  vel .5 release, pan_set cb29", W.lut_voice, 0.5, 29, pan)
case("korg vel .5 release, pan_set cb26", W.korg_voice, 0.5, 26, pan)
case("korg vel .5 release, pan_set cb17 (same launch)", W.korg_voice, 0.5, 17, pan)
case("korg vel .5 release, pan_set cb29, 8 voices", W.korg_voice, 0.5, 29, pan, others=7)
case("korg vel .5 release, pan_set cb29, 512-frame calls", W.korg_voice, 0.5, 29, pan, call=512)
