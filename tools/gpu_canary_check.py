#!/usr/bin/env python3
"""The modulation-group kernels under the exchange canary (engine built -DSKB_CANARY=1, skred_b200/variants/canary): every
voice_sample[] word exchanged between the voices of a frame-lock-step component carries the frame it was written in, and
every modulator read checks that it sees the frame the reference's loop order promises (synth.c:526: the current frame for
m < n, the previous one for m > n).  Renders the `mods` set (feedback pairs, chains, self references: k_render_bins_warp) and
the oversized components (k_render_bins at 100 voices, k_render_bins_huge at 1,502) with events, against the compiled
reference, and reports the canary count (skb_stats.wide_errors), which must be 0.
  SKB_ENGINE_LIB=skred_b200/variants/canary/libskred_b200.so python tools/gpu_canary_check.py [out.txt]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases                               # noqa: E402
from oracle import oracle as O             # noqa: E402
from skred_b200 import workloads as W      # noqa: E402
from skred_b200.host import load_engine_lib  # noqa: E402
import ctypes as C                         # noqa: E402

eng = load_engine_lib()
eng.skb_backend_name.restype = C.c_char_p
backend = eng.skb_backend_name().decode()
lines, ok = ["engine: %s (%s)" % (backend, os.environ.get("SKB_ENGINE_LIB", "default build"))], backend.endswith("canary")
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))
for name, V, wl, tap in (("mods", 64, cases.SYNTHETIC["mods"](luts), False), ("mods + per-voice tap (levels off: every component frame-lock-step)", 64, cases.SYNTHETIC["mods"](luts), True),
                         ("oversized_components", 4096, cases.oversized_components(4096), False)):
    if not O.have_ref(V):
        lines.append("%s: reference for %d voices not built" % (name, V)); ok = False
        continue
    ref, gpu = O.RefSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    outs = []
    for s in (ref, gpu):
        if tap:
            s.enable_tap(512)
        W.install(s, wl)
        outs.append(s.render_with_tap(wl["frames"], events=wl["events"])[0] if tap else s.render(wl["frames"], events=wl["events"]))
    d = float(np.max(np.abs(outs[0].astype(np.float64) - outs[1])))
    st = gpu.engine_stats()
    same = all(np.array_equal(ref.state()[k].view(np.uint32) if ref.state()[k].dtype == np.float32 else ref.state()[k],
                              gpu.state()[k].view(np.uint32) if gpu.state()[k].dtype == np.float32 else gpu.state()[k]) for k in ("phase", "sample", "finished"))
    good = d <= 1e-5 and same and st.wide_errors == 0
    ok = ok and good
    lines.append("%-70s %5d voices in modulation groups, %d bins, %6d frames: max|diff| vs reference %.3g, phase / sample / finished bit-equal %s, "
                 "canary mismatches %d -> %s" % (name, st.n_group_voices, st.n_groups, wl["frames"], d, same, st.wide_errors, "OK" if good else "FAIL"))
txt = "\n".join(lines) + "\n"
print(txt, end="")
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt)
sys.exit(0 if ok else 1)
