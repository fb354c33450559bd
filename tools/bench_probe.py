#!/usr/bin/env python3
"""The bench workload, launch by launch: kernel time, class histogram, active fraction.
python tools/bench_probe.py [voices] [launches] [events 0/1]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skred_b200 import Skred, workloads as W  # noqa: E402
from skred_b200.host import load_engine_lib  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
EV = int(sys.argv[3]) if len(sys.argv) > 3 else 1
BLK = int(sys.argv[4]) if len(sys.argv) > 4 else 512          # frames per synth() call
WORLD = int(sys.argv[5]) if len(sys.argv) > 5 else 1           # render only rank 0's shard of a job cut WORLD ways
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))
sk = Skred(V, max_frames=max(512, BLK), rank=0, world=WORLD)
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=(N * BLK / 44100.0 + 1.0) if EV else 0.0, stationary=True)
W.install(sk, wl)
if EV:
    ev = W.to_skb_events(wl["timed"])
    sk.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    sk.lib.skb_shim_queue_events(ev.ctypes.data, len(ev))
out = np.zeros((BLK, 2), dtype=np.float32)
prev = sk.stats()
base = prev
for k in range(N):
    sk.lib.synth(out.ctypes.data, None, BLK, 2, None)
    st = sk.stats()
    cr = [int(a - b) for a, b in zip(st.class_rows, prev.class_rows)]
    act = (st.active_voice_frames - prev.active_voice_frames) / (V // WORLD * BLK)
    if k < 3 or k % 8 == 0 or k == N - 1:
        print("launch %3d  kernel %.3f ms (A %.3f B %.3f C+reduce %.3f)  active %.3f  rows/class [none none+f pw pw+f pow pow+f mixed generic] = %s" %
              (k, st.last_render_ms, st.last_wide_ms[0], st.last_wide_ms[1], st.last_wide_ms[2], act, cr), flush=True)
    if k == N - 1:
        nb = st.cta_batches - base.cta_batches
        if st.rows_launches:
            print("   k_render_rows us/CTA since launch 16 [A work, G work, C work, A wait, G wait, C wait]:",
                  ["%.1f" % ((a - b) / max(nb, 1) / 1965.0) for a, b in zip(st.phase_cycles[:6], base.phase_cycles[:6])])
        print("   phase us/CTA-pass [compact setup tables prepass render wait rowsum store] since launch 16:",
              ["%.1f" % ((a - b) / max(nb, 1) / 1965.0) for a, b in zip(st.phase_cycles, base.phase_cycles)])
    if k == min(15, N - 3):
        base = st
    prev = st

# per-CTA picture of the last launch
eng = load_engine_lib()
ph = np.zeros((160, 8), dtype=np.uint64)
rows = np.zeros(160 * 64, dtype=np.int32)
cap = C.c_int(0)
eng.skb_debug_cta_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
eng.skb_debug_slot_rank.argtypes = [C.c_void_p, C.c_int]
n = eng.skb_debug_cta_phases(sk.engine, ph.ctypes.data, rows.ctypes.data, 160, C.byref(cap))
us = ph[:n].astype(np.float64) / 1965.0
tot = us.sum(axis=1)
names = "compact setup events+rows prepass render wait rowsum store".split()
print("per-CTA body us: min %.1f  mean %.1f  max %.1f" % (tot.min(), tot.mean(), tot.max()))
for k, nm in enumerate(names):
    print("   %-8s min %6.1f mean %6.1f max %6.1f" % (nm, us[:, k].min(), us[:, k].mean(), us[:, k].max()))
rows = rows[:n * cap.value].reshape(n, cap.value)
for c in list(np.argsort(-tot)[:4]) + list(np.argsort(tot)[:2]):
    rk = [eng.skb_debug_slot_rank(sk.engine, int(r) * 32) if r >= 0 else -1 for r in rows[c]]
    print("   CTA %3d: %.1f us (render %.1f wait %.1f)  row ranks (+8 = one-shot) %s" % (c, tot[c], us[c, 4], us[c, 5], rk))

# per-warp picture of the slowest and fastest CTAs (first batch of the last launch)
wc = np.zeros((160, 14), dtype=np.uint64)
eng.skb_debug_warp_clocks.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
nw = eng.skb_debug_warp_clocks(sk.engine, wc.ctypes.data, 160)
names_c = ["none", "none+f", "pw", "pw+f", "pow", "pow+f", "mixed", "generic"]
for c in list(np.argsort(-tot)[:3]) + list(np.argsort(tot)[:3]):
    row = []
    for w in range(14):
        v = int(wc[c, w])
        cyc, dynb, nl, cl = v & ((1 << 47) - 1), (v >> 47) & 1, (v >> 48) & 0xff, (v >> 56) & 0xff
        row.append("w%d(s%d):%s%s/%d=%.0fus" % (w, w % 4, names_c[cl] if cl < 8 else "-", "*" if dynb else "", nl, cyc / 1965.0))
    print("   CTA %3d %.1f us: %s" % (c, tot[c], "  ".join(row)))
