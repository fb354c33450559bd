#!/usr/bin/env python3
"""One GPU, one voice shard at a time: the engine configured as rank r of `world` against the CPU
restatement configured the same way (no NCCL involved).  python tools/gpu_shard_vs_port.py [V] [blocks] [world]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as OO  # noqa: E402
from skred_b200 import workloads as W  # noqa: E402
from skred_b200.host import load_engine_lib  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 8
WORLD = int(sys.argv[3]) if len(sys.argv) > 3 else 2
luts = dict(np.load(os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")))
wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=NB * 512 / 44100.0 + 1.0, stationary=True)
wl["setup"] += [("freq_mod_set", 0, 3, 2.0), ("amp_mod_set", 6, 9, 0.5)]


def make(path, rank):
    s = OO.HarnessSkred.__new__(OO.HarnessSkred)
    lib = C.CDLL(OO.private_copy(path))
    lib.skb_shim_configure.argtypes = [C.c_int] * 4
    assert lib.skb_shim_configure(0, rank, WORLD, 8192) == 0
    OO.SynthAPI.__init__(s, lib, V)
    s.backend = "port" if "port" in path else "cuda"
    s.run_seq = 0
    s.cpu_seconds = 0.0
    lib.ref_render.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int]
    lib.ref_render.restype = C.c_double
    lib.ref_init()
    return s


load_engine_lib()
for rank in range(WORLD):
    a, b = make(OO.port_lib_path(V), rank), make(OO.cuda_dropin_path(V), rank)
    for s in (a, b):
        W.install(s, wl)
    oa = a.render(NB * 512, events=wl["events"])
    ob = b.render(NB * 512, events=wl["events"])
    per = [float(np.max(np.abs(oa[k * 512:(k + 1) * 512].astype(np.float64) - ob[k * 512:(k + 1) * 512]))) for k in range(NB)]
    sa, sb = a.state(), b.state()
    bad = {k: int(np.count_nonzero(np.asarray(sa[k]).view(np.uint32 if np.asarray(sa[k]).dtype == np.float32 else np.asarray(sa[k]).dtype)
                                   != np.asarray(sb[k]).view(np.uint32 if np.asarray(sb[k]).dtype == np.float32 else np.asarray(sb[k]).dtype)))
           for k in sa}
    print("rank %d/%d: per-block max|diff| %s | words differing per field %s" %
          (rank, WORLD, " ".join("%.2g" % x for x in per), {k: v for k, v in bad.items() if v}), flush=True)
