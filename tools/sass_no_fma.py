#!/usr/bin/env python3
"""Parity guard for the packed fp32 ops (free_kernel.cuh, SKB_F32X2): ptxas contracts mul.rn.f32x2 -> add.rn.f32x2 into
FFMA2 even under --fmad=false, which would round once where the reference rounds twice.  This tool disassembles the built
engine and checks, per function, that (1) no FFMA2 exists at all and (2) the number of scalar FFMA (they come from the IEEE
division / sqrt sequences of -prec-div=true and are the same ones the scalar build has) equals the count in the all-scalar build.

  python tools/sass_no_fma.py [--all]    # the default build (mix pairs packed) — or, --all, skred_b200/variants/x2
                                         # (-DSKB_F32X2=1: every pair) — against skred_b200/variants/scalar
                                         # (-DSKB_F32X2_MIX=0); exits 1 on a difference
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skred_b200 import build as B  # noqa: E402


def counts(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    out = {}
    for b in re.split(r"\n\s+Function : ", txt)[1:]:
        name = b.split("\n", 1)[0].strip()
        c = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", b, re.M):
            c[m.group(1)] += 1
        out[name] = c
    return out


def main():
    full = "--all" in sys.argv      # check the build with EVERY pair packed (-DSKB_F32X2=1) instead of the default one
    eng = B.build_engine_variant("x2", ["-DSKB_F32X2=1"]) if full else B.build_engine()
    ref = B.build_engine_variant("scalar", ["-DSKB_F32X2_MIX=0"])          # no packed op anywhere
    a, b = counts(eng), counts(ref)
    bad = 0
    for fn in sorted(a):
        f2, f1, f1ref = a[fn]["FFMA2"], a[fn]["FFMA"], b.get(fn, {}).get("FFMA", 0)
        packed = a[fn]["FADD2"] + a[fn]["FMUL2"]
        flag = "" if (f2 == 0 and f1 == f1ref) else "   <-- CONTRACTION"
        bad += bool(flag)
        if packed or flag:
            print("%-60s FADD2+FMUL2 %5d  FFMA2 %d  FFMA %4d (scalar build %4d)%s" % (fn[:60], packed, f2, f1, f1ref, flag))
    print("OK: no packed product was contracted" if not bad else "FAILED: %d function(s)" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
