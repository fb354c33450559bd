"""Static guard on the built parity engine (no GPU needed: cuobjdump reads the sm_100a code in the shared library).

The packed fp32 ops of the mix (free_kernel.cuh, SKB_F32X2_MIX: add / mul .f32x2 -> FADD2 / FMUL2) are only parity-safe while
ptxas does not contract a packed product into a packed add: it does exactly that for mul.rn.f32x2 -> add.rn.f32x2, even under
--fmad=false (profiles/r02_ab_f32x2.txt), rounding once where the reference (gcc -ffp-contract=off) rounds twice.  The
sources never feed a packed product to a packed add; this test checks the RESULT: no FFMA2 anywhere in the parity build, and
the packed ops really are there (the build is the one the measurements describe)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENGINE = os.path.join(ROOT, "skred_b200", "libskred_b200.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"), reason="no cuobjdump")
def test_parity_engine_has_no_contracted_packed_product():
    assert os.path.exists(ENGINE), "engine not built: python -m skred_b200.build"
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    txt = subprocess.run([exe, "-sass", ENGINE], capture_output=True, text=True, check=True).stdout
    per = {}
    for b in re.split(r"\n\s+Function : ", txt)[1:]:
        name = b.split("\n", 1)[0].strip()
        c = collections.Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", b, re.M))
        per[name] = c
    assert per, "no device code found in the engine"
    bad = {n: c["FFMA2"] for n, c in per.items() if c["FFMA2"]}
    assert not bad, "packed multiply-add (FFMA2) in the parity build: %r" % bad
    free = [c for n, c in per.items() if "k_render_free" in n and "tap" not in n and "skb_lo" not in n]
    assert free and all(c["FADD2"] > 0 and c["FMUL2"] > 0 for c in free), "k_render_free has no packed mix ops: not the measured build"
