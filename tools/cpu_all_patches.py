#!/usr/bin/env python3
"""Every shipped .sk patch of a skred tree (not only the ones with committed fixtures) through the compiled reference
and through the drop-in shim over the CPU restatement: 2 s of audio with the sequencer running, mix compared bit for
bit and every evolving word of every voice.  Build-container tool: reads $SKRED_REF (default /root/reference), whose
working directory is where `:wN,slot` looks for N.wav (wire.c:409).   python tools/cpu_all_patches.py [frames]"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import oracle as O  # noqa: E402
from tests_util import assert_state_equal  # noqa: E402

REF = os.environ.get("SKRED_REF", "/root/reference")


def run_patch(n, frames):
    lines = open(os.path.join(REF, "%d.sk" % n), errors="replace").read().splitlines()
    a, b = O.RefSkred(64), O.PortSkred(64)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        a.load_lines(lines)
        b.load_lines(lines)
    finally:
        os.chdir(cwd)
    oa, ob = a.render(frames), b.render(frames)
    same_mix = np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    try:
        assert_state_equal(a.state(), b.state())
        same_state = True
    except AssertionError as e:
        same_state = str(e)[:120]
    return same_mix, same_state, float(np.abs(oa).max())


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 88200
    ids = sorted(int(os.path.basename(p)[:-3]) for p in glob.glob(os.path.join(REF, "*.sk")))
    bad = 0
    for n in ids:
        mix, state, peak = run_patch(n, frames)
        ok = mix and state is True
        bad += not ok
        print("%4d.sk  mix %s  state %s  peak %.4f" % (n, "bit-equal" if mix else "DIFFERS", "bit-equal" if state is True else state, peak), flush=True)
    print("%d patches, %d differ" % (len(ids), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
