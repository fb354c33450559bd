#!/usr/bin/env python3
"""SURVEY 8d's second CPU row: the reference compiled with its SHIPPED flags (-O3 -march=native: FMA contraction, not
parity-valid) beside the pinned parity build (-O2 -ffp-contract=off), on the bench's bounded sample (first 4,096 voices
of the config-5 load, 2,048 frames x 3 steps, one core).  Build-container tool: -march=native code must run on the
machine that compiled it, so this row is not part of bench.py; the shipped-flag library goes to a temporary directory."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402

import bench  # noqa: E402
from oracle import oracle as O  # noqa: E402

_spec = importlib.util.spec_from_file_location("build_oracle", os.path.join(ROOT, "oracle", "build_oracle.py"))
B = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(B)

SHARD = 4096


def rate(tag):
    best = 0.0
    for _ in range(2):
        times, active = bench._ref_worker((65536, SHARD, 0, 2048, 3, 3, 30.0))
        best = max(best, sum(active) / sum(times))
    print("%-34s %.3e voice-samples/s on one core" % (tag, best), flush=True)
    return best


def main():
    pinned = rate("gcc " + " ".join(B.PIN_CFLAGS[:2]))
    tmp = tempfile.mkdtemp(prefix="skb_o3_")
    B.REF_OUT = tmp
    B.PIN_CFLAGS = ["-O3", "-march=native", "-fPIC", "-fno-strict-aliasing"]
    amy = os.path.join(B.GEN, "amysamples.o")          # data only: shared with the pinned build
    assert os.path.exists(amy)
    lib = B.build_ref(SHARD, force=True)
    O.ref_lib_path = lambda v: lib
    shipped = rate("gcc -O3 -march=native (shipped)")
    print("shipped / pinned = %.3f" % (shipped / pinned))


if __name__ == "__main__":
    main()
