#!/usr/bin/env python3
"""Many seeds of the random-skode parity test against the CUDA drop-in (tests/test_gpu_fuzz.py runs 8 of them):
python tools/gpu_fuzz_sweep.py [first_seed] [count]   -> one line per failing seed, then a summary."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import test_setter_equivalence as T      # noqa: E402
from oracle import oracle as O           # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = []
for seed in range(first, first + count):
    try:
        T.run_random_wire_streams(seed, O.DropinCuda)
    except AssertionError as e:
        bad.append(seed)
        print("SEED %d: %s" % (seed, str(e)[:900]), flush=True)
print("seeds %d..%d: %d failed %s" % (first, first + count - 1, len(bad), bad))
