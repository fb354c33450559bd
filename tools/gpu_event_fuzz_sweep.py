#!/usr/bin/env python3
"""Many seeds of tests/test_event_fuzz.py against the CUDA drop-in: python tools/gpu_event_fuzz_sweep.py [first] [count] [neg]
(neg = 1: negative amplitudes in the streams, 37,164 frames per seed)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases                                  # noqa: E402
import test_event_fuzz as T                   # noqa: E402
from oracle import oracle as O                # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
NEG = len(sys.argv) > 3 and int(sys.argv[3]) != 0
luts = cases.load_luts()
bad = []
calls = [4096, 8192, 512, 1536, 2048, 4096]
for seed in range(first, first + count):
    call = calls[seed % len(calls)]
    try:
        if NEG:
            T.run(seed, O.DropinCuda, luts, frames=9 * 4096 + 300, n_events=3000, call=call, neg_amp=True)
        else:
            T.run(seed, O.DropinCuda, luts, call=call)
    except AssertionError as e:
        bad.append((seed, call))
        print("SEED %d call %d: %s" % (seed, call, str(e)[:300]), flush=True)
print("seeds %d..%d: %d failed %s" % (first, first + count - 1, len(bad), bad))
