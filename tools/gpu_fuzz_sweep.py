#!/usr/bin/env python3
"""Many seeds of the random-skode parity test against the CUDA drop-in (tests/test_gpu_fuzz.py runs 8 of them):
python tools/gpu_fuzz_sweep.py [first_seed] [count] [dense K]   -> one line per failing seed, then a summary.
`dense K`: every line picks its voices among the first K (long chains of interacting setters and modulation edges on
few voices -> the lock-step bin kernel); the CPU twin tools/cpu_fuzz_sweep.py ran 2,000 such streams in round 1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import test_setter_equivalence as T      # noqa: E402
from oracle import oracle as O           # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dense = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if dense:
    plain = T.rand_line

    def dense_line(rng):
        T.V = dense
        try:
            return plain(rng)
        finally:
            T.V = 64
    T.rand_line = dense_line
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cpu_fuzz_sweep as CF              # noqa: E402  (upstream_ub_armed: streams that reach the reference's out-of-bounds read)
bad = []
for seed in range(first, first + count):
    try:
        T.run_random_wire_streams(seed, O.DropinCuda)
    except AssertionError as e:
        armed = CF.upstream_ub_armed(seed)
        if armed:
            print("seed %d not comparable (upstream UB): %s" % (seed, armed), flush=True)
            continue
        bad.append(seed)
        print("SEED %d: %s" % (seed, str(e)[:900]), flush=True)
print("seeds %d..%d: %d failed %s" % (first, first + count - 1, len(bad), bad))
