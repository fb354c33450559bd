#!/usr/bin/env python3
"""CPU sweep of the two random-stream tests with many seeds (no GPU): the drop-in shim over the CPU restatement
(oracle/skred_port.c behind the engine's C-ABI) against the compiled reference.  It exercises the SAME host code the
product runs (csrc/synth_shim.c: setters, parameter records, ordered ops, the timestamped event queue) — only the
engine under the C-ABI is the port instead of the CUDA one (tools/gpu_fuzz_sweep.py / gpu_event_fuzz_sweep.py are the
GPU twins).   python tools/cpu_fuzz_sweep.py [first_seed] [n_wire] [n_event]"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
import test_event_fuzz as EF  # noqa: E402
import test_setter_equivalence as SE  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n_wire = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    n_event = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    luts = cases.load_luts()
    bad = []
    t0 = time.time()
    for seed in range(first, first + n_wire):
        try:
            SE.run_random_wire_streams(seed, O.PortSkred)
        except Exception:
            bad.append(("wire", seed, traceback.format_exc(limit=2)))
    t1 = time.time()
    print("random skode streams (every array of synth.def word for word after each callback): seeds %d..%d, %d failed, %.0f s"
          % (first, first + n_wire - 1, sum(1 for b in bad if b[0] == "wire"), t1 - t0), flush=True)
    calls = [512, 1536, 4096, 8192]
    for seed in range(first, first + n_event):
        try:
            EF.run(seed, O.PortSkred, luts, call=calls[seed % len(calls)])
        except Exception:
            bad.append(("event", seed, traceback.format_exc(limit=2)))
    print("random timestamped event streams (1,024 voices, every device event code; calls of 512 ... 8,192 frames): seeds %d..%d, "
          "%d failed, %.0f s" % (first, first + n_event - 1, sum(1 for b in bad if b[0] == "event"), time.time() - t1), flush=True)
    for kind, seed, tb in bad:
        print("FAILED", kind, seed, "\n", tb)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
