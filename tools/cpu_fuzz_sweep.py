#!/usr/bin/env python3
"""CPU sweep of the two random-stream tests with many seeds (no GPU): the drop-in shim over the CPU restatement
(oracle/skred_port.c behind the engine's C-ABI) against the compiled reference.  It exercises the SAME host code the
product runs (csrc/synth_shim.c: setters, parameter records, ordered ops, the timestamped event queue) — only the
engine under the C-ABI is the port instead of the CUDA one (tools/gpu_fuzz_sweep.py / gpu_event_fuzz_sweep.py are the
GPU twins).   python tools/cpu_fuzz_sweep.py [first_seed] [n_wire] [n_event] [dense: voices drawn from the first K]"""
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


def upstream_ub_armed(seed):
    """`C<n>` stores any n: cmod_set does not validate (synth.c:646-650), and a rendering CZ voice then reads
    voice_sample[n] (synth.c:263-266) -- past the array for n >= VOICE_MAX, i.e. whatever global the reference's linker
    put there.  The product reads silence instead (csrc/partition.h skb_live_mods).  Replays the seed's lines and
    returns a description of the first callback in which some voice renders with such an edge live (mode, depth and
    amplitude non-zero, not finished), or None if the stream never gets there."""
    import numpy as np
    rng = np.random.RandomState(seed)
    dut = O.PortSkred(SE.V, run_seq=False)
    for step in range(60):
        for _ in range(rng.randint(1, 12)):
            dut.wire(SE.rand_line(rng))
        osc, mode = dut.array("voice_cz_mod_osc", SE.I), dut.array("voice_cz_mode", SE.I)
        depth, amp = dut.array("voice_cz_mod_depth", SE.F), dut.array("voice_amp", SE.F)
        fin = dut.array("voice_finished", SE.I)
        hit = [int(v) for v in range(SE.V)
               if osc[v] >= SE.V and mode[v] != 0 and depth[v] != 0.0 and amp[v] != 0.0 and fin[v] == 0]
        if hit:
            return "from step %d on voice(s) %s render with cz_mode != 0 and cz_mod_osc %s >= VOICE_MAX" % (
                step, hit, [int(osc[v]) for v in hit])
        n = int(rng.choice([512, 512, 512, 64, 300]))
        dut.render(n, block=n)
    return None


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n_wire = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    n_event = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    dense = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    if dense:
        # skode streams whose lines pick their voices (targets of `v`, modulators, links, copies) among the first
        # `dense` voices only: long chains of interacting setters on few voices instead of 64 nearly independent ones
        plain = SE.rand_line

        def dense_line(rng):
            SE.V = dense
            try:
                return plain(rng)
            finally:
                SE.V = 64
        SE.rand_line = dense_line
    luts = cases.load_luts()
    bad = []
    t0 = time.time()
    ub = []
    for seed in range(first, first + n_wire):
        try:
            SE.run_random_wire_streams(seed, O.PortSkred)
        except Exception:
            armed = upstream_ub_armed(seed)
            if armed:
                ub.append((seed, armed))
            else:
                bad.append(("wire", seed, traceback.format_exc(limit=2)))
    t1 = time.time()
    print("random skode streams%s (every array of synth.def word for word after each callback): seeds %d..%d, %d failed, %.0f s"
          % (" on the first %d voices" % dense if dense else "", first, first + n_wire - 1,
             sum(1 for b in bad if b[0] == "wire"), t1 - t0), flush=True)
    for seed, armed in ub:
        print("  seed %d not comparable: upstream undefined behaviour reached -- %s" % (seed, armed), flush=True)
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
