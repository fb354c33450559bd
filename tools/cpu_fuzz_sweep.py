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


def upstream_ub_armed(seed):
    """`C<n>` stores any n: cmod_set does not validate (synth.c:646-650), and a sounding CZ voice then reads
    voice_sample[n] (synth.c:263-266) -- past the array for n >= VOICE_MAX, i.e. whatever global the reference's linker
    put there.  The product reads silence instead (csrc/partition.h skb_live_mods).  Replays the seed's lines and says
    whether some voice has such an edge live (mode, depth and amplitude non-zero) when the first difference shows."""
    import numpy as np
    rng = np.random.RandomState(seed)
    ref, dut = O.RefSkred(SE.V, run_seq=False), O.PortSkred(SE.V, run_seq=False)
    for step in range(60):
        for _ in range(rng.randint(1, 12)):
            ln = SE.rand_line(rng)
            ref.wire(ln)
            dut.wire(ln)
        n = int(rng.choice([512, 512, 512, 64, 300]))
        oa, ob = ref.render(n, block=n), dut.render(n, block=n)
        if not np.array_equal(oa.view(np.uint32), ob.view(np.uint32)):
            osc, mode = dut.array("voice_cz_mod_osc", SE.I), dut.array("voice_cz_mode", SE.I)
            depth, amp = dut.array("voice_cz_mod_depth", SE.F), dut.array("voice_amp", SE.F)
            hit = [int(v) for v in range(SE.V) if osc[v] >= SE.V and mode[v] != 0 and depth[v] != 0.0 and amp[v] != 0.0]
            if hit:
                return "step %d: voice(s) %s sound with cz_mode != 0 and cz_mod_osc %s >= VOICE_MAX" % (
                    step, hit, [int(osc[v]) for v in hit])
            return None
    return None


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n_wire = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    n_event = int(sys.argv[3]) if len(sys.argv) > 3 else 30
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
    print("random skode streams (every array of synth.def word for word after each callback): seeds %d..%d, %d failed, %.0f s"
          % (first, first + n_wire - 1, sum(1 for b in bad if b[0] == "wire"), t1 - t0), flush=True)
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
