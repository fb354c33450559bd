"""BASELINE.json `configs` at their FULL sizes: helpers shared by the GPU tests and their CPU dry run.

The compiled reference is single-threaded and renders ~5e7 voice-samples/s per core, so a full-size
configuration is checked against it the way SURVEY §8d prescribes: the voices of these loads are
independent (no modulation edge), so `workers` independent reference instances (VOICE_MAX = 64 each,
one process per host core) render disjoint voice subsets for the whole duration; their mixes are added
in float64 (the master volume is linear and identical in every instance) and their evolving per-voice
words are compared bit for bit with the product's at checkpoints.  TEST INFRASTRUCTURE: imports oracle/.
"""
import ctypes as C
import os

import numpy as np

from skred_b200 import workloads as W

BLOCK = 512
REF_V = 64          # VOICE_MAX of the reference instances (oracle/_ref/libskred_ref_v64.so)


def select_many(wl, subsets):
    """The sub-workloads of several disjoint voice subsets (each ascending), re-indexed 0..len-1 — what
    W.shard does for a contiguous range — in ONE pass over the setup and event lists.  Valid when no
    modulation edge leaves a subset."""
    where = {}
    for j, sub in enumerate(subsets):
        for i, v in enumerate(sub):
            where[int(v)] = (j, i)
    outs = []
    for sub in subsets:
        o = dict(wl)
        o.update(voices=len(sub), setup=[], timed=[], pcm_tau={})
        outs.append(o)
    for c in wl["setup"]:
        w = where.get(c[1])
        if w is not None:
            outs[w[0]]["setup"].append((c[0], w[1]) + tuple(c[2:]))
    for t, c in wl.get("timed", []):
        w = where.get(c[1])
        if w is not None:
            outs[w[0]]["timed"].append((t, (c[0], w[1]) + tuple(c[2:])))
    for v, tau in wl.get("pcm_tau", {}).items():
        w = where.get(v)
        if w is not None:
            outs[w[0]]["pcm_tau"][w[1]] = tau
    for o in outs:
        o["events"] = W.bucket(o["timed"])
    return outs


def select(wl, voices):
    return select_many(wl, [list(voices)])[0]


def segments(total_frames, n_checkpoints, unit=8192):
    """Split `total_frames` into n segments whose boundaries are multiples of `unit` frames (a whole
    number of product calls and of reference callbacks); the last one takes the ragged tail."""
    per = (total_frames // n_checkpoints) // unit * unit
    if per == 0:
        return [total_frames]
    segs = [per] * (n_checkpoints - 1)
    segs.append(total_frames - per * (n_checkpoints - 1))
    return segs


STATE_KEYS = ("phase", "finished", "sh_hold", "sh_count", "env_active", "env_start", "env_release",
              "sample", "filter_xy", "smoother_gain", "pan_left", "pan_right")


def ref_worker(args):
    """One process: an independent reference instance renders `wl` (<= 64 voices) for sum(segs) frames as
    512-frame callbacks with the events of `wl["events"]` applied before their callback, and
    returns (mix or None, [state after each segment], seconds inside synth())."""
    wl, segs, keep_mix, factory = args[:4]
    ref_v = args[4] if len(args) > 4 else REF_V
    from oracle import oracle as O
    s = getattr(O, factory)(ref_v, run_seq=False)
    n = wl["voices"]
    W.install(s, wl)
    ev = wl["events"]
    keys = sorted(ev)
    total = sum(segs)
    mix = np.zeros((total, 2), dtype=np.float32) if keep_mix else None
    buf = np.zeros((max(segs), 2), dtype=np.float32)
    states, done, ki = [], 0, 0
    for seg in segs:
        seg_end = done + seg
        out = mix[done:seg_end] if keep_mix else buf[:seg]
        pos = 0
        while done < seg_end:
            k = done // BLOCK
            while ki < len(keys) and keys[ki] < k:
                ki += 1
            if ki < len(keys) and keys[ki] == k:
                s.apply(ev[k])
                ki += 1
            # event-free run: up to the next callback that has events, or the end of the segment
            nxt = keys[ki] * BLOCK if ki < len(keys) else seg_end
            run = min(nxt, seg_end) - done
            s.cpu_seconds += s.lib.ref_render(out[pos:pos + run].ctypes.data, run, BLOCK, 0)
            pos += run
            done += run
        # seq() runs right after the callback (skred.c:116-119): the events of the NEXT callback have already
        # fired when anything looks at the state between two callbacks — as they have on the product
        k = done // BLOCK
        if done % BLOCK == 0 and ki < len(keys) and keys[ki] == k:
            s.apply(ev[k])
            ki += 1
        elif done % BLOCK != 0:
            # a ragged last callback of r frames ending at count c: seq(r) fires what is due by c + r (seq.c:171-178),
            # i.e. the head of the NEXT callback's bucket; only valid at the very end of a job (nothing renders after it)
            r = done % BLOCK
            s.apply([c for t, c in sorted(wl.get("timed", []), key=lambda x: x[0])
                     if W.callback_for_time(t) == k + 1 and t <= done + r])
        st = s.state()
        states.append({k_: st[k_][:n].copy() for k_ in STATE_KEYS})
    return mix, states, s.cpu_seconds


def reference_by_subsets(wl, subsets, segs, keep_mix=True, procs=None, factory="RefSkred", ref_v=REF_V):
    """Render every voice subset on its own reference process (an instance compiled for VOICE_MAX = ref_v).
    Returns (float64 sum of the mixes or None, [per checkpoint: {key: array over the concatenated subsets}],
    total CPU seconds)."""
    import multiprocessing as mp
    jobs = [(sub_wl, segs, keep_mix, factory, ref_v) for sub_wl in select_many(wl, subsets)]
    procs = max(1, min(procs or os.cpu_count() or 1, len(jobs)))
    total = None
    states = [dict() for _ in segs]
    parts = [[] for _ in segs]
    cpu = 0.0
    with mp.get_context("spawn").Pool(procs) as pool:
        for mix, sts, sec in pool.imap(ref_worker, jobs, chunksize=1):
            if mix is not None:
                total = mix.astype(np.float64) if total is None else total + mix
            for i, st in enumerate(sts):
                parts[i].append(st)
            cpu += sec
    for i in range(len(segs)):
        states[i] = {k: np.concatenate([p[k] for p in parts[i]]) for k in STATE_KEYS}
    return total, states, cpu


def queue_events(s, timed):
    """Hand the whole timestamped event list to the product's device event queue (skb_shim_queue_events)."""
    dt = np.dtype([("when", "<u8"), ("voice", "<i4"), ("code", "<i4"), ("a0", "<f4"), ("a1", "<f4")])
    timed = sorted(timed, key=lambda x: x[0])
    a = np.zeros(len(timed), dtype=dt)
    if timed:
        codes = W.EVENT_CODES
        cols = [(t, c[1], codes[c[0]], c[2] if len(c) > 2 else 0.0, c[3] if len(c) > 3 else 0.0) for t, c in timed]
        m = np.array(cols, dtype=np.float64)          # sample counts < 2^53: exact in a double
        a["when"], a["voice"], a["code"] = m[:, 0].astype(np.uint64), m[:, 1].astype(np.int32), m[:, 2].astype(np.int32)
        a["a0"], a["a1"] = m[:, 3].astype(np.float32), m[:, 4].astype(np.float32)
    s.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    assert s.lib.skb_shim_queue_events(a.ctypes.data, len(a)) == 0
    return a


def product_render(s, wl, segs, call_frames=8192, keep_mix=True, voices=None):
    """The product side: install, queue every event, then synth() in calls of `call_frames` frames.
    Returns (mix or None, [state after each segment restricted to `voices`], seconds inside synth())."""
    import time
    W.install(s, wl)
    queue_events(s, wl.get("timed", []))
    total = sum(segs)
    mix = np.zeros((total, 2), dtype=np.float32) if keep_mix else None
    buf = np.zeros((call_frames, 2), dtype=np.float32)
    sel = slice(None) if voices is None else np.asarray(voices)
    states, done, sec = [], 0, 0.0
    for seg in segs:
        seg_end = done + seg
        t0 = time.perf_counter()
        while done < seg_end:
            n = min(call_frames, seg_end - done)
            s.lib.synth((mix[done:done + n] if keep_mix else buf[:n]).ctypes.data, None, n, 2, None)
            done += n
        sec += time.perf_counter() - t0
        st = s.state()
        states.append({k: st[k][sel].copy() for k in STATE_KEYS})
    return mix, states, sec


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_checkpoints_equal(ref_states, got_states):
    for i, (a, b) in enumerate(zip(ref_states, got_states)):
        for k in STATE_KEYS:
            same = bits(a[k]) == bits(b[k])
            assert bool(np.all(same)), ("checkpoint %d" % i, k, np.nonzero(~same)[0][:8])


def chunks(voices, n):
    voices = list(voices)
    return [voices[i:i + n] for i in range(0, len(voices), n)]
