#!/usr/bin/env python3
"""bench.py — voice-samples/s of the per-voice render path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE configs[4] — 65,536 voices, the mixed
config-2 / config-3 / config-4 recipe by v%3 (LUT + ADSR + pan, Korg + CZ +
resonant biquad + ADSR, AMY one-shot PCM with pitch shift), amplitudes 40/V,
one (re)trigger per voice per 10 s as timestamped events applied at 512-frame
block boundaries.  One STEP = one batch of `--frames` frames
(default 8192 = 16 reference callbacks, the engine's largest launch) of all voices.

  value   device-resident: params/state/tables/pending events live in HBM; K steps of
          skb_shim_render_mix (+ NCCL reduce of the stereo partial mixes for N > 1),
          CUDA events on the launching stream, max over ranks.  N = 1: the L2 is flushed before
          every timed step (256 MB fill outside the per-step events); the K steps back to back
          are config.value_l2_unflushed.  N > 1: not flushed (18 MB of records per GPU and launch,
          each word touched once; N = 1 measures flushed = unflushed within 1 %).
  e2e     the call a skred host makes: synth(buffer, NULL, frames, 2, NULL) with a HOST
          buffer — event firing, parameter/op/trace H2D, kernels, D2H of the block
          inside the timed region (N > 1: render_mix -> reduce -> finish on rank 0).
  cpu_baseline  the compiled reference (oracle/_ref) on ONE host core over a bounded
          sample (first 4096 voices of the same load).

N > 1 (`scaling`: "strong"): BASELINE configs[4] as written — ONE V-voice job, voice-sharded
over the N GPUs (whole modulation groups per GPU, csrc/partition.h), the stereo partial mixes
summed to rank 0 by the engine's exchange step behind the C-ABI (skb_reduce_mix: ncclReduce
over NVLink; value: overlapped with the next render on a second stream; e2e: render -> reduce
-> master volume -> host buffer out of rank 0, in order, every step).  The companion job with
the per-GPU work fixed (every GPU renders V voices of an N*V-voice render) rides in the same
line as `weak_scaling`; BASELINE names no such config, so it is never the headline.

`roofline` is the BINDING bound of the dominant kernel: individually rounded fp32
ops per rendered voice-sample against 148 SMs x 128 lanes x SM clock, duration measured live
with the engine's CUDA events.  The HBM figure the contract template asks for is `roofline_hbm`
(0.3 % of peak: not the bound).  `roofline.traffic` / `roofline_issue` use ncu counters of one
launch of this workload, and only when profiles/r02_ncu_counters.json was captured from the
kernel sources of this very tree (a hash of csrc/ is checked); otherwise they are null.

`--impl reference` times the reference's own synth.c (oracle/_ref, pinned flags)
voice-sharded over ALL host cores (independent processes — the reference is
single-threaded) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from skred_b200 import workloads as W  # noqa: E402

SR = 44100
METRIC = "voice-samples/sec (osc+filter+env+mix)"
UNIT = "voice-samples/s"
WORKLOAD = "config5: %d voices mixed LUT/Korg+CZ+biquad/PCM by v%%3, one retrigger per voice per 10 s, one-shots started in their long-run state"
# algorithmic traffic 276 B per voice per 512-frame block = 0.54 B per voice-sample
# (224 B parameter + state read, 52 B state write); per launch of F frames: 276 B per voice + 8 B per frame.
BYTES_PER_VOICE_LAUNCH = 276.0
# individually rounded fp32 ops per voice-sample (no FMA in parity mode):
# 15 for a plain LUT/PCM voice, 32 for Korg + CZ + biquad; config 5 mixes them 2:1.
FLOPS_PER_VOICE_SAMPLE = (15.0 * 2 + 32.0) / 3.0
FP32_LANES_PER_SM = 128
N_SM = 148


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


_JSON_FD = None


def quiet_stdout():
    """bench.py's stdout carries ONE JSON line: everything else that writes to fd 1 while it runs — the reference's
    and the shim's `# make w0 sine ...` start-up lines (synth.c:1203-1290 prints them), NCCL's version banner when
    NCCL_DEBUG is set on the box — goes to stderr instead.  Spawned worker processes inherit the redirected fd."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def kernel_source_hash():
    """sha256 over the engine's kernel sources: what an ncu capture must have been taken from to describe this build."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "skred_b200", "csrc")
    for f in ("engine.cu", "free_lo.cu", "voice_kernels.cuh", "free_kernel.cuh", "row_kernel.cuh", "level_kernel.cuh", "partition.h"):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_ncu_counters(path=None):
    """Per-launch counters of the dominant kernel (DRAM bytes for roofline.traffic, warp instructions for the
    issue-slot roofline) from one `ncu --set full` capture of THIS workload, written by tools/ncu_counters.py
    together with the hash of the kernel sources it was taken from.  A capture of other sources is STALE: it is
    refused (None) and the reason is reported in the line instead of a number that no longer describes the kernel."""
    p = path or os.path.join(ROOT, "profiles", "r02_ncu_counters.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("src_hash") != kernel_source_hash():
        sys.stderr.write("bench.py: %s was captured from other kernel sources (%s != %s): traffic / issue counters not reported\n"
                         % (os.path.basename(p), d.get("src_hash"), kernel_source_hash()))
        return None
    return d


def load_luts():
    p = os.path.join(ROOT, "tests", "golden", "notamy_luts.npz")
    return dict(np.load(p)) if os.path.exists(p) else None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- #
# reference arm / CPU baseline: the compiled reference on host cores           #
# --------------------------------------------------------------------------- #
def _ref_worker(args):
    """One process = one independent instance of the reference (VOICE_MAX = shard size)
    rendering voices [v0, v0 + shard) of the V-voice load."""
    V, shard, v0, frames_per_step, steps, warmup, event_seconds = args
    from oracle import oracle as O
    wl = W.shard(W.config5(V, seconds=600.0, luts=load_luts(), event_seconds=event_seconds, stationary=True), v0, shard)
    s = O.RefSkred(shard, run_seq=False)
    W.install(s, wl)
    ev = wl["events"]
    out = np.zeros((frames_per_step, 2), dtype=np.float32)
    fin = s.array("voice_finished", C.c_int)
    amp = s.array("voice_amp")

    def alive():
        # voices the loop renders (not skipped by synth.c:531-542)
        return int(np.count_nonzero((fin[:shard] == 0) & (amp[:shard] != 0.0)))

    times, active = [], []
    k = 0
    for step in range(warmup + steps):
        t_step, a_step = 0.0, 0.0
        done = 0
        while done < frames_per_step:
            if k in ev:
                s.apply(ev[k])
            a0 = alive()                       # untimed bookkeeping
            t0 = time.perf_counter()
            s._synth(out[done:done + 512], 512)
            t_step += time.perf_counter() - t0
            a_step += 0.5 * (a0 + alive()) * 512   # a voice that finishes inside the block counts half
            done += 512
            k += 1
        times.append(t_step)
        active.append(a_step)
    return times[warmup:], active[warmup:]


def run_reference_cpu(V, frames_per_step, steps, warmup, procs, shard=4096, voices_limit=None):
    """Returns (voice-samples/s, seconds per step, voices rendered, procs)."""
    import multiprocessing as mp
    nshards = (voices_limit or V) // shard
    jobs = [(V, shard, i * shard, frames_per_step, steps, warmup, 30.0) for i in range(nshards)]
    procs = max(1, min(procs, nshards))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    # shards run `procs` at a time: step time of the job = sum over waves of the slowest shard
    per_step = np.zeros(steps)
    for w0 in range(0, nshards, procs):
        per_step += np.max(np.array([r[0] for r in res[w0:w0 + procs]]), axis=0)
    sec = float(np.mean(per_step))
    voices = nshards * shard
    active = float(np.mean(np.sum(np.array([r[1] for r in res]), axis=0)))     # rendered voice-frames per step
    return active / sec, sec, voices, procs, wall, active / (voices * frames_per_step)


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    V = a.voices
    shard = 4096 if V >= 4096 else V
    from oracle import oracle as O
    if not O.have_ref(shard):
        emit({"impl": "reference", "unavailable": "oracle/_ref/libskred_ref_v%d.so not built" % shard})
        return 0
    cores = os.cpu_count() or 1
    # bounded sample: 4 callbacks per step; with W >= 3 the warm-up (>= 6,144 frames) gets past the attack + decay
    # (4,851 frames) of the envelopes all triggered at t = 0, like the own arm's warm-up does
    frames = 2048 if a.ref_frames is None else a.ref_frames
    vps, sec, voices, procs, wall, frac = run_reference_cpu(V, frames, a.steps, a.warmup, cores, shard)
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the own arm's config at every N: ONE V-voice job (BASELINE configs[4]); each step of this arm is a bounded sample
        # of one of its steps (the reference's cost is linear in frames: a rate measured on 2,048 frames is the rate of 8,192)
        "config": {"workload": WORKLOAD % V, "voices": voices, "frames_per_step": a.frames, "block_frames": 512,
                   "sample_frames_per_step": frames,
                   "sample": "every step renders %d of the workload step's %d frames, all %d voices (ms_per_step is the sample's)"
                             % (frames, a.frames, voices),
                   "active_fraction": frac,
                   "counting": "rendered voice-frames only: voices skipped by synth.c:531-542 (finished one-shots) do not count",
                   "parallelism": "%d independent reference processes x %d voices (reference is single-threaded)" % (procs, shard)},
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": procs, "kind": "reference",
                         "sample": "%d voices x %d frames per step, %d steps; synth.c compiled gcc -O2 -ffp-contract=off" % (voices, frames, a.steps)},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------- #
# own arm                                                                      #
# --------------------------------------------------------------------------- #
def own_arm(a):
    import torch
    import torch.distributed as dist
    from skred_b200 import Skred
    from skred_b200.host import load_engine_lib
    from skred_b200.sharded import comm_init

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, F = a.voices, a.frames
    hbm_peak, sm_mhz, peak_kind = load_peaks()

    sk = Skred(V, device=local, rank=rank, world=world, max_frames=max(F, 512))
    # steps this arm renders: (W + R*K) device-resident, 3 kernel-alone, K with the L2 flushed, W + K end to end, latency blocks
    # repetitions needed if a step took only 0.4 ms (the fastest it gets, 8 GPUs): the event horizon is sized for them
    reps_guess = max(1, min(64, int(np.ceil(a.min_timed_s / max(a.steps * 0.4e-3 * F / 8192.0, 1e-6))))) if a.min_timed_s > 0 else 1
    total_frames = (2 * a.warmup + (2 * reps_guess + 4) * a.steps + 8) * F + (a.latency_blocks + 64) * 512
    wl = W.config5(V, seconds=600.0, luts=load_luts(), event_seconds=total_frames / SR + 1.0, stationary=True)
    W.install(sk, wl)
    ev = W.to_skb_events(wl["timed"])
    sk.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    sk.lib.skb_shim_queue_events(ev.ctypes.data, len(ev))
    sk.flush()
    eng = load_engine_lib()
    st0 = sk.stats()
    owned = st0.n_owned_voices if world > 1 else V

    # a real stream: the legacy default stream's handle is 0, which the engine reads as "use your own",
    # and then neither the CUDA events below nor the exchange step would be ordered with the kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert sp.value
    d_mix = torch.zeros((F, 2), dtype=torch.float32, device="cuda")
    out = np.zeros((F, 2), dtype=np.float32)
    if world > 1:
        # the exchange step behind the C-ABI: skb_comm_init_rank (ncclCommInitRank) + skb_reduce_mix (ncclReduce);
        # torch.distributed ships the 128-byte id and times / synchronises the ranks
        comm_init(sk, dist, eng)

    LF = a.launch_frames                     # frames per launch: events land on 512-frame boundaries
    mix_ptr = d_mix.data_ptr()

    sk.lib.skb_shim_render_calls.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def render_calls(ptr):
        # one callback-sized segment per skb_shim_render_mix, like a host's audio callback; the engine batches
        # the calls of a step into one launch and applies the events at the 512-frame boundaries itself
        r = sk.lib.skb_shim_render_calls(LF, F // LF, ptr, sp)
        if r != 0:
            raise RuntimeError("render failed %d" % r)

    def reduce_mix(ptr, n, stream_ptr):
        r = eng.skb_reduce_mix(sk.engine, ptr, n, stream_ptr)
        if r != 0:
            raise RuntimeError("skb_reduce_mix failed %d: %s" % (r, eng.skb_error_string(sk.engine).decode()))

    def render_step():
        render_calls(mix_ptr)

    # N > 1, device-resident loop: the reduce of step k runs on its own stream while the engine renders
    # step k + 1 into the other mix buffer (a batch render has no reason to serialise them)
    comm = torch.cuda.Stream() if world > 1 else None
    cp = C.c_void_p(comm.cuda_stream) if world > 1 else None
    d_mix2 = [d_mix, torch.zeros_like(d_mix)] if world > 1 else [d_mix]
    reduced = [torch.cuda.Event() for _ in d_mix2]
    step_no = [0]

    def step_device():
        if world == 1:
            render_step()
            return
        k = step_no[0] % 2
        step_no[0] += 1
        stream.wait_event(reduced[k])                     # the reduce that last read this buffer is done
        render_calls(d_mix2[k].data_ptr())
        reduce_mix(d_mix2[k].data_ptr(), F, cp)           # ordered after the render by an event inside the engine
        reduced[k].record(comm)

    def drain_device():
        if world > 1:
            stream.wait_stream(comm)

    def step_e2e(n=F, buf=out):
        if world == 1:
            sk.lib.synth(buf.ctypes.data, None, n, 2, None)
        else:
            r = sk.lib.skb_shim_render_calls(min(LF, n), max(1, n // LF), mix_ptr, sp)
            if r != 0:
                raise RuntimeError("render failed %d" % r)
            reduce_mix(mix_ptr, n, sp)
            if rank == 0:
                sk.finish(mix_ptr, n, buf, sp)
            else:
                sk.lib.skb_shim_discard_gain()
                eng.skb_sync(sk.engine, sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident value ------------------------------------------------
    for _ in range(a.warmup):
        step_device()
        sk.lib.skb_shim_discard_gain()
    drain_device()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    def timed_k_steps():
        """EXACTLY K steps between a barrier + synchronize on both sides; returns (device ms max over ranks, rendered
        voice-frames over all ranks, rank 0's kernel launches)."""
        barrier()
        eng.skb_sync(sk.engine, sp)
        s_b = sk.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            step_device()
            sk.lib.skb_shim_discard_gain()
        drain_device()
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        eng.skb_sync(sk.engine, sp)
        s_a = sk.stats()
        return ms, sum_over_ranks(s_a.active_voice_frames - s_b.active_voice_frames), int(s_a.kernel_launches - s_b.kernel_launches)

    # The K steps are timed R times (each repetition EXACTLY K steps, bracketed as above) until >= --min-timed-s of
    # device time has been measured (K = 20 steps are 14 ms: too short to be robust on its own); the line reports the
    # MEDIAN repetition and the spread of all of them.
    reps = []
    while True:
        reps.append(timed_k_steps())
        spent = sum(r[0] for r in reps) * 1e-3
        if world > 1:
            spent = max_over_ranks(spent)                 # every rank must take the same decision
        if spent >= a.min_timed_s or len(reps) >= reps_guess:
            break
    order = sorted(range(len(reps)), key=lambda i: reps[i][0])
    dev_ms, act_dev, launches = reps[order[len(order) // 2]]
    rep_ms_per_step = [r[0] / a.steps for r in reps]

    # dominant kernel alone (k_render_free + bins + reduce), CUDA events inside the engine
    kern_ms, kern_active = [], []
    for _ in range(15):                                   # (the median of 15 isolated launches: 3 were too few for a stable frac)
        a_b = sk.stats().active_voice_frames
        step_device()
        sk.lib.skb_shim_discard_gain()
        drain_device()
        eng.skb_sync(sk.engine, sp)
        st_k = sk.stats()
        kern_ms.append(st_k.last_render_ms)
        kern_active.append(st_k.active_voice_frames - a_b)                    # per launch (one batched launch per step)
    barrier()
    sk.lib.skb_shim_discard_gain()

    # the same steps with the L2 flushed before each one (a 256 MB fill, outside the per-step CUDA events):
    # the launch's 18 MB of records would otherwise still sit in the 126 MB L2 from the step before
    flushed_ms, flushed_act = None, 0.0
    if world == 1:
        junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        fl = []
        for _ in range(len(reps)):                        # as many repetitions of EXACTLY K steps as the loop above
            pairs = []
            a_f = sk.stats().active_voice_frames
            for _ in range(a.steps):
                junk.fill_(1)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record(stream)
                step_device()
                sk.lib.skb_shim_discard_gain()
                f1.record(stream)
                pairs.append((f0, f1))
            torch.cuda.synchronize()
            eng.skb_sync(sk.engine, sp)
            fl.append((sum(x.elapsed_time(y) for x, y in pairs), float(sk.stats().active_voice_frames - a_f)))
        fl.sort()
        flushed_ms, flushed_act = fl[len(fl) // 2]       # the median repetition
        del junk

    # ---- end to end through synth() ---------------------------------------------
    for _ in range(a.warmup):
        step_e2e()
    barrier()
    eng.skb_sync(sk.engine, sp)
    s_b = sk.stats()
    tm = (C.c_double * 4)()
    sk.lib.skb_shim_timing.argtypes = [C.c_void_p, C.c_int]
    sk.lib.skb_shim_timing(tm, 1)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sk.lib.skb_shim_timing(tm, 0)
    host_ms = [1e3 * x / a.steps for x in tm]
    eng_us = [(x - y) / a.steps for x, y in zip(sk.stats().host_us, s_b.host_us)]
    eng.skb_sync(sk.engine, sp)
    s_a = sk.stats()
    # counted by the engine from the copies it issues (skb_stats.h2d_bytes / d2h_bytes): parameter records, the
    # per-launch staging block (window list, op lists, wake bits), gain trace; stereo block + counters back
    h2d = (s_a.h2d_bytes - s_b.h2d_bytes) / a.steps
    d2h = (s_a.d2h_bytes - s_b.d2h_bytes) / a.steps

    # rendered voice-frames (voices the loop skips — finished one-shots — do not count)
    act_e2e = sum_over_ranks(s_a.active_voice_frames - s_b.active_voice_frames)
    value = act_dev / (dev_ms * 1e-3)
    e2e = act_e2e / e2e_s
    k_ms = float(np.median(kern_ms))
    k_act = float(np.median(kern_active))
    ncu = load_ncu_counters()
    if ncu and (ncu.get("frames") != F or ncu.get("voices_on_gpu") != owned):
        ncu = None                                        # counters of another launch shape
    roof = roofline_objects(k_act, k_ms, owned, F, hbm_peak, sm_mhz, peak_kind, ncu)

    # p50 host-observed latency of ONE 512-frame block through the same path (N > 1: every rank renders its shard of
    # the block, the exchange step, master volume + D2H on rank 0; timed on rank 0 between two host returns)
    lat, lat_parts = None, None
    if not a.no_latency:
        if world == 1:
            tm2 = (C.c_double * 4)()
            sk.lib.skb_shim_timing.argtypes = [C.c_void_p, C.c_int]
            block_latency(sk, 50)
            sk.lib.skb_shim_timing(tm2, 1)
            s_l0 = sk.stats()
            lat = block_latency(sk, a.latency_blocks)
            sk.lib.skb_shim_timing(tm2, 0)
            s_l1 = sk.stats()
            nb = float(a.latency_blocks)
            eu = [(x - y) / nb * 1e-3 for x, y in zip(s_l1.host_us, s_l0.host_us)]
            lat_parts = {"flush_and_traces": tm2[0] / nb * 1e3, "queue_segment": tm2[1] / nb * 1e3, "fire_events": tm2[2] / nb * 1e3,
                         "finish": tm2[3] / nb * 1e3, "of_which_launch_host": eu[0], "finish_enqueue": eu[1], "stream_wait": eu[2],
                         "device_ms_last_block": s_l1.last_render_ms, "note": "means over the blocks; the p50 is of the whole call"}
        else:
            blk = np.zeros((512, 2), dtype=np.float32)
            ts = []
            for _ in range(20):
                step_e2e(512, blk)
            barrier()
            for _ in range(a.latency_blocks):
                t0 = time.perf_counter()
                step_e2e(512, blk)
                ts.append(time.perf_counter() - t0)
            barrier()
            lat = float(np.median(ts) * 1e3)              # rank 0's is reported (the host that gets the buffer)

    weak = None
    if world > 1 and not a.no_weak:
        weak = weak_scaling_leg(a, local, world, stream, sp, (hbm_peak, sm_mhz, peak_kind))
    clk = clocks.stop() if rank == 0 else None           # sampled over every timed region above, the weak leg included

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
            # BASELINE configs[4] as written: ONE V-voice job.  N = 1 renders all of it; N > 1 cuts the same job N ways.
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % V,
                       "voices": V, "voices_per_gpu": owned, "frames_per_step": F, "frames_per_call": LF, "frames_per_launch": F,
                       "block_frames": 512,
                       "active_fraction": act_dev / (V * F * a.steps),
                       "counting": "rendered voice-frames only: voices skipped by synth.c:531-542 (finished one-shots) do not count",
                       "value_counting_all_voice_slots": V * F * a.steps / (dev_ms * 1e-3),
                       "timed": "EXACTLY %d steps between barrier + synchronize, repeated %d times (%.3f s of device time in all); "
                                "value = the median repetition" % (a.steps, len(reps), sum(r[0] for r in reps) * 1e-3),
                       "ms_per_step_repetitions": {"n": len(reps), "min": min(rep_ms_per_step), "median": dev_ms / a.steps,
                                                   "max": max(rep_ms_per_step)},
                       "parallelism": ("ONE %d-voice job voice-sharded x%d (whole modulation groups per GPU), stereo partial mixes summed "
                                       "to rank 0 by skb_reduce_mix = ncclReduce over NVLink behind the C-ABI (value: the reduce of "
                                       "step k overlaps the render of step k + 1 on a second stream; e2e: render -> reduce -> master "
                                       "volume -> host buffer, in order)" % (V, world)) if world > 1 else "1 GPU",
                       "launches": "the engine renders the %d callbacks of a step in one launch; events are applied in-kernel at the 512-frame boundaries" % (F // LF),
                       "l2": "state+params %.1f MB per launch, each word touched once per launch; not flushed between the timed steps "
                             "(the path is issue bound, DRAM < 1 %% of peak) -- value_l2_flushed is the same loop with a 256 MB fill "
                             "before every step, timed per step" % (owned * 276 / 1e6),
                       "value_l2_flushed": (flushed_act / (flushed_ms * 1e-3)) if flushed_ms else None},
            "roofline": roof["roofline"], "roofline_hbm": roof["roofline_hbm"], "roofline_issue": roof["roofline_issue"],
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / a.steps * 1e3,
                    "api": ("synth(buffer, NULL, %d, 2, NULL)" % F) if world == 1 else
                           ("per rank: %d x skb_shim_render_mix(%d) -> skb_reduce_mix -> rank 0: skb_shim_finish into a host buffer" % (F // LF, LF)),
                    "host_ms_per_step": None if world > 1 else {"flush_and_traces": host_ms[0], "queue_segments": host_ms[1],
                                                                "fire_events": host_ms[2], "finish_launch_kernels_d2h_sync": host_ms[3],
                                                                "of_which_launch_host": eng_us[0] * 1e-3, "finish_enqueue": eng_us[1] * 1e-3,
                                                                "stream_wait": eng_us[2] * 1e-3}},
            "gpu_launches": launches + (a.steps if world > 1 else 0),
            "clocks": clk,
            "block_latency_ms_p50": lat, "block_latency_parts_ms": lat_parts,
        }
        if world == 1 and not a.no_latency:
            line["block_latency_ms_p50_64_voices"] = block_latency_small(local, a.latency_blocks)
        if not a.no_cpu:
            line["cpu_baseline"] = cpu_baseline(V)          # rank 0's host, one core, a bounded sample — at every N
        if world == 1 and not a.no_latency:
            try:
                line["modulation_groups"] = modulation_leg(local)
            except Exception as exc:
                line["modulation_groups"] = {"unavailable": repr(exc)[:200]}
        if world == 1 and not a.no_fast:
            try:
                line["fast_mode"] = fast_mode_leg(a)
            except Exception as exc:
                line["fast_mode"] = {"unavailable": repr(exc)[:200]}
        if world == 1:
            try:
                line = l2_flushed_headline(line, flushed_act, flushed_ms, a.steps, V, F)
            except Exception as exc:                     # (the unflushed line is complete on its own)
                sys.stderr.write("bench.py: flushed headline not applied: %r\n" % (exc,))
        emit(weak_companion(line, weak, V, world, F, a.steps))
    if world > 1:
        dist.barrier()
        eng.skb_comm_destroy(sk.engine)
        dist.destroy_process_group()
    return 0


def roofline_objects(k_act, k_ms, owned, F, hbm_peak, sm_mhz, peak_kind, ncu):
    """roofline (binding: fp32 issue) / roofline_hbm / roofline_issue of one launch of the dominant kernel on ONE GPU:
    k_act rendered voice-frames in k_ms (the engine's CUDA events around k_render_free + bins + reduce, measured live),
    `owned` voices on that GPU.  `roofline_fp32` is kept as an alias of `roofline` for readers of round 1's lines."""
    # launch traffic: every owned voice's amp + first state group are read (32 B) to decide the skip;
    # a rendered voice moves its full 276 B record
    alive_per_launch = k_act / F
    algo_bytes = alive_per_launch * BYTES_PER_VOICE_LAUNCH + (owned - alive_per_launch) * 32.0 + F * 8
    ach_gbs = algo_bytes / (k_ms * 1e-3) / 1e9
    fp32_peak = N_SM * FP32_LANES_PER_SM * sm_mhz * 1e6
    # config 5 by v%3: LUT (15 ops) and Korg+CZ+biquad (32 ops) voices always render; the one-shot PCM third (15 ops)
    # renders only while a sample plays
    alive_pcm = max(0.0, alive_per_launch - 2.0 * owned / 3.0)
    flops_vs = (owned / 3.0 * 15.0 + owned / 3.0 * 32.0 + alive_pcm * 15.0) / max(alive_per_launch, 1.0)
    ach_flops = k_act * flops_vs / (k_ms * 1e-3)
    traffic = float(ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) if ncu else None
    issue_peak = N_SM * 4 * sm_mhz * 1e6                     # warp instructions per second: 4 schedulers per SM
    fp32 = {"bound": "fp32-issue", "achieved": ach_flops / 1e12, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
            "frac": ach_flops / fp32_peak, "traffic": traffic, "peak_kind": "%d SMs x %d fp32 lanes x %.0f MHz (%s sm_max_mhz); "
            "no FMA: parity mode rounds every op, 1 op = 1 lane-issue" % (N_SM, FP32_LANES_PER_SM, sm_mhz, peak_kind),
            "flops_per_voice_sample": flops_vs, "flops_per_launch": k_act * flops_vs,
            "kernel": "k_render_free(+k_render_bins,+k_reduce_rows)", "kernel_ms": k_ms,
            "note": "SURVEY 8d: the path is FP32-issue bound (HBM: roofline_hbm); ops counted from synth.c per class, "
                    "duration = the engine's CUDA events around the launch, measured in this run"}
    return {
        "roofline": fp32,
        "roofline_fp32": fp32,
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": traffic, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": algo_bytes,
                         "kernel_ms": k_ms, "note": "not the bound: state is register-resident for a whole launch"},
        "roofline_issue": None if not ncu else {
            "bound": "warp-instruction issue slots (4 schedulers x %d SMs x SM clock)" % N_SM,
            "achieved": ncu["warp_instructions"] / (k_ms * 1e-3) / 1e12, "peak": issue_peak / 1e12,
            "unit": "T warp-instructions/s", "frac": ncu["warp_instructions"] / (k_ms * 1e-3) / issue_peak,
            "warp_instructions_per_launch": ncu["warp_instructions"],
            "note": "instruction count from the ncu capture of this workload on these kernel sources (src_hash %s), duration measured live"
                    % ncu.get("src_hash")},
    }


def l2_flushed_headline(line, flushed_act, flushed_ms, steps, V, F):
    """N = 1: the headline `value` / `ms_per_step` are the K steps timed with the L2 flushed before each one (a 256 MB
    fill, larger than the 126 MB L2, outside the per-step CUDA events); the same K steps back to back without the flush
    stay in the line as config.value_l2_unflushed.  Pure dict work (tested on CPU)."""
    if not flushed_ms or flushed_ms <= 0.0:
        return line
    out = dict(line)
    cfg = dict(line["config"])
    cfg.pop("value_l2_flushed", None)
    cfg["value_l2_unflushed"] = line["value"]
    cfg["ms_per_step_l2_unflushed"] = line["ms_per_step"]
    cfg["l2"] = ("flushed before every timed step: a 256 MB fill (the L2 holds 126 MB) outside the per-step CUDA events; "
                 "state+params are %.1f MB per launch, each word touched once per launch (the path is issue bound, DRAM < 1 %% "
                 "of peak: value_l2_unflushed is the same K steps back to back)" % (V * 276 / 1e6))
    cfg["active_fraction"] = flushed_act / (float(V) * F * steps)
    cfg["value_counting_all_voice_slots"] = float(V) * F * steps / (flushed_ms * 1e-3)
    out["config"] = cfg
    out["value"] = flushed_act / (flushed_ms * 1e-3)
    out["ms_per_step"] = flushed_ms / steps
    return out


def weak_companion(line, weak, V, world, F, steps):
    """N > 1: the headline stays BASELINE configs[4] as written (ONE V-voice job cut N ways, `scaling`: "strong");
    the job with the per-GPU work fixed (every GPU renders a full V-voice job, an N x V-voice render) is attached
    as `weak_scaling`.  Pure dict work (tested on CPU)."""
    if not weak:
        return line
    out = dict(line)
    w = {"value": weak["act"] / (weak["ms"] * 1e-3), "unit": line["unit"], "ms_per_step": weak["ms"] / weak.get("steps", steps),
         "steps": weak.get("steps", steps), "voices_total": V * world, "voices_per_gpu": V,
         "e2e": weak["e2e"]["value"] if weak.get("e2e") else None,
         "e2e_ms_per_step": weak["e2e"].get("ms_per_step") if weak.get("e2e") else None,
         "gpu_launches": weak["launches"],
         "active_fraction": weak["act"] / (float(V) * world * F * weak.get("steps", steps)),
         "note": "not a BASELINE config: %d GPUs x %d voices each (rank r renders voices r*%d ... of a %d-voice render), "
                 "the same exchange step; reported to show the path scales with the voices a box is given" % (world, V, V, V * world)}
    if weak.get("roof"):
        w["roofline"] = weak["roof"]["roofline"]
    out["weak_scaling"] = w
    return out


def weak_scaling_leg(a, local, world, stream, sp, peaks):
    """N > 1 only, the companion of the strong headline: every GPU renders a full V-voice job (rank r holds voices
    r*V ... (r+1)*V - 1 of an N*V-voice render; the V-voice load is periodic in v by construction, so every rank installs
    the same recipe on a private engine that holds all its voices and joins a communicator of its own), partial mixes
    summed by skb_reduce_mix overlapping the next render like the headline loop.  min(K, 50) steps.  Returns {"act":
    rendered voice-frames over all ranks, "ms": device ms (max over ranks), "steps", "launches", "e2e": the same job with
    a host buffer out of rank 0 every step (or None), "roof"}."""
    import torch
    import torch.distributed as dist
    from skred_b200 import Skred
    from skred_b200.host import load_engine_lib
    from skred_b200.sharded import comm_init
    V, F, LF = a.voices, a.frames, a.launch_frames
    K = min(a.steps, 50)
    eng = load_engine_lib()
    sk = Skred(V, device=local, rank=0, world=1, max_frames=max(F, 512), private=True)
    total_frames = (a.warmup + K) * F * 2 + 4 * F
    wl = W.config5(V, seconds=600.0, luts=load_luts(), event_seconds=total_frames / SR + 1.0, stationary=True)
    W.install(sk, wl)
    ev = W.to_skb_events(wl["timed"])
    sk.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    sk.lib.skb_shim_queue_events(ev.ctypes.data, len(ev))
    sk.flush()
    comm_init(sk, dist, eng)
    sk.lib.skb_shim_render_calls.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    comm = torch.cuda.Stream()
    cp = C.c_void_p(comm.cuda_stream)
    mix = [torch.zeros((F, 2), dtype=torch.float32, device="cuda") for _ in range(2)]
    reduced = [torch.cuda.Event() for _ in mix]
    bad = [0]

    def step(i):
        k = i % 2
        stream.wait_event(reduced[k])
        if sk.lib.skb_shim_render_calls(LF, F // LF, mix[k].data_ptr(), sp) != 0:
            bad[0] = 1
        if eng.skb_reduce_mix(sk.engine, mix[k].data_ptr(), F, cp) != 0:
            bad[0] = 1
        reduced[k].record(comm)
        sk.lib.skb_shim_discard_gain()

    for i in range(a.warmup):
        step(i)
    stream.wait_stream(comm)
    dist.barrier()
    torch.cuda.synchronize()
    eng.skb_sync(sk.engine, sp)
    s_0 = sk.stats()
    before, launches_before = s_0.active_voice_frames, s_0.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(a.warmup + i)
    stream.wait_stream(comm)
    e1.record(stream)
    dist.barrier()
    torch.cuda.synchronize()
    eng.skb_sync(sk.engine, sp)
    s_dev = sk.stats()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(s_dev.active_voice_frames - before)], dtype=torch.float64, device="cuda")
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    res = {"act": float(n.item()), "ms": float(t.item()), "steps": K,
           "launches": int(s_dev.kernel_launches - launches_before) + K, "e2e": None, "roof": None}
    # dominant kernel alone on this GPU (the engine's CUDA events), as in the N = 1 line; rank 0's numbers are reported
    k_ms, k_act = [], []
    for i in range(3):
        a_b = sk.stats().active_voice_frames
        step(a.warmup + K + i)
        stream.wait_stream(comm)
        eng.skb_sync(sk.engine, sp)
        st_k = sk.stats()
        k_ms.append(st_k.last_render_ms)
        k_act.append(st_k.active_voice_frames - a_b)
    if min(k_ms) > 0.0:
        ncu = load_ncu_counters()
        if ncu and (ncu.get("frames") != F or ncu.get("voices_on_gpu") != V):
            ncu = None
        res["roof"] = roofline_objects(float(np.mean(k_act)), float(np.mean(k_ms)), V, F, peaks[0], peaks[1], peaks[2], ncu)

    # end to end, host buffer on rank 0 every step (render -> exchange step -> master volume -> D2H, in order), like the
    # N = 1 e2e through synth().  Every rank issues the same collectives whatever happens locally: a local failure
    # only sets a flag that is summed at the end (a raised exception on one rank would leave the others in a reduce).
    rank = dist.get_rank()
    out = np.zeros((F, 2), dtype=np.float32)

    def step_e2e():
        if sk.lib.skb_shim_render_calls(LF, F // LF, mix[0].data_ptr(), sp) != 0:
            bad[0] = 1
        if eng.skb_reduce_mix(sk.engine, mix[0].data_ptr(), F, sp) != 0:
            bad[0] = 1
        try:
            if rank == 0:
                sk.finish(mix[0].data_ptr(), F, out, sp)
            else:
                sk.lib.skb_shim_discard_gain()
                eng.skb_sync(sk.engine, sp)
        except Exception:
            bad[0] = 1

    for _ in range(a.warmup):
        step_e2e()
    dist.barrier()
    torch.cuda.synchronize()
    eng.skb_sync(sk.engine, sp)
    s_b = sk.stats()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    dist.barrier()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    eng.skb_sync(sk.engine, sp)
    s_a = sk.stats()
    v = torch.tensor([e2e_s, float(bad[0])], dtype=torch.float64, device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(s_a.active_voice_frames - s_b.active_voice_frames)], dtype=torch.float64, device="cuda")
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    if float(v[1].item()) == 0.0 and float(v[0].item()) > 0.0:
        res["e2e"] = {"value": float(n.item()) / float(v[0].item()), "unit": UNIT,
                      "h2d_bytes_per_step": (s_a.h2d_bytes - s_b.h2d_bytes) / K,        # rank 0's engine
                      "d2h_bytes_per_step": (s_a.d2h_bytes - s_b.d2h_bytes) / K,
                      "ms_per_step": float(v[0].item()) / K * 1e3,
                      "api": "per rank: %d x skb_shim_render_mix(%d) -> skb_reduce_mix -> rank 0: skb_shim_finish into a host buffer"
                             % (F // LF, LF)}
    dist.barrier()
    eng.skb_comm_destroy(sk.engine)
    return res


def fast_leg_child(a):
    """Child process of fast_mode_leg: the engine library is whatever SKB_ENGINE_LIB names.  K device-resident steps of the
    bench workload timed with CUDA events, then the master-volume-scaled mix of 2 more steps dumped for the comparison."""
    import torch
    from skred_b200 import Skred
    from skred_b200.host import load_engine_lib
    torch.cuda.set_device(0)
    V, F, LF = a.voices, a.frames, a.launch_frames
    eng = load_engine_lib()
    sk = Skred(V, device=0, max_frames=max(F, 512))
    K = min(a.steps, 40)
    wl = W.config5(V, seconds=600.0, luts=load_luts(), event_seconds=(K + 12) * F / SR + 1.0, stationary=True)
    W.install(sk, wl)
    ev = W.to_skb_events(wl["timed"])
    sk.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    sk.lib.skb_shim_queue_events(ev.ctypes.data, len(ev))
    sk.flush()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    d_mix = torch.zeros((F, 2), dtype=torch.float32, device="cuda")
    sk.lib.skb_shim_render_calls.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    out = np.zeros((2 * F, 2), dtype=np.float32)
    for i in range(2):                        # the first two steps, through synth(), are the ones compared
        sk.lib.synth(out[i * F:(i + 1) * F].ctypes.data, None, F, 2, None)
    for _ in range(3):
        sk.lib.skb_shim_render_calls(LF, F // LF, d_mix.data_ptr(), sp)
        sk.lib.skb_shim_discard_gain()
    torch.cuda.synchronize()
    eng.skb_sync(sk.engine, sp)
    a0 = sk.stats().active_voice_frames
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        sk.lib.skb_shim_render_calls(LF, F // LF, d_mix.data_ptr(), sp)
        sk.lib.skb_shim_discard_gain()
    e1.record(stream)
    torch.cuda.synchronize()
    eng.skb_sync(sk.engine, sp)
    ms = e0.elapsed_time(e1)
    act = float(sk.stats().active_voice_frames - a0)
    eng.skb_backend_name.restype = C.c_char_p
    np.save(a.fast_dump, out)
    emit({"backend": eng.skb_backend_name().decode(), "value": act / (ms * 1e-3), "ms_per_step": ms / K, "steps": K,
          "kernel_ms": float(sk.stats().last_render_ms)})
    return 0


def fast_mode_leg(a):
    """N = 1 only, reported under its own key: the NON-PARITY build of the engine (FMA contraction on,
    linearly interpolating oscillator read; skred_b200/fast/libskred_b200.so) on the bench workload, beside the parity
    build measured the same way in a child process of its own, and the distance between the two renders."""
    import tempfile
    from skred_b200 import build as B
    if not os.path.exists(B.FAST_SO):
        return {"unavailable": "skred_b200/fast/libskred_b200.so not built"}
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for name, lib in (("parity", B.ENGINE_SO), ("fast", B.FAST_SO)):
            dump = os.path.join(td, name + ".npy")
            env = dict(os.environ, SKB_ENGINE_LIB=lib)
            for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
                env.pop(k, None)
            cmd = [sys.executable, os.path.abspath(__file__), "--fast-child", "--fast-dump", dump, "--voices", str(a.voices),
                   "--frames", str(a.frames), "--launch-frames", str(a.launch_frames), "--steps", str(a.steps)]
            r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode != 0 or not line:
                return {"unavailable": "%s child failed (rc %d)" % (name, r.returncode)}
            res[name] = json.loads(line[-1])
            res[name]["mix"] = np.load(dump)
    pa, fa = res["parity"], res["fast"]
    d = np.abs(pa["mix"].astype(np.float64) - fa["mix"].astype(np.float64))
    return {"what": "NON-PARITY build: -fmad=true (a*b+c contracted to FMA) and a linearly interpolating table read instead of "
                    "the reference's truncating one (synth.c:261-274); not used by any parity test or parity number",
            "backend": fa["backend"], "value": fa["value"], "unit": UNIT, "ms_per_step": fa["ms_per_step"], "steps": fa["steps"],
            "parity_build_same_measurement": {"backend": pa["backend"], "value": pa["value"], "ms_per_step": pa["ms_per_step"]},
            "speedup_vs_parity_build": fa["value"] / pa["value"],
            "max_abs_diff_vs_parity_render": float(d.max()), "rms_diff_vs_parity_render": float(np.sqrt((d ** 2).mean())),
            "peak_of_parity_render": float(np.abs(pa["mix"]).max()),
            "compared": "the first %d frames of the workload rendered by both builds through synth()" % len(d)}


def modulation_leg(device):
    """N = 1: the modulated-voice kernels (SURVEY 8d "a sub-variant adds C-modulation pairs to exercise K2"; VERDICT r1
    item 5).  (a) BASELINE configs[0], 0.sk (a two-voice FM pair): device ms of one 512-frame callback, next to the
    compiled reference rendering the same callback on one host core; (b) BASELINE configs[2] at 1,024 voices with every
    voice of a pair CZ-modulating its neighbour: rendered voice-samples/s in 4,096-frame calls.  Both loads are DAGs, so
    they run through k_render_levels (level_kernel.cuh)."""
    from skred_b200 import Skred
    out = {}
    ef_old = os.environ.get("SKB_EARLY_FLUSH")
    os.environ["SKB_EARLY_FLUSH"] = "0"       # one launch per synth() call: last_render_ms then covers the whole call
    try:
        return _modulation_leg(device, out)
    finally:
        if ef_old is None:
            os.environ.pop("SKB_EARLY_FLUSH", None)
        else:
            os.environ["SKB_EARLY_FLUSH"] = ef_old


def _modulation_leg(device, out):
    from skred_b200 import Skred
    sk = Skred(64, device=device, private=True, max_frames=8192)
    sk.apply([("wave_reset", 0, 100), ("wave_set", 0, 0), ("freq_set", 0, 440.0), ("amp_set", 0, 4.0), ("freq_mod_set", 0, 1, 10.0),
              ("wave_set", 1, 0), ("freq_set", 1, 1.0), ("amp_set", 1, 50.0), ("wave_mute", 1, 1)])
    buf = np.zeros((512, 2), dtype=np.float32)
    ms, host = [], []
    for k in range(300):
        t0 = time.perf_counter()
        sk.lib.synth(buf.ctypes.data, None, 512, 2, None)
        host.append(time.perf_counter() - t0)
        ms.append(sk.stats().last_render_ms)
    out["configs0_0sk"] = {"device_ms_per_512_frame_callback": float(np.median(ms[100:])),
                           "synth_call_ms_p50": float(np.median(host[100:]) * 1e3),
                           "levelled_voices": int(sk.stats().n_group_voices), "bins": int(sk.stats().n_groups)}
    sk.lib.synth_free()
    try:
        from oracle import oracle as O
        if O.have_ref(64):
            ref = O.RefSkred(64, run_seq=False)
            ref.load_lines(["S100", "v0 w0 f440 a4 F1,10", "v1 w0 f1 a50 m1"])
            ref.render(50 * 512)
            t0 = ref.cpu_seconds
            ref.render(400 * 512)
            out["configs0_0sk"]["reference_cpu_ms_per_callback_1_core"] = (ref.cpu_seconds - t0) / 400 * 1e3
    except Exception as exc:
        out["configs0_0sk"]["reference_cpu_ms_per_callback_1_core"] = None
        sys.stderr.write("bench.py: 0.sk reference timing skipped: %r\n" % (exc,))
    V, F = 1024, 4096
    sk = Skred(V, device=device, private=True, max_frames=F)
    W.install(sk, W.config3(V, seconds=60.0, cmod_pairs=V // 2))
    buf = np.zeros((F, 2), dtype=np.float32)
    ms = []
    for k in range(24):
        if k == 8:
            a0 = sk.stats().active_voice_frames
        sk.lib.synth(buf.ctypes.data, None, F, 2, None)
        ms.append(sk.stats().last_render_ms)
    act = (sk.stats().active_voice_frames - a0) / 16.0
    med = float(np.median(ms[8:]))
    out["config3_cmod_pairs"] = {"voices": V, "frames_per_call": F, "device_ms_per_call": med, "value": act / (med * 1e-3), "unit": UNIT,
                                 "levelled_voices": int(sk.stats().n_group_voices), "bins": int(sk.stats().n_groups)}
    sk.lib.synth_free()
    return out


def block_latency(sk, nblocks):
    """p50 host-observed time of one synth() call of 512 frames (event flush -> kernels -> 4 KiB D2H)."""
    out = np.zeros((512, 2), dtype=np.float32)
    ts = []
    for _ in range(nblocks):
        t0 = time.perf_counter()
        sk.lib.synth(out.ctypes.data, None, 512, 2, None)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts) * 1e3)


def block_latency_small(device, nblocks):
    """The same for skred as shipped: VOICE_MAX = 64 (BASELINE configs[1]: LUT oscillators + ADSR + pan), one
    512-frame callback per call — what the audio thread of a real-time host would see."""
    from skred_b200 import Skred
    from skred_b200.host import shim_lib_path
    if not os.path.exists(shim_lib_path(64)):
        return None
    sk64 = Skred(64, device=device, max_frames=512)
    W.install(sk64, W.config2(64, seconds=60.0, luts=load_luts()))
    out = np.zeros((512, 2), dtype=np.float32)
    for _ in range(50):
        sk64.lib.synth(out.ctypes.data, None, 512, 2, None)
    return block_latency(sk64, nblocks)


def cpu_baseline(V):
    from oracle import oracle as O
    shard = 4096 if V >= 4096 else V
    if not O.have_ref(shard):
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    frames, steps = 2048, 80                              # ~10 s of one core (the timed steps; set-up and warm-up come on top)
    vps, sec, voices, procs, wall, frac = run_reference_cpu(V, frames, steps, 3, 1, shard, voices_limit=shard)
    return {"value": vps, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": "first %d voices of the same load x %d frames x %d steps (%.1f s of CPU in the timed steps, %.1f s with set-up), "
                      "synth.c gcc -O2 -ffp-contract=off" % (voices, frames, steps, sec * steps, wall)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--min-timed-s", type=float, default=0.5,
                    help="repeat the EXACTLY-K-steps measurement until this much device time has been timed (0 = once)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--voices", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=8192, help="frames per step (multiple of 512; <= 8192)")
    ap.add_argument("--launch-frames", type=int, default=512)
    ap.add_argument("--ref-frames", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling companion measurement")
    ap.add_argument("--no-fast", action="store_true", help="N = 1: skip the non-parity fast_mode leg")
    ap.add_argument("--fast-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--fast-dump", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--latency-blocks", type=int, default=1000)
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if not (a.impl == "own" and a.gpus > 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1):
        quiet_stdout()                       # (the torchrun re-launch convenience leaves it to its children)
    if a.fast_child:
        return fast_leg_child(a)
    if a.impl == "reference":
        return reference_arm(a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return own_arm(a)


if __name__ == "__main__":
    sys.exit(main())
