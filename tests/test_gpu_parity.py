"""GPU: the CUDA engine behind the drop-in API versus the oracle.

Bar (north_star): phase / table index / finished latch / event timing BIT-EXACT;
floating-point outputs within 1e-5 of full scale (1.0) per sample.  The only
float difference allowed by design is the ORDER of the cross-voice sum
(DESIGN.md §Mix); everything per-voice is computed with the reference's own
individually rounded IEEE ops.
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from oracle import oracle as O
from tests_util import PATCH_IDS, WAV_PATCH_IDS, load_wav_patch, patch_lines, trace_render, assert_state_equal, FULL_SCALE_TOL

pytestmark = pytest.mark.gpu

# per-voice evolving words that never see a cross-voice sum: bit-exact unless a
# modulator feeds them, in which case they inherit nothing inexact either (the
# modulator is another voice's exact sample) — so exact everywhere.
EXACT = ("phase", "finished", "sh_hold", "sh_count", "env_active", "env_start", "env_release",
         "sample", "filter_xy", "smoother_gain", "pan_left", "pan_right")


def maxdiff(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


@pytest.mark.parametrize("n", PATCH_IDS)
def test_patch_vs_golden(n, golden_patches):
    s = O.DropinCuda(64)
    s.load_lines(patch_lines(golden_patches, n))
    gold = golden_patches["p%d_out" % n]
    out, ph, fin = trace_render(s, gold.shape[0])
    assert np.array_equal(ph.view(np.uint32), golden_patches["p%d_phase" % n].view(np.uint32)), "phase trace"
    assert np.array_equal(fin, golden_patches["p%d_finished" % n]), "finished trace"
    assert maxdiff(out, gold) <= FULL_SCALE_TOL


@pytest.mark.parametrize("name", list(cases.SYNTHETIC))
def test_synthetic_vs_golden(name, golden_synth, luts):
    wl = cases.SYNTHETIC[name](luts)
    s = O.DropinCuda(wl["voices"])
    cases.drive_setup(s, wl)
    out = cases.drive_render(s, wl)
    st = s.state()
    assert np.array_equal(st["phase"].view(np.uint32), golden_synth[name + "_phase"].view(np.uint32))
    assert np.array_equal(st["finished"], golden_synth[name + "_finished"])
    assert maxdiff(out, golden_synth[name + "_out"]) <= FULL_SCALE_TOL


@pytest.mark.parametrize("name", list(cases.SYNTHETIC))
def test_synthetic_vs_port_all_state(name, luts):
    """Every evolving word of every voice, against the CPU restatement."""
    wl = cases.SYNTHETIC[name](luts)
    a, b = O.PortSkred(wl["voices"]), O.DropinCuda(wl["voices"])
    for s in (a, b):
        cases.drive_setup(s, wl)
    oa, ob = cases.drive_render(a, wl), cases.drive_render(b, wl)
    assert maxdiff(oa, ob) <= FULL_SCALE_TOL
    assert_state_equal(a.state(), b.state(), exact_keys=EXACT)


def test_config1_0sk_full_10s(golden_patches):
    """BASELINE configs[0]: 0.sk, 10 s at 44.1 kHz = 861 callbacks of 512 + a 168-frame tail."""
    lines = patch_lines(golden_patches, 0)
    ref = O.RefSkred(64) if O.have_ref(64) else O.PortSkred(64)
    gpu = O.DropinCuda(64)
    ref.load_lines(lines)
    gpu.load_lines(lines)
    a, pa, _ = trace_render(ref, 441000)
    b, pb, _ = trace_render(gpu, 441000)
    assert np.array_equal(pa[:, :2].view(np.uint32), pb[:, :2].view(np.uint32)), "voice_phase[0..1] at every block end"
    assert maxdiff(a, b) <= FULL_SCALE_TOL
    assert abs(float(np.abs(a).max()) - 0.05) < 1e-3


@pytest.mark.parametrize("n", [1, 15, 23, 26, 42, 64, 73])
def test_patch_2s_vs_oracle(n, golden_patches):
    lines = patch_lines(golden_patches, n)
    ref = O.RefSkred(64) if O.have_ref(64) else O.PortSkred(64)
    gpu = O.DropinCuda(64)
    ref.load_lines(lines)
    gpu.load_lines(lines)
    a, b = ref.render(88200), gpu.render(88200)
    assert maxdiff(a, b) <= FULL_SCALE_TOL
    assert_state_equal(ref.state(), gpu.state(), exact_keys=EXACT)


def test_big_callbacks_equal_small_callbacks(luts, monkeypatch):
    """synth(4096) == 8 x synth(512) when no event falls inside (batch mode): identical bits when both
    run the sequential kernel; the time-split launch (SKB_WIDE=1, >= 4 windows) regroups the cross-voice
    sum (filtered rows are mixed by pass C), so there the mix is compared within 1e-7 and the state bits."""
    wl = cases.SYNTHETIC["korg_cz_filter"](luts)
    for wide in ("0", "1"):
        monkeypatch.setenv("SKB_WIDE", wide)
        a, b = O.DropinCuda(64, run_seq=False), O.DropinCuda(64, run_seq=False)
        for s in (a, b):
            cases.drive_setup(s, wl)
        oa = a.render(8192, block=512)
        ob = b.render(8192, block=4096)
        if wide == "0":
            assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
        else:
            assert b.engine_stats().wide_launches == 2
            assert maxdiff(oa, ob) <= 1e-7
        assert_state_equal(a.state(), b.state())


@pytest.mark.parametrize("name", ["lut_adsr", "korg_cz_filter", "pcm_retrigger", "misc"])
def test_specialised_path_equals_generic_path(name, luts, monkeypatch):
    """The pipelined path and the generic per-frame path of k_render_free must leave identical
    bits in every evolving word of every voice (same ops on the same operands, only hoisted).
    The mix is compared within the float budget: the two runs pack different warps, which
    regroups the cross-voice sum."""
    wl = cases.SYNTHETIC[name](luts)
    a = O.DropinCuda(wl["voices"])
    cases.drive_setup(a, wl)
    oa = cases.drive_render(a, wl)
    monkeypatch.setenv("SKB_FORCE_GENERIC", "1")
    b = O.DropinCuda(wl["voices"])
    cases.drive_setup(b, wl)
    ob = cases.drive_render(b, wl)
    assert maxdiff(oa, ob) <= 1e-6
    assert_state_equal(a.state(), b.state())


def test_rendered_voice_frame_count_matches_port(luts):
    """skb_stats.active_voice_frames (the numerator of voice-samples/s: voices the loop does not
    skip, synth.c:531-542) is the same integer on the GPU as in the CPU restatement."""
    for name in ("pcm_retrigger", "misc", "lut_adsr", "mods"):
        wl = cases.SYNTHETIC[name](luts)
        a, b = O.PortSkred(wl["voices"]), O.DropinCuda(wl["voices"])
        for s in (a, b):
            cases.drive_setup(s, wl)
            cases.drive_render(s, wl)
        na, nb = a.engine_stats().active_voice_frames, b.engine_stats().active_voice_frames
        assert na == nb and na > 0, (name, na, nb)


def test_config5_stationary_1024_vs_reference(luts):
    """The bench workload at V = 1,024 (mixed LUT / Korg+CZ+biquad / one-shot PCM started mid-sample
    through voice_phase[] + skb_shim_restore_range), 1 s with its retrigger events, against the
    compiled reference: every evolving word bit-exact, mix within 1e-5."""
    from skred_b200 import workloads as W
    V, frames = 1024, 86 * 512
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
    ref = O.RefSkred(V) if O.have_ref(V) else O.PortSkred(V)
    gpu = O.DropinCuda(V)
    outs = []
    for s in (ref, gpu):
        W.install(s, wl)
        outs.append(s.render(frames, events=wl["events"]))
    assert maxdiff(outs[0], outs[1]) <= FULL_SCALE_TOL
    assert_state_equal(ref.state(), gpu.state(), exact_keys=EXACT)
    fin = gpu.state()["finished"]
    assert 0 < int(fin.sum()) < V          # some one-shots are over, some still play


def test_long_launch_equals_callbacks_with_envelopes_and_one_shots(luts):
    """One synth() call of 8,492 frames (17 envelope windows inside one launch, one-shots ending
    inside it) against 17 callbacks: every evolving word of every voice bit-identical.  The mix is
    compared within the float budget only: a voice that ended is dropped from the warps at the
    next launch, which regroups the (fixed-order, but launch-dependent) cross-voice sum."""
    from skred_b200 import workloads as W
    wl = W.config5(1024, seconds=600.0, luts=luts, event_seconds=0.0, stationary=True)
    a, b = O.DropinCuda(1024, run_seq=False), O.DropinCuda(1024, run_seq=False)
    for s in (a, b):
        W.install(s, wl)
    oa = a.render(8192 + 300, block=512)
    ob = b.render(8192 + 300, block=8192 + 300)
    assert maxdiff(oa, ob) <= 1e-7
    assert_state_equal(a.state(), b.state())


def _queue(s, timed):
    import ctypes as C
    from skred_b200 import workloads as W
    ev = W.to_skb_events(timed)
    s.lib.skb_shim_queue_events.argtypes = [C.c_void_p, C.c_int]
    assert s.lib.skb_shim_queue_events(ev.ctypes.data, len(ev)) == 0
    return ev


@pytest.mark.parametrize("which", ["config4_64_dense", "config5_1024"])
def test_batched_launch_applies_events_in_kernel(which, luts):
    """synth(4096) with timestamped events queued: the engine renders the 8 callbacks in ONE launch
    and applies triggers / envelope on-off at the 512-frame boundaries inside the kernel.  Against
    the CPU restatement fed the same queue one callback at a time: every evolving word of every
    voice bit-identical (event timing included), mix within the float budget."""
    from skred_b200 import workloads as W
    if which == "config4_64_dense":
        V, frames = 64, 6 * 4096
        wl = W.config4(V, seconds=frames / 44100.0, rate_hz=60.0)       # ~45 events per callback on 64 voices
    else:
        V, frames = 1024, 6 * 4096
        wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
        # make it busy: every voice is re-triggered several times inside the test
        wl["timed"] = sorted(wl["timed"] + [(int((0.05 + 0.37 * (v % 7) / 7.0 + 0.11 * k) * 44100),
                                             ("voice_trigger", v) if v % 3 == 2 else ("envelope_velocity", v, float(k % 2)))
                                            for v in range(V) for k in range(5)], key=lambda x: x[0])
    a, b = O.PortSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    for s in (a, b):
        W.install(s, wl)
        _queue(s, wl["timed"])
    oa = a.render(frames, block=512)
    st0 = b.engine_stats()
    ob = b.render(frames, block=4096)
    st1 = b.engine_stats()
    assert maxdiff(oa, ob) <= FULL_SCALE_TOL
    assert_state_equal(a.state(), b.state(), exact_keys=EXACT)
    # it really was batched: far fewer launches than callbacks (3 kernels per launch: ops, render, reduce)
    assert st1.kernel_launches - st0.kernel_launches <= 10 * (frames // 4096)      # unbatched: 24 per call


def test_batched_equals_unbatched(luts, monkeypatch):
    """Same engine with batching off (one launch per callback, events by k_apply_ops): identical state."""
    from skred_b200 import workloads as W
    V, frames = 1024, 4 * 4096
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
    wl["timed"] = sorted(wl["timed"] + [(int((0.03 + 0.29 * (v % 5) / 5.0) * 44100), ("voice_trigger", v) if v % 3 == 2
                                         else ("envelope_velocity", v, 0.0)) for v in range(V)], key=lambda x: x[0])
    a = O.DropinCuda(V, run_seq=False)
    W.install(a, wl)
    _queue(a, wl["timed"])
    oa = a.render(frames, block=4096)
    monkeypatch.setenv("SKB_NO_BATCH", "1")
    b = O.DropinCuda(V, run_seq=False)
    W.install(b, wl)
    _queue(b, wl["timed"])
    ob = b.render(frames, block=4096)
    assert maxdiff(oa, ob) <= 1e-6
    assert_state_equal(a.state(), b.state())


def test_ragged_call_sizes_vs_port(luts):
    """synth() with arbitrary frame counts (1 ... 1,500, not multiples of anything): every
    evolving word equals the CPU restatement's after the same sequence of calls."""
    rng = np.random.RandomState(7)
    sizes = [1, 7, 15, 16, 17, 33, 511, 513, 1000] + [int(x) for x in rng.randint(1, 1500, size=12)]
    for name in ("korg_cz_filter", "pcm_retrigger", "lut_adsr"):
        wl = cases.SYNTHETIC[name](luts)
        a, b = O.PortSkred(wl["voices"], run_seq=False), O.DropinCuda(wl["voices"], run_seq=False)
        for s in (a, b):
            cases.drive_setup(s, wl)
        oa = np.concatenate([a.render(n, block=n) for n in sizes])
        ob = np.concatenate([b.render(n, block=n) for n in sizes])
        assert maxdiff(oa, ob) <= FULL_SCALE_TOL, name
        assert_state_equal(a.state(), b.state(), exact_keys=EXACT)


def test_full_size_65536_vs_reference_shards(luts):
    """BASELINE configs[4] at FULL width: 65,536 voices, the first 1,024 frames (two callbacks with
    their events), against 16 instances of the compiled reference rendering 4,096 voices each,
    their mixes added in float64 (voices are independent in this load).  Per sample <= 1e-5."""
    from skred_b200 import workloads as W
    V, shard, frames = 65536, 4096, 1024
    if not O.have_ref(shard):
        pytest.skip("compiled reference for VOICE_MAX = 4096 not present")
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
    gpu = O.DropinCuda(V, run_seq=False)
    W.install(gpu, wl)
    got = gpu.render(frames, events=wl["events"])
    want = np.zeros((frames, 2), dtype=np.float64)
    for k in range(V // shard):
        sub = W.shard(wl, k * shard, shard)
        ref = O.RefSkred(shard, run_seq=False)
        W.install(ref, sub)
        # the master volume is applied per instance: undo nothing, it is linear and identical in every instance
        want += ref.render(frames, events=sub["events"]).astype(np.float64)
    assert float(np.max(np.abs(got.astype(np.float64) - want))) <= FULL_SCALE_TOL
    assert float(np.abs(want).max()) > 1e-3


def test_run_to_run_deterministic(luts):
    """Two engines, same calls: identical output bits (every sum has a fixed order)."""
    from skred_b200 import workloads as W
    wl = W.config5(1024, seconds=600.0, luts=luts, event_seconds=1.0, stationary=True)
    outs = []
    for _ in range(2):
        s = O.DropinCuda(1024, run_seq=False)
        W.install(s, wl)
        _queue(s, wl["timed"])
        outs.append(s.render(3 * 4096, block=4096))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))


def test_direct_finish_equals_staged_finish(luts, monkeypatch):
    """skb_finish writes the frames straight into mapped host memory and the host polls a flag (k_finish_host); with
    SKB_FINISH_COPY=1 it stages them on the device and copies (k_finish + cudaMemcpyAsync + stream synchronise).  Same
    products, same bits — ragged call sizes and a master-volume change included (a trace that is not flat takes the
    per-frame gain path of k_finish_host)."""
    from skred_b200 import workloads as W
    wl = W.config5(1024, seconds=600.0, luts=luts, event_seconds=1.0, stationary=True)
    outs, stats = [], []
    for staged in (False, True):
        if staged:
            monkeypatch.setenv("SKB_FINISH_COPY", "1")
        s = O.DropinCuda(1024, run_seq=False)
        W.install(s, wl)
        _queue(s, wl["timed"])
        parts = [s.render(4096, block=4096), s.render(1500, block=700)]
        s.lib.volume_set.argtypes = [C.c_float]
        s.lib.volume_set(0.37)                             # the one-pole of the master volume starts moving
        parts += [s.render(2 * 4096, block=4096), s.render(3, block=3)]
        outs.append(np.concatenate(parts))
        stats.append(s.engine_stats().active_voice_frames)
    monkeypatch.delenv("SKB_FINISH_COPY")
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert stats[0] == stats[1] and stats[0] > 0
    assert float(np.abs(outs[0]).max()) > 1e-3


@pytest.mark.parametrize("which", ["config5_events", "korg_filter", "pcm_dense", "lut_release"])
def test_time_split_launch_equals_sequential(which, luts, monkeypatch):
    """With SKB_WIDE=1 a launch of >= 4 windows is rendered TIME-SPLIT (pass A: light advance + per-window snapshots,
    pass B: one CTA per (rows, window), pass C: sequential biquad from the x scratch).  Against the
    same engine with the split off (the default: one thread walks all windows of its voice): every
    evolving word of every voice bit-identical, the mix within the regrouping budget of the sum."""
    from skred_b200 import workloads as W
    frames = 5 * 4096 + 1024 + 300                       # 8-window launches, a 2-window launch, a ragged tail
    timed = None
    if which == "config5_events":
        V = 1024
        wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
        timed = sorted(wl["timed"] + [(int((0.02 + 0.31 * (v % 7) / 7.0 + 0.09 * k) * 44100),
                                       ("voice_trigger", v) if v % 3 == 2 else ("envelope_velocity", v, float(k % 2)))
                                      for v in range(V) for k in range(4)], key=lambda x: x[0])
    elif which == "korg_filter":
        V = 1024
        wl = W.config3(V, seconds=1.0)
    elif which == "pcm_dense":
        V = 1024
        wl = W.config4(V, seconds=frames / 44100.0, rate_hz=40.0)
        timed = [(t, c) for t, c in wl["timed"] if c[0] == "voice_trigger"]     # pure state edits: they ride in the launch
    else:
        V = 1024
        wl = W.config2(V, seconds=1.0, luts=luts)
        timed = [(int(0.1 * 44100) + 97 * v, ("envelope_velocity", v, 0.0)) for v in range(V)] + \
                [(int(0.3 * 44100) + 53 * v, ("envelope_velocity", v, 1.0)) for v in range(0, V, 2)]
    outs, states, stats = [], [], []
    for wide in ("0", "1"):
        monkeypatch.setenv("SKB_WIDE", wide)
        s = O.DropinCuda(V, run_seq=False)
        W.install(s, wl)
        if timed:
            _queue(s, timed)
        outs.append(s.render(frames, block=4096))
        states.append(s.state())
        stats.append(s.engine_stats())
    assert stats[0].wide_launches == 0 and stats[1].wide_launches >= 5
    assert stats[1].wide_errors == 0
    assert maxdiff(outs[0], outs[1]) <= 1e-6
    assert float(np.abs(outs[0]).max()) > 1e-4
    assert_state_equal(states[0], states[1])
    assert stats[0].active_voice_frames == stats[1].active_voice_frames


@pytest.mark.parametrize("which", ["patch0_fm", "patch42", "korg_cz_filter", "pcm_retrigger", "misc"])
def test_voice_tap_matches_reference(which, luts, golden_patches):
    """synth()'s `user` buffer (synth.c:503-511, 533-611; skred.c:120-131 records from it): per frame and
    voice the (left, right) added to the mix, zeros for skipped / disconnected voices.  Each value is one
    voice's own arithmetic, so the tap is compared BIT FOR BIT — pipelined rows, generic rows and
    modulation bins (patch 0 is FM) alike; the second half of the run uses 2,048-frame calls."""
    if which.startswith("patch"):
        V = 64
        ref, gpu = O.RefSkred(V), O.DropinCuda(V)
        n = 0 if which == "patch0_fm" else 42
        for s in (ref, gpu):
            s.load_lines(patch_lines(golden_patches, n))
        drive = None
    else:
        wl = cases.SYNTHETIC[which](luts)
        V = wl["voices"]
        ref, gpu = O.RefSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
        for s in (ref, gpu):
            cases.drive_setup(s, wl)
        drive = wl.get("events")
    ref.enable_tap(512)
    gpu.enable_tap(2048)
    oa, ta = ref.render_with_tap(8 * 512, events=drive)
    ob, tb = gpu.render_with_tap(8 * 512, events=drive)
    assert maxdiff(oa, ob) <= FULL_SCALE_TOL
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    if drive is not None and all(k < 8 for k in drive):
        # (not the patches: their sequencer fires per callback, so the callback size is part of the input, F8)
        oa2, ta2 = ref.render_with_tap(4096)
        ob2, tb2 = gpu.render_with_tap(4096, block=2048)
        assert maxdiff(oa2, ob2) <= FULL_SCALE_TOL
        assert np.array_equal(ta2.view(np.uint32), tb2.view(np.uint32))
    assert float(np.abs(ta).max()) > 0.0


@pytest.mark.parametrize("n", WAV_PATCH_IDS)
def test_wav_patch_vs_golden(n, golden_wav_patches, tmp_path):
    """Shipped patches that play user samples loaded with `:wN,slot` (wire.c:406-441): one-shot
    tables of 10-35 k samples at their own rate, re-triggered by the sequencer, FM / pan-mod / S&H on top.
    Against the reference render: mix within the budget, phase and finished traces bit for bit."""
    s = O.DropinCuda(64)
    load_wav_patch(s, golden_wav_patches, n, tmp_path)
    gold = golden_wav_patches["p%d_out" % n]
    out, ph, fin = trace_render(s, gold.shape[0])
    assert maxdiff(out, gold) <= FULL_SCALE_TOL
    assert np.array_equal(ph.view(np.uint32), golden_wav_patches["p%d_phase" % n].view(np.uint32)), "phase trace"
    assert np.array_equal(fin, golden_wav_patches["p%d_finished" % n]), "finished trace"


def test_voice_tap_1024_voices_batched_vs_port(luts):
    """The tap at 1,024 voices with 4,096- and 8,192-frame calls (8 windows per launch, events applied in-kernel
    at the boundaries, voices skipped / woken inside the launch; the 8,192-frame calls are rendered by two
    launches, the first handed to the GPU early): bit for bit against the CPU restatement fed the same
    timestamped queue one callback at a time."""
    from skred_b200 import workloads as W
    V, frames = 1024, 4096 + 2 * 8192
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
    wl["timed"] = sorted(wl["timed"] + [(int((0.04 + 0.33 * (v % 7) / 7.0) * 44100), ("voice_trigger", v) if v % 3 == 2
                                         else ("envelope_velocity", v, float(v % 2))) for v in range(V)], key=lambda x: x[0])
    a, b = O.PortSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    a.enable_tap(512)
    b.enable_tap(8192)
    for s in (a, b):
        W.install(s, wl)
        _queue(s, wl["timed"])
    oa, ta = a.render_with_tap(frames, block=512)
    ob1, tb1 = b.render_with_tap(4096, block=4096)
    ob2, tb2 = b.render_with_tap(2 * 8192, block=8192)
    ob, tb = np.concatenate([ob1, ob2]), np.concatenate([tb1, tb2])
    assert maxdiff(oa, ob) <= FULL_SCALE_TOL
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    silent = float(np.mean(np.abs(ta).sum(axis=2) == 0.0))         # (frame, voice) entries of skipped voices
    assert float(np.abs(ta).max()) > 0.0 and 0.02 < silent < 0.98
    assert_state_equal(a.state(), b.state(), exact_keys=EXACT)


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_render_n_gpus_vs_reference(world, tmp_path):
    """N > 1 on real GPUs: one process per GPU under torch.distributed.run, every engine renders its voice
    shard (whole modulation groups, two modulated pairs included), the exchange step runs behind the C-ABI
    (skb_comm_init_rank + skb_reduce_mix: ncclReduce, and the rank-ordered gather + sum), rank 0 applies the master
    volume; against the compiled reference <= 1e-5 per sample, all voices owned exactly once, and two runs from scratch
    bit-identical in each mode (tools/gpu_sharded_check.py).  Skipped when the box has fewer GPUs."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, the box has %d" % (world, torch.cuda.device_count()))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out", "sharded_parity_n%d.txt" % world) if os.path.isdir(os.path.join(root, "gpurun_out")) \
        else str(tmp_path / "sharded.txt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29541 + world),
           os.path.join(root, "tools", "gpu_sharded_check.py"), "4096", "24", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:]
    txt = open(out).read()
    assert txt.count("-> OK") == 2 and "FAIL" not in txt, txt


def test_wex_in_place_table_edit_vs_reference():
    """`/wex<slot>` (wave_table_dynamic_expand, wire.c:553-586) rescales a loaded user sample in place while voices play
    it; the CUDA drop-in re-uploads the slot when its fingerprint changes (no host cooperation needed).  Against the
    compiled reference: mix <= 1e-5, every evolving word bit for bit."""
    if not O.have_ref(64):
        pytest.skip("compiled reference not present")
    a, sa = cases.wex_scenario(O.RefSkred(64))
    b, sb = cases.wex_scenario(O.DropinCuda(64))
    assert maxdiff(a, b) <= FULL_SCALE_TOL
    assert_state_equal(sa, sb, exact_keys=EXACT)


@pytest.mark.parametrize("selective", [0, 1])
def test_recording_and_save_wav_vs_reference(selective, tmp_path, monkeypatch):
    """SURVEY 8f N3 on the GPU: `:r` / `<sec` / `*` through the unmodified wire.c over the CUDA drop-in: save_wav's file
    equals the reference's byte for byte, with the full per-voice tap and with the selective read-back (only the recorded
    voices' columns cross PCIe: skb_read_tap_selected)."""
    if not O.have_ref(64):
        pytest.skip("compiled reference not present")
    want = cases.recording_scenario(O.RefSkred(64), tmp_path)
    monkeypatch.setenv("SKB_TAP_SELECTIVE", str(selective))
    got = cases.recording_scenario(O.DropinCuda(64), tmp_path)
    assert want[:4] == b"RIFF" and len(want) == 44 + 2048 * 3 * 2 * 2
    assert got == want


@pytest.mark.parametrize("n", [23, 26, 41, 64])
def test_sequencer_patch_batched_equals_callbacks(n, golden_patches):
    """SURVEY 8f N1 on the GPU: the shipped sequencer patches rendered as 8,192-frame calls with seq() walked ahead of the
    audio (skb_shim_synth_between; the engine batches the 16 callbacks of a call into as few launches as the sequencer's
    parameter changes allow) against the callback loop on a second engine: state bit for bit, mix within the regrouping
    of the cross-voice sum, and fewer kernel launches."""
    lines = patch_lines(golden_patches, n)
    a, b = O.DropinCuda(64), O.DropinCuda(64)
    a.load_lines(lines)
    b.load_lines(lines)
    frames = 2 * 8192 + 700
    oa = a.render(frames)
    la = a.engine_stats().kernel_launches
    ob = b.render_batched(frames, 8192)
    lb = b.engine_stats().kernel_launches
    assert maxdiff(oa, ob) <= 1e-6
    assert_state_equal(a.state(), b.state(), exact_keys=EXACT)
    assert lb < la, (la, lb)


def test_oversized_modulation_components_vs_reference():
    """Modulation components above 1,024 voices: a CYCLIC one of 1,502 voices (feedback pair + 1,500 AM readers:
    k_render_bins_huge) and an ACYCLIC one of 1,601 voices (one LFO pan- and CZ-modulating 1,600 carriers: k_render_levels),
    plus a cyclic one of 100 voices (one CTA: k_render_bins), with events, against the compiled reference: every evolving word
    bit-exact, mix within 1e-5."""
    from skred_b200 import workloads as W
    V = 4096
    if not O.have_ref(V):
        pytest.skip("compiled reference for 4,096 voices not present")
    wl = cases.oversized_components(V)
    ref, gpu = O.RefSkred(V, run_seq=False), O.DropinCuda(V, run_seq=False)
    outs = []
    for s in (ref, gpu):
        W.install(s, wl)
        outs.append(s.render(wl["frames"], events=wl["events"]))
    assert float(np.abs(outs[0]).max()) > 1e-3
    assert maxdiff(outs[0], outs[1]) <= FULL_SCALE_TOL
    assert_state_equal(ref.state(), gpu.state(), exact_keys=EXACT)
    st = gpu.engine_stats()
    assert st.n_group_voices >= 1502 + 1601 + 100


def test_exchange_canary(tmp_path):
    """compute-sanitizer's racecheck is closed on the GPU pool this was developed on (it refuses to start), so the
    shared-memory / global exchange of the frame-lock-step modulation kernels (k_render_bins_warp, k_render_bins,
    k_render_bins_huge) is checked in-kernel instead: the engine built -DSKB_CANARY=1 tags every exchanged voice_sample[]
    with the frame it was written in and counts modulator reads that see another frame than synth.c:526's loop order
    promises.  tools/gpu_canary_check.py renders the `mods` set (with and without the per-voice tap) and the oversized
    components against the reference with that build: parity as usual, canary count 0."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "skred_b200", "variants", "canary", "libskred_b200.so")
    if not os.path.exists(lib):
        pytest.skip("canary build missing: python -m skred_b200.build")
    out = os.path.join(root, "gpurun_out", "exchange_canary.txt") if os.path.isdir(os.path.join(root, "gpurun_out")) else str(tmp_path / "canary.txt")
    env = dict(os.environ, SKB_ENGINE_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_canary_check.py"), out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "canary mismatches 0" in r.stdout and "FAIL" not in r.stdout, r.stdout
