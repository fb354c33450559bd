"""CPU: pin the oracle.  The restatement (oracle/skred_port.c behind the product's
host shim) must equal the compiled reference (oracle/_ref) BIT FOR BIT, and both
must equal the committed golden fixtures generated from the reference."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle as O
from tests_util import PATCH_IDS, WAV_PATCH_IDS, load_wav_patch, patch_lines, trace_render, assert_state_equal

needs_ref = pytest.mark.skipif(not O.have_ref(64), reason="compiled reference (oracle/_ref) not present")


@pytest.mark.parametrize("n", PATCH_IDS)
def test_port_matches_golden_patch(n, golden_patches):
    s = O.PortSkred(64)
    s.load_lines(patch_lines(golden_patches, n))
    out, ph, fin = trace_render(s, golden_patches["p%d_out" % n].shape[0])
    assert np.array_equal(out.view(np.uint32), golden_patches["p%d_out" % n].view(np.uint32))
    assert np.array_equal(ph.view(np.uint32), golden_patches["p%d_phase" % n].view(np.uint32))
    assert np.array_equal(fin, golden_patches["p%d_finished" % n])


@pytest.mark.parametrize("name", list(cases.SYNTHETIC))
def test_port_matches_golden_synthetic(name, golden_synth, luts):
    wl = cases.SYNTHETIC[name](luts)
    s = O.PortSkred(wl["voices"])
    cases.drive_setup(s, wl)
    out = cases.drive_render(s, wl)
    st = s.state()
    assert np.array_equal(out.view(np.uint32), golden_synth[name + "_out"].view(np.uint32))
    assert np.array_equal(st["phase"].view(np.uint32), golden_synth[name + "_phase"].view(np.uint32))
    assert np.array_equal(st["finished"], golden_synth[name + "_finished"])


@needs_ref
@pytest.mark.parametrize("n", PATCH_IDS)
def test_reference_matches_golden_patch(n, golden_patches):
    s = O.RefSkred(64)
    s.load_lines(patch_lines(golden_patches, n))
    out, ph, fin = trace_render(s, golden_patches["p%d_out" % n].shape[0])
    assert np.array_equal(out.view(np.uint32), golden_patches["p%d_out" % n].view(np.uint32))
    assert np.array_equal(ph.view(np.uint32), golden_patches["p%d_phase" % n].view(np.uint32))


@needs_ref
@pytest.mark.parametrize("n", [0, 15, 26, 42, 64])
def test_port_matches_reference_long(n, golden_patches):
    """2 s of audio incl. sequencer-driven events, every evolving word compared."""
    a, b = O.RefSkred(64), O.PortSkred(64)
    lines = patch_lines(golden_patches, n)
    a.load_lines(lines)
    b.load_lines(lines)
    oa, ob = a.render(88200), b.render(88200)
    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert_state_equal(a.state(), b.state())


@needs_ref
@pytest.mark.parametrize("name", list(cases.SYNTHETIC))
def test_port_matches_reference_synthetic_state(name, luts):
    wl = cases.SYNTHETIC[name](luts)
    a, b = O.RefSkred(64), O.PortSkred(64)
    for s in (a, b):
        cases.drive_setup(s, wl)
    oa, ob = cases.drive_render(a, wl), cases.drive_render(b, wl)
    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert_state_equal(a.state(), b.state())


@pytest.mark.skipif(not O.have_ref(1024), reason="compiled reference (oracle/_ref, V=1024) not present")
def test_port_matches_reference_bench_workload_1024(luts):
    """The bench workload (config 5, one-shots started mid-sample through voice_phase[] and
    skb_shim_restore_range) at V = 1,024: shim + port == reference bit for bit, and the
    rendered-voice-frame counter counts exactly the voices the reference loop does not skip."""
    from skred_b200 import workloads as W
    V, frames = 1024, 12 * 512
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / 44100.0 + 1.0, stationary=True)
    a, b = O.RefSkred(V), O.PortSkred(V)
    for s in (a, b):
        W.install(s, wl)
    oa = np.zeros((frames, 2), np.float32)
    ob = np.zeros((frames, 2), np.float32)
    for k in range(frames // 512):
        for s, o in ((a, oa), (b, ob)):
            if k in wl["events"]:
                s.apply(wl["events"][k])
            s._synth(o[k * 512:(k + 1) * 512], 512)
    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert_state_equal(a.state(), b.state())
    n = b.engine_stats().active_voice_frames
    # two thirds of the voices always render, the one-shot third only while a sample plays
    assert 2 * V // 3 * frames <= n < V * frames


@needs_ref
@pytest.mark.parametrize("n", [0, 15, 42])
def test_port_tap_matches_reference(n, golden_patches):
    """The per-voice tap `user` of synth() (synth.c:533-611): [frame][voice][L,R] before the master
    volume, zeros for skipped / disconnected voices — bit for bit, through the drop-in shim."""
    a, b = O.RefSkred(64), O.PortSkred(64)
    a.enable_tap(512)
    b.enable_tap(512)
    lines = patch_lines(golden_patches, n)
    a.load_lines(lines)
    b.load_lines(lines)
    oa, ta = a.render_with_tap(20 * 512)
    ob, tb = b.render_with_tap(20 * 512)
    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32))
    assert float(np.abs(ta).max()) > 0.0


@pytest.mark.parametrize("n", WAV_PATCH_IDS)
def test_port_matches_golden_wav_patch(n, golden_wav_patches, tmp_path):
    """Shipped patches that load user samples (`:wN,slot`, wire.c:406-441): the table arrives through
    wave_table_data[] / wave_size[] / wave_one_shot[] ... exactly as wire.c leaves them, the shim uploads it."""
    s = O.PortSkred(64)
    load_wav_patch(s, golden_wav_patches, n, tmp_path)
    gold = golden_wav_patches["p%d_out" % n]
    out, ph, fin = trace_render(s, gold.shape[0])
    assert np.array_equal(out.view(np.uint32), gold.view(np.uint32))
    assert np.array_equal(ph.view(np.uint32), golden_wav_patches["p%d_phase" % n].view(np.uint32))
    assert np.array_equal(fin, golden_wav_patches["p%d_finished" % n])
    assert float(np.abs(gold).max()) > 0.0


@needs_ref
@pytest.mark.parametrize("n", WAV_PATCH_IDS)
def test_reference_matches_golden_wav_patch(n, golden_wav_patches, tmp_path):
    s = O.RefSkred(64)
    load_wav_patch(s, golden_wav_patches, n, tmp_path)
    gold = golden_wav_patches["p%d_out" % n]
    out, ph, fin = trace_render(s, gold.shape[0])
    assert np.array_equal(out.view(np.uint32), gold.view(np.uint32))


def _shipped_patch_ids():
    import glob
    import os
    ref = os.environ.get("SKRED_REF", "/root/reference")
    return sorted(int(os.path.basename(p)[:-3]) for p in glob.glob(os.path.join(ref, "*.sk")))


@needs_ref
@pytest.mark.skipif(not _shipped_patch_ids(), reason="no skred tree here (the GPU box): the committed fixtures cover 29 patches")
def test_every_shipped_patch_shim_over_port_equals_reference():
    """All 64 `.sk` patches of the skred tree (not only the 29 with committed fixtures), read where they lie with their
    wav files, 0.5 s with the sequencer running: the drop-in shim over the restatement leaves the same mix and the same
    evolving words as the compiled reference, bit for bit (tools/cpu_all_patches.py 441000 runs the full 10 s of BASELINE
    configs[0] on each and prints the table: profiles/r01_s8_all_patches_cpu.txt)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import cpu_all_patches as AP
    ids = _shipped_patch_ids()
    assert len(ids) >= 60
    sounding = 0
    for n in ids:
        mix, state, peak = AP.run_patch(n, 22050)
        assert mix and state is True, (n, mix, state)
        sounding += peak > 0.0
    assert sounding >= 40


def test_wex_in_place_table_edit_port_vs_reference():
    """`/wex` edits a user sample in place (wire.c:553-586); the drop-in notices the changed floats at the next block
    boundary (synth_shim.c: refresh_edited_tables) and uploads the table again.  Shim over the CPU restatement vs the
    compiled reference: mix and every evolving word equal."""
    import cases
    from tests_util import assert_state_equal, FULL_SCALE_TOL
    if not (O.have_ref(64) and os.path.exists(O.port_lib_path(64))):
        pytest.skip("oracle libraries not built")
    a, sa = cases.wex_scenario(O.RefSkred(64))
    b, sb = cases.wex_scenario(O.PortSkred(64))
    assert float(np.max(np.abs(a.astype(np.float64) - b))) <= FULL_SCALE_TOL
    assert float(np.abs(a[1024:]).max()) > 1e-4 and not np.allclose(a[:1024], a[1024:2048])
    assert_state_equal(sa, sb)


@pytest.mark.parametrize("selective", [0, 1])
def test_recording_and_save_wav_port_vs_reference(selective, tmp_path, monkeypatch):
    """`:r` / `<sec` / `*` through the unmodified wire.c on top of the drop-in.  The WAV file save_wav
    writes must equal the reference's byte for byte — with the full tap, and with the SELECTIVE read-back in which only
    the recorded voices' columns (and the extremes of the others) come back from the engine."""
    import cases
    if not (O.have_ref(64) and os.path.exists(O.port_lib_path(64))):
        pytest.skip("oracle libraries not built")
    want = cases.recording_scenario(O.RefSkred(64), tmp_path)
    monkeypatch.setenv("SKB_TAP_SELECTIVE", str(selective))
    got = cases.recording_scenario(O.PortSkred(64), tmp_path)
    assert len(want) == 44 + 2048 * 3 * 2 * 2 and want[:4] == b"RIFF"
    assert got == want


@pytest.mark.parametrize("n", [23, 26, 41, 64])
def test_sequencer_patch_batched_equals_callbacks_port(n, golden_patches):
    """a patch driven by the pattern sequencer / deferred wire strings (seq.c:164-213) rendered as ONE
    8,192-frame call with seq() as the per-callback hook (skb_shim_synth_between: the timeline is walked ahead of the
    audio) equals the callback loop `synth(512); seq(512);` bit for bit — mix and evolving state.  Shim over the CPU
    restatement; the GPU twin is in tests/test_gpu_parity.py."""
    if not os.path.exists(O.port_lib_path(64)):
        pytest.skip("oracle libraries not built")
    lines = patch_lines(golden_patches, n)
    a, b = O.PortSkred(64), O.PortSkred(64)
    a.load_lines(lines)
    b.load_lines(lines)
    frames = 2 * 8192 + 700
    oa = a.render(frames)                       # 512-frame callbacks, seq() after each
    ob = b.render_batched(frames, 8192)
    assert np.array_equal(oa.view(np.uint32), ob.view(np.uint32))
    assert_state_equal(a.state(), b.state())
    assert n == 64 or float(np.abs(oa).max()) > 0.0       # (64.sk only arms its pattern: it is never started)
