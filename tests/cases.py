"""Per-feature synthetic voice sets (V = 64) shared by the golden generator, the
oracle tests and the GPU parity tests.  Every case is a workload dict as in
skred_b200/workloads.py plus `gold_frames` (how much the fixture holds)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from skred_b200 import workloads as W  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load_luts():
    return dict(np.load(os.path.join(GOLD, "notamy_luts.npz")))


def _wl(name, setup, events=None, tables=None, frames=8 * 512, voices=64):
    return {"name": name, "voices": voices, "tables": tables or {}, "setup": setup,
            "events": events or {}, "frames": frames, "gold_frames": frames}


def case_lut_adsr(luts):
    wl = W.config2(64, seconds=1.0, luts=luts)
    ev = {}
    for v in range(64):
        ev.setdefault(3, []).append(("envelope_velocity", v, 0.0))
        ev.setdefault(6, []).append(("envelope_velocity", v, 1.0))
    wl["events"] = ev
    wl["frames"] = wl["gold_frames"] = 10 * 512 + 168      # ragged tail like config 1
    wl["name"] = "lut_adsr"
    return wl


def case_fxpt_lut(luts):
    """fxpt LUT convention: (float)int16 / 32767.0f, installed as looping tables."""
    tables = {}
    for i, k in enumerate(("sine_fxpt_0", "triangle_fxpt_0", "impulse_fxpt_0")):
        tables[210 + i] = ((luts[k].astype(np.float32) / np.float32(32767.0)).astype(np.float32), {})
    setup = []
    for v in range(64):
        setup += [("wave_set", v, 210 + v % 3), ("freq_set", v, 27.5 * 2.0 ** (v / 9.0)),
                  ("amp_set", v, 0.5), ("pan_set", v, (v % 5 - 2) / 2.0)]
    return _wl("fxpt_lut", setup, tables=tables)


def case_korg_cz_filter(luts):
    wl = W.config3(64, seconds=1.0)
    wl["frames"] = wl["gold_frames"] = 8 * 512
    wl["name"] = "korg_cz_filter"
    # cover CZ on non power-of-two tables and extreme distortion / resonance too
    wl["setup"] += [("cz_set", 60, 6, 1.5), ("cz_set", 61, 7, -0.5), ("cz_set", 62, 1, 0.0),
                    ("mmf_set_res", 63, 40.0), ("wave_set", 59, 120), ("cz_set", 59, 3, 0.7),
                    ("voice_trigger", 59)]
    return wl


def case_pcm_retrigger(luts):
    wl = W.config4(64, seconds=16 * 512 / 44100.0, rate_hz=60.0)
    wl["frames"] = wl["gold_frames"] = 16 * 512
    wl["name"] = "pcm_retrigger"
    return wl


def case_mods(luts):
    s = []
    # forward-index FM (0.sk style), backward-index FM, 3-level chain, feedback pair
    s += [("wave_set", 0, 0), ("freq_set", 0, 440.0), ("amp_set", 0, 4.0), ("freq_mod_set", 0, 1, 10.0),
          ("wave_set", 1, 0), ("freq_set", 1, 1.0), ("amp_set", 1, 50.0), ("wave_mute", 1, 1)]
    s += [("wave_set", 5, 0), ("freq_set", 5, 3.0), ("amp_set", 5, 20.0), ("wave_mute", 5, 1),
          ("wave_set", 7, 1), ("freq_set", 7, 110.0), ("amp_set", 7, 3.0), ("freq_mod_set", 7, 5, 2.0)]
    s += [("wave_set", 10, 0), ("freq_set", 10, 0.7), ("amp_set", 10, 1.0), ("wave_mute", 10, 1),
          ("wave_set", 11, 4), ("freq_set", 11, 55.0), ("amp_set", 11, 2.0), ("amp_mod_set", 11, 10, 1.0),
          ("wave_set", 12, 36), ("freq_set", 12, 220.0), ("amp_set", 12, 2.0), ("pan_mod_set", 12, 11, 0.4),
          ("cz_set", 12, 2, 0.3), ("cmod_set", 12, 10, 0.5)]
    s += [("wave_set", 20, 0), ("freq_set", 20, 100.0), ("amp_set", 20, 1.0), ("freq_mod_set", 20, 21, 0.5),
          ("wave_set", 21, 0), ("freq_set", 21, 150.0), ("amp_set", 21, 1.0), ("freq_mod_set", 21, 20, 0.5)]
    # self references: AM on itself, CZ-mod on itself, pan-mod on itself, FM on itself (= no FM)
    s += [("wave_set", 30, 0), ("freq_set", 30, 330.0), ("amp_set", 30, 1.5), ("amp_mod_set", 30, 30, 0.8),
          ("wave_set", 31, 40), ("freq_set", 31, 82.0), ("amp_set", 31, 1.5), ("cz_set", 31, 1, 0.2), ("cmod_set", 31, 31, 0.3),
          ("wave_set", 32, 2), ("freq_set", 32, 61.0), ("amp_set", 32, 1.5), ("pan_mod_set", 32, 32, 0.5),
          ("wave_set", 33, 3), ("freq_set", 33, 77.0), ("amp_set", 33, 1.5), ("freq_mod_set", 33, 33, 5.0)]
    # CZ with a negative modulator (the +1.0f branch, synth.c:264) and the default depth-0 edge
    s += [("wave_set", 40, 0), ("freq_set", 40, 123.0), ("amp_set", 40, 1.0), ("cz_set", 40, 1, 0.4), ("cmod_set", 40, -1, 0.0),
          ("wave_set", 41, 0), ("freq_set", 41, 124.0), ("amp_set", 41, 1.0), ("cz_set", 41, 5, 0.4)]
    # a modulator that is silent (amp 0) and one that finishes (one-shot)
    s += [("wave_set", 50, 0), ("freq_set", 50, 2.0),
          ("wave_set", 51, 0), ("freq_set", 51, 200.0), ("amp_set", 51, 1.0), ("amp_mod_set", 51, 50, 1.0),
          ("wave_set", 52, 101), ("freq_midi", 52, 60.0), ("amp_set", 52, 8.0), ("wave_mute", 52, 1), ("voice_trigger", 52),
          ("wave_set", 53, 0), ("freq_set", 53, 300.0), ("amp_set", 53, 1.0), ("freq_mod_set", 53, 52, 3.0)]
    ev = {4: [("freq_mod_set", 0, 5, 4.0), ("amp_set", 50, 0.5)],        # regroup mid-stream
          6: [("voice_trigger", 52), ("wave_reset", 0, 20)]}
    return _wl("mods", s, events=ev, frames=10 * 512)


def case_misc(luts):
    s = []
    for v in range(8):                         # sample & hold + quantize
        s += [("wave_set", v, v % 5), ("freq_set", v, 50.0 + 31.0 * v), ("amp_set", v, 0.4),
              ("hold", v, 1 + 3 * v), ("wave_quant", v, 2 + v)]
    for v in range(8, 16):                     # reverse / loop / reverse one-shot
        s += [("wave_set", v, 32 + v), ("freq_set", v, 70.0 + 13.0 * v), ("amp_set", v, 0.4),
              ("wave_dir", v, 1), ("wave_loop", v, v % 2)]
    s += [("wave_set", 16, 103), ("freq_midi", 16, 62.0), ("amp_set", 16, 1.0), ("wave_dir", 16, 1), ("voice_trigger", 16),
          ("wave_set", 17, 104), ("freq_midi", 17, 50.0), ("amp_set", 17, 1.0), ("wave_loop", 17, 1), ("voice_trigger", 17),
          ("wave_set", 18, 105), ("freq_midi", 18, 70.0), ("amp_set", 18, 1.0), ("wave_loop", 18, 1), ("wave_dir", 18, 1), ("voice_trigger", 18)]
    for v in range(20, 24):                    # noise sources w5 (table) and w6 (shared draw)
        s += [("wave_set", v, 5 + v % 2), ("freq_set", v, 500.0 * (v - 19)), ("amp_set", v, 0.2),
              ("filter_mode", v, 1 + v % 5), ("mmf_set_freq", v, 900.0), ("mmf_set_res", v, 2.0)]
    for v in range(24, 30):                    # smoother off / custom, disconnect, all five filter modes
        s += [("wave_set", v, 1), ("freq_set", v, 90.0 + v), ("amp_set", v, 0.3), ("smoother", v, 0.0 if v % 2 else 0.3),
              ("filter_mode", v, 1 + v % 5), ("mmf_set_freq", v, 300.0 + 100 * v), ("mmf_set_res", v, 0.9)]
    s += [("wave_mute", 29, 1), ("voice_copy", 3, 40), ("voice_copy", 25, 41), ("voice_copy", 12, 42)]
    s += [("wave_set", 45, 0), ("freq_set", 45, 20000.0), ("amp_set", 45, 0.2),     # inc > table/2
          ("wave_set", 46, 36), ("freq_set", 46, 44000.0), ("amp_set", 46, 0.2),    # multi-wrap fmodf
          ("wave_set", 47, 0), ("freq_set", 47, 0.0), ("amp_set", 47, 0.2)]         # frozen phase
    ev = {2: [("volume_set", 3.0), ("amp_set", 4, 0.0), ("pan_set", 5, -1.0)],
          3: [("voice_copy", 0, 43), ("wave_reset", 0, 2), ("wave_set", 6, 110), ("voice_trigger", 6)],
          5: [("wave_reset", 0, 100)],            # invalid index: reset ALL voices (synth.c:1140-1144)
          6: [("wave_set", 1, 0), ("freq_set", 1, 440.0), ("amp_set", 1, 1.0)]}
    return _wl("misc", s, events=ev, frames=8 * 512 + 100)


SYNTHETIC = {
    "lut_adsr": case_lut_adsr,
    "fxpt_lut": case_fxpt_lut,
    "korg_cz_filter": case_korg_cz_filter,
    "pcm_retrigger": case_pcm_retrigger,
    "mods": case_mods,
    "misc": case_misc,
}


def drive_setup(s, wl):
    W.install(s, wl)


def drive_render(s, wl, frames=None):
    return s.render(wl["frames"] if frames is None else frames, events=wl["events"])


def wex_scenario(s):
    """`/wex<slot>` (wave_table_dynamic_expand, wire.c:553-586) rescales a loaded user sample IN PLACE — same pointer, same
    size — while voices play it; the reference reads the edited floats from the next frame on.  A user table with peaks
    beyond +-1 in slot 300, six voices on it (plain, CZ, filtered), two callbacks, `/wex300`, three more callbacks.
    Returns (mix, state)."""
    from skred_b200.host import install_table
    rng = np.random.RandomState(11)
    tbl = (2.5 * np.sin(np.linspace(0, 6 * np.pi, 1024)) + 0.3 * rng.randn(1024)).astype(np.float32)
    install_table(s, 300, tbl)
    s._wex_table = tbl                        # (the library reads this memory: keep it alive)
    for v in range(6):
        s.wire("v%d w300 f%d a0.1 p%.1f" % (v, 110 * (v + 1), (v - 3) / 4.0))
    s.wire("v2 c1,0.4")
    s.wire("v3 J1 K900 Q2")
    out = [s.render(1024)]
    s.wire("/wex300")
    out.append(s.render(1536))
    return np.concatenate(out), s.state()


def recording_scenario(s, tmpdir):
    """`v.. r1` marks voices for recording, `<0.5` starts a recording, `*` stops it and writes skred-<pid>-<ms>.wav through
    wire.c's save_wav (wire.c:94-185, 698, 816-848; the copy of the per-voice tap into the recording is skred.c:120-131).
    Eight sounding voices (one FM pair, a muted one), three of them recorded.  Returns the bytes of the WAV file."""
    import glob
    s.record_init(1.0)
    for v in range(8):
        s.wire("v%d w%d f%d a%.2f p%.1f" % (v, v % 5, 110 * (v + 1), 0.05 + 0.4 * (v % 3), (v - 4) / 5.0))
    s.wire("v1 F2,3.0")
    s.wire("v5 m1")
    s.wire("v6 a2.5")                       # the loudest voice is NOT recorded: save_wav's scale still depends on it
    for v in (0, 1, 4):
        s.wire("v%d r1" % v)
    s.render_recording(1024)
    s.wire("<0.5")
    s.render_recording(4 * 512)
    cwd = os.getcwd()
    os.chdir(str(tmpdir))
    try:
        for f in glob.glob("skred-*.wav"):
            os.remove(f)
        s.wire("*")
        files = glob.glob("skred-*.wav")
        assert len(files) == 1, files
        data = open(files[0], "rb").read()
        os.remove(files[0])
    finally:
        os.chdir(cwd)
    return data


def oversized_components(V=4096):
    """Modulation components above the 1,024 voices one CTA holds (legal in the reference, used to be refused).
    (a) CYCLIC: a feedback FM pair (voices 0 <-> 1) whose voice 0 also AM-modulates 1,500 carriers -> one component of 1,502
        voices that has to stay frame-lock-step (k_render_bins_huge);
    (b) ACYCLIC: one LFO (voice 2000) pan- and CZ-modulating 1,600 carriers -> levels (k_render_levels), no size limit;
    (c) CYCLIC, 100 voices: one CTA, frame-lock-step (k_render_bins)."""
    s = [("wave_set", 0, 0), ("freq_set", 0, 100.0), ("amp_set", 0, 1.0), ("freq_mod_set", 0, 1, 0.5),
         ("wave_set", 1, 0), ("freq_set", 1, 150.0), ("amp_set", 1, 1.0), ("freq_mod_set", 1, 0, 0.5)]
    for i in range(1500):
        v = 2 + i
        s += [("wave_set", v, i % 5), ("freq_set", v, 60.0 + 0.37 * i), ("amp_set", v, 0.02), ("amp_mod_set", v, 0, 0.5 + (i % 4) * 0.1),
              ("pan_set", v, (i % 21 - 10) / 10.0)]
        if i % 3 == 0:
            s += [("filter_mode", v, 1), ("mmf_set_freq", v, 500.0 + 3.0 * i)]
    s += [("wave_set", 2000, 0), ("freq_set", 2000, 1.5), ("amp_set", 2000, 1.0), ("wave_mute", 2000, 1)]
    for i in range(1600):
        v = 2001 + i
        s += [("wave_set", v, 32 + i % 12), ("freq_set", v, 50.0 + 0.21 * i), ("amp_set", v, 0.02), ("pan_mod_set", v, 2000, 0.3),
              ("cz_set", v, 1 + i % 5, 0.2 + 0.01 * (i % 30)), ("cmod_set", v, 2000, 0.2)]
    # (c) a mid-size CYCLIC component: feedback pair 3700 <-> 3701 + 98 readers = 100 voices -> one CTA (k_render_bins)
    s += [("wave_set", 3700, 0), ("freq_set", 3700, 80.0), ("amp_set", 3700, 1.0), ("freq_mod_set", 3700, 3701, 0.4),
          ("wave_set", 3701, 1), ("freq_set", 3701, 120.0), ("amp_set", 3701, 1.0), ("amp_mod_set", 3701, 3700, 0.6)]
    for i in range(98):
        v = 3702 + i
        s += [("wave_set", v, i % 5), ("freq_set", v, 90.0 + 1.3 * i), ("amp_set", v, 0.05), ("freq_mod_set", v, 3700 + i % 2, 1.0 + 0.1 * i),
              ("pan_set", v, (i % 9 - 4) / 4.0)]
    ev = {3: [("amp_set", 0, 0.7), ("freq_set", 2000, 2.5), ("pan_set", 700, -0.5), ("amp_set", 3701, 0.5)]}
    return _wl("oversized_components", s, events=ev, frames=6 * 512, voices=V)
