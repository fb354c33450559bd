"""CPU: the drop-in's host side (skred_b200/csrc/synth_shim.c — the synth.h setters re-implemented over the engine
C-ABI) against the compiled reference: SURVEY §8c pin (4) and the §8b error convention.

Random skode lines go through the reference's own wire() (wire.c:591-867) into (a) the reference's synth.c and
(b) the product shim (here over the CPU restatement: the shim is the same C file that the CUDA drop-in links);
after every few lines both render a callback, and EVERY array of synth.def (synth.def:1-89) is compared word for
word — floats by their bits — together with the filter and envelope structs.  Direct setter calls with out-of-range
arguments must return the reference's codes (100 / 101, synth.c:661, 834, 844, 860, 894, 1087) and leave the same
state behind.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O

F, I = C.c_float, C.c_int
# synth.def:1-89 minus the pointer arrays (voice_table / wave_table_data: compared through the sizes and the
# rendered audio) and the timespec marks
VOICE_ARRAYS = {
    "voice_phase": F, "voice_phase_inc": F, "voice_table_size": I, "voice_one_shot": I, "voice_finished": I,
    "voice_loop_enabled": I, "voice_table_rate": F, "voice_loop_start": I, "voice_loop_end": I, "voice_midi_note": F,
    "voice_midi_transpose": F, "voice_link_midi_a": F, "voice_link_midi_b": F, "voice_link_velo_a": F,
    "voice_link_velo_b": F, "voice_link_trig": F, "voice_offset_hz": F, "voice_freq": F, "voice_note": F,
    "voice_sample": F, "voice_sample_hold": F, "voice_sample_hold_count": I, "voice_sample_hold_max": I,
    "voice_amp": F, "voice_user_amp": F, "voice_pan_left": F, "voice_pan_right": F, "voice_pan": F,
    "voice_use_amp_envelope": I, "voice_freq_mod_osc": I, "voice_freq_mod_depth": F, "voice_freq_scale": F,
    "voice_pan_mod_osc": I, "voice_amp_mod_osc": I, "voice_cz_mod_osc": I, "voice_pan_mod_depth": F,
    "voice_amp_mod_depth": F, "voice_cz_mod_depth": F, "voice_disconnect": I, "voice_quantize": I,
    "voice_direction": I, "voice_phase_reset": I, "voice_record": I, "voice_wave_table_index": I,
    "voice_cz_mode": I, "voice_cz_distortion": F, "voice_smoother_enable": I, "voice_smoother_gain": F,
    "voice_smoother_smoothing": F, "voice_glissando_enable": I, "voice_glissando_speed": F,
    "voice_glissando_target": F, "voice_filter_freq": F, "voice_filter_res": F, "voice_filter_mode": I,
    "voice_loop_valid": I, "voice_loop_length": I, "voice_loop_start_f": F, "voice_loop_end_f": F, "voice_mark_go": I,
}
WAVE_ARRAYS = {"wave_size": I, "wave_rate": F, "wave_one_shot": I, "wave_loop_enabled": I, "wave_loop_start": I,
               "wave_loop_end": I, "wave_midi_note": F, "wave_offset_hz": F, "wave_is_miniwav": I}
SCALARS = {"volume_user": F, "volume_final": F, "volume_smoother_gain": F, "synth_frames_per_callback": I}
V = 64


def _need():
    if not (O.have_ref(V) and os.path.exists(O.port_lib_path(V))):
        pytest.skip("oracle libraries not built")


def snapshot(s):
    s.sync_state()
    out = {}
    for n, t in VOICE_ARRAYS.items():
        out[n] = s.array(n, t).copy()
    for n, t in WAVE_ARRAYS.items():
        out[n] = s.array(n, t, 1200).copy()
    for n, t in SCALARS.items():
        out[n] = np.array([s.scalar(n, t).value])
    filt = np.zeros((V, 9), dtype=np.float32)
    envf = np.zeros((V, 9), dtype=np.float32)
    envu = np.zeros((V, 2), dtype=np.uint64)
    act = np.zeros(V, dtype=np.int32)
    for v in range(V):
        s.lib.ref_get_filter(v, filt[v].ctypes.data)
        s.lib.ref_get_envelope(v, envf[v].ctypes.data, envu[v].ctypes.data, act[v:v + 1].ctypes.data)
    out.update(filter=filt, env_f=envf, env_u=envu, env_active=act, ssc=np.array([s.sample_count], dtype=np.uint64))
    return out


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_same(a, b, ctx):
    for k in a:
        same = bits(a[k]) == bits(b[k])
        if a[k].dtype == np.float32:
            same |= np.isnan(a[k]) & np.isnan(b[k])      # a NaN is a NaN: the payload/sign an invalid op leaves is the FPU's
        if not np.all(same):
            i = np.argwhere(~same)[0]
            raise AssertionError("%s differs at %s: ref %r, shim %r  [%s]" % (k, tuple(i), a[k][tuple(i)], b[k][tuple(i)], ctx))


def rand_line(rng):
    """One line of skode: a voice select plus 1-4 parameter atoms (wire.c:591-867)."""
    def f(lo, hi, nd=3):
        return ("%." + str(nd) + "f") % rng.uniform(lo, hi)

    def voice(p_bad=0.06):
        return str(rng.randint(-3, 70)) if rng.rand() < p_bad else str(rng.randint(0, V))
    waves = [0, 1, 2, 3, 4, 5, 6, 32, 40, 62, 63, 100, 120, 166, 167, 199, 1199, 1200, -1]
    atoms = [
        lambda: "a" + f(-0.2, 2.0), lambda: "a0",
        lambda: "A" + voice() + "," + f(-1, 2), lambda: "A" + voice(),
        lambda: "b" + str(rng.randint(0, 2)), lambda: "b", lambda: "B" + str(rng.randint(0, 2)), lambda: "B",
        lambda: "c" + str(rng.randint(-1, 9)) + "," + f(-0.2, 1.3), lambda: "c" + str(rng.randint(0, 8)), lambda: "c",
        lambda: "C" + voice() + "," + f(-1, 1), lambda: "C" + voice(), lambda: "C-1",
        lambda: "f" + f(-10, 46000, 2), lambda: "f" + f(20, 4000, 2), lambda: "f0",
        lambda: "F" + voice() + "," + f(-2, 10), lambda: "F" + voice(), lambda: "F-1",
        # link targets and the copy destination are not range-checked by wire.c (wire.c:648-665, 851) and the
        # fan-out indexes the arrays with them: only valid voices, anything else is undefined behaviour upstream
        lambda: "g" + f(0, 2), lambda: "G" + voice(0) + "," + voice(0), lambda: "H" + voice(0), lambda: "L" + voice(0),
        lambda: "h" + str(rng.randint(0, 40)),
        lambda: "J" + str(rng.randint(0, 7)), lambda: "K" + f(-100, 25000, 1), lambda: "Q" + f(-1, 12),
        lambda: "l" + f(0, 1.5), lambda: "l0", lambda: "l1",
        lambda: "m" + str(rng.randint(0, 2)),
        lambda: "n" + f(-5, 135, 2), lambda: "n" + str(rng.randint(24, 96)), lambda: "N" + str(rng.randint(-12, 13)),
        lambda: "p" + f(-1.3, 1.3), lambda: "P" + voice() + "," + f(-1, 1), lambda: "P" + voice(),
        lambda: "q" + str(rng.randint(0, 18)),
        lambda: "s" + f(-0.1, 0.5), lambda: "s0",
        lambda: "t" + ",".join(f(0, 0.3) for _ in range(4)), lambda: "t0,0,1,0",
        lambda: "T", lambda: "T",
        lambda: "w" + str(waves[rng.randint(len(waves))]), lambda: "w" + str(rng.randint(32, 63)),
        lambda: "w" + str(rng.randint(100, 167)),
        lambda: "V" + f(0, 3), lambda: "/", lambda: ">" + voice(0),
    ]
    line = "v" + voice(0.03)
    for _ in range(rng.randint(1, 5)):
        line += " " + atoms[rng.randint(len(atoms))]()
    if rng.rand() < 0.01:
        line += " S" + voice(0.3)
    return line


def run_random_wire_streams(seed, make_dut):
    rng = np.random.RandomState(seed)
    ref, dut = O.RefSkred(V, run_seq=False), make_dut(V, run_seq=False)
    assert_same(snapshot(ref), snapshot(dut), "after init")
    lines = []
    for step in range(60):
        for _ in range(rng.randint(1, 12)):
            ln = rand_line(rng)
            lines.append(ln)
            ref.wire(ln)
            dut.wire(ln)
        # host-visible arrays right after the setters ran (before any render) ...
        assert_same(snapshot(ref), snapshot(dut), "step %d before render; last lines: %s" % (step, lines[-12:]))
        n = int(rng.choice([512, 512, 512, 64, 300]))      # <= 512: the harness sizes the reference tap for 512 frames
        oa, ob = ref.render(n, block=n), dut.render(n, block=n)
        both_finite = np.isfinite(oa) & np.isfinite(ob)
        ctx = "step %d mix; last lines: %s" % (step, lines[-12:])
        assert np.array_equal(np.isfinite(oa), np.isfinite(ob)), "finite-ness of the mix differs, " + ctx
        with np.errstate(invalid="ignore"):
            err = np.abs(oa.astype(np.float64) - ob)
        err[~both_finite] = 0.0
        # 1e-5 of full scale (1.0); a mix driven beyond full scale by the random gains is held to the same RELATIVE bound
        bound = 1e-5 * max(1.0, float(np.max(np.abs(oa[both_finite]), initial=0.0)))
        assert float(err.max(initial=0.0)) <= bound, "mix differs by %g (peak %g) at %s, %s" % (
            err.max(), np.abs(oa[both_finite]).max(), np.unravel_index(np.argmax(err), err.shape), ctx)
        # ... and after the callback (evolving words included)
        assert_same(snapshot(ref), snapshot(dut), "step %d after render; last lines: %s" % (step, lines[-12:]))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_wire_streams_leave_identical_arrays(seed):
    _need()
    run_random_wire_streams(seed, O.PortSkred)


def test_setter_return_codes_and_range_rules():
    """0 ok / 100 invalid voice or out of range / 101 frequency out of range."""
    _need()
    ref, dut = O.RefSkred(V, run_seq=False), O.PortSkred(V, run_seq=False)
    calls = []
    # only these check the voice index (synth.c:883, 899, 906, 1147; wave_reset checks its second argument, 1141);
    # every other setter indexes its arrays with whatever it is given (undefined behaviour in the reference:
    # wire.c only ever passes the `v` it validated in voice_set, synth.c:656-662)
    for v in (-1, -7, V, V + 1, 1000):
        calls += [("amp_mod_set", v, 3, 0.5), ("freq_mod_set", v, 2, 1.0), ("pan_mod_set", v, 1, 0.3),
                  ("envelope_velocity", v, 1.0), ("envelope_velocity", v, 0.0)]
    calls += [("wave_reset", 0, 5), ("amp_set", 9, 0.7), ("wave_reset", 0, -1)]       # invalid n: every voice is reset
    for v in (0, 5, V - 1):
        calls += [("amp_set", v, 0.5), ("amp_set", v, -0.1), ("amp_set", v, 0.0),
                  ("pan_set", v, -1.0), ("pan_set", v, 1.0), ("pan_set", v, 1.01), ("pan_set", v, -1.5),
                  ("freq_set", v, 0.0), ("freq_set", v, 440.0), ("freq_set", v, 44099.9), ("freq_set", v, 44100.0),
                  ("freq_set", v, -1.0), ("freq_midi", v, 0.0), ("freq_midi", v, 127.0), ("freq_midi", v, 127.5),
                  ("freq_midi", v, -0.5), ("wave_set", v, 0), ("wave_set", v, 6), ("wave_set", v, 7),
                  ("wave_set", v, 40), ("wave_set", v, 63), ("wave_set", v, 100), ("wave_set", v, 166),
                  ("wave_set", v, 167), ("wave_set", v, 1199), ("wave_set", v, 1200), ("wave_set", v, -1),
                  ("mmf_set_res", v, 0.0), ("mmf_set_res", v, -1.0), ("mmf_set_res", v, 2.0),
                  ("mmf_set_freq", v, 1000.0), ("mmf_set_freq", v, 0.0), ("mmf_set_freq", v, 30000.0),
                  ("amp_mod_set", v, 3, 0.5), ("amp_mod_set", v, -1, 0.0), ("amp_mod_set", v, V, 0.5),
                  ("freq_mod_set", v, 2, 1.0), ("freq_mod_set", v, -1, 0.0), ("freq_mod_set", v, V + 3, 1.0),
                  ("pan_mod_set", v, 1, 0.3), ("pan_mod_set", v, -1, 0.3), ("pan_mod_set", v, 999, 0.3),
                  ("cmod_set", v, 4, 0.2), ("cmod_set", v, -1, 0.2), ("cz_set", v, 3, 0.4), ("cz_set", v, 9, 2.0),
                  ("wave_mute", v, 1), ("wave_mute", v, 0), ("wave_mute", v, -1), ("wave_dir", v, 1), ("wave_dir", v, -1),
                  ("wave_dir", v, 0), ("wave_loop", v, 1), ("wave_loop", v, -1), ("wave_quant", v, 8), ("wave_quant", v, 0),
                  ("wave_quant", v, 12),
                  ("envelope_set", v, 0.01, 0.1, 0.5, 0.2), ("envelope_set", v, 0.0, 0.0, 1.0, 0.0),
                  ("envelope_velocity", v, 1.0), ("envelope_velocity", v, 0.0), ("voice_trigger", v),
                  ("voice_copy", v, 7), ("voice_copy", 7, v), ("wave_default", v), ("wave_reset", 0, v)]
    calls += [("volume_set", 1.0), ("volume_set", -1.0), ("volume_set", 0.0), ("volume_set", 100.0)]
    codes = set()
    for i, c in enumerate(calls):
        ra, rb = ref.call(*c), dut.call(*c)
        assert ra == rb, (c, ra, rb)
        codes.add(ra)
        if i % 40 == 39:
            assert_same(snapshot(ref), snapshot(dut), "after %r" % (c,))
    assert_same(snapshot(ref), snapshot(dut), "end")
    assert {0, 100, 101} <= codes
    oa, ob = ref.render(1024), dut.render(1024)
    assert float(np.max(np.abs(oa.astype(np.float64) - ob))) <= 1e-5
    assert_same(snapshot(ref), snapshot(dut), "after render")
