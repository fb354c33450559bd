"""CPU: the C-ABI libraries load and export every symbol the headers declare
(no compute calls — there is no GPU here), and the engine refuses to run
without a B200 instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SYNTH_H_FUNCTIONS = """synth synth_init synth_free audio_rng_init audio_rng_next audio_rng_float osc_get_phase_inc
osc_set_freq cz_phasor osc_next osc_set_wave_table_index osc_trigger quantize_bits_int mmf_init mmf_set_params
mmf_set_freq mmf_set_res mmf_process envelope_init amp_envelope_trigger amp_envelope_release amp_envelope_step
volume_set envelope_is_flat cz_set cmod_set amp_set pan_set wave_quant freq_set voice_set voice_copy wave_set
wave_mute wave_dir freq_midi amp_mod_set envelope_velocity envelope_set wave_reset freq_mod_set pan_mod_set
voice_format voice_show voice_show_all voice_trigger wave_default wave_loop midi2hz wave_table_init wave_free
voice_init synth_stats synth_voice_bench""".split()

SYNTH_H_DATA = """requested_synth_frames_per_callback synth_frames_per_callback synth_sample_count volume_user
volume_final volume_smoother_gain volume_smoother_smoothing volume_threshold volume_smoother_higher_smoothing
wave_table_data wave_size wave_rate wave_one_shot wave_loop_enabled wave_loop_start wave_loop_end wave_midi_note
wave_offset_hz wave_is_miniwav voice_phase voice_phase_inc voice_sample voice_amp voice_user_amp voice_pan
voice_pan_left voice_pan_right voice_table voice_table_size voice_table_rate voice_finished voice_one_shot
voice_direction voice_loop_enabled voice_filter voice_filter_mode voice_amp_envelope voice_use_amp_envelope
voice_cz_mode voice_cz_distortion voice_cz_mod_osc voice_cz_mod_depth voice_freq_mod_osc voice_freq_mod_depth
voice_amp_mod_osc voice_amp_mod_depth voice_pan_mod_osc voice_pan_mod_depth voice_smoother_enable
voice_smoother_gain voice_smoother_smoothing voice_sample_hold voice_sample_hold_count voice_sample_hold_max
voice_quantize voice_disconnect voice_record voice_mark_a voice_mark_b voice_mark_go""".split()


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(skb_[a-z_0-9]+)\s*\(", txt)))


def test_engine_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(ROOT, "skred_b200", "libskred_b200.so"))
    names = declared("skred_b200.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    lib.skb_backend_name.restype = C.c_char_p
    assert lib.skb_backend_name() == b"cuda-sm100a"


def test_port_exports_the_same_abi():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libskred_port.so"))
    for n in declared("skred_b200.h"):
        assert hasattr(lib, n), n


@pytest.mark.parametrize("v", [64, 65536])
def test_shim_exports_synth_h_and_additions(v):
    C.CDLL(os.path.join(ROOT, "skred_b200", "libskred_b200.so"), mode=C.RTLD_GLOBAL)
    lib = C.CDLL(os.path.join(ROOT, "skred_b200", "libskred_shim_v%d.so" % v))
    for n in SYNTH_H_FUNCTIONS + SYNTH_H_DATA + [x for x in declared("skred_b200_shim.h") if x.startswith("skb_shim")]:
        assert hasattr(lib, n), n
    assert C.c_int.in_dll(lib, "voice_phase__len__").value == v          # synth.c:30-32


def test_no_cpu_fallback():
    """Without a B200, skb_create must fail with SKB_ERR_NO_DEVICE (-3)."""
    import shutil
    if shutil.which("nvidia-smi"):
        pytest.skip("a GPU is present")
    lib = C.CDLL(os.path.join(ROOT, "skred_b200", "libskred_b200.so"))

    class Cfg(C.Structure):
        _fields_ = [(k, C.c_int32) for k in ("abi", "device", "n_voices", "max_frames", "rank", "world")] + [
            ("flags", C.c_uint32), ("_r", C.c_int32)]
    e = C.c_void_p()
    cfg = Cfg(1, 0, 64, 512, 0, 1, 0, 0)
    assert lib.skb_create(C.byref(e), C.byref(cfg)) == -3
    assert not e.value


def test_product_never_references_the_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "skred_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")) :
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt, os.path.join(d, f)
