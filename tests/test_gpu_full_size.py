"""GPU: BASELINE.json configs[1..4] at their FULL sizes against the compiled reference.

configs[1] 64 voices x 60 s, configs[2] 1,024 voices x 60 s, configs[3] 4,096 voices x 30 s with the dense
retrigger stream: the product renders the whole job through synth() in 8,192-frame calls (16 callbacks per
launch, events applied in-kernel at the 512-frame boundaries); the reference renders it as voice subsets on
every host core (tests/full_size.py).  Compared: the stereo mix SAMPLE FOR SAMPLE over the whole duration
(<= 1e-5 of full scale) and every evolving per-voice word of every voice BIT FOR BIT at checkpoints.
configs[4] 65,536 voices x 10 min: the product renders the full job; 256 voices drawn with seed 0x5EED0003
(SURVEY §8d parity (ii)) are rendered solo by the reference for the full 10 minutes and their evolving words
compared bit for bit at ten checkpoints.
"""
import json
import os
import time

import numpy as np
import pytest

import full_size as FS
from oracle import oracle as O
from tests_util import FULL_SCALE_TOL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 44100


def _record(name, **kw):
    """Job times of the full-size runs, kept beside the other GPU evidence (gpurun_out/ travels back)."""
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "full_size_times.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **kw)) + "\n")


def _need_ref():
    if not O.have_ref(FS.REF_V):
        pytest.skip("compiled reference (oracle/_ref/libskred_ref_v64.so) not present")


def _full_job(wl, V, n_checkpoints):
    segs = FS.segments(wl["frames"], n_checkpoints)
    gpu = O.DropinCuda(V, run_seq=False)
    got, got_states, t_gpu = FS.product_render(gpu, wl, segs)
    t0 = time.perf_counter()
    want, ref_states, cpu = FS.reference_by_subsets(wl, FS.chunks(range(V), FS.REF_V), segs)
    t_ref = time.perf_counter() - t0
    err = float(np.max(np.abs(got.astype(np.float64) - want)))
    _record(wl["name"], voices=V, frames=wl["frames"], product_synth_s=t_gpu, reference_wall_s=t_ref,
            reference_cpu_s=cpu, cores=os.cpu_count(), max_abs_err=err, peak=float(np.abs(want).max()))
    assert err <= FULL_SCALE_TOL, err
    assert float(np.abs(want).max()) > 1e-3
    FS.assert_checkpoints_equal(ref_states, got_states)
    return got_states


def test_config2_full_64_voices_60s(luts):
    """configs[1]: 64 voices, notamy LUT oscillators + ADSR + pan, 60 s (l0 at 30 s, l1 at 31 s)."""
    from skred_b200 import workloads as W
    _need_ref()
    wl = W.config2(64, seconds=60.0, luts=luts)
    assert wl["frames"] == 2646000
    st = _full_job(wl, 64, 6)
    assert int(st[-1]["env_active"].sum()) == 64          # re-triggered at 31 s, sustaining at 60 s


def test_config3_full_1024_voices_60s():
    """configs[2]: 1,024 voices, CZ phase distortion over the Korg tables + resonant biquad, 60 s."""
    from skred_b200 import workloads as W
    _need_ref()
    wl = W.config3(1024, seconds=60.0)
    _full_job(wl, 1024, 6)


def test_config4_full_4096_voices_30s_dense_retrigger():
    """configs[3]: 4,096 one-shot AMY sample voices pitch-shifted by midi note, 30 s, ~8 triggers per voice per
    second (~380 events per 512-frame callback, every 4th voice gated by its envelope instead)."""
    from skred_b200 import workloads as W
    _need_ref()
    wl = W.config4(4096, seconds=30.0)
    assert len(wl["timed"]) > 900000
    st = _full_job(wl, 4096, 3)
    fin = st[-1]["finished"]
    assert 0 < int(fin.sum()) < 4096


def test_config5_full_65536_voices_10min_subset_vs_reference(luts):
    """configs[4] at full width AND full length (26,460,000 frames, ~3.9 million timestamped events)."""
    from skred_b200 import workloads as W
    _need_ref()
    V = 65536
    wl = W.config5(V, seconds=600.0, luts=luts, stationary=True)
    assert wl["frames"] == 26460000
    rng = np.random.RandomState(0x5EED0003)
    sel = np.sort(rng.choice(V, size=256, replace=False))
    segs = FS.segments(wl["frames"], 10)
    gpu = O.DropinCuda(V, run_seq=False)
    _, got_states, t_gpu = FS.product_render(gpu, wl, segs, keep_mix=False, voices=sel)
    st = gpu.engine_stats()
    t0 = time.perf_counter()
    _, ref_states, cpu = FS.reference_by_subsets(wl, FS.chunks(sel, 16), segs, keep_mix=False)
    t_ref = time.perf_counter() - t0
    _record(wl["name"], voices=V, frames=wl["frames"], events=len(wl["timed"]), product_synth_s=t_gpu,
            rendered_voice_frames=int(st.active_voice_frames), reference_wall_s=t_ref, reference_cpu_s=cpu,
            reference_voices=256, cores=os.cpu_count())
    FS.assert_checkpoints_equal(ref_states, got_states)
    assert int(st.active_voice_frames) > 0.5 * V * wl["frames"]


def test_config5_full_width_first_5s_vs_16_reference_shards(luts):
    """configs[4], SURVEY 8d parity (i): ALL 65,536 voices, the first 5 s (220,500 frames = 430 callbacks and a
    ragged 340-frame tail, ~33,000 timestamped events) against 16 instances of the compiled reference of 4,096
    voices each (one per host core, voices of this load are independent), their mixes added in float64: the stereo
    mix sample for sample <= 1e-5 of full scale, and every evolving word of every voice bit for bit at 2.5 s and 5 s."""
    from skred_b200 import workloads as W
    V, shard_v, frames = 65536, 4096, 5 * SR
    if not O.have_ref(shard_v):
        pytest.skip("compiled reference for VOICE_MAX = 4096 not present")
    wl = W.config5(V, seconds=600.0, luts=luts, event_seconds=frames / float(SR) + 1.0, stationary=True)
    wl["frames"] = frames
    segs = FS.segments(frames, 2)
    assert sum(segs) == frames and segs[0] % 8192 == 0
    gpu = O.DropinCuda(V, run_seq=False)
    got, got_states, t_gpu = FS.product_render(gpu, wl, segs)
    t0 = time.perf_counter()
    want, ref_states, cpu = FS.reference_by_subsets(wl, FS.chunks(range(V), shard_v), segs, ref_v=shard_v)
    t_ref = time.perf_counter() - t0
    err = float(np.max(np.abs(got.astype(np.float64) - want)))
    _record("config5_full_width_5s", voices=V, frames=frames, events=len(wl["timed"]), product_synth_s=t_gpu,
            reference_wall_s=t_ref, reference_cpu_s=cpu, cores=os.cpu_count(), max_abs_err=err,
            peak=float(np.abs(want).max()))
    assert err <= FULL_SCALE_TOL, err
    assert float(np.abs(want).max()) > 1e-3
    FS.assert_checkpoints_equal(ref_states, got_states)
