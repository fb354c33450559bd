"""CPU dry run of the full-size parity harness (tests/full_size.py): the same code path as
tests/test_gpu_full_size.py at reduced sizes, with the CPU restatement behind the drop-in shim standing in
for the CUDA engine.  Pins the harness itself (voice subsets, checkpoints, queue-vs-callback event timing,
ragged tail) against the compiled reference before it is trusted on the GPU box."""
import numpy as np
import pytest

import full_size as FS
from oracle import oracle as O
from tests_util import FULL_SCALE_TOL
from skred_b200 import workloads as W


def _need():
    import os
    if not (O.have_ref(64) and os.path.exists(O.port_lib_path(64))):
        pytest.skip("oracle libraries not built")


def test_segments_cover_the_job():
    for total, n in ((2646000, 6), (26460000, 10), (1323000, 3), (5000, 4), (8192 * 3, 3)):
        segs = FS.segments(total, n)
        assert sum(segs) == total and all(s > 0 for s in segs)
        assert all(s % 8192 == 0 for s in segs[:-1])


def test_select_reindexes_an_arbitrary_subset(luts):
    wl = W.config5(96, seconds=600.0, luts=luts, event_seconds=3.0, stationary=True)
    sub = FS.select(wl, [5, 17, 50, 95])
    assert sub["voices"] == 4 and {c[1] for c in sub["setup"]} == {0, 1, 2, 3}
    assert [c for c in sub["setup"] if c[1] == 2] == [(c[0], 2) + tuple(c[2:]) for c in wl["setup"] if c[1] == 50]
    assert sorted(sub["pcm_tau"]) == sorted(i for i, v in enumerate([5, 17, 50, 95]) if v in wl["pcm_tau"])


@pytest.mark.parametrize("which", ["config4_dense", "config4_dense_ragged_end", "config5_subset", "config2_release"])
def test_full_size_harness_port_vs_reference_subsets(which, luts):
    _need()
    if which.startswith("config4_dense"):
        wl = W.config4(64, seconds=1.3, rate_hz=40.0)
        if which.endswith("ragged_end"):
            # a last callback of 340 frames with triggers due right after it: seq(340) fires what is due by c + 340
            wl["frames"] = 6 * 8192 + 340
        V, subsets, n_ck = 64, FS.chunks(range(64), 16), 3
    elif which == "config2_release":
        wl = W.config2(64, seconds=0.9, luts=luts)
        wl["timed"] = [(int(0.3 * 44100), ("envelope_velocity", v, 0.0)) for v in range(64)] + \
                      [(int(0.5 * 44100), ("envelope_velocity", v, 1.0)) for v in range(64)]
        wl["events"] = W.bucket(wl["timed"])
        V, subsets, n_ck = 64, FS.chunks(range(64), 32), 2
    else:
        wl = W.config5(64, seconds=600.0, luts=luts, event_seconds=2.0, stationary=True)
        wl["frames"] = 3 * 8192 + 100
        rng = np.random.RandomState(3)
        V, subsets, n_ck = 64, FS.chunks(np.sort(rng.choice(64, size=24, replace=False)), 8), 3
    segs = FS.segments(wl["frames"], n_ck)
    port = O.PortSkred(V, run_seq=False)
    sel = np.concatenate([np.asarray(s) for s in subsets])
    full = len(sel) == V
    got, got_states, _ = FS.product_render(port, wl, segs, voices=None if full else sel)
    want, ref_states, cpu = FS.reference_by_subsets(wl, subsets, segs, keep_mix=full, procs=4)
    FS.assert_checkpoints_equal(ref_states, got_states)
    if full:
        assert float(np.max(np.abs(got.astype(np.float64) - want))) <= FULL_SCALE_TOL
        assert float(np.abs(want).max()) > 1e-4
    assert cpu > 0.0
