import numpy as np

PATCH_IDS = [0, 1, 3, 5, 7, 8, 15, 16, 17, 21, 23, 26, 29, 30, 31, 41, 42, 64, 71, 73]
FULL_SCALE_TOL = 1e-5      # north_star: within 1e-5 of full scale (1.0) per sample on float paths


WAV_PATCH_IDS = [10, 13, 20, 36, 38, 43, 44, 47, 69]     # shipped patches that load user samples with `:wN,slot` (wire.c:406-441)


def load_wav_patch(s, golden_wav, n, tmpdir):
    """Feed a user-sample patch to `s` with the wav files it names (fixtures) in the working directory,
    which is where wave_load looks for "N.wav" (wire.c:409)."""
    import os
    for wn in golden_wav["p%d_wavs" % n]:
        with open(os.path.join(str(tmpdir), "%d.wav" % int(wn)), "wb") as f:
            f.write(golden_wav["wav%d" % int(wn)].tobytes())
    cwd = os.getcwd()
    os.chdir(str(tmpdir))
    try:
        s.load_lines(str(golden_wav["p%d_text" % n]).splitlines())
    finally:
        os.chdir(cwd)


def patch_lines(golden, n):
    return str(golden["p%d_text" % n]).splitlines()


def trace_render(s, nframes, block=512):
    out = np.zeros((nframes, 2), dtype=np.float32)
    phases, fin = [], []
    for k in range(0, nframes, block):
        n = min(block, nframes - k)
        s.render(n, block=block, out=out[k:k + n])
        st = s.state()
        phases.append(st["phase"].copy())
        fin.append(st["finished"].copy())
    return out, np.array(phases), np.array(fin)


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def assert_state_equal(a, b, exact_keys=None, tol_keys=(), tol=FULL_SCALE_TOL):
    for k in a:
        if k in tol_keys:
            d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
            assert np.nanmax(d) <= tol, (k, float(np.nanmax(d)))
        elif exact_keys is None or k in exact_keys:
            assert np.array_equal(bits(a[k]), bits(b[k])), (k, np.nonzero(bits(a[k]) != bits(b[k])))
