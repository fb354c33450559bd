#!/usr/bin/env python3
"""Generate the golden fixtures from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (needs /root/reference):
    python oracle/build_oracle.py --voices 64 && python tests/golden/make_golden.py

Writes
  tests/golden/notamy_luts.npz   the float / fxpt LUTs of notamy/*_lutset*.h that config 2 installs
                                 into user wave slots (the reference ships them as dead data)
  tests/golden/patches.npz       for every .wav-free shipped patch: its text (the test INPUT) and the
                                 reference render of the first GOLD_FRAMES frames + end-of-block
                                 voice_phase / voice_finished traces (bit-exact targets)
  tests/golden/synthetic.npz     reference renders of the per-feature synthetic sets in
                                 tests/cases.py
  tests/golden/wav_patches.npz   four shipped patches that load user samples (`:wN,slot`), their wav files
                                 (bytes) and the reference render of 16 callbacks
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("SKRED_REF", "/root/reference")

GOLD_FRAMES = 8 * 512
PATCHES = [0, 1, 3, 5, 7, 8, 15, 16, 17, 21, 23, 26, 29, 30, 31, 41, 42, 64, 71, 73]
# patches that load user samples with `:wN,slot` (wire.c:406-441): the wav files they name travel as fixtures
WAV_PATCHES = {13: [13, 16], 36: [3], 38: [17, 9], 44: [19], 10: [24, 23, 22, 21], 20: [24, 23], 43: [32, 33, 3],
               47: [24, 30, 31, 32], 69: [17, 9]}


def parse_tables(path, ctype):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"const\s+%s\s+(\w+)\[(\d+)\]\s*(?:PROGMEM\s*)?=\s*\{([^}]*)\}" % ctype, txt):
        name, n, body = m.group(1), int(m.group(2)), m.group(3)
        vals = [x for x in body.replace("\n", " ").split(",") if x.strip()]
        assert len(vals) == n, (name, len(vals), n)
        if ctype == "float":
            out[name] = np.array([np.float32(x) for x in vals], dtype=np.float32)
        else:
            out[name] = np.array([int(x) for x in vals], dtype=np.int16)
    return out


def make_luts():
    d = {}
    for base in ("sine", "triangle", "impulse"):
        t = parse_tables(os.path.join(REF, "notamy", "%s_lutset.h" % base), "float")
        d["%s_lutable_0" % base] = t["%s_lutable_0" % base]
        tf = parse_tables(os.path.join(REF, "notamy", "%s_lutset_fxpt.h" % base), "int16_t")
        key = [k for k in tf if k.endswith("_0")][0]
        d["%s_fxpt_0" % base] = tf[key]
    np.savez_compressed(os.path.join(HERE, "notamy_luts.npz"), **d)
    return d


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


def main():
    from oracle.oracle import RefSkred
    import cases
    luts = make_luts()
    d = {}
    for n in PATCHES:
        p = os.path.join(REF, "%d.sk" % n)
        if not os.path.exists(p):
            continue
        text = open(p, "rb").read().decode("latin-1")
        s = RefSkred(64)
        s.load_lines(text.splitlines())
        out, ph, fin = trace_render(s, GOLD_FRAMES)
        d["p%d_text" % n] = np.array(text)
        d["p%d_out" % n] = out
        d["p%d_phase" % n] = ph
        d["p%d_finished" % n] = fin
        print("patch %d: peak %.5f" % (n, np.abs(out).max()))
    np.savez_compressed(os.path.join(HERE, "patches.npz"), **d)
    # user-sample patches: rendered with the cwd at the reference tree (wave_load opens "N.wav"); the wav bytes
    # are stored next to the render so that the tests can write them into a scratch directory
    d = {}
    cwd = os.getcwd()
    for n, wavs in WAV_PATCHES.items():
        text = open(os.path.join(REF, "%d.sk" % n), "rb").read().decode("latin-1")
        s = RefSkred(64)
        os.chdir(REF)
        try:
            s.load_lines(text.splitlines())
        finally:
            os.chdir(cwd)
        out, ph, fin = trace_render(s, 2 * GOLD_FRAMES)
        d["p%d_text" % n] = np.array(text)
        d["p%d_out" % n] = out
        d["p%d_phase" % n] = ph
        d["p%d_finished" % n] = fin
        d["p%d_wavs" % n] = np.array(wavs, dtype=np.int32)
        for wn in wavs:
            d["wav%d" % wn] = np.frombuffer(open(os.path.join(REF, "%d.wav" % wn), "rb").read(), dtype=np.uint8)
        print("wav patch %d: peak %.5f" % (n, np.abs(out).max()))
    np.savez_compressed(os.path.join(HERE, "wav_patches.npz"), **d)
    d = {}
    for name, fn in cases.SYNTHETIC.items():
        wl = fn(luts)
        s = RefSkred(wl["voices"])
        cases.drive_setup(s, wl)
        out = cases.drive_render(s, wl, wl["gold_frames"])
        st = s.state()
        d[name + "_out"] = out
        d[name + "_phase"] = st["phase"]
        d[name + "_finished"] = st["finished"]
        print("case %s: peak %.5f" % (name, np.abs(out).max()))
    np.savez_compressed(os.path.join(HERE, "synthetic.npz"), **d)


if __name__ == "__main__":
    main()
