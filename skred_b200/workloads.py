"""Synthetic voice/event loads of BASELINE.json `configs`, expressed as
lists of synth.h setter calls so the same load drives the product and (in
tests) the reference.  Fully deterministic: closed formulae + xorshift64.

A workload is a dict:
  voices   VOICE_MAX it is defined for
  tables   {slot: (float32 array, kwargs for install_table)} user wave slots to fill first
  setup    [(setter, *args), ...] applied before the first callback
  events   {callback_index: [(setter, *args), ...]} applied BEFORE that 512-frame callback,
           which is where the reference's seq() fires them (seq.c:170-178)
  frames   total frames of the full configuration
"""
import math

BLOCK = 512
SR = 44100


def callback_for_time(when_samples, block=BLOCK):
    """F8: seq() after the callback ending at count c fires items with
    when <= c + block; they are audible from callback index c/block.  The first
    seq() runs after callback 0, so nothing queued can land before callback 1."""
    return max(1, -(-int(when_samples) // block) - 1)


EVENT_CODES = {"voice_trigger": 1, "envelope_velocity": 2, "freq_set": 3, "freq_midi": 4, "amp_set": 5,
               "pan_set": 6, "wave_set": 7, "cz_set": 8, "mmf_set_freq": 9, "mmf_set_res": 10, "wave_mute": 11}


def bucket(timed, block=BLOCK):
    """[(when, call)] -> {callback_index: [call, ...]} under the F8 rule."""
    ev = {}
    for when, call in sorted(timed, key=lambda x: x[0]):
        ev.setdefault(callback_for_time(when, block), []).append(call)
    return ev


def to_skb_events(timed):
    """[(when, call)] -> structured array matching `skb_event` (include/skred_b200_shim.h)."""
    import numpy as np
    dt = np.dtype([("when", "<u8"), ("voice", "<i4"), ("code", "<i4"), ("a0", "<f4"), ("a1", "<f4")])
    timed = sorted(timed, key=lambda x: x[0])
    a = np.zeros(len(timed), dtype=dt)
    for i, (when, call) in enumerate(timed):
        a[i] = (when, call[1], EVENT_CODES[call[0]], call[2] if len(call) > 2 else 0.0,
                call[3] if len(call) > 3 else 0.0)
    return a


class XorShift64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF or 1

    def next(self):
        s = self.s
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        self.s = s
        return s

    def uniform(self):
        return (self.next() >> 11) / float(1 << 53)


def _freq(v):
    return 55.0 * 2.0 ** ((v % 48) / 12.0)


def lut_voice(v, V, luts_installed=True):
    """config-2 recipe: LUT oscillator + ADSR + pan (no filter / CZ / mods)."""
    k = v % 3
    amp = 40.0 / V
    if luts_installed:
        wave = 200 + k
        if k == 2:
            amp = amp / 128.0          # impulse_lutable_0 peaks at 128
    else:
        wave = (0, 4, 1)[k]            # built-in sine / triangle / square
    return [("wave_set", v, wave), ("freq_set", v, _freq(v)), ("amp_set", v, amp),
            ("pan_set", v, (v % 21 - 10) / 10.0), ("envelope_set", v, 0.01, 0.1, 0.5, 0.2),
            ("envelope_velocity", v, 1.0)]


def korg_voice(v, V):
    """config-3 recipe: Korg table + CZ phase distortion + resonant biquad + ADSR."""
    return [("wave_set", v, 32 + (v % 31)), ("freq_set", v, _freq(v)), ("amp_set", v, 40.0 / V),
            ("pan_set", v, (v % 21 - 10) / 10.0),
            ("cz_set", v, 1 + v % 7, 0.1 + 0.8 * (v % 11) / 11.0),
            ("filter_mode", v, 1 + v % 4), ("mmf_set_freq", v, 200.0 + 100.0 * (v % 50)),
            ("mmf_set_res", v, 0.5 + (v % 9)),
            ("envelope_set", v, 0.01, 0.1, 0.5, 0.2), ("envelope_velocity", v, 1.0)]


def pcm_voice(v, V):
    """config-4 recipe: AMY one-shot sample, pitch-shifted by midi note."""
    return [("wave_set", v, 100 + (v % 67)), ("freq_midi", v, 36.0 + (v % 48)), ("amp_set", v, 40.0 / V),
            ("pan_set", v, (v % 21 - 10) / 10.0)]


def config2(V=64, seconds=60.0, luts=None):
    tables = {}
    if luts is not None:
        tables = {200: (luts["sine_lutable_0"], {}), 201: (luts["triangle_lutable_0"], {}),
                  202: (luts["impulse_lutable_0"], {})}
    setup = []
    for v in range(V):
        setup += lut_voice(v, V, luts is not None)
    timed = []
    for v in range(V):
        if seconds > 30.0:
            timed.append((30 * SR, ("envelope_velocity", v, 0.0)))
        if seconds > 31.0:
            timed.append((31 * SR, ("envelope_velocity", v, 1.0)))
    timed.sort(key=lambda x: x[0])
    return {"name": "config2_lut_adsr_pan", "voices": V, "tables": tables, "setup": setup,
            "timed": timed, "events": bucket(timed), "frames": int(seconds * SR)}


def config3(V=1024, seconds=60.0, cmod_pairs=0):
    setup = []
    for v in range(V):
        setup += korg_voice(v, V)
    # sub-variant: C-modulation pairs (2k+1 modulates 2k) exercise the lock-step bins
    for i in range(cmod_pairs):
        setup.append(("cmod_set", 2 * i, 2 * i + 1, 0.5))
    return {"name": "config3_korg_cz_filter", "voices": V, "tables": {}, "setup": setup,
            "events": {}, "frames": int(seconds * SR)}


def config4(V=4096, seconds=30.0, rate_hz=8.0, seed=0x5EED0002):
    setup = []
    for v in range(V):
        setup += pcm_voice(v, V)
    timed = []
    rng = XorShift64(seed)
    for v in range(V):
        t = 0.0
        held = False
        while True:
            t += -math.log(1.0 - rng.uniform()) / rate_hz
            if t >= seconds:
                break
            if v % 4 == 0:
                timed.append((int(t * SR), ("envelope_velocity", v, 0.0 if held else 1.0)))
                held = not held
            else:
                timed.append((int(t * SR), ("voice_trigger", v)))
    for v in range(V):
        setup.append(("voice_trigger", v))        # one-shots are silent until triggered
    timed.sort(key=lambda x: x[0])
    return {"name": "config4_pcm_retrigger", "voices": V, "tables": {}, "setup": setup,
            "timed": timed, "events": bucket(timed), "frames": int(seconds * SR)}


def config5(V=65536, seconds=600.0, luts=None, event_seconds=None, seed=0x5EED0005, stationary=False):
    """Mixed load: v%3 -> config-2 / config-3 / config-4 recipe, amplitudes 40/V,
    one (re)trigger per voice per 10 s.

    stationary=True starts the one-shot PCM third in its long-run state instead of all
    triggered at t = 0: voice v was last triggered tau_v ~ U(0, 10 s) ago, so it is either
    somewhere inside its sample (phase = tau * inc, written to voice_phase[] by `install`) or
    already finished, and its next trigger comes at 10 s - tau_v.  A short benchmark window
    then sees the same mix of rendering and skipped voices as minute 5 of the 10 min job."""
    tables = {}
    if luts is not None:
        tables = {200: (luts["sine_lutable_0"], {}), 201: (luts["triangle_lutable_0"], {}),
                  202: (luts["impulse_lutable_0"], {})}
    setup = []
    for v in range(V):
        k = v % 3
        u = v // 3
        if k == 0:
            c = lut_voice(u, V, luts is not None)
        elif k == 1:
            c = korg_voice(u, V)
        else:
            c = pcm_voice(u, V)
        # the recipes are written for voice index u: retarget to v, keep amp = 40/V
        setup += [(x[0], v) + tuple(x[2:]) for x in c]
    timed = []
    horizon = seconds if event_seconds is None else min(seconds, event_seconds)
    rng = XorShift64(seed)
    pcm_tau = {}
    for v in range(V):
        t = 10.0 * rng.uniform()
        if v % 3 == 2 and stationary:
            pcm_tau[v] = int(t * SR)               # frames since the last trigger
            t = 10.0 - t
        while t < horizon:
            if v % 3 == 2:
                timed.append((int(t * SR), ("voice_trigger", v)))
            else:
                timed.append((int(t * SR), ("envelope_velocity", v, 1.0)))
            t += 10.0
        if v % 3 == 2 and not stationary:
            setup.append(("voice_trigger", v))     # one-shots are silent until triggered
    timed.sort(key=lambda x: x[0])
    return {"name": "config5_mixed_%d" % V, "voices": V, "tables": tables, "setup": setup,
            "timed": timed, "events": bucket(timed), "frames": int(seconds * SR), "pcm_tau": pcm_tau}


def shard(wl, v0, n):
    """The sub-workload of voices [v0, v0 + n), re-indexed from 0 (an independent instance of
    the reference with VOICE_MAX = n renders it; valid when no modulation edge leaves the range)."""
    def mv(c):
        return (c[0], c[1] - v0) + tuple(c[2:])
    inside = lambda c: v0 <= c[1] < v0 + n      # noqa: E731
    out = dict(wl)
    out["voices"] = n
    out["setup"] = [mv(c) for c in wl["setup"] if inside(c)]
    out["timed"] = [(t, mv(c)) for t, c in wl.get("timed", []) if inside(c)]
    out["events"] = {k: [mv(c) for c in calls if inside(c)] for k, calls in wl["events"].items()}
    out["events"] = {k: c for k, c in out["events"].items() if c}
    out["pcm_tau"] = {v - v0: t for v, t in wl.get("pcm_tau", {}).items() if v0 <= v < v0 + n}
    return out


def install(api, wl):
    """Apply tables + setup of a workload to a SynthAPI."""
    from .host import install_table
    for slot, (data, kw) in wl["tables"].items():
        install_table(api, slot, data, **kw)
    api.apply(wl["setup"])
    if wl.get("pcm_tau"):
        import ctypes as C
        import numpy as np
        inc = api.array("voice_phase_inc")
        size = api.array("voice_table_size", C.c_int)
        for v, tau in wl["pcm_tau"].items():
            ph = np.float32(tau) * np.float32(inc[v])
            if ph < np.float32(size[v] - 1):
                api.set_phase(v, float(ph), 0)     # mid-sample
            # else: finished (wave_set of a one-shot leaves it finished, synth.c:281-282)
        api.commit_state()
