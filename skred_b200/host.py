"""ctypes glue over the drop-in library (skred's synth.h API, synth.h:8-85).

`SynthAPI` binds the reference's setter names on ANY library that exports them
(the product shim, or — in tests only — the compiled reference), so the same
list of calls drives both sides of a parity test.  `Skred` is the product:
``libskred_shim_v<V>.so`` over the CUDA engine.
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BLOCK = 512          # SYNTH_FRAMES_PER_CALLBACK, skred.h:12 (event granularity)
SAMPLE_RATE = 44100  # MAIN_SAMPLE_RATE, skred.h:6


class NativeLibraryMissing(RuntimeError):
    pass


def engine_lib_path():
    # SKB_ENGINE_LIB: an alternative build of the same engine (tuning experiments, tools/)
    return os.environ.get("SKB_ENGINE_LIB") or os.path.join(HERE, "libskred_b200.so")


def shim_lib_path(voice_max):
    return os.path.join(HERE, "libskred_shim_v%d.so" % voice_max)


_engine_lib = None


def load_engine_lib():
    """dlopen the CUDA engine (RTLD_GLOBAL so the shim's DT_NEEDED resolves to it)."""
    global _engine_lib
    if _engine_lib is None:
        p = engine_lib_path()
        if not os.path.exists(p):
            raise NativeLibraryMissing(
                "%s not built: run `python -m skred_b200.build` (there is no CPU fallback)" % p)
        _engine_lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
        _engine_lib.skb_backend_name.restype = C.c_char_p
        _engine_lib.skb_error_string.restype = C.c_char_p
        _engine_lib.skb_error_string.argtypes = [C.c_void_p]
        _engine_lib.skb_last_error.argtypes = [C.c_void_p]
        _engine_lib.skb_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        _engine_lib.skb_sync.argtypes = [C.c_void_p, C.c_void_p]
    return _engine_lib


def private_copy(path):
    """Every drop-in library keeps its state in globals (the reference's design,
    synth.def).  A second, independent instance in one process = a private copy
    of the .so under a fresh name."""
    d = tempfile.mkdtemp(prefix="skb_")
    q = os.path.join(d, os.path.basename(path))
    shutil.copy(path, q)
    return q


i32, f32 = C.c_int, C.c_float

# name -> argtypes (all return int unless noted); synth.h:24-85
SETTERS = {
    "amp_set": (i32, f32), "pan_set": (i32, f32), "freq_set": (i32, f32), "freq_midi": (i32, f32),
    "wave_set": (i32, i32), "wave_mute": (i32, i32), "wave_dir": (i32, i32), "wave_loop": (i32, i32),
    "wave_quant": (i32, i32), "wave_reset": (i32, i32), "wave_default": (i32,),
    "cz_set": (i32, i32, f32), "cmod_set": (i32, i32, f32),
    "amp_mod_set": (i32, i32, f32), "freq_mod_set": (i32, i32, f32), "pan_mod_set": (i32, i32, f32),
    "envelope_set": (i32, f32, f32, f32, f32), "envelope_velocity": (i32, f32),
    "voice_trigger": (i32,), "voice_copy": (i32, i32),
    "mmf_set_freq": (i32, f32), "mmf_set_res": (i32, f32), "mmf_set_params": (i32, f32, f32),
    "volume_set": (f32,),
}


class skb_stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("frames_rendered", C.c_uint64),
                ("ops_applied", C.c_uint64), ("params_uploaded", C.c_uint64), ("replans", C.c_uint64),
                ("n_free_voices", C.c_int32), ("n_group_voices", C.c_int32), ("n_groups", C.c_int32),
                ("n_owned_voices", C.c_int32), ("last_render_ms", C.c_float), ("_pad", C.c_int32),
                ("active_voice_frames", C.c_uint64), ("class_rows", C.c_uint64 * 8),
                ("phase_cycles", C.c_uint64 * 8), ("cta_batches", C.c_uint64),
                ("wide_launches", C.c_uint64), ("wide_errors", C.c_uint64), ("last_wide_ms", C.c_float * 3),
                ("_pad2", C.c_int32), ("host_us", C.c_double * 4),
                ("rows_launches", C.c_uint64), ("migrated_voices", C.c_uint64), ("lo_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class SynthAPI:
    """The reference's public voice/parameter API on a loaded library."""

    def __init__(self, lib, voice_max):
        self.lib = lib
        self.voice_max = voice_max
        for name, at in SETTERS.items():
            fn = getattr(lib, name)
            fn.argtypes = list(at)
            fn.restype = None if name == "mmf_set_params" else C.c_int
        lib.synth.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.synth.restype = None
        self._mark = getattr(lib, "skb_shim_mark_dirty", None)
        if self._mark is not None:
            self._mark.argtypes = [C.c_int]
            self._mark.restype = None

    # ---- direct views of the exported state arrays (synth.def) -------------
    def array(self, name, ctype=C.c_float, n=None):
        n = self.voice_max if n is None else n
        a = (ctype * n).in_dll(self.lib, name)
        return np.ctypeslib.as_array(a)

    def scalar(self, name, ctype):
        return ctype.in_dll(self.lib, name)

    @property
    def sample_count(self):
        return C.c_uint64.in_dll(self.lib, "synth_sample_count").value

    # ---- the atoms wire.c implements by writing arrays directly ------------
    def filter_mode(self, v, mode):            # `J`, wire.c:666-672
        self.array("voice_filter_mode", C.c_int)[v] = mode
        self.lib.mmf_set_params(v, float(self.array("voice_filter_freq")[v]), float(self.array("voice_filter_res")[v]))

    def hold(self, v, n):                      # `h`, wire.c:653
        self.array("voice_sample_hold_max", C.c_int)[v] = n
        if self._mark:
            self._mark(v)

    def smoother(self, v, k):                  # `s`, wire.c:699-707
        if k <= 0:
            self.array("voice_smoother_enable", C.c_int)[v] = 0
        else:
            self.array("voice_smoother_enable", C.c_int)[v] = 1
            self.array("voice_smoother_smoothing")[v] = k
        if self._mark:
            self._mark(v)

    # ---- evolving state written directly (checkpoint resume / stationary workloads) ----
    def set_phase(self, v, phase, finished):
        self.array("voice_phase")[v] = phase
        self.array("voice_finished", C.c_int)[v] = finished
        self._state_dirty = True

    def commit_state(self):
        """On the reference the arrays ARE the state; the drop-in pushes them to the device."""
        if getattr(self, "_state_dirty", False) and hasattr(self.lib, "skb_shim_engine"):
            self.lib.skb_shim_restore_range.argtypes = [C.c_int, C.c_int]
            r = self.lib.skb_shim_restore_range(0, self.voice_max)
            if r != 0:
                raise RuntimeError("skb_shim_restore_range failed: %d" % r)
        self._state_dirty = False

    def call(self, name, *args):
        if name in SETTERS:
            return getattr(self.lib, name)(*args)
        return getattr(self, name)(*args)

    def apply(self, calls):
        """calls: iterable of (name, *args) — setter names of synth.h or the
        three direct-write atoms above."""
        for c in calls:
            self.call(c[0], *c[1:])

    # ---- rendering -----------------------------------------------------------
    def _synth(self, out, nframes):
        self.lib.synth(out.ctypes.data, None, nframes, 2, None)

    def render(self, nframes, block=BLOCK, events=None, out=None):
        """Render `nframes` frames as callbacks of `block` frames.  `events` maps a
        callback index k to a list of calls applied BEFORE callback k — where the
        reference's seq() would fire them (after callback k-1, seq.c:170-178)."""
        if out is None:
            out = np.zeros((nframes, 2), dtype=np.float32)
        done, k = 0, 0
        while done < nframes:
            n = min(block, nframes - done)
            if events and k in events:
                self.apply(events[k])
            self._synth(out[done:done + n], n)
            done += n
            k += 1
        return out


class Skred(SynthAPI):
    """The product: skred's synth.h API on the B200 engine.  One instance per
    process per VOICE_MAX unless `private=True` (independent globals)."""

    def __init__(self, voice_max=64, device=None, rank=0, world=1, max_frames=8192, private=False):
        load_engine_lib()
        p = shim_lib_path(voice_max)
        if not os.path.exists(p):
            raise NativeLibraryMissing(
                "%s not built: run `python -m skred_b200.build --voices %d`" % (p, voice_max))
        lib = C.CDLL(private_copy(p) if private else p)
        super().__init__(lib, voice_max)
        lib.skb_shim_configure.argtypes = [C.c_int] * 4
        lib.skb_shim_engine.restype = C.c_void_p
        lib.skb_shim_render_mix.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        lib.skb_shim_finish.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        if device is None:
            device = int(os.environ.get("SKB_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        r = lib.skb_shim_configure(device, rank, world, max_frames)
        if r != 0:
            raise RuntimeError("skb_shim_configure failed: %d" % r)
        # same order as main(): skred.c:232-236
        lib.synth_init()
        lib.wave_table_init()
        lib.voice_init()
        self.engine = lib.skb_shim_engine()    # aborts the process if no B200 (no fallback)

    def stats(self):
        s = skb_stats()
        load_engine_lib().skb_get_stats(self.engine, C.byref(s))
        return s

    def flush(self):
        r = self.lib.skb_shim_flush()
        if r != 0:
            raise RuntimeError("engine error %d: %s" % (r, load_engine_lib().skb_error_string(self.engine).decode()))

    def snapshot(self):
        self.lib.skb_shim_snapshot()

    def render_mix(self, nframes, d_mix_ptr, stream=None):
        r = self.lib.skb_shim_render_mix(nframes, d_mix_ptr, stream)
        if r != 0:
            raise RuntimeError("render_mix failed %d: %s" % (r, load_engine_lib().skb_error_string(self.engine).decode()))

    def finish(self, d_mix_ptr, nframes, out, stream=None):
        r = self.lib.skb_shim_finish(d_mix_ptr, nframes, out.ctypes.data, 2, stream)
        if r != 0:
            raise RuntimeError("finish failed %d" % r)

    def install_table(self, slot, data, rate=44100.0, one_shot=0, loop_start=0, loop_end=None,
                      midi_note=69.0, offset_hz=440.0):
        """Put a caller-owned float table into a user wave slot the way `data_load`
        does (wire.c:374-404); the array must stay alive."""
        return install_table(self, slot, data, rate, one_shot, loop_start, loop_end, midi_note, offset_hz)


def install_table(api, slot, data, rate=44100.0, one_shot=0, loop_start=0, loop_end=None,
                  midi_note=69.0, offset_hz=440.0):
    WAVE_TABLE_MAX = 1200
    data = np.ascontiguousarray(data, dtype=np.float32)
    if not hasattr(api, "_tables"):
        api._tables = []
    api._tables.append(data)
    n = len(data)
    api.array("wave_table_data", C.c_uint64, WAVE_TABLE_MAX)[slot] = data.ctypes.data
    api.array("wave_size", C.c_int, WAVE_TABLE_MAX)[slot] = n
    api.array("wave_rate", C.c_float, WAVE_TABLE_MAX)[slot] = rate
    api.array("wave_one_shot", C.c_int, WAVE_TABLE_MAX)[slot] = one_shot
    api.array("wave_loop_enabled", C.c_int, WAVE_TABLE_MAX)[slot] = 0
    api.array("wave_loop_start", C.c_int, WAVE_TABLE_MAX)[slot] = loop_start
    api.array("wave_loop_end", C.c_int, WAVE_TABLE_MAX)[slot] = (n - 1) if loop_end is None else loop_end
    api.array("wave_midi_note", C.c_float, WAVE_TABLE_MAX)[slot] = midi_note
    api.array("wave_offset_hz", C.c_float, WAVE_TABLE_MAX)[slot] = offset_hz
    return 0
