"""ctypes loaders for the oracle libraries.  TEST INFRASTRUCTURE (oracle/README.md).

  RefSkred(V)            oracle/_ref/libskred_ref_v<V>.so — the reference's own synth.c / wire.c /
                         seq.c / skode.c compiled with the pinned flags
  PortSkred(V)           oracle/_build/libskred_dropin_port_v<V>.so — reference wire/seq/skode +
                         the product's host shim rendering through the CPU restatement
                         (oracle/skred_port.c); pins shim + port against the reference
  DropinCuda(V)          oracle/_ref/libskred_dropin_cuda_v<V>.so — reference wire/seq/skode on
                         top of the PRODUCT (shim + CUDA engine): the drop-in proof

All three share oracle/ref_harness.c, so they are driven identically: wire()
lines or synth.h setter calls, then `synth(); seq();` callbacks of 512 frames.
"""
import ctypes as C
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from skred_b200.host import SynthAPI, private_copy, load_engine_lib, BLOCK  # noqa: E402


def ref_lib_path(v):
    return os.path.join(HERE, "_ref", "libskred_ref_v%d.so" % v)


def port_lib_path(v):
    return os.path.join(HERE, "_build", "libskred_dropin_port_v%d.so" % v)


def cuda_dropin_path(v):
    return os.path.join(HERE, "_ref", "libskred_dropin_cuda_v%d.so" % v)


class HarnessSkred(SynthAPI):
    def __init__(self, path, voice_max, run_seq=True):
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run `python oracle/build_oracle.py --voices %d`" % (path, voice_max))
        copy = private_copy(path)
        lib = C.CDLL(copy)
        # the mapping outlives the file: without this every instance of a sweep leaves a 3 MB copy in /tmp
        shutil.rmtree(os.path.dirname(copy), ignore_errors=True)
        super().__init__(lib, voice_max)
        assert lib.ref_voice_max() == voice_max
        lib.ref_render.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int]
        lib.ref_render.restype = C.c_double
        lib.ref_wire.argtypes = [C.c_char_p]
        lib.ref_sample_count.restype = C.c_uint64
        lib.ref_get_filter.argtypes = [C.c_int, C.c_void_p]
        lib.ref_get_envelope.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.run_seq = 1 if run_seq else 0
        self.backend = "ref" if "libskred_ref" in path else ("port" if "dropin_port" in path else "cuda")
        self.cpu_seconds = 0.0
        lib.ref_init()

    def wire(self, line):
        if isinstance(line, str):
            line = line.encode()
        return self.lib.ref_wire(line)

    def load_lines(self, lines):
        """What sk_load does (wire.c:342-368): one wire() per line, shared context."""
        for ln in lines:
            self.wire(ln)

    def _synth(self, out, nframes):
        self.cpu_seconds += self.lib.ref_render(out.ctypes.data, nframes, nframes, self.run_seq)

    def sync_state(self):
        self.lib.ref_sync_state()

    def enable_tap(self, frames=BLOCK):
        """Per-voice tap `user` of synth() (synth.c:503-511, 533-611) sized for callbacks of up to
        `frames` frames; call before the first render.  Returns a [frames][voice][L,R] view that every
        callback overwrites from frame 0."""
        self.lib.ref_enable_tap.restype = C.c_void_p
        self.lib.ref_enable_tap.argtypes = [C.c_int]
        p = self.lib.ref_enable_tap(frames)
        buf = (C.c_float * (frames * self.voice_max * 2)).from_address(p)
        self._tap = np.ctypeslib.as_array(buf).reshape(frames, self.voice_max, 2)
        return self._tap

    def render_with_tap(self, nframes, block=BLOCK, events=None):
        """render(), collecting the tap of every callback: (out[nframes][2], tap[nframes][voice][2])."""
        out = np.zeros((nframes, 2), dtype=np.float32)
        taps = np.zeros((nframes, self.voice_max, 2), dtype=np.float32)
        done, k = 0, 0
        while done < nframes:
            n = min(block, nframes - done)
            if events and k in events:
                self.apply(events[k])
            self._synth(out[done:done + n], n)
            taps[done:done + n] = self._tap[:n]
            done += n
            k += 1
        return out, taps

    def record_init(self, max_sec=1.0, block=BLOCK):
        """skred.c's synth_callback_init + the per-voice tap: call before the first render."""
        self.lib.ref_record_init.argtypes = [C.c_float, C.c_int]
        self.lib.ref_enable_tap.restype = C.c_void_p
        self.lib.ref_enable_tap.argtypes = [C.c_int]
        self.lib.ref_record_init(max_sec, block)
        self.lib.ref_render_recording.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int]
        self.lib.ref_render_recording.restype = C.c_double
        self.lib.ref_rec_ptr.restype = C.c_long

    def render_recording(self, nframes, block=BLOCK):
        """render() through the body of synth_callback (skred.c:116-131): while `<sec` is recording, every callback's
        per-voice tap is appended to the recording buffer that `*` writes out with save_wav (wire.c:94-185)."""
        out = np.zeros((nframes, 2), dtype=np.float32)
        self.lib.ref_render_recording(out.ctypes.data, nframes, block, self.run_seq)
        return out

    def render_batched(self, nframes, call_frames=8192):
        """The job of render(nframes) with the sequencer running (seq() after every 512-frame callback), handed to the
        drop-in in calls of `call_frames` frames through skb_shim_synth_between: seq() is walked ahead of the audio."""
        out = np.zeros((nframes, 2), dtype=np.float32)
        self.lib.ref_render_batched.argtypes = [C.c_void_p, C.c_long, C.c_int]
        self.lib.ref_render_batched.restype = C.c_double
        self.lib.ref_render_batched(out.ctypes.data, nframes, call_frames)
        return out

    def engine_stats(self):
        """skb_stats of the engine behind a drop-in build (port: linked in; cuda: libskred_b200.so)."""
        from skred_b200.host import skb_stats
        self.lib.skb_shim_engine.restype = C.c_void_p
        eng = self.lib.skb_shim_engine()
        owner = self.lib if self.backend == "port" else load_engine_lib()
        owner.skb_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        st = skb_stats()
        owner.skb_get_stats(eng, C.byref(st))
        return st

    def state(self):
        """Evolving per-voice state as a dict of arrays."""
        self.sync_state()
        n = self.voice_max
        filt = np.zeros((n, 9), dtype=np.float32)
        envf = np.zeros((n, 9), dtype=np.float32)
        envu = np.zeros((n, 2), dtype=np.uint64)
        act = np.zeros(n, dtype=np.int32)
        for v in range(n):
            self.lib.ref_get_filter(v, filt[v].ctypes.data)
            self.lib.ref_get_envelope(v, envf[v].ctypes.data, envu[v].ctypes.data, act[v:v + 1].ctypes.data)
        return {
            "phase": self.array("voice_phase").copy(),
            "finished": self.array("voice_finished", C.c_int).copy(),
            "sample": self.array("voice_sample").copy(),
            "sh_hold": self.array("voice_sample_hold").copy(),
            "sh_count": self.array("voice_sample_hold_count", C.c_int).copy(),
            "filter_xy": filt[:, :4].copy(),
            "env_active": act,
            "env_start": envu[:, 0].copy(),
            "env_release": envu[:, 1].copy(),
            "smoother_gain": self.array("voice_smoother_gain").copy(),
            "pan_left": self.array("voice_pan_left").copy(),
            "pan_right": self.array("voice_pan_right").copy(),
        }


def have_ref(v=64):
    return os.path.exists(ref_lib_path(v))


def RefSkred(v=64, run_seq=True):
    return HarnessSkred(ref_lib_path(v), v, run_seq)


def PortSkred(v=64, run_seq=True):
    return HarnessSkred(port_lib_path(v), v, run_seq)


def DropinCuda(v=64, run_seq=True):
    load_engine_lib()           # RTLD_GLOBAL: satisfies the private copy's DT_NEEDED
    return HarnessSkred(cuda_dropin_path(v), v, run_seq)
