"""skred_b200 — B200-native implementation of skred's per-voice render loop.

The product is native code: ``libskred_b200.so`` (CUDA engine, sm_100a) and
``libskred_shim_v<VOICE_MAX>.so`` (skred's synth.h / synth.def API in C on top of
it).  This package is only the thin ctypes glue tests and ``bench.py`` use to
drive those libraries; it contains no render code and has no CPU fallback —
loading fails loudly when the CUDA library is missing.
"""
from .host import (SynthAPI, Skred, load_engine_lib, engine_lib_path, shim_lib_path,  # noqa: F401
                   NativeLibraryMissing)

__all__ = ["SynthAPI", "Skred", "load_engine_lib", "engine_lib_path", "shim_lib_path",
           "NativeLibraryMissing"]
