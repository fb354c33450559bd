#!/usr/bin/env python3
"""Build the product libraries in-tree (they are git-ignored but travel to the GPU box).

  python -m skred_b200.build [--voices 64,1024,...] [--force]

  skred_b200/libskred_b200.so         the CUDA engine (csrc/engine.cu + voice_kernels.cuh),
                                      sm_100a only, parity flags
  skred_b200/libskred_shim_v<V>.so    the drop-in host library: skred's synth.h / synth.def API
                                      (csrc/synth_shim.c) for VOICE_MAX = V, linked against the
                                      engine.  It is compiled AGAINST a skred source tree
                                      ($SKRED_SRC, default /root/reference) the way any skred
                                      translation unit is: skred.h / synth.h / synth.def /
                                      retro/korg.h / amysamples.c supply the API and the table
                                      data; nothing is copied.  Without a skred tree (the GPU
                                      box) the prebuilt file is used.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
GEN = os.path.join(ROOT, "build", "gen")
SKRED_SRC = os.environ.get("SKRED_SRC", os.environ.get("SKRED_REF", "/root/reference"))

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# Every float op individually rounded, IEEE div/sqrt, denormals kept: the same
# arithmetic `gcc -O2 -ffp-contract=off` gives the reference on x86-64.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
              "-Xlinker", "-soname=libskred_b200.so", "-ldl"]
HOST_CFLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-fno-strict-aliasing", "-g1"]
DEFAULT_VOICES = [64, 1024, 4096, 65536]

ENGINE_SO = os.path.join(HERE, "libskred_b200.so")


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("command failed: %s" % cmd[0])
    return r.stdout


def newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def shim_so(v):
    return os.path.join(HERE, "libskred_shim_v%d.so" % v)


def build_engine(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "free_lo.cu", "voice_kernels.cuh", "free_kernel.cuh", "row_kernel.cuh", "level_kernel.cuh", "partition.h")] + [
        os.path.join(INC, "skred_b200.h"), __file__]
    if not force and newer(ENGINE_SO, srcs):
        return ENGINE_SO
    out = run([NVCC] + NVCC_FLAGS + ["-I" + INC, "-I" + CSRC, srcs[0], srcs[1], "-o", ENGINE_SO])
    log = os.path.join(ROOT, "build", "ptxas_engine.log")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    open(log, "w").write(out)
    if verbose:
        print(out)
    return ENGINE_SO


FAST_SO = os.path.join(HERE, "fast", "libskred_b200.so")


def build_engine_fast(force=False):
    """The NON-PARITY build: FMA contraction on, interpolating oscillator read.  Same soname, its own
    directory; selected with SKB_ENGINE_LIB (bench.py's fast_mode leg).  No parity test ever loads it."""
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "free_lo.cu", "voice_kernels.cuh", "free_kernel.cuh", "row_kernel.cuh", "level_kernel.cuh", "partition.h")] + [
        os.path.join(INC, "skred_b200.h"), __file__]
    if not force and newer(FAST_SO, srcs):
        return FAST_SO
    os.makedirs(os.path.dirname(FAST_SO), exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-fmad=false"] + ["-fmad=true", "-DSKB_FAST_MODE=1"]
    run([NVCC] + flags + ["-I" + INC, "-I" + CSRC, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "free_lo.cu"), "-o", FAST_SO])
    return FAST_SO


CANARY_SO = os.path.join(HERE, "variants", "canary", "libskred_b200.so")


def build_engine_canary(force=False):
    """The self-check build (-DSKB_CANARY=1, voice_kernels.cuh): the modulation-group kernels tag every exchanged voice_sample
    with its frame and count reads that see another frame than the index rule promises (tests: test_exchange_canary)."""
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "free_lo.cu", "voice_kernels.cuh", "free_kernel.cuh", "row_kernel.cuh", "level_kernel.cuh", "partition.h")] + [
        os.path.join(INC, "skred_b200.h"), __file__]
    if not force and newer(CANARY_SO, srcs):
        return CANARY_SO
    return build_engine_variant("canary", ["-DSKB_CANARY=1"])


def build_engine_variant(name, defines):
    """A tuning build of the engine (same soname) under skred_b200/variants/<name>/:
    select it with SKB_ENGINE_LIB=<path>.  defines: e.g. ["-DSKB_SUB=4", "-DSKB_CTA_WARPS=12"]."""
    d = os.path.join(HERE, "variants", name)
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, "libskred_b200.so")
    run([NVCC] + NVCC_FLAGS + list(defines) + ["-I" + INC, "-I" + CSRC, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "free_lo.cu"), "-o", out])
    return out


def have_skred_src():
    return os.path.exists(os.path.join(SKRED_SRC, "synth.def"))


def _gen_skred_h(v):
    """VOICE_MAX is a hard #define in skred.h:9: a generated,
    force-included copy of that one header carries the override."""
    d = os.path.join(GEN, "v%d" % v)
    os.makedirs(d, exist_ok=True)
    src = open(os.path.join(SKRED_SRC, "skred.h")).read()
    out, n = re.subn(r"#define\s+VOICE_MAX\s+\(\d+\)", "#define VOICE_MAX (%d)" % v, src)
    assert n == 1
    p = os.path.join(d, "skred.h")
    if not os.path.exists(p) or open(p).read() != out:
        open(p, "w").write(out)
    return p


def _amysamples_o(v):
    """The AMY PCM map + blob of the skred tree (amysamples.c).  The 1.17 M-sample
    pcm[] header is missing from this checkout; tools/gen_pcm_stub.py
    writes the seeded synthetic stand-in both sides of every parity test share."""
    hdr = os.path.join(GEN, "pcm_samples_large.h")
    if not os.path.exists(hdr):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import gen_pcm_stub
        gen_pcm_stub.main(SKRED_SRC, GEN)
    amy_o = os.path.join(GEN, "amysamples.o")
    if not os.path.exists(amy_o):
        run(["gcc"] + HOST_CFLAGS + ["-include", _gen_skred_h(v), "-I" + GEN, "-I" + SKRED_SRC, "-w",
                                     "-c", os.path.join(SKRED_SRC, "amysamples.c"), "-o", amy_o])
    return amy_o


def build_shim(v, force=False):
    out = shim_so(v)
    if not have_skred_src():
        return out if os.path.exists(out) else None
    srcs = [os.path.join(CSRC, "synth_shim.c"), os.path.join(CSRC, "shim_host.c")]
    deps = srcs + [os.path.join(INC, "skred_b200.h"), os.path.join(INC, "skred_b200_shim.h"), ENGINE_SO, __file__]
    if not force and newer(out, deps):
        return out
    amy_o = _amysamples_o(v)
    run(["gcc"] + HOST_CFLAGS + ["-shared", "-include", _gen_skred_h(v), "-I" + GEN, "-I" + SKRED_SRC,
                                 "-I" + INC, "-Wno-multichar", "-w"] + srcs + [amy_o, "-o", out,
                                 "-L" + HERE, "-lskred_b200", "-Wl,-rpath,$ORIGIN", "-lm"])
    return out


def build_all(voices=None, force=False, verbose=False):
    # the three engine builds (parity, non-parity fast, exchange canary) are independent nvcc runs of ~1.5 min each: side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=3) as pool:
        f_main = pool.submit(build_engine, force, verbose)
        f_fast = pool.submit(build_engine_fast, force)
        f_can = pool.submit(build_engine_canary, force)
        outs = [f_main.result()]
        f_fast.result()
        f_can.result()
    for v in voices or DEFAULT_VOICES:
        outs.append(build_shim(v, force))
    return outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voices", default=",".join(map(str, DEFAULT_VOICES)))
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    for o in build_all([int(x) for x in a.voices.split(",") if x], a.force, a.verbose):
        print("built ->", o)


if __name__ == "__main__":
    main()
