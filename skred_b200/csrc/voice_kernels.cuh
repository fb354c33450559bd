/* voice_kernels.cuh — sm_100a device code of the voice-render path.
 *
 * Replaces the frame-outer / voice-inner loop of synth() (synth.c:520-613).
 * Exactness contract: this TU is compiled with
 *   -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
 * so every float op below is ONE individually rounded IEEE-754 binary32 op in
 * the reference's evaluation order — the same thing `gcc -O2
 * -ffp-contract=off` emits for synth.c on x86-64/SSE2.  The phase
 * recurrence, table index and `finished` latch are therefore bit-identical
 * to the reference; only the ORDER OF THE CROSS-VOICE SUM differs (fixed,
 * documented in DESIGN.md), which is inside the 1e-5 float budget.
 *
 * Layout.  Parameters and evolving state live in HBM as float4-packed
 * structure-of-arrays indexed by SLOT (q[k][slot]); a warp's 32 threads read
 * 32 consecutive float4 = 512 contiguous bytes per field group.  Slots are a
 * permutation of voices chosen by the host planner (engine.cu): free voices
 * first, sorted so that a warp is homogeneous in CZ mode / filter / table,
 * then modulation groups.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>
#include "skred_b200.h"

#define SKB_NPQ 9          /* float4 groups per voice: parameters (32 words + the words of a levelled modulated voice) */
#define SKB_NSQ 5          /* float4 groups per voice: evolving state (18 of 20 words) */
#define SKB_REF_NONE (-1)      /* no modulator: the reference's literal for that site (1.0f / no FM) */
#define SKB_REF_SELF (-2)      /* the voice reads its own voice_sample */
#define SKB_REF_ZERO (-3)      /* modulator contributes an identical 0.0f (depth-0 CZ default, out-of-range osc) */
#define SKB_REF_CUR  (1 << 20)   /* read the modulator's CURRENT-frame value (m < n) */
#define SKB_REF_TRACE (1 << 21)  /* the modulator is rendered by an EARLIER launch: read its per-frame trace (row = ref & MASK) */
#define SKB_REF_MASK ((1 << 20) - 1)

struct VoiceP {
  float amp, inc, fscale, fm_depth;
  int fm_ref, toff, tsize; unsigned flags;
  float lo, hi, cz_dist, cz_depth;
  int cz_mode, cz_ref, sh_max, quant;
  float b0, b1, b2, a1;
  float a2, envA, envD, envS;
  float envR, am_depth, sm_k, pm_depth;
  int fmode, am_ref, pm_ref, level;
  float fm_minc;             /* levelled voice: voice_phase_inc of its FM modulator (synth.c:553: inc[mod]) */
  int trace_out;             /* levelled voice: trace row its voice_sample goes to (somebody reads it), or -1 */
};

struct VoiceS {
  float phase; int finished; float sample, sh_hold;
  int sh_count; float x1, x2, y1;
  float y2; int env_active; float env_vel, sm_gain;
  float panL, panR; unsigned long long env_start;
  unsigned long long env_rel;
  int aux;                   /* spare word of the record: 0 in the state proper; in a window snapshot of a time-split
                                launch bit 0 = "mmf_init ran at this boundary" (free_kernel.cuh) */
  int nact;                  /* frames actually rendered (not skipped) in this launch; not stored */
};

__device__ __forceinline__ float4 ldq(const float4 *base, int k, int cap, int slot) {
  return base[(size_t)k * cap + slot];
}

__device__ __forceinline__ void load_params(const float4 *__restrict__ q, int cap, int slot, VoiceP &p) {
  float4 a;
  a = ldq(q, 0, cap, slot); p.amp = a.x; p.inc = a.y; p.fscale = a.z; p.fm_depth = a.w;
  a = ldq(q, 1, cap, slot); p.fm_ref = __float_as_int(a.x); p.toff = __float_as_int(a.y); p.tsize = __float_as_int(a.z); p.flags = __float_as_uint(a.w);
  a = ldq(q, 2, cap, slot); p.lo = a.x; p.hi = a.y; p.cz_dist = a.z; p.cz_depth = a.w;
  a = ldq(q, 3, cap, slot); p.cz_mode = __float_as_int(a.x); p.cz_ref = __float_as_int(a.y); p.sh_max = __float_as_int(a.z); p.quant = __float_as_int(a.w);
  a = ldq(q, 4, cap, slot); p.b0 = a.x; p.b1 = a.y; p.b2 = a.z; p.a1 = a.w;
  a = ldq(q, 5, cap, slot); p.a2 = a.x; p.envA = a.y; p.envD = a.z; p.envS = a.w;
  a = ldq(q, 6, cap, slot); p.envR = a.x; p.am_depth = a.y; p.sm_k = a.z; p.pm_depth = a.w;
  a = ldq(q, 7, cap, slot); p.fmode = __float_as_int(a.x); p.am_ref = __float_as_int(a.y); p.pm_ref = __float_as_int(a.z); p.level = __float_as_int(a.w);
  a = ldq(q, 8, cap, slot); p.fm_minc = a.x; p.trace_out = __float_as_int(a.y) - 1;      /* (stored + 1: a cleared record reads "none") */
}

__device__ __forceinline__ void load_state(const float4 *__restrict__ q, int cap, int slot, VoiceS &s) {
  float4 a;
  a = ldq(q, 0, cap, slot); s.phase = a.x; s.finished = __float_as_int(a.y); s.sample = a.z; s.sh_hold = a.w;
  a = ldq(q, 1, cap, slot); s.sh_count = __float_as_int(a.x); s.x1 = a.y; s.x2 = a.z; s.y1 = a.w;
  a = ldq(q, 2, cap, slot); s.y2 = a.x; s.env_active = __float_as_int(a.y); s.env_vel = a.z; s.sm_gain = a.w;
  a = ldq(q, 3, cap, slot); s.panL = a.x; s.panR = a.y;
  s.env_start = ((unsigned long long)__float_as_uint(a.w) << 32) | __float_as_uint(a.z);
  a = ldq(q, 4, cap, slot);
  s.env_rel = ((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x);
  s.aux = __float_as_int(a.z);
  s.nact = 0;
}

__device__ __forceinline__ void store_state(float4 *__restrict__ q, int cap, int slot, const VoiceS &s) {
  q[(size_t)0 * cap + slot] = make_float4(s.phase, __int_as_float(s.finished), s.sample, s.sh_hold);
  q[(size_t)1 * cap + slot] = make_float4(__int_as_float(s.sh_count), s.x1, s.x2, s.y1);
  q[(size_t)2 * cap + slot] = make_float4(s.y2, __int_as_float(s.env_active), s.env_vel, s.sm_gain);
  q[(size_t)3 * cap + slot] = make_float4(s.panL, s.panR, __uint_as_float((unsigned)(s.env_start & 0xffffffffull)),
                                          __uint_as_float((unsigned)(s.env_start >> 32)));
  q[(size_t)4 * cap + slot] = make_float4(__uint_as_float((unsigned)(s.env_rel & 0xffffffffull)),
                                          __uint_as_float((unsigned)(s.env_rel >> 32)), __int_as_float(s.aux), 0.0f);
}

/* C's (int)float on x86-64 is cvttss2si: anything unrepresentable (NaN, +-Inf,
 * |v| >= 2^31) yields INT_MIN, whereas cvt.rzi.s32.f32 saturates and maps NaN
 * to 0.  Needed where the reference can overflow: fast_pow (synth.c:145) and
 * the CZ-warped index (synth.c:265). */
__device__ __forceinline__ int c_f2i(float v) {
  return (v < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
}
__device__ __forceinline__ int c_d2i(double v) {
  return (v < 2147483648.0) ? __double2int_rz(v) : (int)0x80000000;
}

/* fast_pow, synth.c:140-147 */
__device__ __forceinline__ float dev_fast_pow(float a, float b) {
  /* straight-line: an early return for a <= 0 becomes a branch per frame, which ends the scheduling
   * region of the pipelined body; the discarded lanes compute on wrapped integers, harmlessly */
  const int ai = (int)((unsigned)__float_as_int(a) - 1065353216u);
  float t = b * __int2float_rn(ai);
  t = t + 1065353216.0f;
  const float r = __int_as_float(c_f2i(t));
  return (a <= 0.0f) ? 0.0f : r;
}

/* Per-block constants of one voice that the reference recomputes every sample
 * from unchanged inputs (identical bits, fewer instructions). */
struct VoiceK {
  float lo, hi, span, span2;  /* wrap window, synth.c:235-239 */
  float size_f, inv_size;     /* inv_size != 0 iff table_size is a power of two (x/2^k == x*2^-k exactly) */
  bool stop_at_end;           /* one_shot && !loop_enabled, synth.c:243,250 */
  int one_shot;
};

__device__ __forceinline__ void derive_consts(const VoiceP &p, VoiceK &k) {
  const bool loop_on = (p.flags & SKB_F_LOOP_ENABLED) != 0;
  const bool use_loop = loop_on && (p.flags & SKB_F_LOOP_VALID);
  k.size_f = __int2float_rn(p.tsize);
  k.lo = use_loop ? p.lo : 0.0f;
  k.hi = use_loop ? p.hi : k.size_f;
  k.span = k.hi - k.lo;
  k.span2 = k.span + k.span;
  k.one_shot = (p.flags & SKB_F_ONE_SHOT) ? 1 : 0;
  k.stop_at_end = k.one_shot && !loop_on;
  const bool pow2 = p.tsize > 0 && (p.tsize & (p.tsize - 1)) == 0;
  k.inv_size = pow2 ? (1.0f / k.size_f) : 0.0f;
}

/* fmodf(x, y) for x >= 0, y > 0.  fmodf is exact by definition; when
 * y <= x < 2y the result x - y is exact too (Sterbenz), so the common
 * single-wrap case needs one subtraction. */
__device__ __forceinline__ float wrap_mod(float x, float y, float y2) {
  if (x < y) return x;
  if (x < y2) return x - y;
  return fmodf(x, y);
}

/* cz_phasor, synth.c:149-215.  `d` already includes the modulator term. */
__device__ __forceinline__ float dev_cz_phasor(int mode, float ph_tab, float d, const VoiceK &k) {
  float x = (k.inv_size != 0.0f) ? ph_tab * k.inv_size : ph_tab / k.size_f;     /* :151 */
  d = (d < 0.0f) ? 0.0f : (d > 0.999f ? 0.999f : d);                              /* :154 */
  switch (mode) {
    case 1: {
      if (x < d) x = x * (0.5f / d);
      else x = 0.5f + (x - d) * (0.5f / (1.0f - d));
      break;
    }
    case 2: {
      const float sc = 0.5f / (0.5f - d * 0.5f);
      x = (x < 0.5f) ? x * sc : 1.0f - (1.0f - x) * sc;
      break;
    }
    case 3: {
      const float sc = 0.5f / (0.5f - d * 0.5f);
      x = (x < 0.5f) ? x * sc : 0.5f + (x - 0.5f) * sc;
      break;
    }
    case 4: x = fmodf(x * 2.0f, 1.0f); break;
    case 5: {
      const float hd = d * 0.5f;
      x = (x < 0.5f) ? x * (0.5f / (0.5f - hd)) : 0.5f + (x - 0.5f) * (0.5f / (0.5f + hd));
      break;
    }
    case 6: x = dev_fast_pow(x, 1.0f + 4.0f * d); break;
    case 7: x = dev_fast_pow(x, 1.0f + 8.0f * d); break;
    default: return ph_tab;
  }
  return x * k.size_f;                                                            /* :214 */
}

/* amp_envelope_step, synth.c:398-431 */
__device__ __forceinline__ float dev_env_step(const VoiceP &p, VoiceS &s, unsigned long long ssc) {
  if (!s.env_active) return 0.0f;
  const float t = __ull2float_rn(ssc - s.env_start);
  if (t < p.envA) return t / p.envA;
  if (t < p.envA + p.envD) {
    const float prog = (t - p.envA) / p.envD;
    return 1.0f - prog * (1.0f - p.envS);
  }
  if (s.env_rel == 0ull) return p.envS;
  const float tr = __ull2float_rn(ssc - s.env_rel);
  if (tr < p.envR) {
    const float prog = tr / p.envR;
    return p.envS * (1.0f - prog);
  }
  s.env_active = 0;
  return 0.0f;
}

/* quantize_bits_int, synth.c:341-345 (double-precision +0.5, x86 shift-count masking) */
__device__ __forceinline__ float dev_quantize(float v, int bits) {
  const int levels = (1 << (bits & 31)) - 1;
  const float lf = __int2float_rn(levels);
  const int iv = c_d2i((double)(v * lf) + 0.5);
  return __int2float_rn(iv) * (1.0f / lf);
}

/* Modulator access policy for voices without live cross-voice reads. */
struct NoMods {
  __device__ __forceinline__ float read(int) const { return 0.0f; }
  __device__ __forceinline__ float inc_of(int) const { return 0.0f; }
};

/* Modulators of a LEVELLED voice: every voice it reads was rendered by an earlier launch of the same step and left its
 * voice_sample of every frame in a trace row: trace[row][0] = the value before the launch, [1 + f] = after frame f.  A
 * modulator with a smaller voice index is read at the current frame, a larger one at the previous frame (synth.c:526). */
struct TraceMods {
  const float *trace; int stride; int frame; float minc;
  __device__ __forceinline__ float read(int ref) const {
    return trace[(size_t)(ref & SKB_REF_MASK) * stride + frame + ((ref & SKB_REF_CUR) ? 1 : 0)];
  }
  __device__ __forceinline__ float inc_of(int) const { return minc; }
};

template <bool MODS, class Mod>
__device__ __forceinline__ float mod_value(int ref, float self_value, const Mod &mod) {
  if (ref == SKB_REF_SELF) return self_value;
  if (ref == SKB_REF_ZERO) return 0.0f;
  return MODS ? mod.read(ref) : 0.0f;
}

/* One voice, one frame: the body of the `for n` loop, synth.c:526-612.
 * Returns the voice's stereo contribution (0,0 when skipped/disconnected).
 * `self_prev` semantics: a self-referencing CZ read sees last frame's final
 * sample (synth.c:264 runs before :570), a self AM read sees the filtered
 * sample (:586 after :577), a self pan read the final one (:599 after :593). */
template <bool MODS, class Mod>
__device__ __forceinline__ float2 voice_frame(const VoiceP &p, const VoiceK &k, VoiceS &s,
                                              unsigned long long ssc, float white,
                                              const float *__restrict__ tables, const Mod &mod) {
  if (s.finished) { s.sample = 0.0f; return make_float2(0.0f, 0.0f); }          /* :531-536 */
  if (p.amp == 0.0f) { s.sample = 0.0f; return make_float2(0.0f, 0.0f); }       /* :537-542 */
  s.nact++;
  float f;
  if (p.flags & SKB_F_NOISE) {                                                    /* :543-546 */
    f = white;
  } else {
    float inc = p.inc;
    if (MODS && p.fm_ref >= 0) {                                                  /* :548-555 (mod != n) */
      const float g = mod.read(p.fm_ref) * p.fm_depth;
      inc = p.inc + ((mod.inc_of(p.fm_ref) * p.fscale) * g);
    }
    /* osc_next, synth.c:217-275 */
    if (p.flags & SKB_F_REVERSE) inc = -inc;                                      /* :224 */
    float ph = s.phase + inc;                                                     /* :226 */
    if (!(fabsf(ph) < CUDART_INF_F)) {                                            /* :228-232 !isfinite */
      s.phase = 0.0f;
      s.finished = k.one_shot;
      f = 0.0f;
    } else {
      if (ph >= k.hi) {                                                           /* :242-248 */
        if (k.stop_at_end) { ph = k.hi - 1e-6f; s.finished = 1; }
        else ph = k.lo + wrap_mod(ph - k.lo, k.span, k.span2);
      } else if (ph < k.lo) {                                                     /* :249-256 */
        if (k.stop_at_end) { ph = k.lo; s.finished = 1; }
        else ph = k.hi - wrap_mod(k.lo - ph, k.span, k.span2);
      }
      s.phase = ph;                                                               /* :258 */
      int idx;
      if (p.cz_mode) {                                                            /* :262-266 */
        const float dm = (p.cz_ref == SKB_REF_NONE)
                             ? 1.0f
                             : mod_value<MODS>(p.cz_ref, s.sample, mod) * p.cz_depth;
        idx = c_f2i(dev_cz_phasor(p.cz_mode, ph, p.cz_dist + dm, k));
      } else {
        idx = __float2int_rz(ph);                                                 /* :268 */
      }
      if (idx >= p.tsize) idx = p.tsize - 1;                                      /* :271-272 */
      if (idx < 0) idx = 0;
      f = (p.toff >= 0) ? __ldg(tables + p.toff + idx) : 0.0f;                    /* :274 */
    }
  }
  float x;
  if (p.sh_max) {                                                                 /* :560-571 */
    if (s.sh_count == 0) s.sh_hold = f;
    x = s.sh_hold;
    s.sh_count++;
    if (s.sh_count >= p.sh_max) s.sh_count = 0;
  } else {
    x = f;
  }
  if (p.quant) x = dev_quantize(x, p.quant);                                      /* :574 */
  if (p.fmode) {                                                                  /* :577, :349-364 */
    const float y = p.b0 * x + p.b1 * s.x1 + p.b2 * s.x2 - p.a1 * s.y1 - p.a2 * s.y2;
    s.x2 = s.x1; s.x1 = x; s.y2 = s.y1; s.y1 = y;
    x = y;
  }
  float env = 1.0f;                                                               /* :581-582 */
  if (p.flags & SKB_F_USE_ENV) env = dev_env_step(p, s, ssc) * s.env_vel;
  float am = 1.0f;                                                                /* :583-587 */
  if (p.am_ref != SKB_REF_NONE) am = mod_value<MODS>(p.am_ref, x, mod) * p.am_depth;
  float gain = p.amp * env * am;                                                  /* :588 */
  if (p.flags & SKB_F_SMOOTHER) {                                                 /* :589-592 */
    s.sm_gain = s.sm_gain + p.sm_k * (gain - s.sm_gain);
    gain = s.sm_gain;
  }
  const float out = x * gain;                                                     /* :593 */
  s.sample = out;
  if (p.flags & SKB_F_DISCONNECT) return make_float2(0.0f, 0.0f);                 /* :595,609-612 */
  if (p.pm_ref != SKB_REF_NONE) {                                                 /* :597-602 */
    const float q = mod_value<MODS>(p.pm_ref, out, mod) * p.pm_depth;
    s.panL = (1.0f - q) / 2.0f;
    s.panR = (1.0f + q) / 2.0f;
  }
  return make_float2(out * s.panL, out * s.panR);                                 /* :603-604 */
}

/* ops are sorted (stably) by slot on the host; one thread replays one slot's
 * run in order.  op.voice holds the SLOT here. */
__device__ __forceinline__ void dev_apply_op(VoiceS &s, const skb_op &op) {
  switch (op.code) {
    case SKB_OP_TRIGGER:      s.finished = 0; s.phase = op.f0; break;                       /* synth.c:316-339 */
    case SKB_OP_SET_FINISHED: s.finished = op.i0; break;                                    /* synth.c:281-282 */
    case SKB_OP_ENV_ON:       s.env_start = op.u0; s.env_rel = 0ull; s.env_vel = op.f0; s.env_active = 1; break; /* :383-388 */
    case SKB_OP_ENV_OFF:      if (s.env_active) s.env_rel = op.u0; break;                   /* :391-395 */
    case SKB_OP_ENV_RESET:    s.env_start = 0ull; s.env_rel = 0ull; s.env_active = 0; break; /* :377-379 */
    case SKB_OP_FILTER_CLEAR: s.x1 = s.x2 = s.y1 = s.y2 = 0.0f; break;                      /* :1017-1018 */
    case SKB_OP_VOICE_CLEAR:  s.sample = 0.0f; s.sm_gain = 0.0f; break;                     /* :1094,1124 */
    case SKB_OP_SET_PAN:      s.panL = op.f0; s.panR = op.f1; break;                        /* :841-842 */
    case SKB_OP_SET_SH:       s.sh_count = op.i0; s.sh_hold = op.f0; break;                 /* :1045-1046 */
    case SKB_OP_SET_PHASE:    s.phase = op.f0; break;
    default: break;
  }
}


/* K1  render_free: free_kernel.cuh; K1b render_rows (few voices per GPU): row_kernel.cuh */
#include "free_kernel.cuh"
#ifndef SKB_LO_VARIANT        /* (free_lo.cu compiles k_render_free a second time with other tuning constants: nothing below) */
#include "row_kernel.cuh"
#include "level_kernel.cuh"

/* ======================================================================== */
/* K2  render_bins: modulation groups, frame-lock-step                       */
/* ======================================================================== */
/* A bin is one or more connected components of the F/A/P/C graph
 * packed by the planner, voices in ascending voice-index order, one thread per
 * voice.  voice_sample[] of the bin lives in shared memory, double buffered:
 * a read of modulator m by voice n sees cur[m] when m < n (already rendered
 * this frame) and prev[m] when m > n (synth.c:526 loop order).  `level` is the
 * depth of a voice in the same-frame dependency DAG; levels run one after the
 * other with a CTA barrier in between. */
struct skb_bin_desc { int slot0, size, nlevels, row; };

/* SKB_CANARY=1 (skred_b200/variants/canary, tools/gpu_canary_check.py): every voice_sample[] word of the exchange carries the
 * frame it was written in, and every modulator read checks that it sees the frame the index rule promises — the current
 * frame for m < n, the previous one for m > n.  A missing or misplaced barrier (a reader ahead of its writer, a writer
 * already one frame further) shows up as a count in skb_stats.wide_errors.  compute-sanitizer's racecheck is closed on
 * the pool this was developed on; this is the in-kernel check instead. */
#ifndef SKB_CANARY
#define SKB_CANARY 0
#endif
struct BinMods {
  const float *prev, *cur, *inc;
#if SKB_CANARY
  const int *ptag, *ctag;
  int frame;
  unsigned long long *bad;
#endif
  __device__ __forceinline__ float read(int ref) const {
    const int l = ref & SKB_REF_MASK;
#if SKB_CANARY
    const int tg = (ref & SKB_REF_CUR) ? ctag[l] : ptag[l];
    if (tg != ((ref & SKB_REF_CUR) ? frame : frame - 1)) atomicAdd(bad, 1ull);
#endif
    return (ref & SKB_REF_CUR) ? cur[l] : prev[l];
  }
  __device__ __forceinline__ float inc_of(int ref) const { return inc[ref & SKB_REF_MASK]; }
};

__global__ void __launch_bounds__(1024)
k_render_bins(const float4 *__restrict__ pq, float4 *__restrict__ sq, int cap,
              const skb_bin_desc *__restrict__ bins,
              const float *__restrict__ tables, const float *__restrict__ noise,
              int nframes, unsigned long long ssc_before,
              float2 *__restrict__ partials, int row_stride, unsigned long long *__restrict__ counter,
              float2 *__restrict__ tap, const int *__restrict__ voice_of_slot, int tap_n) {
  extern __shared__ float bsm[];
  const skb_bin_desc bd = bins[blockIdx.x];
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  float *vs0 = bsm, *vs1 = bsm + nt, *incs = bsm + 2 * nt;
  float2 *wsum = (float2 *)(bsm + 3 * nt);        /* [2][nwarps] */
#if SKB_CANARY
  int *tg0 = (int *)(wsum + 2 * (nt >> 5)), *tg1 = tg0 + nt;
  tg0[tid] = -1; tg1[tid] = -2;                   /* vs0 holds "the frame before the launch", vs1 nothing yet */
#endif
  const bool live = tid < bd.size;
  const int slot = bd.slot0 + (live ? tid : 0);
  VoiceP p; VoiceS s; VoiceK k;
  load_params(pq, cap, slot, p);
  load_state(sq, cap, slot, s);
  derive_consts(p, k);
  vs0[tid] = live ? s.sample : 0.0f;
  vs1[tid] = 0.0f;
  incs[tid] = live ? p.inc : 0.0f;
  const bool wants_noise = live && (p.flags & SKB_F_NOISE);
  float2 *out_row = partials + (size_t)bd.row * row_stride;
  float2 *tap_lane = nullptr;                      /* per-voice tap, synth.c:533-611 */
  if (tap_n && live) { const int tv = __ldg(voice_of_slot + slot); if (tv >= 0) tap_lane = tap + tv; }
  __syncthreads();
  for (int f = 0; f < nframes; f++) {
    BinMods mod;
    mod.prev = (f & 1) ? vs1 : vs0;
    float *cur = (f & 1) ? vs0 : vs1;
    mod.cur = cur;
    mod.inc = incs;
#if SKB_CANARY
    int *ctagw = (f & 1) ? tg0 : tg1;
    mod.ptag = (f & 1) ? tg1 : tg0; mod.ctag = ctagw; mod.frame = f; mod.bad = counter + 18;
#endif
    const float white = wants_noise ? __ldg(noise + f) : 0.0f;
    float2 o = make_float2(0.0f, 0.0f);
    for (int lvl = 0; lvl < bd.nlevels; lvl++) {
      if (live && p.level == lvl) {
        o = voice_frame<true>(p, k, s, ssc_before + (unsigned long long)(f + 1), white, tables, mod);
        cur[tid] = s.sample;
#if SKB_CANARY
        ctagw[tid] = f;
#endif
      }
      __syncthreads();
    }
    if (tap_lane) tap_lane[(size_t)f * tap_n] = o;
    /* fixed-order sum: xor butterfly inside the warp, then warps in order */
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      o.x += __shfl_xor_sync(0xffffffffu, o.x, d);
      o.y += __shfl_xor_sync(0xffffffffu, o.y, d);
    }
    if (lane == 0) wsum[(f & 1) * nwarps + warp] = o;
    if (f > 0 && tid == 0) {
      float L = 0.0f, R = 0.0f;
      const float2 *w = wsum + ((f - 1) & 1) * nwarps;
      for (int i = 0; i < nwarps; i++) { L += w[i].x; R += w[i].y; }
      out_row[f - 1] = make_float2(L, R);
    }
  }
  __syncthreads();
  if (tid == 0 && nframes > 0) {
    float L = 0.0f, R = 0.0f;
    const float2 *w = wsum + ((nframes - 1) & 1) * nwarps;
    for (int i = 0; i < nwarps; i++) { L += w[i].x; R += w[i].y; }
    out_row[nframes - 1] = make_float2(L, R);
  }
  if (live) store_state(sq, cap, slot, s);
  int na = live ? s.nact : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) na += __shfl_xor_sync(0xffffffffu, na, d);
  if (lane == 0 && na) atomicAdd(counter, (unsigned long long)na);
}

/* K2h  render_bins_huge: a CYCLIC modulation component of more than 1,024 voices (one shared feedback loop with thousands
 * of readers).  Legal in the reference (any F / A / P / C graph is), rare, and inherently frame-lock-step — so this is the
 * fallback that keeps it rendering instead of refusing it (round 1: SKB_ERR_CAPACITY): one CTA of 1,024 threads, a thread
 * walks voices tid, tid + 1024, ... of the component per dependency level, every voice's record is re-read from and
 * written back to HBM each frame, and voice_sample[] of the component is exchanged through a global buffer
 * xs[3][cap] (previous frame, current frame, phase increments) under the same index rule as k_render_bins. */
__global__ void __launch_bounds__(1024)
k_render_bins_huge(const float4 *__restrict__ pq, float4 *sq, int cap,
                   const skb_bin_desc *__restrict__ bins,
                   const float *__restrict__ tables, const float *__restrict__ noise,
                   int nframes, unsigned long long ssc_before,
                   float2 *__restrict__ partials, int row_stride, unsigned long long *__restrict__ counter,
                   float2 *__restrict__ tap, const int *__restrict__ voice_of_slot, int tap_n, float *xs) {
  __shared__ float2 wsum[2][32];
  const skb_bin_desc bd = bins[blockIdx.x];
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  float *vs0 = xs + bd.slot0, *vs1 = xs + (size_t)cap + bd.slot0, *incs = xs + 2 * (size_t)cap + bd.slot0;
#if SKB_CANARY
  int *tg0 = (int *)(xs + 3 * (size_t)cap) + bd.slot0, *tg1 = tg0 + cap;
#endif
  for (int i = tid; i < bd.size; i += nt) {
    const float4 st0 = ldq(sq, 0, cap, bd.slot0 + i);
    const float4 pa0 = ldq(pq, 0, cap, bd.slot0 + i);
    vs0[i] = st0.z; vs1[i] = 0.0f; incs[i] = pa0.y;
#if SKB_CANARY
    tg0[i] = -1; tg1[i] = -2;
#endif
  }
  float2 *out_row = partials + (size_t)bd.row * row_stride;
  int na = 0;
  __syncthreads();
  for (int f = 0; f < nframes; f++) {
    BinMods mod;
    mod.prev = (f & 1) ? vs1 : vs0;
    float *cur = (f & 1) ? vs0 : vs1;
    mod.cur = cur;
    mod.inc = incs;
#if SKB_CANARY
    int *ctagw = (f & 1) ? tg0 : tg1;
    mod.ptag = (f & 1) ? tg1 : tg0; mod.ctag = ctagw; mod.frame = f; mod.bad = counter + 18;
#endif
    float2 acc = make_float2(0.0f, 0.0f);
    for (int lvl = 0; lvl < bd.nlevels; lvl++) {
      for (int i = tid; i < bd.size; i += nt) {
        const int slot = bd.slot0 + i;
        if (__float_as_int(ldq(pq, 7, cap, slot).w) != lvl) continue;
        VoiceP p; VoiceS s; VoiceK k;
        load_params(pq, cap, slot, p);
        load_state(sq, cap, slot, s);
        derive_consts(p, k);
        const float white = (p.flags & SKB_F_NOISE) ? __ldg(noise + f) : 0.0f;
        const float2 o = voice_frame<true>(p, k, s, ssc_before + (unsigned long long)(f + 1), white, tables, mod);
        cur[i] = s.sample;
#if SKB_CANARY
        ctagw[i] = f;
#endif
        store_state(sq, cap, slot, s);
        na += s.nact;
        if (tap_n) { const int tv = __ldg(voice_of_slot + slot); if (tv >= 0) tap[(size_t)f * tap_n + tv] = o; }
        acc.x += o.x; acc.y += o.y;
      }
      __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, d);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, d);
    }
    if (lane == 0) wsum[f & 1][warp] = acc;
    if (f > 0 && tid == 0) {
      float L = 0.0f, R = 0.0f;
      for (int i = 0; i < nwarps; i++) { L += wsum[(f - 1) & 1][i].x; R += wsum[(f - 1) & 1][i].y; }
      out_row[f - 1] = make_float2(L, R);
    }
  }
  __syncthreads();
  if (tid == 0 && nframes > 0) {
    float L = 0.0f, R = 0.0f;
    for (int i = 0; i < nwarps; i++) { L += wsum[(nframes - 1) & 1][i].x; R += wsum[(nframes - 1) & 1][i].y; }
    out_row[nframes - 1] = make_float2(L, R);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) na += __shfl_xor_sync(0xffffffffu, na, d);
  if (lane == 0 && na) atomicAdd(counter, (unsigned long long)na);
}

/* K2w  render_bins_warp: modulation groups of <= 32 voices, ONE WARP per bin.
 * The same frame-lock-step rule as k_render_bins (synth.c:526 loop order: modulator m < n is read at the current
 * frame, m > n at the previous one), but the voices of a bin are the lanes of one warp: the same-frame dependency
 * levels are separated by __syncwarp() instead of CTA barriers, the bin's voice_sample[] exchange is 64 floats of
 * shared memory private to the warp, and the stereo sum goes through the warp's tile 16 frames at a time (the
 * transposed reduce of k_render_free) instead of ten shuffles and a thread-0 loop per frame.  Because nothing of a
 * bin is shared with another warp, bins ride inside BATCHED launches like free voices: the launch renders `nwin`
 * consecutive callbacks and the lane that owns a voice replays the ops queued for the boundary before each of them
 * (seq.c:170-178) on its registers.  Ops of bin voices sit in bucket `ob_bucket` of the per-window CSR. */
#define SKB_BINW_WARPS 4
struct BinWarpArgs {
  const float4 *pq; float4 *sq; int cap;
  const skb_bin_desc *bins; int nbins;
  const float *tables; const float *noise;
  int nframes; unsigned long long ssc_before;
  const int *win_frames; const int *win_ob; int nwin, ob_stride, ob_bucket;
  const skb_op *bops;
  float2 *partials; int row_stride;
  unsigned long long *counter;
  float2 *tap; const int *voice_of_slot; int tap_n;
};

__global__ void __launch_bounds__(SKB_BINW_WARPS * 32) k_render_bins_warp(const __grid_constant__ BinWarpArgs a) {
  __shared__ float s_vs[SKB_BINW_WARPS][2][32];
  __shared__ float s_inc[SKB_BINW_WARPS][32];
  __shared__ float2 s_tile[SKB_BINW_WARPS][SKB_TILE_FLOAT2];
  __shared__ float2 s_row[SKB_BINW_WARPS][SKB_UNIT];
#if SKB_CANARY
  __shared__ int s_tag[SKB_BINW_WARPS][2][32];
#endif
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * SKB_BINW_WARPS + warp;
  if (b >= a.nbins) return;                                   /* (warps are independent: no CTA barrier below) */
  const skb_bin_desc bd = a.bins[b];
  const bool live = lane < bd.size;
  const int slot = bd.slot0 + (live ? lane : 0);
  VoiceP p; VoiceS s; VoiceK k;
  load_params(a.pq, a.cap, slot, p);
  load_state(a.sq, a.cap, slot, s);
  if (!live) { p.amp = 0.0f; p.flags = SKB_F_SMOOTHER; p.cz_mode = 0; p.fmode = 0; p.level = 0;
               p.fm_ref = SKB_REF_NONE; p.am_ref = SKB_REF_NONE; p.pm_ref = SKB_REF_NONE; p.cz_ref = SKB_REF_NONE; p.sh_max = 0; p.quant = 0; }
  derive_consts(p, k);
  float *vs0 = s_vs[warp][0], *vs1 = s_vs[warp][1];
  float2 *mytile = s_tile[warp], *myrow = s_row[warp];
  vs0[lane] = live ? s.sample : 0.0f;
  vs1[lane] = 0.0f;
#if SKB_CANARY
  int *tg0 = s_tag[warp][0], *tg1 = s_tag[warp][1];
  tg0[lane] = -1; tg1[lane] = -2;
#endif
  s_inc[warp][lane] = live ? p.inc : 0.0f;
  const bool wants_noise = live && (p.flags & SKB_F_NOISE);
  float2 *out_row = a.partials + (size_t)bd.row * a.row_stride;
  float2 *tap_lane = nullptr;                                 /* per-voice tap, synth.c:533-611 */
  if (a.tap_n && live) { const int tv = __ldg(a.voice_of_slot + slot); if (tv >= 0) tap_lane = a.tap + tv; }
  __syncwarp();
  int fr = 0;                                                 /* launch-relative frame */
  for (int w = 0; w < a.nwin; w++) {
    /* ops of the boundary before window w, replayed by the lane that owns the voice */
    const int ob = a.win_ob ? __ldg(a.win_ob + (size_t)w * a.ob_stride + a.ob_bucket) : 0;
    const int oe = a.win_ob ? __ldg(a.win_ob + (size_t)w * a.ob_stride + a.ob_bucket + 1) : 0;
    if (oe > ob) {
      if (live) {
        int lo = ob, hi = oe;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&a.bops[mid].voice) < slot) lo = mid + 1; else hi = mid; }
        for (int i = lo; i < oe; i++) {
          const skb_op op = a.bops[i];
          if (op.voice != slot) break;
          dev_apply_op(s, op);
        }
        ((fr & 1) ? vs1 : vs0)[lane] = s.sample;              /* what modulators read as "previous frame" (voice_reset clears it) */
#if SKB_CANARY
        ((fr & 1) ? tg1 : tg0)[lane] = fr - 1;
#endif
      }
      __syncwarp();
    }
    const int wn = __ldg(a.win_frames + w);
    for (int f0 = 0; f0 < wn; f0 += SKB_UNIT) {
      const int cnt = min(SKB_UNIT, wn - f0);
      for (int j = 0; j < cnt; j++, fr++) {
        BinMods mod;
        mod.prev = (fr & 1) ? vs1 : vs0;
        float *cur = (fr & 1) ? vs0 : vs1;
        mod.cur = cur;
        mod.inc = s_inc[warp];
#if SKB_CANARY
        int *ctagw = (fr & 1) ? tg0 : tg1;
        mod.ptag = (fr & 1) ? tg1 : tg0; mod.ctag = ctagw; mod.frame = fr; mod.bad = a.counter + 18;
#endif
        const float white = wants_noise ? __ldg(a.noise + fr) : 0.0f;
        float2 o = make_float2(0.0f, 0.0f);
        for (int lvl = 0; lvl < bd.nlevels; lvl++) {
          if (live && p.level == lvl) {
            o = voice_frame<true>(p, k, s, a.ssc_before + (unsigned long long)(fr + 1), white, a.tables, mod);
            cur[lane] = s.sample;
#if SKB_CANARY
            ctagw[lane] = fr;
#endif
          }
          __syncwarp();
        }
        if (tap_lane) tap_lane[(size_t)fr * a.tap_n] = o;
        mytile[j * SKB_TILE_STRIDE + lane] = o;
      }
      for (int j = cnt; j < SKB_UNIT; j++) mytile[j * SKB_TILE_STRIDE + lane] = make_float2(0.0f, 0.0f);
      __syncwarp();
      reduce_unit(mytile, myrow, lane, cnt);
      __syncwarp();
      if (lane < cnt) out_row[fr - cnt + lane] = myrow[lane];
      __syncwarp();
    }
  }
  if (live) store_state(a.sq, a.cap, slot, s);
  int na = live ? s.nact : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) na += __shfl_xor_sync(0xffffffffu, na, d);
  if (lane == 0 && na) atomicAdd(a.counter, (unsigned long long)na);
}

/* Selected columns of the per-voice tap (only the voices being recorded, voice_record[], cross PCIe).
 * Thread (frame f, voice v): a selected voice's (L, R) goes to the compact out[f][col]; the others only feed the
 * extremes ext[0] = min(0, samples), ext[1] = max(0, samples) that wire.c's save_wav derives its scale from
 * (wire.c:150-166 scans EVERY voice of the recording, recorded or not).  Bit patterns of floats >= 0 order like ints,
 * those of floats <= 0 like reversed unsigned ints: one atomic each. */
__global__ void k_tap_select(const float2 *__restrict__ tap, int n, int nframes, const int *__restrict__ col_of_voice, int nsel,
                             float2 *__restrict__ out, unsigned int *__restrict__ ext) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float big = 0.0f, small = 0.0f;
  if (i < (size_t)n * nframes) {
    const int v = (int)(i % n);
    const size_t f = i / n;
    const float2 t = tap[i];
    const int c = col_of_voice[v];
    if (c >= 0) out[f * nsel + c] = t;
    else {
      if (t.x > big) big = t.x; if (t.y > big) big = t.y;            /* (a NaN never wins: `g > fbig` is false, wire.c:152) */
      if (t.x < small) small = t.x; if (t.y < small) small = t.y;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    big = fmaxf(big, __shfl_xor_sync(0xffffffffu, big, d));
    small = fminf(small, __shfl_xor_sync(0xffffffffu, small, d));
  }
  if ((threadIdx.x & 31) == 0) {
    if (big > 0.0f) atomicMax(ext + 1, __float_as_uint(big));
    if (small < 0.0f) atomicMax(ext + 0, __float_as_uint(small));     /* more negative = larger bit pattern */
  }
}

/* ======================================================================== */
/* K4  reduce_rows: partial rows -> raw stereo mix, fixed order              */
/* ======================================================================== */
/* One partial row per (CTA, batch) of k_render_free plus one per modulation bin.  Grid =
 * frame tiles of 32 x row chunks.  Chunk c sums its rows in index order (y-strided, then the
 * 8 y partials in order) into part2[c][f]; the LAST chunk CTA of a frame tile to arrive
 * (atomic ticket) adds the chunk partials in chunk order — the result does not depend on
 * which CTA that is, so the sum is run-to-run deterministic. */
#define SKB_RED_X 32
#define SKB_RED_Y 8
#define SKB_RED_CHUNKS 16
__global__ void __launch_bounds__(SKB_RED_X * SKB_RED_Y)
k_reduce_rows(const float2 *__restrict__ partials, int nrows,
              int nframes, int row_stride, float2 *__restrict__ part2, unsigned int *__restrict__ tickets,
              float2 *__restrict__ mix) {
  __shared__ float2 acc[SKB_RED_Y][SKB_RED_X];
  __shared__ unsigned int s_ticket;
  const int f = blockIdx.x * SKB_RED_X + threadIdx.x;
  const int nch = gridDim.y, c = blockIdx.y;
  const int per = (nrows + nch - 1) / nch;
  const int r0 = c * per, r1 = min(nrows, r0 + per);
  float L = 0.0f, R = 0.0f;
  if (f < nframes) {
#pragma unroll 4
    for (int r = r0 + threadIdx.y; r < r1; r += SKB_RED_Y) {
      const float2 v = partials[(size_t)r * row_stride + f];
      L += v.x; R += v.y;
    }
  }
  acc[threadIdx.y][threadIdx.x] = make_float2(L, R);
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 1; y < SKB_RED_Y; y++) { L += acc[y][threadIdx.x].x; R += acc[y][threadIdx.x].y; }
    if (nch == 1) {
      if (f < nframes) mix[f] = make_float2(L, R);
    } else {
      if (f < nframes) part2[(size_t)c * row_stride + f] = make_float2(L, R);
      __threadfence();
    }
  }
  if (nch == 1) return;
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_ticket = atomicAdd(&tickets[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != (unsigned)(nch - 1)) return;
  __threadfence();
  if (threadIdx.y == 0 && f < nframes) {
    float2 v[SKB_RED_CHUNKS];
#pragma unroll
    for (int i = 0; i < SKB_RED_CHUNKS; i++)
      v[i] = (i < nch) ? __ldcg(&part2[(size_t)i * row_stride + f]) : make_float2(0.0f, 0.0f);
    float l = 0.0f, r = 0.0f;
#pragma unroll
    for (int i = 0; i < SKB_RED_CHUNKS; i++) if (i < nch) { l += v[i].x; r += v[i].y; }
    mix[f] = make_float2(l, r);
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) tickets[blockIdx.x] = 0u;     /* ready for the next launch */
}

/* K5  master volume (synth.c:619-624): the one-pole trace `gain` is computed
 * by the host in the reference's arithmetic; the device scales and stores. */
__global__ void k_finish(const float2 *__restrict__ mix, const float *__restrict__ gain,
                         float2 *__restrict__ out, int nframes) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < nframes) {
    const float2 m = mix[f];
    const float g = gain[f];
    out[f] = make_float2(m.x * g, m.y * g);
  }
}

/* The same, straight into the HOST's buffers (round 2, session 3): `out`, `counters_out` and `flag` are pinned host memory
 * mapped into the device's address space, so the scaled frames and the engine's counters cross PCIe as the kernel's own
 * posted writes — no device staging buffer, no two device-to-host copies queued behind the kernel, no stream
 * synchronisation: the last block to finish (ticket) publishes `seq` in `flag` after a system-wide fence, and the host,
 * which polls that word, finds every frame in place when it sees it.  `gain` = the per-frame trace, or nullptr for the
 * usual case of a master volume that sits on its fixed point (`gain_const`; the products are the same either way). */
__global__ void k_finish_host(const float2 *__restrict__ mix, const float *__restrict__ gain, float gain_const,
                              float2 *__restrict__ out, int nframes,
                              const unsigned long long *__restrict__ counters, unsigned long long *__restrict__ counters_out,
                              int ncounters, unsigned int *ticket, volatile unsigned int *flag, unsigned int seq) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < nframes) {
    const float2 m = mix[f];
    const float g = gain ? gain[f] : gain_const;
    out[f] = make_float2(m.x * g, m.y * g);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ncounters) counters_out[threadIdx.x] = counters[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0u;
      __threadfence_system();
      *flag = seq;
    }
  }
}

/* ======================================================================== */
/* K3  state edits at a block boundary                                       */
/* ======================================================================== */
__global__ void k_apply_ops(float4 *__restrict__ sq, int cap, const skb_op *__restrict__ ops,
                            const int2 *__restrict__ runs, int nruns) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nruns) return;
  const int2 r = runs[i];
  const int slot = ops[r.x].voice;
  VoiceS s;
  load_state(sq, cap, slot, s);
  for (int j = r.x; j < r.y; j++) dev_apply_op(s, ops[j]);
  store_state(sq, cap, slot, s);
}

/* parameter records: recs[i][SKB_NPQ] float4 -> pq[k][slots[i]] */
__global__ void k_scatter_params(float4 *__restrict__ pq, int cap, const int *__restrict__ slots,
                                 const float4 *__restrict__ recs, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * SKB_NPQ) return;
  const int i = t / SKB_NPQ, k = t % SKB_NPQ;
  pq[(size_t)k * cap + slots[i]] = recs[t];
}

/* re-plan: dst[k][s] = src[k][old_slot[s]] (zeros for a fresh slot) */
__global__ void k_permute_state(const float4 *__restrict__ src, float4 *__restrict__ dst, int cap,
                                const int *__restrict__ old_slot, int nslots) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots * SKB_NSQ) return;
  const int k = t / nslots, sidx = t % nslots;
  const int o = old_slot[sidx];
  dst[(size_t)k * cap + sidx] = (o >= 0) ? src[(size_t)k * cap + o] : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void k_gather_state(const float4 *__restrict__ sq, int cap, const int *__restrict__ slots,
                               int n, skb_voice_state *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  skb_voice_state o;
  const int slot = slots[i];
  if (slot < 0) {
    memset(&o, 0, sizeof(o));
  } else {
    VoiceS s;
    load_state(sq, cap, slot, s);
    o.phase = s.phase; o.finished = s.finished; o.sample = s.sample; o.sh_hold = s.sh_hold;
    o.sh_count = s.sh_count; o.x1 = s.x1; o.x2 = s.x2; o.y1 = s.y1; o.y2 = s.y2;
    o.env_active = s.env_active; o.env_velocity = s.env_vel; o.smoother_gain = s.sm_gain;
    o.pan_left = s.panL; o.pan_right = s.panR; o.env_start = s.env_start; o.env_release = s.env_rel;
  }
  out[i] = o;
}

__global__ void k_scatter_state(float4 *__restrict__ sq, int cap, const int *__restrict__ slots,
                                int n, const skb_voice_state *__restrict__ in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = slots[i];
  if (slot < 0) return;
  const skb_voice_state o = in[i];
  VoiceS s;
  s.phase = o.phase; s.finished = o.finished; s.sample = o.sample; s.sh_hold = o.sh_hold;
  s.sh_count = o.sh_count; s.x1 = o.x1; s.x2 = o.x2; s.y1 = o.y1; s.y2 = o.y2;
  s.env_active = o.env_active; s.env_vel = o.env_velocity; s.sm_gain = o.smoother_gain;
  s.panL = o.pan_left; s.panR = o.pan_right; s.env_start = o.env_start; s.env_rel = o.env_release; s.aux = 0;
  store_state(sq, cap, slot, s);
}
#endif /* !SKB_LO_VARIANT */
