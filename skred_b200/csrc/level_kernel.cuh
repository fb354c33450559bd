/* level_kernel.cuh — K2L `k_render_levels`: modulated voices whose modulation graph is a DAG.
 *
 * Replaces synth.c:520-613 for voices that read voice_sample[] of other voices (FM :548-555, CZ-mod :263-266, AM
 * :584-587, pan-mod :597-602) when nothing reads back: the component's dependency graph has no cycle and no voice
 * modulates itself.  The reference walks the voices in index order inside every frame, so a voice sees a modulator
 * with a smaller index at the CURRENT frame and one with a larger index at the PREVIOUS frame.  Either way
 * the modulator's samples do not depend on the reader — so the modulators are rendered FIRST, for the whole launch, and
 * leave voice_sample of every frame in a TRACE row (trace[row][0] = the value before the launch, [1 + f] = after frame
 * f); the readers are rendered by a later launch of the same step and pick `trace[1 + f]` or `trace[f]` by the index
 * rule.  The planner (engine.cu) cuts a component into LEVELS (longest path from a source) and issues one launch of this
 * kernel per level.  Components with a cycle or a self-reference stay in the frame-lock-step bins (k_render_bins*).
 *
 * What that buys: a modulated voice is no longer walked frame by frame through the generic voice_frame<> (~200
 * dependent instructions per frame for a lone warp: 0.59 ms per 512-frame callback for the two voices of 0.sk); it
 * goes through the stage pipeline of k_render_rows (row_kernel.cuh): ONE CTA per 32 voices of a level, lane = voice,
 *     warp 0        A   phase (FM increment from the trace, osc_next's general wrap + one-shot end + isfinite guard)
 *     warps 1..4    G   CZ warp (distortion + trace * depth per frame), index, table read; envelope gains;
 *                       pan (pan-mod from the trace per frame), row sum -> HBM, voice_sample -> own trace row
 *     warp 5        C   biquad, gain = (amp * env) * AM from the trace, amp smoother, voice_sample
 * over blocks of RP_FB frames with one CTA barrier per step.  Every op is the reference's on the reference's operands
 * (the per-frame expressions are those of voice_frame<>): every evolving word is bit-identical.
 * A row with a lane the stages cannot render (S&H, quantize, noise, reverse, smoother off, envelope timers beyond
 * 2^31 samples) falls back to voice_frame<true, TraceMods> on warp 0 for that segment.
 */
#pragma once

struct LevMods {                        /* per lane */
  const float *fm, *cz, *am;            /* the modulator's trace row, already offset by the index rule; nullptr = none */
  float fm_depth, fm_prod, cz_depth, am_depth;
  bool cz_one;                          /* negative cz_mod_osc: the literal 1.0f (synth.c:264) */
};

struct LevSmem {
  RowSmem r;
  int tro[32];                          /* trace row the lane's voice_sample goes to, or -1 */
  int pm_row[32];                       /* pan-mod: trace row (+ offset) or -1, depth */
  int pm_off[32];
  float pm_depth[32];
  int lastf[32];                        /* launch-relative index of the lane's last rendered frame of the segment, -1 = none */
};

__host__ __device__ inline size_t skb_levels_smem_bytes() { return sizeof(LevSmem); }

/* the rare branch of osc_next (synth.c:228-256), out of line so that the 32 unrolled frames of a block stay small:
 * a non-finite phase (reset to 0, the sample of that frame is the literal 0.0f: flag 2), a wrap in either direction, or
 * the end of a one-shot (flag 1). */
__device__ __noinline__ float lv_phase_rare(float p, float lo, float hi, float span, float span2, int stop_at_end, int one_shot, int *flags) {
  if (!(fabsf(p) < CUDART_INF_F)) { *flags = 2 | (one_shot ? 1 : 0); return 0.0f; }     /* :228-232 */
  if (p >= hi) {                                                                           /* :242-248 */
    if (stop_at_end) { *flags = 1; return hi - 1e-6f; }
    return lo + wrap_mod(p - lo, span, span2);
  }
  if (stop_at_end) { *flags = 1; return lo; }                                              /* :249-256 */
  return hi - wrap_mod(lo - p, span, span2);
}

/* A: phases of one block, osc_next (synth.c:217-258) in full: FM increment, !isfinite guard, both wrap directions,
 * one-shot end.  A NaN in ph[] marks a frame whose oscillator output is the literal 0.0f (:228-232). */
__device__ __forceinline__ void lv_phase_block(RowSmem &S, int bi, int nf, int frame0, const VoiceK &k, float inc0, const LevMods &m,
                                               float &phase, bool &dead, int lane) {
  float (*ph)[32] = S.ph[bi];
  int nv = 0;
  bool ended = false;
  /* the increments of the block's frames do not depend on the phase: all of the modulator's samples are requested first
   * (one L2 round trip per block, not per 8 frames), the increments follow off the chain */
  float inc[RP_FB];
#pragma unroll
  for (int t = 0; t < RP_FB; t++) inc[t] = (m.fm != nullptr && !dead && t < nf) ? __ldg(m.fm + frame0 + t) : 0.0f;
#pragma unroll
  for (int t = 0; t < RP_FB; t++) {
    const float g = inc[t] * m.fm_depth;                              /* :553 */
    inc[t] = (m.fm != nullptr) ? inc0 + (m.fm_prod * g) : inc0;       /* :554 */
  }
#pragma unroll
  for (int t = 0; t < RP_FB; t++) {
    float out = 0.0f;
    if (!dead && t < nf) {
      float p = phase + inc[t];                                       /* :226 */
      out = p;
      if (!(p >= k.lo && p < k.hi)) {                                 /* rare: a wrap, an end, or a non-finite phase */
        int fl = 0;
        p = lv_phase_rare(p, k.lo, k.hi, k.span, k.span2, k.stop_at_end ? 1 : 0, k.one_shot, &fl);
        out = (fl & 2) ? CUDART_NAN_F : p;
        if (fl & 1) { ended = true; dead = true; }
      }
      phase = p;                                                      /* :258 */
      nv = t + 1;
    }
    ph[t][lane] = out;
  }
  S.nval[bi][lane] = nv | (ended ? RP_ENDED : 0);
  const unsigned eb = __ballot_sync(0xffffffffu, ended);
  if (lane == 0) { S.bfr[bi] = nf; S.bslow[bi] = eb != 0u; }
}

/* G: oscillator output of RP_FB / RP_G frames of one block (synth.c:262-274, the expressions of voice_frame<>) */
__device__ __forceinline__ void lv_gather_block(RowSmem &S, int bi, int j0, int nf, int frame0, const VoiceP &p, const VoiceK &k,
                                                const LevMods &m, bool setup, int lane, const float *__restrict__ tables) {
  constexpr int NF = RP_FB / RP_G;
  float q[NF], dmv[NF], x[NF];
#pragma unroll
  for (int t = 0; t < NF; t++) q[t] = S.ph[bi][j0 + t][lane];
#pragma unroll
  for (int t = 0; t < NF; t++) dmv[t] = (m.cz != nullptr && setup && j0 + t < nf) ? __ldg(m.cz + frame0 + j0 + t) : 0.0f;
#pragma unroll
  for (int t = 0; t < NF; t++) {
    float v = 0.0f;
    if (setup && !(q[t] != q[t])) {
      int idx;
      if (p.cz_mode) {                                                /* :262-266 */
        const float dm = m.cz_one ? 1.0f : dmv[t] * m.cz_depth;
        idx = c_f2i(dev_cz_phasor(p.cz_mode, q[t], p.cz_dist + dm, k));
      } else {
        idx = __float2int_rz(q[t]);                                   /* :268 */
      }
      if (idx >= p.tsize) idx = p.tsize - 1;                          /* :271-272 */
      if (idx < 0) idx = 0;
      v = __ldg(tables + p.toff + idx);                               /* :274 */
    }
    x[t] = v;
  }
#pragma unroll
  for (int t = 0; t < NF; t++) S.xs[bi][j0 + t][lane] = x[t];
}

/* G: pan (pan-mod per frame), sum over the row's voices, partial row -> HBM, voice_sample -> the voice's trace row */
__device__ __forceinline__ void lv_mix_block(const LevSmem &L, int bi, int j0, int nf, int frame0, float2 *orow_at,
                                             float *__restrict__ trace, int tstride, int lane) {
  constexpr int FPW = RP_FB / RP_G, NG = 32 / FPW, VPG = 32 / NG;
  const RowSmem &S = L.r;
  const int f = lane % FPW, h = lane / FPW, j = j0 + f;
  float Ls = 0.0f, Rs = 0.0f;
  if (j < nf) {
#pragma unroll
    for (int vv = 0; vv < VPG; vv++) {
      const int v = VPG * h + vv;
      float2 p = S.pan[v];
      const float o = S.out[bi][j][v];
      if (L.pm_row[v] >= 0) {                                         /* :597-602 */
        const float qq = __ldg(trace + (size_t)L.pm_row[v] * tstride + L.pm_off[v] + frame0 + j) * L.pm_depth[v];
        p.x = (1.0f - qq) / 2.0f; p.y = (1.0f + qq) / 2.0f;
      }
      Ls += o * p.x; Rs += o * p.y;                                   /* :603-606 */
      if (L.tro[v] >= 0) trace[(size_t)L.tro[v] * tstride + 1 + frame0 + j] = o;
    }
  }
#pragma unroll
  for (int d = FPW; d < 32; d <<= 1) { Ls += __shfl_xor_sync(0xffffffffu, Ls, d); Rs += __shfl_xor_sync(0xffffffffu, Rs, d); }
  if (h == 0 && j < nf) orow_at[j] = make_float2(Ls, Rs);
}

/* C: biquad, gain (AM from the trace), amp smoother, voice_sample of one block, frame by frame */
__device__ __forceinline__ void lv_out_block(RowSmem &S, int bi, int nf, int frame0, const FastK &c, FastS &s, const LevMods &m,
                                             bool dead, int nv, int lane) {
  for (int j0 = 0; j0 < nf; j0 += 8) {
    float v8[8], a8[8], g8[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const bool on = !dead && j0 + t < nv;
      v8[t] = on ? S.xs[bi][j0 + t][lane] : 0.0f;
      g8[t] = (on && c.is_buf) ? S.gv[bi][j0 + t][lane] : c.gc;
      a8[t] = (on && m.am != nullptr) ? __ldg(m.am + frame0 + j0 + t) : 0.0f;
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int j = j0 + t;
      if (j >= nf) break;
      float o = 0.0f;
      if (!dead && j < nv) {
        float w = v8[t];
        if (c.has_f) {                                                /* :349-364 */
          const float y = c.b0 * w + c.b1 * s.x1 + c.b2 * s.x2 - c.a1 * s.y1 - c.a2 * s.y2;
          s.x2 = s.x1; s.x1 = w; s.y2 = s.y1; s.y1 = y;
          w = y;
        }
        float gain = g8[t];                                           /* amp * env, :580-582 */
        if (m.am != nullptr) gain = gain * (a8[t] * m.am_depth);      /* :583-588 */
        s.g = s.g + c.sm_k * (gain - s.g);                            /* :589-592 */
        o = w * s.g;                                                  /* :593 */
        s.sample = o;
      }
      S.out[bi][j][lane] = o;
    }
  }
}

__global__ void __launch_bounds__(RP_THREADS) k_render_levels(const __grid_constant__ FreeArgs a) {
  extern __shared__ float4 smem_raw[];
  LevSmem &L = *reinterpret_cast<LevSmem *>(smem_raw);
  RowSmem &S = L.r;
  const float4 *__restrict__ pq = a.pq;
  float4 *__restrict__ sq = a.sq;
  const float *__restrict__ tables = a.tables;
  float *__restrict__ trace = a.trace;
  const int tstride = a.trace_stride;
  const int cap = a.cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = a.row0 + blockIdx.x;
  const bool roleA = warp == 0, roleC = warp == RP_WARPS - 1, roleG = !roleA && !roleC;
  const int slot = row * 32 + lane;
  const bool inr = slot < a.n_free;
  float2 *orow = a.ctarows + (size_t)(a.group0 + blockIdx.x) * a.row_stride;
  const int *__restrict__ obrow = a.win_ob + a.ob_bucket0 + blockIdx.x;
  const TblCtx tb = {nullptr, nullptr, nullptr};
  const int my_tro = inr ? (__float_as_int(pq[(size_t)8 * cap + slot].y) - 1) : -1;

  int nact = 0;

  int w = 0, f0 = 0;
  while (w < a.nwin) {
    {
      const int ob = __ldg(obrow + (size_t)w * a.ob_stride), oe = __ldg(obrow + (size_t)w * a.ob_stride + 1);
      if (oe > ob && roleC && inr) {
        int lo = ob, hi = oe;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&a.bops[mid].voice) < slot) lo = mid + 1; else hi = mid; }
        if (lo < oe && __ldg(&a.bops[lo].voice) == slot) {
          VoiceS s;
          load_state(sq, cap, slot, s);
          for (int i = lo; i < oe; i++) {
            const skb_op op = a.bops[i];
            if (op.voice != slot) break;
            dev_apply_op(s, op);
          }
          store_state(sq, cap, slot, s);
        }
      }
    }
    __syncthreads();
    if (w == 0) {
      /* ---- launch start (after the ops queued before the first window): the value readers with a LARGER index... no:
       * a reader that comes EARLIER in the voice loop sees at frame 0 what this voice last left — it is zeroed only when
       * its own turn comes (synth.c:534,540) — so the leftover goes to trace[0] first; then the skip rule ---- */
      bool renders = false, woken = false;
      if (inr) {
        const float amp = pq[slot].x;
        float4 s0 = sq[slot];
        renders = (__float_as_int(s0.y) == 0) && (amp != 0.0f);
        woken = a.wake != nullptr && ((__ldg(a.wake + (slot >> 5)) >> (slot & 31)) & 1u);
        if (roleA && my_tro >= 0) trace[(size_t)my_tro * tstride] = s0.z;
        if (roleA && !renders && __float_as_uint(s0.z) != 0u) { s0.z = 0.0f; sq[slot] = s0; }
      }
      const bool any = __any_sync(0xffffffffu, renders || woken);
      __syncthreads();                                                /* (the zeroed samples are visible to every warp) */
      if (!any) {
        for (int f = tid; f < a.nframes; f += RP_THREADS) orow[f] = make_float2(0.0f, 0.0f);
        if (roleA && my_tro >= 0) for (int f = 0; f < a.nframes; f++) trace[(size_t)my_tro * tstride + 1 + f] = 0.0f;
        return;
      }
    }
    int w_end = w + 1, nfr = __ldg(a.win_frames + w);
    while (w_end < a.nwin && __ldg(obrow + (size_t)w_end * a.ob_stride) == __ldg(obrow + (size_t)w_end * a.ob_stride + 1)) {
      nfr += __ldg(a.win_frames + w_end);
      w_end++;
    }
    const unsigned long long ssc_seg = a.ssc_before + (unsigned long long)f0;

    /* ---- every warp sets its lane up from the record ---- */
    VoiceP p; VoiceS st; VoiceK kk;
    FastK c; FastS fs;
    LevMods m;
    m.fm = m.cz = m.am = nullptr; m.fm_depth = m.fm_prod = m.cz_depth = m.am_depth = 0.0f; m.cz_one = false;
    bool dead = true, varying = false, setup = false;
    int cls = -1;
    EnvRec er;
    er.A = er.D = er.S = er.R = er.vel = er.amp = 0.0f; er.t0 = er.tr0 = er.flags = 0;
    fast_neutral(c, fs, tables);
    load_params(pq, cap, inr ? slot : 0, p);
    load_state(sq, cap, inr ? slot : 0, st);
    if (!inr) { p.amp = 0.0f; p.flags = SKB_F_SMOOTHER; p.cz_mode = 0; p.fmode = 0; p.toff = 0; p.tsize = 1;
                p.fm_ref = p.am_ref = p.pm_ref = p.cz_ref = SKB_REF_NONE; p.sh_max = 0; p.quant = 0; p.trace_out = -1; }
    derive_consts(p, kk);
    if (inr && !st.finished && p.amp != 0.0f) {
      cls = a.force_generic ? 7 : lane_class(p, kk, st, a.nframes - f0, ssc_seg, false, true);
      dead = false;
      if (cls != 7) {
        setup = true;
        varying = env_varying(p, st, ssc_seg);
        /* the constants of the stages (the pipelined path's own set-up; a traced CZ lane does not use its CZ constants) */
        fast_setup(p, kk, st, varying, tables, tb, c, fs);
        if (p.flags & SKB_F_DISCONNECT) { c.panL = 0.0f; c.panR = 0.0f; }   /* :595, 609-612: not mixed, stored pan untouched */
        if (varying) {
          er.A = p.envA; er.D = p.envD; er.S = p.envS; er.R = p.envR; er.vel = st.env_vel; er.amp = p.amp;
          er.t0 = (int)(unsigned)(ssc_seg - st.env_start);
          er.tr0 = (int)(unsigned)(ssc_seg - st.env_rel);
          er.flags = (st.env_active ? 1 : 0) | (st.env_rel != 0ull ? 2 : 0);
        }
        auto tr = [&](int ref) -> const float * {
          return (ref >= 0 && (ref & SKB_REF_TRACE)) ? trace + (size_t)(ref & SKB_REF_MASK) * tstride + ((ref & SKB_REF_CUR) ? 1 : 0) + f0 : nullptr; };
        m.fm = tr(p.fm_ref); m.fm_depth = p.fm_depth; m.fm_prod = p.fm_minc * p.fscale;   /* (inc[mod] * freq_scale[n]), :554 */
        m.cz = p.cz_mode ? tr(p.cz_ref) : nullptr; m.cz_depth = p.cz_depth; m.cz_one = p.cz_ref == SKB_REF_NONE;
        m.am = tr(p.am_ref); m.am_depth = p.am_depth;
      }
    }
    const bool any_live = __any_sync(0xffffffffu, !dead);
    const bool generic = __any_sync(0xffffffffu, !dead && cls == 7);
    if (tid == 0 && any_live) atomicAdd(a.counters + 2 + (generic ? 7 : 6), 1ull);   /* diagnostics: segments per path (class_rows[6] staged, [7] generic) */

    if (!any_live) {
      for (int f = tid; f < nfr; f += RP_THREADS) orow[f0 + f] = make_float2(0.0f, 0.0f);
      if (roleA && my_tro >= 0) for (int f = 0; f < nfr; f++) trace[(size_t)my_tro * tstride + 1 + f0 + f] = 0.0f;
    } else if (generic) {
      /* ---- voice_frame<> with the modulators read from the traces, warp 0 alone ---- */
      if (roleA) {
        TraceMods tm; tm.trace = trace; tm.stride = tstride; tm.frame = 0; tm.minc = p.fm_minc;
        const bool wants_noise = (p.flags & SKB_F_NOISE) != 0;
        for (int f = 0; f < nfr; f += SKB_UNIT) {
          const int cnt = min(SKB_UNIT, nfr - f);
          for (int j = 0; j < cnt; j++) {
            tm.frame = f0 + f + j;
            const float white = wants_noise ? __ldg(a.noise + f0 + f + j) : 0.0f;
            S.gtile[j * SKB_TILE_STRIDE + lane] = voice_frame<true>(p, kk, st, a.ssc_before + (unsigned long long)(f0 + f + j + 1), white, tables, tm);
            if (my_tro >= 0) trace[(size_t)my_tro * tstride + 1 + f0 + f + j] = st.sample;
          }
          for (int j = cnt; j < SKB_UNIT; j++) S.gtile[j * SKB_TILE_STRIDE + lane] = make_float2(0.0f, 0.0f);
          __syncwarp();
          reduce_unit(S.gtile, S.grow, lane, cnt);
          __syncwarp();
          if (lane < cnt) orow[f0 + f + lane] = S.grow[lane];
          __syncwarp();
        }
        if (inr) store_state(sq, cap, slot, st);
        nact += inr ? st.nact : 0;
      }
    } else {
      /* ---- the pipeline ---- */
      if (roleC) {
        S.pan[lane] = make_float2(c.panL, c.panR);
        S.envover[lane] = 0;
        if (varying) S.er[lane] = er;
        const unsigned vm = __ballot_sync(0xffffffffu, varying);
        if (lane == 0) S.varmask = vm;
        L.tro[lane] = my_tro;
        const bool pmod = setup && !(p.flags & SKB_F_DISCONNECT) && p.pm_ref >= 0 && (p.pm_ref & SKB_REF_TRACE);
        L.pm_row[lane] = pmod ? (p.pm_ref & SKB_REF_MASK) : -1;
        L.pm_off[lane] = (p.pm_ref >= 0 && (p.pm_ref & SKB_REF_CUR)) ? 1 : 0;
        L.pm_depth[lane] = p.pm_depth;
        L.lastf[lane] = -1;
      }
      const bool has_rows = __any_sync(0xffffffffu, varying);
      /* which stages have to take their general form: FM drives the phase (A, and G for the 0.0f marker), a traced CZ
       * modulator or a table that is not a power of two (G), AM (C); the others run the fast forms of k_render_rows */
      const bool row_fm = __any_sync(0xffffffffu, m.fm != nullptr);
      const bool row_cz = __any_sync(0xffffffffu, setup && p.cz_mode != 0 && (m.cz != nullptr || kk.inv_size == 0.0f));
      const bool row_am = __any_sync(0xffffffffu, m.am != nullptr);
      const bool cz_any = __any_sync(0xffffffffu, !dead && c.czT != CUDART_INF_F);
      bool dyn = false;
      int filt = 0;
      if (roleC) {
        const bool stn = !varying && smoother_settled(fs.g, c.sm_k, c.gc);
        dyn = !__all_sync(0xffffffffu, stn);
        const bool anyf = __any_sync(0xffffffffu, c.has_f), allf = __all_sync(0xffffffffu, c.has_f || dead);
        filt = !anyf ? 0 : (allf ? 1 : 2);
      }
      const int nb = (nfr + RP_FB - 1) / RP_FB;
      float phase = fs.phase;
      bool deadA = dead;
      __syncthreads();
      for (int t = 0; t < nb + 3; t++) {
        const long long t_s0 = clock64();
        if (roleA) {
          if (t < nb) {
            if (row_fm) lv_phase_block(S, t & (RP_NBUF - 1), min(RP_FB, nfr - t * RP_FB), t * RP_FB, kk, p.inc, m, phase, deadA, lane);
            else { fs.phase = phase; rp_phase_block(S, t & (RP_NBUF - 1), min(RP_FB, nfr - t * RP_FB), c, fs, deadA, lane); phase = fs.phase; }
          }
        } else if (roleG) {
          const int g = warp - 1;
          const int bg = t - 1, bm = t - 3;
          if (bg >= 0 && bg < nb) {
            const int nf = min(RP_FB, nfr - bg * RP_FB);
            if (row_fm || row_cz) lv_gather_block(S, bg & (RP_NBUF - 1), g * (RP_FB / RP_G), nf, bg * RP_FB, p, kk, m, setup, lane, tables);
            else if (cz_any) rp_gather_block<1>(S, bg & (RP_NBUF - 1), g * (RP_FB / RP_G), c, lane, tables);
            else rp_gather_block<0>(S, bg & (RP_NBUF - 1), g * (RP_FB / RP_G), c, lane, tables);
            if (has_rows) rp_env_block(S, bg & (RP_NBUF - 1), nf, bg * RP_FB, g * 32 + lane);
          }
          if (bm >= 0 && bm < nb)
            lv_mix_block(L, bm & (RP_NBUF - 1), g * (RP_FB / RP_G), min(RP_FB, nfr - bm * RP_FB), f0 + bm * RP_FB, orow + f0 + bm * RP_FB,
                         trace, tstride, lane);
        } else {
          const int b = t - 2;
          if (b >= 0 && b < nb) {
            const int bi = b & (RP_NBUF - 1);
            const int nf = S.bfr[bi];
            const int nvw = S.nval[bi][lane];
            const int nv = nvw & (RP_ENDED - 1);
            if (!row_am && !S.bslow[bi]) {
              if (dyn) {
                switch (filt) {
                  case 0: rp_out_block<0, 1>(S, bi, nf, c, fs, lane); break;
                  case 1: rp_out_block<1, 1>(S, bi, nf, c, fs, lane); break;
                  default: rp_out_block<2, 1>(S, bi, nf, c, fs, lane); break;
                }
                if (!has_rows) {
                  const bool stn = smoother_settled(fs.g, c.sm_k, c.gc);
                  dyn = !__all_sync(0xffffffffu, stn);
                }
              } else {
                switch (filt) {
                  case 0: rp_out_block<0, 0>(S, bi, nf, c, fs, lane); break;
                  case 1: rp_out_block<1, 0>(S, bi, nf, c, fs, lane); break;
                  default: rp_out_block<2, 0>(S, bi, nf, c, fs, lane); break;
                }
              }
            } else {
              lv_out_block(S, bi, nf, b * RP_FB, c, fs, m, dead, nv, lane);
            }
            if (!dead && nv > 0) { nact += nv; L.lastf[lane] = f0 + b * RP_FB + nv - 1; }
            if (!dead && (nvw & RP_ENDED)) {
              /* the voice's state is final (one-shot end, or the !isfinite reset of a one-shot): cold words from HBM */
              const bool skipped_later = f0 + b * RP_FB + nv < a.nframes;
              VoiceS s2;
              load_state(sq, cap, slot, s2);
              const float pl = S.ph[bi][nv - 1][lane];
              s2.phase = (pl != pl) ? 0.0f : pl;
              s2.finished = 1; s2.sm_gain = fs.g; s2.sample = skipped_later ? 0.0f : fs.sample;
              if (c.has_f) { s2.x1 = fs.x1; s2.x2 = fs.x2; s2.y1 = fs.y1; s2.y2 = fs.y2; }
              if (c.is_buf && S.envover[lane] != 0) s2.env_active = 0;
              if (L.pm_row[lane] >= 0) {
                const float qq = __ldg(trace + (size_t)L.pm_row[lane] * tstride + L.pm_off[lane] + L.lastf[lane]) * L.pm_depth[lane];
                s2.panL = (1.0f - qq) / 2.0f; s2.panR = (1.0f + qq) / 2.0f;
              }
              store_state(sq, cap, slot, s2);
              dead = true;
              fast_neutral(c, fs, tables);                            /* (later blocks: exact zeros from the fast forms too) */
            }
          }
        }
        const long long t_s1 = clock64();
        __syncthreads();
        if (lane == 0 && (warp == 0 || warp == 1 || roleC)) {      /* diagnostics (skb_stats.phase_cycles): work and wait per stage */
          const int kx = roleA ? 0 : (roleC ? 2 : 1);
          atomicAdd(a.counters + 10 + kx, (unsigned long long)(t_s1 - t_s0));
          atomicAdd(a.counters + 13 + kx, (unsigned long long)(clock64() - t_s1));
        }
      }
      if (roleA) S.fphase[lane] = phase;
      __syncthreads();
      if (roleC && inr && !dead) {
        VoiceS s2;
        load_state(sq, cap, slot, s2);
        s2.phase = S.fphase[lane]; s2.sm_gain = fs.g; s2.sample = fs.sample;
        if (c.has_f) { s2.x1 = fs.x1; s2.x2 = fs.x2; s2.y1 = fs.y1; s2.y2 = fs.y2; }
        if (c.is_buf && S.envover[lane] != 0) s2.env_active = 0;
        if (L.pm_row[lane] >= 0 && L.lastf[lane] >= 0) {               /* pan-mod overwrites the stored gains, :600-601 */
          const float qq = __ldg(trace + (size_t)L.pm_row[lane] * tstride + L.pm_off[lane] + L.lastf[lane]) * L.pm_depth[lane];
          s2.panL = (1.0f - qq) / 2.0f; s2.panR = (1.0f + qq) / 2.0f;
        }
        store_state(sq, cap, slot, s2);
      }
    }
    __syncthreads();
    f0 += nfr;
    w = w_end;
  }
  if (roleA || roleC) {
    int na = nact;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) na += __shfl_xor_sync(0xffffffffu, na, d);
    if (lane == 0 && na) atomicAdd(a.counters + 1, (unsigned long long)na);
  }
}
