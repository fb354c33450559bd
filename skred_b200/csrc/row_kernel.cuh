/* row_kernel.cuh — K1b `k_render_rows`: free voices when a launch leaves the SMs nearly empty.
 *
 * Replaces synth.c:520-613 for voices with no live cross-voice read, like k_render_free, for the case that
 * kernel cannot help: ONE 65,536-voice job cut 8 ways (BASELINE configs[4] as written) leaves 8,192 voices =
 * 256 rows of 32 on a GPU, 1-2 warps per SM.  One thread per voice then walks the launch alone and a launch takes
 * (frames x 50 ... 93 cycles), the pace of ONE warp's instruction stream, however idle the GPU is.  The recurrences
 * that force this are short — phase (synth.c:226-258: add, compare, wrap), biquad (:349-364), amp smoother
 * (:589-592) — everything between them (CZ warp, index, table read, envelope, pan, mix) is not a recurrence.
 *
 * Shape.  ONE CTA PER ROW of 32 voices, lane = voice in every warp, the warps are STAGES of a pipeline over blocks
 * of RP_FB frames, handed over through shared memory with one CTA barrier per step:
 *
 *   warp 0          A   phase recurrence of block t (+ one-shot end, synth.c:242-245)        -> ph[t]
 *   warps 1..RP_G   G   block t-1: CZ warp, index, table read (:149-215, 262-274) of 8 frames each,
 *                       envelope gain amp * (env * velocity) of lanes on a moving segment (:398-431, 582, 588)
 *                                                                                            -> xs[t-1], gv[t-1]
 *                       block t-3: pan, sum over the row's voices, partial row to HBM (:603-606)
 *   warp RP_G+1     C   block t-2: biquad, amp smoother, voice_sample (:349-364, 589-593)     -> out[t-2]
 *
 * so the only serial work per frame is A's 5 and C's 10-15 instructions on two different warps: ~25 cycles per
 * frame instead of 50-93.  Every op is the reference's, on the operands the sequential loop has (the helpers are
 * the ones of the pipelined path of k_render_free: frame_x, env_gain_at, fast_setup): every evolving word is
 * bit-identical (tests run the whole parity suite through this kernel with SKB_ROWS=1), the cross-voice sum has
 * its own fixed order.  Boundary events: the launch is cut into SEGMENTS at the window boundaries
 * at which this row has ops; between segments the state goes back to its HBM record, the ops are replayed on it and
 * the lanes are set up again.  A row with a voice the pipelined path cannot render (S&H, quantize, noise, reverse,
 * ...: lane_needs_generic) is rendered by warp 0 alone through voice_frame<>, like a generic warp of k_render_free.
 */
#pragma once

#ifndef RP_FB
#define RP_FB 32                       /* frames per block */
#endif
#ifndef RP_G
#define RP_G 4                         /* gather warps; each takes RP_FB / RP_G frames of a block */
#endif
#define RP_WARPS (RP_G + 2)
#define RP_THREADS (RP_WARPS * 32)
#define RP_NBUF 4                      /* blocks in flight: t (A), t-1 (G), t-2 (C), t-3 (mix) */
#define RP_OSTRIDE 33
#define RP_ENDED 0x100                 /* nval flag: the lane's one-shot ended inside the block */

struct RowSmem {
  float ph[RP_NBUF][RP_FB][32];
  float xs[RP_NBUF][RP_FB][32];
  float gv[RP_NBUF][RP_FB][32];
  float out[RP_NBUF][RP_FB][RP_OSTRIDE];
  int nval[RP_NBUF][32];               /* rendered frames of the lane in the block (| RP_ENDED) */
  int bfr[RP_NBUF];                    /* frames of the block */
  int bslow[RP_NBUF];                  /* some lane's one-shot ends inside the block */
  float2 pan[32];
  int envover[32];                     /* the lane's envelope ended at a rendered frame of this segment (synth.c:429) */
  EnvRec er[32];                       /* envelope records of the lanes on a time-varying segment */
  unsigned varmask;                    /* ... and which lanes those are */
  float fphase[32];
  float2 gtile[SKB_TILE_FLOAT2];       /* generic rows: stereo tile and 16 frames of row */
  float2 grow[SKB_UNIT];
};

__host__ __device__ inline size_t skb_rows_smem_bytes() { return sizeof(RowSmem); }

/* A: phases of one block.  Returns with fs.phase advanced; `dead` set when the lane's one-shot ended. */
__device__ __forceinline__ void rp_phase_block(RowSmem &S, int bi, int nf, const FastK &c, FastS &fs, bool &dead, int lane) {
  float (*ph)[32] = S.ph[bi];
  int nv = 0;
  bool ended = false;
  if (!dead) {
    /* no lane reaches its table end within H frames (the conservative horizon of k_render_free) */
    int H = 0x7fffffff;
    if (c.stop && c.inc > 0.0f) {
      const float n = ((c.hi - fs.phase) / (c.inc + c.hi * 1.1920929e-7f)) * 0.999f - 3.0f;
      H = (n < 1.0e9f) ? max(__float2int_rz(n), 0) : 0x7fffffff;
    }
    if (H >= nf) {
      float phase = fs.phase;
      int j = 0;
      for (; j + 8 <= nf; j += 8) {
        float p8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const float q = phase + c.inc;                              /* :226 */
          const float w = q - c.hi_wrap;                              /* :247 (exact, see stage_phase) */
          phase = wrap_pick(q, w, c.hi_wrap);
          p8[k] = phase;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) ph[j + k][lane] = p8[k];          /* :258 */
      }
      for (; j < nf; j++) {
        const float q = phase + c.inc;
        const float w = q - c.hi_wrap;
        phase = wrap_pick(q, w, c.hi_wrap);
        ph[j][lane] = phase;
      }
      fs.phase = phase;
      nv = nf;
    } else {
      bool fin = false;
      for (int j = 0; j < nf; j++) {
        if (!fin) { frame_phase(c, fs, fin); nv = j + 1; }           /* :226-258 incl. :243-245 */
        ph[j][lane] = fs.phase;
      }
      ended = fin;
    }
  } else {
    for (int j = 0; j < nf; j++) ph[j][lane] = fs.phase;             /* a finite phase: G computes a value nobody uses */
  }
  S.nval[bi][lane] = nv | (ended ? RP_ENDED : 0);
  const unsigned eb = __ballot_sync(0xffffffffu, ended);
  if (lane == 0) { S.bfr[bi] = nf; S.bslow[bi] = eb != 0u; }
  if (ended) dead = true;
}

/* G: table reads of RP_FB / RP_G frames of one block, branch-free: the phases of all frames are loaded first, then the
 * indices, then the table words, so the loads of the frames overlap.  Frames past the end of a ragged block hold a
 * stale phase: the index is clamped into the table, the value is never used.
 * CZ: 0 none (the index is the truncated phase), 1 the lane's piecewise / fast_pow form (frame_x). */
template <int CZ>
__device__ __forceinline__ void rp_gather_block(RowSmem &S, int bi, int j0, const FastK &c, int lane,
                                                const float *__restrict__ tables) {
  constexpr int NF = RP_FB / RP_G;
  float q[NF], x[NF];
  unsigned idx[NF];
#pragma unroll
  for (int k = 0; k < NF; k++) q[k] = S.ph[bi][j0 + k][lane];
#pragma unroll
  for (int k = 0; k < NF; k++) {
    if (CZ == 0) {
      idx[k] = min(trunc_small_u(q[k]), (unsigned)c.imax);           /* :268 */
    } else {
      const float u = q[k] * c.inv_size;                              /* :151 */
      const float r = c.is_pow ? dev_fast_pow(u, c.k1) : ((u < c.czT) ? u * c.k1 : c.czC + (u - c.czU) * c.k2);
      idx[k] = (unsigned)max(min(c_f2i(r * c.size_f), c.imax), 0);    /* :214, 265, 271-272 */
    }
  }
#pragma unroll
  for (int k = 0; k < NF; k++) x[k] = __ldg(c.tp + idx[k]);           /* :274 */
#pragma unroll
  for (int k = 0; k < NF; k++) S.xs[bi][j0 + k][lane] = x[k];
}

/* G: envelope gains amp * (amp_envelope_step() * velocity) of one block (synth.c:398-431, 582, 588), FRAME-PARALLEL: the
 * envelope is a closed form of the sample counter, so thread (q, f) evaluates frame f of the q-th lane that sits on a
 * time-varying segment — the IEEE divisions cost (varying lanes x frames) / 128 threads, not 8 frames x a whole warp. */
__device__ __forceinline__ void rp_env_block(RowSmem &S, int bi, int nf, int frame0, int gtid) {
  const unsigned mask = S.varmask;
  const int items = __popc(mask) * nf;
  for (int i = gtid; i < items; i += RP_G * 32) {
    const int q = i / nf, f = i - q * nf;
    const int v = __fns(mask, 0, q + 1);                              /* the q-th varying lane */
    if (f < (S.nval[bi][v] & (RP_ENDED - 1))) {
      const EnvRec r = S.er[v];
      bool done;
      S.gv[bi][f][v] = env_gain_at(r, r.t0 + frame0 + f + 1, r.tr0 + frame0 + f + 1, &done);
      if (done) S.envover[v] = 1;                                     /* :429 */
    }
  }
}

/* G: pan, sum over the row's 32 voices, partial row -> HBM.  Lane (f, h), f = lane % FPW, h = lane / FPW, adds voices
 * VPG h .. VPG h + VPG - 1 of frame j0 + f left to right, the lane groups are joined by shuffles: a fixed order. */
__device__ __forceinline__ void rp_mix_block(const RowSmem &S, int bi, int j0, int nf, float2 *orow_at, int lane) {
  constexpr int FPW = RP_FB / RP_G;              /* frames per gather warp */
  constexpr int NV = 32 / FPW * 0 + (32 * FPW / 32) * 0 + (32 / (32 / FPW));   /* = FPW ... voices per lane group: 32 / (32 / FPW) */
  static_assert(FPW == 4 || FPW == 8 || FPW == 16, "RP_FB / RP_G must be 4, 8 or 16");
  constexpr int NG = 32 / FPW;                   /* lane groups; each adds 32 / NG voices */
  constexpr int VPG = 32 / NG;
  (void)NV;
  const int f = lane % FPW, h = lane / FPW, j = j0 + f;
  float L = 0.0f, R = 0.0f;
  if (j < nf) {
#pragma unroll
    for (int v = 0; v < VPG; v++) {
      const float2 p = S.pan[VPG * h + v];
      const float o = S.out[bi][j][VPG * h + v];
      L += o * p.x; R += o * p.y;                                     /* :603-606 */
    }
  }
#pragma unroll
  for (int d = FPW; d < 32; d <<= 1) { L += __shfl_xor_sync(0xffffffffu, L, d); R += __shfl_xor_sync(0xffffffffu, R, d); }
  if (h == 0 && j < nf) orow_at[j] = make_float2(L, R);
}

/* C: biquad, smoother, voice_sample of one block in which no lane ends.  The inputs of 8 frames are loaded into
 * registers BEFORE any of their results is stored: left to itself the compiler keeps every shared-memory load behind
 * the store of the frame before (it cannot tell the arrays of RowSmem apart once the block index is dynamic), which
 * puts the 30-cycle load latency on the chain of every frame (measured: 51 cycles per frame for a plain voice). */
template <int FILT, int DYN>
__device__ __forceinline__ void rp_out_block(RowSmem &S, int bi, int nf, const FastK &c, FastS &s, int lane) {
  float x1 = s.x1, x2 = s.x2, y1 = s.y1, y2 = s.y2, g = s.g, last = s.sample;
  const float *xs = &S.xs[bi][0][lane];
  const float *gv = &S.gv[bi][0][lane];
  float *out = &S.out[bi][0][lane];
  int j = 0;
  for (; j + 8 <= nf; j += 8) {
    float v[8], gn[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = xs[(j + k) * 32];
    if (DYN) {
#pragma unroll
      for (int k = 0; k < 8; k++) gn[k] = c.is_buf ? gv[(j + k) * 32] : c.gc;     /* :580-588 */
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float w = v[k];
      if (FILT) {                                                     /* :349-364 */
        const float y = c.b0 * w + c.b1 * x1 + c.b2 * x2 - c.a1 * y1 - c.a2 * y2;
        x2 = x1; x1 = w; y2 = y1; y1 = y;
        w = (FILT == 2 && !c.has_f) ? w : y;
      }
      if (DYN) g = g + c.sm_k * (gn[k] - g);                          /* :589-592 */
      o[k] = w * g;                                                   /* :593 */
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[(j + k) * RP_OSTRIDE] = o[k];
    last = o[7];
  }
  for (; j < nf; j++) {
    float w = xs[j * 32];
    if (FILT) {
      const float y = c.b0 * w + c.b1 * x1 + c.b2 * x2 - c.a1 * y1 - c.a2 * y2;
      x2 = x1; x1 = w; y2 = y1; y1 = y;
      w = (FILT == 2 && !c.has_f) ? w : y;
    }
    if (DYN) { const float gain = c.is_buf ? gv[j * 32] : c.gc; g = g + c.sm_k * (gain - g); }
    last = w * g;
    out[j * RP_OSTRIDE] = last;
  }
  if (FILT) { s.x1 = x1; s.x2 = x2; s.y1 = y1; s.y2 = y2; }
  s.g = g; s.sample = last;
}

__global__ void __launch_bounds__(RP_THREADS) k_render_rows(const __grid_constant__ FreeArgs a) {
  extern __shared__ float4 smem_raw[];
  RowSmem &S = *reinterpret_cast<RowSmem *>(smem_raw);
  const float4 *__restrict__ pq = a.pq;
  float4 *__restrict__ sq = a.sq;
  const float *__restrict__ tables = a.tables;
  const int cap = a.cap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row = blockIdx.x;
  const bool roleA = warp == 0, roleC = warp == RP_WARPS - 1, roleG = !roleA && !roleC;
  const int slot = row * 32 + lane;
  const bool inr = slot < a.n_free;
  float2 *orow = a.ctarows + (size_t)(a.group0 + row) * a.row_stride;
  const int *__restrict__ obrow = a.win_ob + row;                     /* CSR [nwin][ob_stride], one row per free row */
  const TblCtx tb = {nullptr, nullptr, nullptr};

  /* ---- a row nobody renders and no event of the batch touches: zeros ---- */
  {
    bool renders = false, woken = false;
    if (inr) {
      const float amp = pq[slot].x;
      float4 s0 = sq[slot];
      renders = (__float_as_int(s0.y) == 0) && (amp != 0.0f);
      woken = a.wake != nullptr && ((__ldg(a.wake + (slot >> 5)) >> (slot & 31)) & 1u);
      if (roleA && !renders && __float_as_uint(s0.z) != 0u) { s0.z = 0.0f; sq[slot] = s0; }      /* skipped voice: voice_sample = 0, :534,540 */
    }
    if (!__any_sync(0xffffffffu, renders || woken)) {
      for (int f = tid; f < a.nframes; f += RP_THREADS) orow[f] = make_float2(0.0f, 0.0f);
      return;
    }
  }
  if (roleA && lane == 0) atomicAdd(a.counters + 18, 1ull);
  int nact = 0;

  int w = 0, f0 = 0;
  while (w < a.nwin) {
    /* ---- ops of the boundary before window w (trigger, envelope on / off, pan ...: seq.c:170-178), replayed on the HBM
     * record by the lane of warp C that owns the voice ---- */
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
    __syncthreads();                                                  /* the HBM records are current for every warp */
    int w_end = w + 1, nfr = __ldg(a.win_frames + w);
    while (w_end < a.nwin && __ldg(obrow + (size_t)w_end * a.ob_stride) == __ldg(obrow + (size_t)w_end * a.ob_stride + 1)) {
      nfr += __ldg(a.win_frames + w_end);
      w_end++;
    }
    const unsigned long long ssc_seg = a.ssc_before + (unsigned long long)f0;     /* count before the segment's first frame */

    /* ---- every warp sets its lane up from the record ---- */
    FastK c; FastS fs;
    bool dead = true, varying = false;
    int cls = -1;
    EnvRec er;
    er.A = er.D = er.S = er.R = er.vel = er.amp = 0.0f; er.t0 = er.tr0 = er.flags = 0;
    fast_neutral(c, fs, tables);
    if (inr) {
      const float amp = pq[slot].x;
      const float4 s0 = sq[slot];
      if (__float_as_int(s0.y) == 0 && amp != 0.0f) {
        VoiceP p; VoiceS s; VoiceK kk;
        load_params(pq, cap, slot, p);
        load_state(sq, cap, slot, s);
        derive_consts(p, kk);
        cls = a.force_generic ? 7 : lane_class(p, kk, s, a.nframes - f0, ssc_seg);
        dead = false;
        if (cls != 7) {
          varying = env_varying(p, s, ssc_seg);
          fast_setup(p, kk, s, varying, tables, tb, c, fs);
          if (varying) {
            er.A = p.envA; er.D = p.envD; er.S = p.envS; er.R = p.envR; er.vel = s.env_vel; er.amp = p.amp;
            er.t0 = (int)(unsigned)(ssc_seg - s.env_start);
            er.tr0 = (int)(unsigned)(ssc_seg - s.env_rel);
            er.flags = (s.env_active ? 1 : 0) | (s.env_rel != 0ull ? 2 : 0);
          }
        }
      }
    }
    const bool any_live = __any_sync(0xffffffffu, !dead);
    const bool generic = __any_sync(0xffffffffu, !dead && cls == 7);

    if (!any_live) {
      for (int f = tid; f < nfr; f += RP_THREADS) orow[f0 + f] = make_float2(0.0f, 0.0f);
    } else if (generic) {
      /* ---- the literal per-frame restatement, warp 0 alone ---- */
      if (roleA) {
        VoiceP p; VoiceK kk; VoiceS s;
        load_params(pq, cap, inr ? slot : 0, p);
        load_state(sq, cap, inr ? slot : 0, s);
        if (!inr) { p.amp = 0.0f; p.flags = SKB_F_SMOOTHER; p.cz_mode = 0; p.fmode = 0; }
        derive_consts(p, kk);
        for (int f = 0; f < nfr; f += SKB_UNIT) {
          const int cnt = min(SKB_UNIT, nfr - f);
          generic_frames(p, kk, s, f0 + f, 0, cnt, a.ssc_before, tables, a.noise, S.gtile, S.grow, lane, nullptr, 0);
          if (lane < cnt) orow[f0 + f + lane] = S.grow[lane];
          __syncwarp();
        }
        if (inr) store_state(sq, cap, slot, s);
        nact += inr ? s.nact : 0;
      }
    } else {
      /* ---- the pipeline ---- */
      bool dyn = false;
      int filt = 0;
      if (roleC) {
        S.pan[lane] = make_float2(c.panL, c.panR);
        S.envover[lane] = 0;
        if (varying) S.er[lane] = er;
        const unsigned vm = __ballot_sync(0xffffffffu, varying);
        if (lane == 0) S.varmask = vm;
        const bool st = !varying && smoother_settled(fs.g, c.sm_k, c.gc);     /* smoother on its fixed point: not stepped */
        dyn = !__all_sync(0xffffffffu, st);
        const bool anyf = __any_sync(0xffffffffu, c.has_f), allf = __all_sync(0xffffffffu, c.has_f || dead);
        filt = !anyf ? 0 : (allf ? 1 : 2);
      }
      const bool has_rows = __any_sync(0xffffffffu, varying);
      const bool cz_any = __any_sync(0xffffffffu, !dead && c.czT != CUDART_INF_F);   /* some lane warps its phase (cz_setup) */
      const int nb = (nfr + RP_FB - 1) / RP_FB;
      __syncthreads();
      for (int t = 0; t < nb + 3; t++) {
        const long long t_s0 = clock64();
        if (roleA) {
          if (t < nb) rp_phase_block(S, t & (RP_NBUF - 1), min(RP_FB, nfr - t * RP_FB), c, fs, dead, lane);
        } else if (roleG) {
          const int g = warp - 1;
          const int bg = t - 1, bm = t - 3;
          if (bg >= 0 && bg < nb) {
            const int nf = min(RP_FB, nfr - bg * RP_FB);
            const int j0 = g * (RP_FB / RP_G);
            if (cz_any) rp_gather_block<1>(S, bg & (RP_NBUF - 1), j0, c, lane, tables);
            else rp_gather_block<0>(S, bg & (RP_NBUF - 1), j0, c, lane, tables);
            if (has_rows) rp_env_block(S, bg & (RP_NBUF - 1), nf, bg * RP_FB, g * 32 + lane);
          }
          if (bm >= 0 && bm < nb)
            rp_mix_block(S, bm & (RP_NBUF - 1), g * (RP_FB / RP_G), min(RP_FB, nfr - bm * RP_FB), orow + f0 + bm * RP_FB, lane);
        } else {
          const int b = t - 2;
          if (b >= 0 && b < nb) {
            const int bi = b & (RP_NBUF - 1);
            const int nf = S.bfr[bi];
            if (!S.bslow[bi]) {
              if (dyn) {
                switch (filt) {
                  case 0: rp_out_block<0, 1>(S, bi, nf, c, fs, lane); break;
                  case 1: rp_out_block<1, 1>(S, bi, nf, c, fs, lane); break;
                  default: rp_out_block<2, 1>(S, bi, nf, c, fs, lane); break;
                }
                if (!has_rows) {                                      /* every smoother settled on its constant target? */
                  const bool st = smoother_settled(fs.g, c.sm_k, c.gc);
                  dyn = !__all_sync(0xffffffffu, st);
                }
              } else {
                switch (filt) {
                  case 0: rp_out_block<0, 0>(S, bi, nf, c, fs, lane); break;
                  case 1: rp_out_block<1, 0>(S, bi, nf, c, fs, lane); break;
                  default: rp_out_block<2, 0>(S, bi, nf, c, fs, lane); break;
                }
              }
              if (!dead) nact += nf;
            } else {
              /* a one-shot ends in this block: frame by frame, the end included (synth.c:242-245, 531-536) */
              const int nvw = S.nval[bi][lane];
              const int nv = nvw & (RP_ENDED - 1);
              for (int j = 0; j < nf; j++) {
                float o = 0.0f;
                if (!dead && j < nv) o = frame_out(c, fs, S.xs[bi][j][lane], &S.gv[bi][0][lane], j * 32);
                S.out[bi][j][lane] = o;
              }
              if (!dead) nact += nv;
              if (!dead && (nvw & RP_ENDED)) {
                /* the voice's state is final: the cold words are still in HBM as loaded */
                const bool skipped_later = f0 + b * RP_FB + nv < a.nframes;
                FastS fin = fs;
                fin.phase = S.ph[bi][nv - 1][lane];
                fast_retire(sq, cap, slot, c, fin, skipped_later, c.is_buf && S.envover[lane] != 0);
                dead = true;
                const float2 pan = S.pan[lane];                       /* (earlier frames of the block are still to be mixed) */
                fast_neutral(c, fs, tables);
                c.panL = pan.x; c.panR = pan.y;
              }
            }
          }
        }
        const long long t_s1 = clock64();
        __syncthreads();
        if (lane == 0 && (warp == 0 || warp == 1 || roleC)) {      /* diagnostics (skb_stats.phase_cycles): work and wait per stage */
          const int k = roleA ? 0 : (roleC ? 2 : 1);
          atomicAdd(a.counters + 10 + k, (unsigned long long)(t_s1 - t_s0));
          atomicAdd(a.counters + 13 + k, (unsigned long long)(clock64() - t_s1));
        }
      }
      /* ---- registers -> HBM record ---- */
      if (roleA) S.fphase[lane] = fs.phase;
      __syncthreads();
      if (roleC && inr && !dead) {
        VoiceS s;
        FastS fin = fs;
        fin.phase = S.fphase[lane];
        fast_writeback(sq, cap, slot, c, fin, c.is_buf && S.envover[lane] != 0, s);
        store_state(sq, cap, slot, s);
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
    if (lane == 0 && na) atomicAdd(a.counters, (unsigned long long)na);
  }
}
