/* oracle/skred_port.c — CPU RESTATEMENT of skred's per-voice render path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing in the product may
 * link, import or execute this file.  It implements the engine C-ABI of
 * include/skred_b200.h on the CPU, frame-outer / voice-inner in voice index
 * order exactly like the reference loop, so that
 *   (1) it can be pinned against the compiled reference (oracle/_ref) and the
 *       golden vectors generated from it (tests/golden/), and
 *   (2) the CUDA engine can be compared with it record-for-record on machines
 *       where /root/reference does not exist.
 *
 * Parity status: PINNED — tests/test_oracle_port.py checks it bit-for-bit
 * against oracle/_ref (reference synth.c compiled with the pinned flags
 * `gcc -O2 -ffp-contract=off`) and against tests/golden/*.npz.
 * The reference itself ships no tests or golden vectors.
 *
 * Build: gcc -O2 -ffp-contract=off (same pinned flags; every float op is a
 * single individually-rounded IEEE op, evaluated in the reference's order).
 *
 * Each function cites the reference lines it restates.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "skred_b200.h"
#include "../skred_b200/csrc/partition.h"

typedef struct {
  float *data;
  int size;
} port_table;

struct skb_engine {
  skb_config cfg;
  int n;
  skb_voice_params *par;
  skb_voice_state *st;
  int32_t *owner;
  int owner_valid;
  port_table *tables;
  int n_tables, cap_tables;
  skb_op *ops;
  int n_ops, cap_ops;
  float *mix;           /* scratch for skb_render */
  float *tap;           /* per-voice tap [max_frames][n][2] (synth.c:533-611), NULL = off */
  int tap_cursor;       /* frames rendered since the last skb_finish */
  int32_t *tap_sel;     /* voices whose tap is read back by skb_read_tap_selected (n_tap_sel = 0: all) */
  int n_tap_sel;
  int err;
  char errtxt[256];
  skb_stats stats;
};

const char *skb_backend_name(void) { return "cpu-port"; }

static int fail(skb_engine *e, int code, const char *what) {
  if (e && e->err == SKB_OK) {
    e->err = code;
    snprintf(e->errtxt, sizeof(e->errtxt), "%s", what);
  }
  return code;
}

int skb_create(skb_engine **out, const skb_config *cfg) {
  if (!out || !cfg || cfg->abi_version != SKB_ABI_VERSION || cfg->n_voices <= 0 ||
      cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world)
    return SKB_ERR_ARG;
  skb_engine *e = (skb_engine *)calloc(1, sizeof(*e));
  e->cfg = *cfg;
  e->n = cfg->n_voices;
  e->par = (skb_voice_params *)calloc((size_t)e->n, sizeof(skb_voice_params));
  e->st = (skb_voice_state *)calloc((size_t)e->n, sizeof(skb_voice_state));
  e->owner = (int32_t *)calloc((size_t)e->n, sizeof(int32_t));
  for (int v = 0; v < e->n; v++) {
    e->par[v].table_id = -1;
    e->par[v].freq_mod_osc = e->par[v].amp_mod_osc = e->par[v].pan_mod_osc = -1;
  }
  e->mix = (float *)calloc((size_t)(cfg->max_frames > 0 ? cfg->max_frames : 512) * 2, sizeof(float));
  *out = e;
  return SKB_OK;
}

void skb_destroy(skb_engine *e) {
  if (!e) return;
  for (int i = 0; i < e->n_tables; i++) free(e->tables[i].data);
  free(e->tables); free(e->par); free(e->st); free(e->owner); free(e->ops); free(e->mix); free(e->tap); free(e->tap_sel);
  free(e);
}

int skb_last_error(const skb_engine *e) { return e ? e->err : SKB_ERR_ARG; }
const char *skb_error_string(const skb_engine *e) { return e ? e->errtxt : "null engine"; }

int skb_table_upload(skb_engine *e, const float *data, int size) {
  if (!e || !data || size <= 0) return fail(e, SKB_ERR_ARG, "table_upload: bad argument");
  if (e->n_tables == e->cap_tables) {
    e->cap_tables = e->cap_tables ? e->cap_tables * 2 : 64;
    e->tables = (port_table *)realloc(e->tables, (size_t)e->cap_tables * sizeof(port_table));
  }
  port_table *t = &e->tables[e->n_tables];
  t->data = (float *)malloc((size_t)size * sizeof(float));
  memcpy(t->data, data, (size_t)size * sizeof(float));
  t->size = size;
  return e->n_tables++;
}

int skb_set_params(skb_engine *e, int voice, const skb_voice_params *p) {
  if (!e || !p || voice < 0 || voice >= e->n) return fail(e, SKB_ERR_ARG, "set_params: bad voice");
  e->par[voice] = *p;
  e->owner_valid = 0;
  e->stats.params_uploaded++;
  return SKB_OK;
}

int skb_push_ops(skb_engine *e, const skb_op *ops, int n) {
  if (!e || (n > 0 && !ops)) return fail(e, SKB_ERR_ARG, "push_ops: bad argument");
  if (e->n_ops + n > e->cap_ops) {
    e->cap_ops = (e->n_ops + n) * 2 + 64;
    e->ops = (skb_op *)realloc(e->ops, (size_t)e->cap_ops * sizeof(skb_op));
  }
  memcpy(e->ops + e->n_ops, ops, (size_t)n * sizeof(skb_op));
  e->n_ops += n;
  return SKB_OK;
}

/* The host setters whose effect lands on evolving state, replayed in order at
 * the block boundary (SURVEY App. B class "D"). */
static void apply_op(skb_voice_state *s, const skb_op *op) {
  switch (op->code) {
    case SKB_OP_TRIGGER:      /* osc_trigger, synth.c:316-339 (phase value chosen by host) */
      s->finished = 0; s->phase = op->f0; break;
    case SKB_OP_SET_FINISHED: /* osc_set_wave_table_index, synth.c:281-282 */
      s->finished = op->i0; break;
    case SKB_OP_ENV_ON:       /* amp_envelope_trigger, synth.c:383-388 */
      s->env_start = op->u0; s->env_release = 0; s->env_velocity = op->f0; s->env_active = 1; break;
    case SKB_OP_ENV_OFF:      /* amp_envelope_release, synth.c:391-395 */
      if (s->env_active) s->env_release = op->u0;
      break;
    case SKB_OP_ENV_RESET:    /* envelope_init tail, synth.c:377-379 */
      s->env_start = 0; s->env_release = 0; s->env_active = 0; break;
    case SKB_OP_FILTER_CLEAR: /* mmf_init, synth.c:1017-1018 */
      s->x1 = s->x2 = s->y1 = s->y2 = 0.0f; break;
    case SKB_OP_VOICE_CLEAR:  /* voice_reset, synth.c:1094,1124 */
      s->sample = 0.0f; s->smoother_gain = 0.0f; break;
    case SKB_OP_SET_PAN:      /* pan_set, synth.c:841-842 */
      s->pan_left = op->f0; s->pan_right = op->f1; break;
    case SKB_OP_SET_SH:       /* voice_copy, synth.c:1045-1046 */
      s->sh_count = op->i0; s->sh_hold = op->f0; break;
    case SKB_OP_SET_PHASE:
      s->phase = op->f0; break;
    default: break;
  }
}

static void flush_ops(skb_engine *e) {
  for (int i = 0; i < e->n_ops; i++) {
    const skb_op *op = &e->ops[i];
    if (op->voice < 0 || op->voice >= e->n) continue;
    apply_op(&e->st[op->voice], op);
  }
  e->stats.ops_applied += (uint64_t)e->n_ops;
  e->n_ops = 0;
}

/* fast_pow, synth.c:140-147 — the int/float punning is part of the result */
static inline float port_fast_pow(float a, float b) {
  if (a <= 0.0f) return 0.0f;
  union { float f; int i; } u = { a };
  u.i = (int)(b * (u.i - 1065353216) + 1065353216);
  return u.f;
}

/* cz_phasor, synth.c:149-215: warp the normalised phase, 7 modes */
static float port_cz_phasor(int mode, float p, float d, int table_size) {
  const float size_f = (float)table_size;
  float ph = p / size_f;
  d = (d < 0.0f) ? 0.0f : (d > 0.999f ? 0.999f : d);            /* :154 */
  if (mode == 1) {                                                /* :157-166 */
    const float k_lo = 0.5f / d;
    const float k_hi = 0.5f / (1.0f - d);
    if (ph < d) ph *= k_lo; else ph = 0.5f + (ph - d) * k_hi;
  } else if (mode == 2) {                                         /* :167-176 */
    const float hd = d * 0.5f;
    const float k = 0.5f / (0.5f - hd);
    if (ph < 0.5f) ph *= k; else ph = 1.0f - (1.0f - ph) * k;
  } else if (mode == 3) {                                         /* :177-186 */
    const float hd = d * 0.5f;
    const float k = 0.5f / (0.5f - hd);
    if (ph < 0.5f) ph *= k; else ph = 0.5f + (ph - 0.5f) * k;
  } else if (mode == 4) {                                         /* :187-192 */
    ph = fmodf(ph * 2.0f, 1.0f);
  } else if (mode == 5) {                                         /* :193-203 */
    const float hd = d * 0.5f;
    const float k1 = 0.5f / (0.5f - hd);
    const float k2 = 0.5f / (0.5f + hd);
    if (ph < 0.5f) ph *= k1; else ph = 0.5f + (ph - 0.5f) * k2;
  } else if (mode == 6) {                                         /* :204-206 */
    ph = port_fast_pow(ph, 1.0f + 4.0f * d);
  } else if (mode == 7) {                                         /* :207-209 */
    ph = port_fast_pow(ph, 1.0f + 8.0f * d);
  } else {
    return p;                                                     /* :210-211 */
  }
  return ph * size_f;                                             /* :214 */
}

static inline float mod_sample(const skb_engine *e, int m) {
  return (m >= 0 && m < e->n) ? e->st[m].sample : 0.0f;
}

/* osc_next, synth.c:217-275 */
static float port_osc_next(skb_engine *e, int n, float inc) {
  const skb_voice_params *p = &e->par[n];
  skb_voice_state *s = &e->st[n];
  if (s->finished) return 0.0f;                                   /* :218 */
  const int size = p->table_size;
  const int one_shot = (p->flags & SKB_F_ONE_SHOT) != 0;
  const int loop_on = (p->flags & SKB_F_LOOP_ENABLED) != 0;
  if (p->flags & SKB_F_REVERSE) inc = -inc;                       /* :224 */
  float ph = s->phase + inc;                                      /* :226 */
  if (!isfinite(ph)) {                                            /* :228-232 */
    s->phase = 0.0f;
    s->finished = one_shot;
    return 0.0f;
  }
  const int use_loop = loop_on && (p->flags & SKB_F_LOOP_VALID);
  const float lo = use_loop ? p->loop_start_f : 0.0f;             /* :235-239 */
  const float hi = use_loop ? p->loop_end_f : (float)size;
  const float span = hi - lo;
  if (ph >= hi) {                                                 /* :242-248 */
    if (one_shot && !loop_on) { ph = hi - 1e-6f; s->finished = 1; }
    else ph = lo + fmodf(ph - lo, span);
  } else if (ph < lo) {                                           /* :249-256 */
    if (one_shot && !loop_on) { ph = lo; s->finished = 1; }
    else ph = hi - fmodf(lo - ph, span);
  }
  s->phase = ph;                                                  /* :258 */
  int idx;
  if (p->cz_mode) {                                               /* :262-266 */
    const int dv = p->cz_mod_osc;
    const float dm = (dv >= 0) ? mod_sample(e, dv) * p->cz_mod_depth : 1.0f;
    idx = (int)port_cz_phasor(p->cz_mode, ph, p->cz_distortion + dm, size);
  } else {
    idx = (int)ph;                                                /* :268 */
  }
  if (idx >= size) idx = size - 1;                                /* :271-272 */
  if (idx < 0) idx = 0;
  if (p->table_id < 0 || p->table_id >= e->n_tables) return 0.0f;
  const port_table *t = &e->tables[p->table_id];
  if (idx >= t->size) idx = t->size - 1;                          /* defensive; equal sizes in practice */
  return t->data[idx];                                            /* :274 */
}

/* quantize_bits_int, synth.c:341-345 (note the double-precision +0.5) */
static inline float port_quantize(float v, int bits) {
  int levels = (1 << bits) - 1;
  int iv = (int)(v * (float)levels + 0.5);
  return (float)iv * (1.0f / (float)levels);
}

/* amp_envelope_step, synth.c:398-431 */
static float port_env_step(const skb_voice_params *p, skb_voice_state *s, uint64_t ssc) {
  if (!s->env_active) return 0;
  float t = (float)(ssc - s->env_start);
  if (t < p->env_attack) return t / p->env_attack;
  float dstart = p->env_attack;
  if (t < dstart + p->env_decay) {
    float in_decay = t - dstart;
    float prog = in_decay / p->env_decay;
    return 1.0f - prog * (1.0f - p->env_sustain);
  }
  if (s->env_release == 0) return p->env_sustain;
  float tr = (float)(ssc - s->env_release);
  if (tr < p->env_release) {
    float prog = tr / p->env_release;
    return p->env_sustain * (1.0f - prog);
  }
  s->env_active = 0;
  return 0.0f;
}

static void ensure_owner(skb_engine *e) {
  if (e->owner_valid) return;
  if (e->cfg.world > 1) {
    int32_t *comp = (int32_t *)malloc((size_t)e->n * sizeof(int32_t));
    skb_components(e->par, e->n, comp);
    skb_partition(comp, e->n, e->cfg.world, e->owner);
    free(comp);
  } else {
    memset(e->owner, 0, (size_t)e->n * sizeof(int32_t));
  }
  e->owner_valid = 1;
}

/* The exchange step (include/skred_b200.h, skb_comm_*): the CPU restatement has no devices and no NCCL.
 * Hosts of the port sum the partial mixes themselves (tests: torch.distributed over gloo); the calls exist so
 * that both implementations export the same ABI, and say so when used with more than one shard. */
int skb_comm_unique_id(void *id_out) { (void)id_out; return SKB_ERR_STATE; }
int skb_comm_init_rank(skb_engine *e, const void *id, int rank, int nranks) { (void)e; (void)id; (void)rank; (void)nranks; return SKB_ERR_STATE; }
int skb_comm_init_all(skb_engine *const *engines, int n) { (void)engines; (void)n; return SKB_ERR_STATE; }
int skb_comm_set_mode(skb_engine *e, int mode) { (void)e; (void)mode; return SKB_OK; }
int skb_comm_size(const skb_engine *e) { (void)e; return 0; }
int skb_comm_destroy(skb_engine *e) { (void)e; return SKB_OK; }
int skb_reduce_mix_all(skb_engine *const *engines, int n, int nframes) { (void)engines; (void)nframes; return n == 1 ? SKB_OK : SKB_ERR_STATE; }
int skb_reduce_mix(skb_engine *e, float *d_mix, int nframes, void *stream) {
  (void)d_mix; (void)nframes; (void)stream;
  if (!e) return SKB_ERR_ARG;
  return e->cfg.world == 1 ? SKB_OK : SKB_ERR_STATE;
}

int skb_owns_voice(skb_engine *e, int voice) {
  if (!e || voice < 0 || voice >= e->n) return 0;
  ensure_owner(e);
  return e->owner[voice] == e->cfg.rank;
}

/* The body of synth(), synth.c:520-613, minus the master volume (:616-624)
 * which skb_finish applies.  mix: nframes*2 floats, HOST memory in the port. */
int skb_render_mix(skb_engine *e, int nframes, uint64_t ssc, const float *noise,
                   float *mix, void *stream) {
  (void)stream;
  if (!e || nframes < 0 || !mix) return fail(e, SKB_ERR_ARG, "render_mix: bad argument");
  ensure_owner(e);
  flush_ops(e);
  const int rank = e->cfg.rank;
  if (e->tap && e->tap_cursor + nframes > e->cfg.max_frames) return fail(e, SKB_ERR_ARG, "render_mix: tap overflow");
  for (int i = 0; i < nframes; i++) {
    ssc++;                                                        /* :521 */
    float L = 0.0f, R = 0.0f;
    float *tap = e->tap ? e->tap + (size_t)(e->tap_cursor + i) * e->n * 2 : NULL;
    if (tap) memset(tap, 0, (size_t)e->n * 2 * sizeof(float));   /* :534-541, 609-611 */
    const float white = noise ? noise[i] : 0.0f;                  /* :525 */
    for (int n = 0; n < e->n; n++) {                              /* :526 */
      if (e->owner[n] != rank) continue;                          /* other shard (components never straddle) */
      const skb_voice_params *p = &e->par[n];
      skb_voice_state *s = &e->st[n];
      if (s->finished) { s->sample = 0.0f; continue; }            /* :531-536 */
      if (p->amp == 0) { s->sample = 0.0f; continue; }            /* :537-542 */
      e->stats.active_voice_frames++;                             /* rendered, not skipped */
      float f;
      if (p->flags & SKB_F_NOISE) {                               /* :543-546 */
        f = white;
      } else {
        const int m = p->freq_mod_osc;
        if (m >= 0 && m != n) {                                   /* :548-555 */
          float g = mod_sample(e, m) * p->freq_mod_depth;
          float minc = (m < e->n) ? e->par[m].phase_inc : 0.0f;
          float inc = p->phase_inc + (minc * p->freq_scale * g);
          f = port_osc_next(e, n, inc);
        } else {
          f = port_osc_next(e, n, p->phase_inc);                  /* :557 */
        }
      }
      float x;
      if (p->sample_hold_max) {                                   /* :560-571 */
        if (s->sh_count == 0) s->sh_hold = f;
        x = s->sh_hold;
        s->sh_count++;
        if (s->sh_count >= p->sample_hold_max) s->sh_count = 0;
      } else {
        x = f;
      }
      if (p->quantize) x = port_quantize(x, p->quantize);         /* :574 */
      if (p->filter_mode) {                                       /* :577, mmf_process :349-364 */
        float y = p->b0 * x + p->b1 * s->x1 + p->b2 * s->x2 - p->a1 * s->y1 - p->a2 * s->y2;
        s->x2 = s->x1; s->x1 = x; s->y2 = s->y1; s->y1 = y;
        x = y;
      }
      s->sample = x;          /* visible to a self-referencing AM read, :586 */
      float env = 1.0f;                                           /* :580-582 */
      if (p->flags & SKB_F_USE_ENV) env = port_env_step(p, s, ssc) * s->env_velocity;
      float am = 1.0f;                                            /* :583-587 */
      if (p->amp_mod_osc >= 0) am = mod_sample(e, p->amp_mod_osc) * p->amp_mod_depth;
      float gain = p->amp * env * am;                             /* :588 */
      if (p->flags & SKB_F_SMOOTHER) {                            /* :589-592 */
        s->smoother_gain += p->smoother_k * (gain - s->smoother_gain);
        gain = s->smoother_gain;
      }
      s->sample = x * gain;                                       /* :593 */
      if (!(p->flags & SKB_F_DISCONNECT)) {                       /* :595-608 */
        if (p->pan_mod_osc >= 0) {
          float q = mod_sample(e, p->pan_mod_osc) * p->pan_mod_depth;
          s->pan_left = (1.0f - q) / 2.0f;
          s->pan_right = (1.0f + q) / 2.0f;
        }
        float l = s->sample * s->pan_left;
        float r = s->sample * s->pan_right;
        L += l;
        R += r;
        if (tap) { tap[2 * n] = l; tap[2 * n + 1] = r; }          /* :607-608 */
      }
    }
    mix[2 * i + 0] = L;
    mix[2 * i + 1] = R;
  }
  e->stats.frames_rendered += (uint64_t)nframes;
  if (e->tap) e->tap_cursor += nframes;
  return e->err;
}

int skb_set_tap(skb_engine *e, int enable) {
  if (!e) return SKB_ERR_ARG;
  if (enable && !e->tap) e->tap = (float *)calloc((size_t)e->cfg.max_frames * e->n * 2, sizeof(float));
  if (!enable) { free(e->tap); e->tap = NULL; }
  e->tap_cursor = 0;
  return e->err;
}

int skb_read_tap(skb_engine *e, int frame0, int nframes, float *out) {
  if (!e || !out || !e->tap || frame0 < 0 || nframes < 0 || frame0 + nframes > e->cfg.max_frames) return SKB_ERR_ARG;
  memcpy(out, e->tap + (size_t)frame0 * e->n * 2, (size_t)nframes * e->n * 2 * sizeof(float));
  return e->err;
}

int skb_set_tap_voices(skb_engine *e, const int32_t *voices, int n) {
  if (!e || n < 0 || (n > 0 && !voices)) return SKB_ERR_ARG;
  free(e->tap_sel);
  e->tap_sel = n ? (int32_t *)malloc((size_t)n * sizeof(int32_t)) : NULL;
  if (n) memcpy(e->tap_sel, voices, (size_t)n * sizeof(int32_t));
  e->n_tap_sel = n;
  return e->err;
}

/* the selected columns, and min(0, x) / max(0, x) over the samples of every other voice (what save_wav's scale needs
 * of them, wire.c:150-166) */
int skb_read_tap_selected(skb_engine *e, int frame0, int nframes, float *out, float *extremes) {
  if (!e || !out || !e->tap || frame0 < 0 || nframes < 0 || frame0 + nframes > e->cfg.max_frames) return SKB_ERR_ARG;
  if (extremes) extremes[0] = extremes[1] = 0.0f;
  if (e->n_tap_sel == 0) return skb_read_tap(e, frame0, nframes, out);
  int *col = (int *)malloc((size_t)e->n * sizeof(int));
  for (int v = 0; v < e->n; v++) col[v] = -1;
  for (int i = 0; i < e->n_tap_sel; i++) col[e->tap_sel[i]] = i;
  float small = 0.0f, big = 0.0f;
  for (int f = 0; f < nframes; f++) {
    const float *src = e->tap + (size_t)(frame0 + f) * e->n * 2;
    for (int v = 0; v < e->n; v++) {
      if (col[v] >= 0) {
        out[((size_t)f * e->n_tap_sel + col[v]) * 2 + 0] = src[2 * v];
        out[((size_t)f * e->n_tap_sel + col[v]) * 2 + 1] = src[2 * v + 1];
      } else {
        for (int k = 0; k < 2; k++) { const float g = src[2 * v + k]; if (g > big) big = g; if (g < small) small = g; }
      }
    }
  }
  free(col);
  if (extremes) { extremes[0] = small; extremes[1] = big; }
  return e->err;
}

/* master volume, synth.c:616-624: the smoother trace `gain` is computed by the
 * host (it is voice independent); here only the scaling and the store. */
int skb_finish(skb_engine *e, const float *mix, int nframes, const float *gain,
               float *out, int num_channels, void *stream) {
  (void)stream;
  if (!e || !mix || !gain || !out || num_channels < 2) return fail(e, SKB_ERR_ARG, "finish: bad argument");
  for (int i = 0; i < nframes; i++) {
    out[(size_t)i * num_channels + 0] = mix[2 * i + 0] * gain[i];
    out[(size_t)i * num_channels + 1] = mix[2 * i + 1] * gain[i];
  }
  e->tap_cursor = 0;
  return e->err;
}

int skb_render(skb_engine *e, int nframes, uint64_t ssc, const float *gain, const float *noise,
               float *out, int num_channels) {
  if (!e) return SKB_ERR_ARG;
  int max = e->cfg.max_frames > 0 ? e->cfg.max_frames : 512;
  int done = 0;
  while (done < nframes) {
    int n = nframes - done < max ? nframes - done : max;
    int r = skb_render_mix(e, n, ssc + (uint64_t)done, noise ? noise + done : NULL, e->mix, NULL);
    if (r) return r;
    r = skb_finish(e, e->mix, n, gain + done, out + (size_t)done * num_channels, num_channels, NULL);
    if (r) return r;
    done += n;
  }
  return e->err;
}

int skb_sync(skb_engine *e, void *stream) { (void)stream; return e ? e->err : SKB_ERR_ARG; }

float *skb_mix_buffer(skb_engine *e) { return e ? e->mix : NULL; }
int skb_flush(skb_engine *e) { return e ? e->err : SKB_ERR_ARG; }      /* the port renders synchronously */

/* no kernels, no phases */
int skb_debug_cta_phases(skb_engine *e, uint64_t *phases, int32_t *rows, int max_ctas, int *rows_cap) {
  (void)e; (void)phases; (void)rows; (void)max_ctas; (void)rows_cap;
  return 0;
}
int skb_debug_slot_rank(skb_engine *e, int slot) { (void)e; (void)slot; return -1; }
int skb_debug_warp_clocks(skb_engine *e, uint64_t *out, int max_ctas) { (void)e; (void)out; (void)max_ctas; return 0; }

int skb_snapshot(skb_engine *e, int first, int n, skb_voice_state *out) {
  if (!e || first < 0 || n < 0 || first + n > e->n || !out) return fail(e, SKB_ERR_ARG, "snapshot: bad range");
  flush_ops(e);
  memcpy(out, e->st + first, (size_t)n * sizeof(skb_voice_state));
  return SKB_OK;
}

int skb_restore(skb_engine *e, int first, int n, const skb_voice_state *in) {
  if (!e || first < 0 || n < 0 || first + n > e->n || !in) return fail(e, SKB_ERR_ARG, "restore: bad range");
  flush_ops(e);
  memcpy(e->st + first, in, (size_t)n * sizeof(skb_voice_state));
  return SKB_OK;
}

int skb_get_stats(skb_engine *e, skb_stats *out) {
  if (!e || !out) return SKB_ERR_ARG;
  ensure_owner(e);
  e->stats.n_owned_voices = 0;
  for (int v = 0; v < e->n; v++) e->stats.n_owned_voices += (e->owner[v] == e->cfg.rank);
  *out = e->stats;
  return SKB_OK;
}
