/* synth_shim.c — host side of the drop-in: skred's synth.h / synth.def API in C,
 * rendering through the B200 engine (include/skred_b200.h).
 *
 * Built against a skred source tree (-I$SKRED_SRC: skred.h, synth.h,
 * synth.def, synth-types.h, retro/korg.h, amysamples.h are used as the API /
 * data contract, nothing is copied), with -DVOICE_MAX override honoured the
 * way the reference does it (a compile-time constant, skred.h:9).
 *
 * What it is: a re-statement of the COLD half of synth.c — every setter and
 * table builder (synth.c:96-136, 277-345, 367-395, 632-1307) keeps its name,
 * arguments, return codes and side effects on the public `voice_*` / `wave_*`
 * arrays, because wire.c reads and writes those arrays directly
 * (wire.c:381-399, 639-708).  What changes is where the HOT half runs:
 * synth() (synth.c:502-630) no longer loops over voices; it
 *   1. turns every voice whose arrays changed since the last block into one
 *      packed parameter record (skb_voice_params),
 *   2. forwards the ordered list of device ops recorded by the setters that
 *      touch evolving state (trigger, envelope on/off/reset, filter clear,
 *      voice clear, pan, finished) — SURVEY App. B class "D",
 *   3. computes the two voice-independent per-frame traces on the host in
 *      the reference's exact arithmetic: the shared noise draw (synth.c:525)
 *      and the master-volume smoother (synth.c:616-617),
 *   4. calls the engine, which renders on the GPU and returns the block.
 * There is no CPU render path: if the engine cannot be created the first
 * synth() call aborts with a message (the reference's synth() returns void).
 *
 * This TU must be compiled with the parity-pinned host flags
 * `-O2 -ffp-contract=off`.
 */
#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "skred.h"
#include "synth-types.h"
#include "miniwav.h"
#include "skred_b200.h"
#include "skred_b200_shim.h"

/* ---- the public state arrays, one definition per synth.def line ---------- */
#define ARRAY(type, name, size, init) type name[size] = init;
#include "synth.def"
#undef ARRAY
#define ARRAY(type, name, size, init) int name##__len__ = size;
#include "synth.def"
#undef ARRAY

#include "synth.h"

int requested_synth_frames_per_callback = SYNTH_FRAMES_PER_CALLBACK;
int synth_frames_per_callback = 0;
volatile uint64_t synth_sample_count = 0;

#define SHIM_SMOOTH_DEFAULT (0.02f)
#define SHIM_INVALID (100)
#define SHIM_BAD_FREQ (101)

float volume_user = 1.0f;
float volume_final = AMY_FACTOR;
float volume_smoother_gain = 0.0f;
float volume_smoother_smoothing = 0.002f;
float volume_threshold = 0.05f;
float volume_smoother_higher_smoothing = 0.3f;

/* ======================================================================== */
/* engine binding, dirty tracking, op queue                                  */
/* ======================================================================== */
static skb_engine *g_engine = NULL;
static int g_cfg_device = 0, g_cfg_rank = 0, g_cfg_world = 1, g_cfg_max_frames = 8192;
static int g_early_flush_frames = 2048;   /* $SKB_EARLY_FLUSH: frames after which synth() launches what it has queued when at least
                                             g_early_flush_rest are still to come, so the GPU renders the head of a long call while the
                                             host fires and queues the events of the rest (0 = one launch per chunk).  Measured on
                                             synth(8192), e2e ms per call: no split 0.848 (round 1), split at 4096 0.784, at 2048 0.768,
                                             at 1024 0.768; a second launch costs ~45 us, so short calls are not split (round 1:
                                             synth(4096) split at 2048 0.529 vs 0.496 ms) */
static int g_early_flush_rest = 4096;     /* $SKB_EARLY_FLUSH_REST */
static int g_scan_all = (VOICE_MAX <= 4096);

static uint8_t g_dirty[VOICE_MAX];
static int32_t g_dirty_list[VOICE_MAX];
static int g_ndirty = 0;
static int32_t g_voice_tid[VOICE_MAX];          /* engine table id of voice_table[v] */
static skb_voice_params g_sent[VOICE_MAX];      /* last record the engine has */
static uint8_t g_sent_valid[VOICE_MAX];

static skb_op *g_ops = NULL;
static int g_nops = 0, g_capops = 0;

/* wave slot -> engine table id cache (keyed by pointer+size: a replaced slot
 * gets a fresh upload while voices that still hold the old pointer keep the
 * old id, mirroring the graveyard in wire.c:370-390) */
static float *g_slot_ptr[WAVE_TABLE_MAX];
static int g_slot_size[WAVE_TABLE_MAX];
static int32_t g_slot_tid[WAVE_TABLE_MAX];

static uint64_t g_rng = 0;
static int g_rng_seeded = 0;
static float *g_gain = NULL, *g_noise = NULL;
static int g_trace_cap = 0;

static void shim_die(const char *what) {
  fprintf(stderr, "skred_b200: FATAL: %s (%s)\n", what,
          g_engine ? skb_error_string(g_engine) : "no engine");
  abort();
}

int skb_shim_configure(int device, int rank, int world, int max_frames) {
  if (g_engine) return SKB_ERR_STATE;
  if (world < 1 || rank < 0 || rank >= world) return SKB_ERR_ARG;
  g_cfg_device = device; g_cfg_rank = rank; g_cfg_world = world;
  if (max_frames >= 512) g_cfg_max_frames = max_frames;
  return SKB_OK;
}

static int g_user_slots[WAVE_TABLE_MAX], g_n_user_slots = 0;   /* user slots that have a device copy (the ones to watch) */
static unsigned char g_user_slot_listed[WAVE_TABLE_MAX];
static skb_engine *engine(void) {
  if (g_engine) return g_engine;
  skb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = SKB_ABI_VERSION;
  const char *s;
  if ((s = getenv("SKB_DEVICE"))) g_cfg_device = atoi(s);
  cfg.device = g_cfg_device;
  cfg.n_voices = VOICE_MAX;
  cfg.max_frames = g_cfg_max_frames;
  cfg.rank = g_cfg_rank;
  cfg.world = g_cfg_world;
  if ((s = getenv("SKB_FORCE_GENERIC")) && atoi(s)) cfg.flags |= SKB_CFG_FORCE_GENERIC;
  if ((s = getenv("SKB_NO_BATCH")) && atoi(s)) cfg.flags |= SKB_CFG_NO_BATCH;
  if ((s = getenv("SKB_WIDE")) && atoi(s)) cfg.flags |= SKB_CFG_WIDE;
  if ((s = getenv("SKB_EARLY_FLUSH"))) g_early_flush_frames = atoi(s);
  if ((s = getenv("SKB_EARLY_FLUSH_REST"))) g_early_flush_rest = atoi(s);
  if ((s = getenv("SKB_NO_AFFINE")) && atoi(s)) cfg.flags |= SKB_CFG_NO_AFFINE;
  int r = skb_create(&g_engine, &cfg);
  if (r != SKB_OK || !g_engine) {
    fprintf(stderr, "skred_b200: FATAL: cannot create %s engine (error %d); there is no CPU fallback\n",
            skb_backend_name(), r);
    abort();
  }
  for (int i = 0; i < WAVE_TABLE_MAX; i++) g_slot_tid[i] = -1;
  g_n_user_slots = 0;
  memset(g_user_slot_listed, 0, sizeof(g_user_slot_listed));
  for (int v = 0; v < VOICE_MAX; v++) g_voice_tid[v] = -1;
  return g_engine;
}

skb_engine *skb_shim_engine(void) { return engine(); }
int skb_shim_last_error(void) { return g_engine ? skb_last_error(g_engine) : SKB_OK; }
void skb_shim_scan_all(int on) { g_scan_all = on; }

static inline void touch(int v) {
  if (v < 0 || v >= VOICE_MAX || g_dirty[v]) return;
  g_dirty[v] = 1;
  g_dirty_list[g_ndirty++] = v;
}
void skb_shim_mark_dirty(int v) { touch(v); }

static void push_op(int v, int code, int i0, float f0, float f1, uint64_t u0) {
  if (g_nops == g_capops) {
    g_capops = g_capops ? g_capops * 2 : 1024;
    g_ops = (skb_op *)realloc(g_ops, (size_t)g_capops * sizeof(skb_op));
  }
  skb_op *o = &g_ops[g_nops++];
  o->voice = v; o->code = code; o->i0 = i0; o->f0 = f0; o->f1 = f1; o->_pad = 0; o->u0 = u0;
}

/* Fingerprint of a user wave slot: 64 samples spread over the table (their bit patterns, position-mixed).  wire.c edits a
 * loaded sample IN PLACE (`/wex` -> wave_table_dynamic_expand, wire.c:553-586, scales every sample; same pointer, same
 * size) without telling synth.c; the device copy is refreshed when the fingerprint of a slot in use has changed. */
static uint64_t g_slot_fp[WAVE_TABLE_MAX];
static uint64_t slot_fingerprint(int wave) {
  const float *d = wave_table_data[wave];
  const int n = wave_size[wave];
  uint64_t h = 1469598103934665603ull;
  if (!d || n <= 0) return h;
  for (int k = 0; k < 64; k++) {
    const int i = (int)(((long long)k * (n - 1)) / 63);
    uint32_t b;
    memcpy(&b, &d[i], 4);
    h = (h ^ (uint64_t)b ^ ((uint64_t)i << 32)) * 1099511628211ull;
  }
  return h;
}

static int32_t slot_table_id(int wave) {
  engine();
  if (g_slot_tid[wave] >= 0 && g_slot_ptr[wave] == wave_table_data[wave] &&
      g_slot_size[wave] == wave_size[wave])
    return g_slot_tid[wave];
  int tid = skb_table_upload(g_engine, wave_table_data[wave], wave_size[wave]);
  if (tid < 0) shim_die("wave table upload failed");
  g_slot_ptr[wave] = wave_table_data[wave];
  g_slot_size[wave] = wave_size[wave];
  g_slot_tid[wave] = tid;
  g_slot_fp[wave] = slot_fingerprint(wave);
  if (wave >= EXT_SAMPLE_000 && !g_user_slot_listed[wave]) { g_user_slot_listed[wave] = 1; g_user_slots[g_n_user_slots++] = wave; }
  return tid;
}

/* Called at every block boundary: a user slot (EXT_SAMPLE_000 ... 999) whose data changed under the same pointer is
 * uploaded again and every voice that plays it is re-pointed — the reference reads the edited floats from the next
 * frame on (its table IS the edited array).  One 64-sample fingerprint per uploaded user slot per call. */
static void refresh_edited_tables(void) {
  for (int k = 0; k < g_n_user_slots; k++) {
    const int wave = g_user_slots[k];
    if (g_slot_tid[wave] < 0 || g_slot_ptr[wave] != wave_table_data[wave] || g_slot_size[wave] != wave_size[wave]) continue;
    if (slot_fingerprint(wave) == g_slot_fp[wave]) continue;
    g_slot_tid[wave] = -1;
    const int32_t tid = slot_table_id(wave);
    for (int v = 0; v < VOICE_MAX; v++)
      if (voice_table[v] == wave_table_data[wave] && voice_wave_table_index[v] == wave) { g_voice_tid[v] = tid; touch(v); }
  }
}

/* Force a re-upload of a slot whose contents were edited in place
 * (wave_table_dynamic_expand, wire.c:553-586). */
void skb_shim_wave_touch(int wave) {
  if (wave >= 0 && wave < WAVE_TABLE_MAX) g_slot_tid[wave] = -1;
}

static void pack_params(int v, skb_voice_params *p) {
  memset(p, 0, sizeof(*p));
  p->amp = voice_amp[v];
  p->phase_inc = voice_phase_inc[v];
  p->freq_scale = voice_freq_scale[v];
  p->freq_mod_depth = voice_freq_mod_depth[v];
  p->freq_mod_osc = voice_freq_mod_osc[v];
  p->table_id = voice_table[v] ? g_voice_tid[v] : -1;
  p->table_size = voice_table_size[v];
  p->loop_start_f = voice_loop_start_f[v];
  p->loop_end_f = voice_loop_end_f[v];
  uint32_t f = 0;
  if (voice_one_shot[v]) f |= SKB_F_ONE_SHOT;
  if (voice_loop_enabled[v]) f |= SKB_F_LOOP_ENABLED;
  if (voice_loop_valid[v]) f |= SKB_F_LOOP_VALID;
  if (voice_direction[v]) f |= SKB_F_REVERSE;
  if (voice_use_amp_envelope[v]) f |= SKB_F_USE_ENV;
  if (voice_smoother_enable[v]) f |= SKB_F_SMOOTHER;
  if (voice_disconnect[v]) f |= SKB_F_DISCONNECT;
  if (voice_wave_table_index[v] == WAVE_TABLE_NOISE_ALT) f |= SKB_F_NOISE;
  p->flags = f;
  p->cz_mode = voice_cz_mode[v];
  p->cz_distortion = voice_cz_distortion[v];
  p->cz_mod_osc = voice_cz_mod_osc[v];
  p->cz_mod_depth = voice_cz_mod_depth[v];
  p->sample_hold_max = voice_sample_hold_max[v];
  p->quantize = voice_quantize[v];
  p->filter_mode = voice_filter_mode[v];
  p->b0 = voice_filter[v].b0; p->b1 = voice_filter[v].b1; p->b2 = voice_filter[v].b2;
  p->a1 = voice_filter[v].a1; p->a2 = voice_filter[v].a2;
  p->env_attack = voice_amp_envelope[v].attack_time;
  p->env_decay = voice_amp_envelope[v].decay_time;
  p->env_sustain = voice_amp_envelope[v].sustain_level;
  p->env_release = voice_amp_envelope[v].release_time;
  p->amp_mod_osc = voice_amp_mod_osc[v];
  p->amp_mod_depth = voice_amp_mod_depth[v];
  p->smoother_k = voice_smoother_smoothing[v];
  p->pan_mod_osc = voice_pan_mod_osc[v];
  p->pan_mod_depth = voice_pan_mod_depth[v];
}

static void send_if_changed(int v) {
  skb_voice_params p;
  pack_params(v, &p);
  if (g_sent_valid[v] && memcmp(&p, &g_sent[v], sizeof(p)) == 0) return;
  if (skb_set_params(g_engine, v, &p) != SKB_OK) shim_die("skb_set_params");
  g_sent[v] = p;
  g_sent_valid[v] = 1;
}

/* Everything the setters (and wire.c's direct array writes) did
 * since the previous block becomes visible to the device here. */
int skb_shim_flush(void) {
  engine();
  refresh_edited_tables();
  if (g_scan_all) {
    for (int v = 0; v < VOICE_MAX; v++) send_if_changed(v);
    for (int i = 0; i < g_ndirty; i++) g_dirty[g_dirty_list[i]] = 0;
  } else {
    for (int i = 0; i < g_ndirty; i++) {
      int v = g_dirty_list[i];
      g_dirty[v] = 0;
      send_if_changed(v);
    }
  }
  g_ndirty = 0;
  if (g_nops) {
    if (skb_push_ops(g_engine, g_ops, g_nops) != SKB_OK) shim_die("skb_push_ops");
    g_nops = 0;
  }
  return skb_last_error(g_engine);
}

/* ======================================================================== */
/* RNG (synth.c:102-123) and volume (synth.c:96-100)                         */
/* ======================================================================== */
void audio_rng_init(uint64_t *rng, uint64_t seed) { *rng = seed ? seed : 1; }

uint64_t audio_rng_next(uint64_t *rng) {
  *rng = *rng * 6364136223846793005ULL + 1442695040888963407ULL;   /* MMIX LCG */
  return *rng;
}

float audio_rng_float(uint64_t *rng) {
  uint32_t hi = (uint32_t)(audio_rng_next(rng) >> 32);
  return (float)((int32_t)hi) / 2147483648.0f;
}

int volume_set(float v) {
  volume_user = v;
  volume_final = v * AMY_FACTOR;
  return 0;
}

/* ======================================================================== */
/* oscillator parameters (synth.c:125-136, 277-339)                          */
/* ======================================================================== */
float osc_get_phase_inc(int v, float f) {
  /* table samples per output sample; association as synth.c:130 (App. A-8) */
  float g = f;
  if (voice_one_shot[v]) g /= voice_offset_hz[v];
  return (g * (float)voice_table_size[v]) / voice_table_rate[v] * (voice_table_rate[v] / MAIN_SAMPLE_RATE);
}

void osc_set_freq(int v, float f) {
  voice_phase_inc[v] = osc_get_phase_inc(v, f);
  touch(v);
}

static int g_noise_voices = 0;      /* voices on WAVE_TABLE_NOISE_ALT: the only consumers of the per-frame draw */

void osc_set_wave_table_index(int voice, int wave) {
  if (!(wave_table_data[wave] && wave_size[wave] && wave_rate[wave] > 0.0)) return;   /* :278 */
  g_noise_voices += (wave == WAVE_TABLE_NOISE_ALT) - (voice_wave_table_index[voice] == WAVE_TABLE_NOISE_ALT);
  voice_wave_table_index[voice] = wave;
  voice_finished[voice] = wave_one_shot[wave] ? 1 : 0;                                   /* :281-282 */
  push_op(voice, SKB_OP_SET_FINISHED, voice_finished[voice], 0, 0, 0);
  const int refreq = voice_table_rate[voice] != wave_rate[wave] ||
                     voice_table_size[voice] != wave_size[wave];                          /* :283-286 */
  voice_table_rate[voice] = wave_rate[wave];
  voice_table_size[voice] = wave_size[wave];
  voice_table[voice] = wave_table_data[wave];
  g_voice_tid[voice] = slot_table_id(wave);
  voice_one_shot[voice] = wave_one_shot[wave];
  voice_loop_start[voice] = wave_loop_start[wave];
  voice_loop_enabled[voice] = wave_loop_enabled[wave];
  voice_loop_end[voice] = wave_loop_end[wave];
  voice_midi_note[voice] = wave_midi_note[wave];
  voice_offset_hz[voice] = wave_offset_hz[wave];
  const int a = voice_loop_start[voice], b = voice_loop_end[voice];                       /* :297-307 */
  voice_loop_start_f[voice] = (float)a;
  voice_loop_end_f[voice] = (float)b;
  voice_loop_valid[voice] = b > a;
  voice_loop_length[voice] = (b > a) ? (float)(b - a) : (float)voice_table_size[voice];
  if (refreq) osc_set_freq(voice, voice_freq[voice]);                                     /* :310-312 */
  touch(voice);
}

void osc_trigger(int voice) {
  /* synth.c:316-339: restart position depends on one-shot / direction / loop */
  float start;
  if (voice_one_shot[voice]) {
    start = voice_direction[voice] ? (float)(voice_table_size[voice] - 1) : 0.0f;
  } else if (voice_direction[voice]) {
    start = voice_loop_enabled[voice] ? (float)voice_loop_end[voice] - 1e-6f
                                      : (float)(voice_table_size[voice] - 1);
  } else {
    start = voice_loop_enabled[voice] ? (float)voice_loop_start[voice] : 0.0f;
  }
  voice_finished[voice] = 0;
  voice_phase[voice] = start;
  push_op(voice, SKB_OP_TRIGGER, 0, start, 0, 0);
}

/* ---- reference-arithmetic helpers kept for API completeness -------------- */
/* (the device implements the same functions; these are the host's exported
 *  symbols cz_phasor / osc_next / quantize_bits_int / mmf_process /
 *  amp_envelope_step of synth.h:28-45.  They operate on the HOST mirror only
 *  and are not used by synth().) */
static inline float shim_fast_pow(float a, float b) {            /* synth.c:140-147 */
  if (a <= 0.0f) return 0.0f;
  union { float f; int i; } u = { a };
  u.i = (int)(b * (u.i - 1065353216) + 1065353216);
  return u.f;
}

float cz_phasor(int n, float p, float d, int table_size) {      /* synth.c:149-215 */
  const float sz = (float)table_size;
  float x = p / sz;
  if (d < 0.0f) d = 0.0f; else if (d > 0.999f) d = 0.999f;
  float hd, k, k2;
  switch (n) {
    case 1: if (x < d) x *= 0.5f / d; else { k = 0.5f / (1.0f - d); x = 0.5f + (x - d) * k; } break;
    case 2: hd = d * 0.5f; k = 0.5f / (0.5f - hd);
            x = (x < 0.5f) ? x * k : 1.0f - (1.0f - x) * k; break;
    case 3: hd = d * 0.5f; k = 0.5f / (0.5f - hd);
            x = (x < 0.5f) ? x * k : 0.5f + (x - 0.5f) * k; break;
    case 4: x = fmodf(x * 2.0f, 1.0f); break;
    case 5: hd = d * 0.5f; k = 0.5f / (0.5f - hd); k2 = 0.5f / (0.5f + hd);
            x = (x < 0.5f) ? x * k : 0.5f + (x - 0.5f) * k2; break;
    case 6: x = shim_fast_pow(x, 1.0f + 4.0f * d); break;
    case 7: x = shim_fast_pow(x, 1.0f + 8.0f * d); break;
    default: return p;
  }
  return x * sz;
}

float quantize_bits_int(float v, int bits) {                    /* synth.c:341-345 */
  int levels = (1 << bits) - 1;
  int iv = (int)(v * (float)levels + 0.5);
  return (float)iv * (1.0f / (float)levels);
}

float osc_next(int voice, float phase_inc) {
  /* Host-mirror evaluation of synth.c:217-275; NOT on the render path. */
  if (voice_finished[voice]) return 0.0f;
  const int size = voice_table_size[voice];
  if (voice_direction[voice]) phase_inc = -phase_inc;
  float ph = voice_phase[voice] + phase_inc;
  if (!isfinite(ph)) { voice_phase[voice] = 0.0f; voice_finished[voice] = voice_one_shot[voice]; return 0.0f; }
  const int lp = voice_loop_enabled[voice] && voice_loop_valid[voice];
  const float lo = lp ? voice_loop_start_f[voice] : 0.0f, hi = lp ? voice_loop_end_f[voice] : (float)size;
  const int stop = voice_one_shot[voice] && !voice_loop_enabled[voice];
  if (ph >= hi) { if (stop) { ph = hi - 1e-6f; voice_finished[voice] = 1; } else ph = lo + fmodf(ph - lo, hi - lo); }
  else if (ph < lo) { if (stop) { ph = lo; voice_finished[voice] = 1; } else ph = hi - fmodf(lo - ph, hi - lo); }
  voice_phase[voice] = ph;
  int idx;
  if (voice_cz_mode[voice]) {
    int dv = voice_cz_mod_osc[voice];
    float dm = (dv >= 0) ? voice_sample[dv] * voice_cz_mod_depth[voice] : 1.0f;
    idx = (int)cz_phasor(voice_cz_mode[voice], ph, voice_cz_distortion[voice] + dm, size);
  } else idx = (int)ph;
  if (idx >= size) idx = size - 1;
  if (idx < 0) idx = 0;
  return voice_table[voice][idx];
}

float mmf_process(int n, float x) {                             /* synth.c:349-364 */
  mmf_t *f = &voice_filter[n];
  float y = f->b0 * x + f->b1 * f->x1 + f->b2 * f->x2 - f->a1 * f->y1 - f->a2 * f->y2;
  f->x2 = f->x1; f->x1 = x; f->y2 = f->y1; f->y1 = y;
  return y;
}

/* ======================================================================== */
/* envelope (synth.c:367-431)                                                */
/* ======================================================================== */
void envelope_init(int v, float attack_time, float decay_time, float sustain_level, float release_time) {
  envelope_t *e = &voice_amp_envelope[v];
  e->a = attack_time; e->d = decay_time; e->s = sustain_level; e->r = release_time;
  e->attack_time = attack_time * MAIN_SAMPLE_RATE;      /* seconds -> samples */
  e->decay_time = decay_time * MAIN_SAMPLE_RATE;
  e->sustain_level = fmaxf(0, fminf(1.0f, sustain_level));
  e->release_time = release_time * MAIN_SAMPLE_RATE;
  e->sample_start = 0; e->sample_release = 0; e->is_active = 0;
  push_op(v, SKB_OP_ENV_RESET, 0, 0, 0, 0);
  touch(v);
}

void amp_envelope_trigger(int v, float f) {
  envelope_t *e = &voice_amp_envelope[v];
  e->sample_start = synth_sample_count;
  e->sample_release = 0;
  e->velocity = f;
  e->is_active = 1;
  push_op(v, SKB_OP_ENV_ON, 0, f, 0, e->sample_start);
}

void amp_envelope_release(int v) {
  /* The reference tests is_active, which the render loop clears when the
   * release tail ends (synth.c:429): the test therefore runs on the device. */
  if (voice_amp_envelope[v].is_active) voice_amp_envelope[v].sample_release = synth_sample_count;
  push_op(v, SKB_OP_ENV_OFF, 0, 0, 0, synth_sample_count);
}

float amp_envelope_step(int v) {
  /* Host-mirror evaluation of synth.c:398-431; NOT on the render path. */
  envelope_t *e = &voice_amp_envelope[v];
  if (!e->is_active) return 0;
  float t = (float)(synth_sample_count - e->sample_start);
  if (t < e->attack_time) return t / e->attack_time;
  if (t < e->attack_time + e->decay_time)
    return 1.0f - ((t - e->attack_time) / e->decay_time) * (1.0f - e->sustain_level);
  if (e->sample_release == 0) return e->sustain_level;
  float tr = (float)(synth_sample_count - e->sample_release);
  if (tr < e->release_time) return e->sustain_level * (1.0f - tr / e->release_time);
  e->is_active = 0;
  return 0.0f;
}

int envelope_is_flat(int v) {
  const envelope_t *e = &voice_amp_envelope[v];
  return e->a == 0.0f && e->d == 0.0f && e->s == 1.0f && e->r == 0.0f;
}

/* ======================================================================== */
/* setters (synth.c:640-650, 829-926, 1033-1169)                             */
/* ======================================================================== */
static int bad_voice(int v) { return v < 0 || v >= VOICE_MAX; }

int cz_set(int v, int n, float f) { voice_cz_mode[v] = n; voice_cz_distortion[v] = f; touch(v); return 0; }
int cmod_set(int voice, int o, float f) { voice_cz_mod_osc[voice] = o; voice_cz_mod_depth[voice] = f; touch(voice); return 0; }

int amp_set(int voice, float f) {
  if (!(f >= 0)) return SHIM_INVALID;
  voice_use_amp_envelope[voice] = 0;
  voice_amp[voice] = f;
  voice_user_amp[voice] = f;
  touch(voice);
  return 0;
}

int pan_set(int voice, float f) {
  if (!(f >= -1.0f && f <= 1.0f)) return SHIM_INVALID;
  voice_pan[voice] = f;
  voice_pan_left[voice] = (1.0f - f) / 2.0f;
  voice_pan_right[voice] = (1.0f + f) / 2.0f;
  push_op(voice, SKB_OP_SET_PAN, 0, voice_pan_left[voice], voice_pan_right[voice], 0);
  return 0;
}

int wave_quant(int voice, int n) { voice_quantize[voice] = n; touch(voice); return 0; }

int freq_set(int voice, float f) {
  if (!(f >= 0 && f < (double)MAIN_SAMPLE_RATE)) return SHIM_BAD_FREQ;
  voice_freq[voice] = f;
  osc_set_freq(voice, f);
  return 0;
}

static int toggle(int cur, int state) { return state < 0 ? (cur == 0) : state; }

int wave_mute(int voice, int state) { voice_disconnect[voice] = toggle(voice_disconnect[voice], state); touch(voice); return 0; }
int wave_dir(int voice, int state) { voice_direction[voice] = toggle(voice_direction[voice], state); touch(voice); return 0; }
int wave_loop(int voice, int state) { voice_loop_enabled[voice] = toggle(voice_loop_enabled[voice], state); touch(voice); return 0; }

int pan_mod_set(int voice, int o, float f) {
  if (bad_voice(voice) || bad_voice(o)) return SHIM_INVALID;
  voice_pan_mod_osc[voice] = o; voice_pan_mod_depth[voice] = f; touch(voice);
  return 0;
}

int amp_mod_set(int voice, int o, float f) {
  if (bad_voice(voice) || bad_voice(o)) return SHIM_INVALID;
  voice_amp_mod_osc[voice] = o; voice_amp_mod_depth[voice] = f; touch(voice);
  return 0;
}

int freq_mod_set(int voice, int o, float f) {
  if (bad_voice(voice) || bad_voice(o)) return SHIM_INVALID;
  voice_freq_mod_osc[voice] = o; voice_freq_mod_depth[voice] = f;
  voice_freq_scale[voice] = (float)voice_table_size[voice] / (float)voice_table_size[o];   /* frozen at F time */
  touch(voice);
  return 0;
}

int wave_set(int voice, int wave) {
  if (wave < 0 || wave >= WAVE_TABLE_MAX) return SHIM_INVALID;
  osc_set_wave_table_index(voice, wave);
  return 0;
}

int envelope_set(int voice, float a, float d, float s, float r) { envelope_init(voice, a, d, s, r); return 0; }

/* biquad coefficients (RBJ cookbook forms as synth.c:929-1008); only
 * recomputed when freq / resonance / mode changed */
void mmf_set_params(int n, float f, float resonance) {
  mmf_t *q = &voice_filter[n];
  if (f == q->last_freq && resonance == q->last_resonance && voice_filter_mode[n] == q->last_mode) return;
  q->last_freq = f; q->last_resonance = resonance; q->last_mode = voice_filter_mode[n];
  const float omega = 2.0f * (float)M_PI * f / (float)MAIN_SAMPLE_RATE;
  const float sn = sinf(omega), cs = cosf(omega);
  const float alpha = sn / (2.0f * resonance);
  const float a0 = 1.0f + alpha, a1 = -2.0f * cs, a2 = 1.0f - alpha;
  float b0, b1, b2;
  touch(n);                                    /* mode may have been written directly (wire.c:666-672) */
  switch (voice_filter_mode[n]) {
    case 0: return;
    default:
    case FILTER_LOWPASS:  b0 = (1.0f - cs) / 2.0f; b1 = 1.0f - cs;    b2 = (1.0f - cs) / 2.0f; break;
    case FILTER_HIGHPASS: b0 = (1.0f + cs) / 2.0f; b1 = -(1.0f + cs); b2 = (1.0f + cs) / 2.0f; break;
    case FILTER_BANDPASS: b0 = alpha;              b1 = 0.0f;         b2 = -alpha;             break;
    case FILTER_NOTCH:    b0 = 1.0f;               b1 = -2.0f * cs;   b2 = 1.0f;               break;
    case FILTER_ALL_PASS: b0 = 1.0f - alpha;       b1 = -2.0f * cs;   b2 = 1.0f + alpha;       break;
  }
  q->b0 = b0 / a0; q->b1 = b1 / a0; q->b2 = b2 / a0; q->a1 = a1 / a0; q->a2 = a2 / a0;
  voice_filter_freq[n] = f;
  voice_filter_res[n] = resonance;
}

void mmf_init(int n, float f, float resonance) {
  mmf_t *q = &voice_filter[n];
  q->x1 = q->x2 = q->y1 = q->y2 = 0.0f;
  push_op(n, SKB_OP_FILTER_CLEAR, 0, 0, 0, 0);
  q->last_freq = -1.0f; q->last_resonance = -1.0f; q->last_mode = -1;    /* force recompute */
  voice_filter_freq[n] = f;
  voice_filter_res[n] = resonance;
  mmf_set_params(n, f, resonance);
}

int mmf_set_freq(int n, float f) { mmf_set_params(n, f, voice_filter_res[n]); return 0; }
int mmf_set_res(int n, float res) { if (res > 0) mmf_set_params(n, voice_filter_freq[n], res); return 0; }

float midi2hz(float f) { return 440.0f * powf(2.0f, (f - 69.0f) / 12.0f); }

int voice_set(int n, int *old_voice) {
  if (bad_voice(n)) return SHIM_INVALID;
  if (old_voice) *old_voice = n;
  return 0;
}

int voice_trigger(int voice) { osc_trigger(voice); return 0; }

int wave_default(int voice) {
  float g = midi2hz((float)voice_midi_note[voice]);
  voice_freq[voice] = g;
  voice_note[voice] = (float)voice_midi_note[voice];
  osc_set_freq(voice, g);
  return 0;
}

int freq_midi(int voice, float f) {
  if (!(f >= 0.0 && f <= 127.0)) return SHIM_INVALID;
  if (voice_midi_transpose[voice]) f += voice_midi_transpose[voice];
  return freq_set(voice, midi2hz(f));
}

int voice_copy(int v, int n) {
  /* synth.c:1033-1054.  The S&H latch is evolving state: fetch the source
   * voice's current value from the device first. */
  skb_shim_snapshot_range(v, 1);
  wave_set(n, voice_wave_table_index[v]);
  amp_set(n, voice_user_amp[v]);
  freq_set(n, voice_freq[v]);
  pan_set(n, voice_pan[v]);
  amp_mod_set(n, voice_amp_mod_osc[v], voice_amp_mod_depth[v]);
  freq_mod_set(n, voice_freq_mod_osc[v], voice_freq_mod_depth[v]);
  pan_mod_set(n, voice_pan_mod_osc[v], voice_pan_mod_depth[v]);
  wave_loop(n, voice_loop_enabled[v]);
  wave_dir(n, voice_direction[v]);
  wave_quant(n, voice_quantize[v]);
  voice_sample_hold_max[n] = voice_sample_hold_max[v];
  voice_sample_hold_count[n] = voice_sample_hold_count[v];
  voice_sample_hold[n] = voice_sample_hold[v];
  push_op(n, SKB_OP_SET_SH, voice_sample_hold_count[n], voice_sample_hold[n], 0, 0);
  envelope_set(n, voice_amp_envelope[v].a, voice_amp_envelope[v].d, voice_amp_envelope[v].s, voice_amp_envelope[v].r);
  cz_set(n, voice_cz_mode[v], voice_cz_distortion[v]);
  cmod_set(n, voice_cz_mod_osc[v], voice_cz_mod_depth[v]);
  voice_filter_mode[n] = voice_filter_mode[v];
  mmf_init(n, voice_filter_freq[v], voice_filter_res[v]);
  touch(n);
  return 0;
}

void voice_reset(int i) {
  /* synth.c:1090-1132.  Deliberately NOT reset (App. B `S`): phase, S&H
   * state, cz_*, amp/pan mod depth, envelope velocity. */
  g_noise_voices -= (voice_wave_table_index[i] == WAVE_TABLE_NOISE_ALT);
  voice_wave_table_index[i] = 0;
  voice_table_rate[i] = 0;
  voice_table_size[i] = 0;
  voice_sample[i] = 0;
  voice_amp[i] = 0;
  voice_user_amp[i] = 0;
  voice_pan[i] = 0;
  voice_pan_left[i] = 0.5f;
  voice_pan_right[i] = 0.5f;
  push_op(i, SKB_OP_SET_PAN, 0, 0.5f, 0.5f, 0);
  voice_use_amp_envelope[i] = 0;
  voice_amp_mod_osc[i] = -1;
  voice_freq_mod_osc[i] = -1;
  voice_freq_mod_depth[i] = 0.0f;
  voice_freq_scale[i] = 1.0f;
  voice_pan_mod_osc[i] = -1;
  voice_disconnect[i] = 0;
  voice_quantize[i] = 0;
  voice_direction[i] = 0;
  envelope_init(i, 0.0f, 0.0f, 1.0f, 0.0f);
  voice_freq[i] = 440.0f;
  voice_midi_note[i] = 69.0f;
  voice_midi_transpose[i] = 0;
  voice_link_midi_a[i] = voice_link_midi_b[i] = -1;
  voice_link_velo_a[i] = voice_link_velo_b[i] = -1;
  voice_link_trig[i] = -1;
  osc_set_wave_table_index(i, WAVE_TABLE_SINE);
  voice_filter_mode[i] = 0;
  mmf_init(i, 8000.0f, 0.707f);
  voice_smoother_enable[i] = 1;
  voice_smoother_gain[i] = 0.0f;
  voice_smoother_smoothing[i] = SHIM_SMOOTH_DEFAULT;
  push_op(i, SKB_OP_VOICE_CLEAR, 0, 0, 0, 0);
  voice_glissando_enable[i] = 0;
  voice_glissando_speed[i] = 0.0f;
  voice_glissando_target[i] = voice_freq[i];
  voice_record[i] = 0;
  touch(i);
}

void voice_init(void) { for (int i = 0; i < VOICE_MAX; i++) voice_reset(i); }

int wave_reset(int voice, int n) {
  (void)voice;
  if (bad_voice(n)) voice_init(); else voice_reset(n);      /* `S<bad>` resets ALL voices */
  return 0;
}

int envelope_velocity(int voice, float f) {
  if (bad_voice(voice)) return SHIM_INVALID;
  if (f == 0) {
    amp_envelope_release(voice);
  } else {
    if (!voice_use_amp_envelope[voice]) { voice_use_amp_envelope[voice] = 1; touch(voice); }   /* (a re-trigger changes no parameter) */
    if (voice_one_shot[voice]) osc_trigger(voice);
    amp_envelope_trigger(voice, f);
  }
  return 0;
}

/* ======================================================================== */
/* display (synth.c:663-827): needs the evolving state back from the device  */
/* ======================================================================== */
int skb_shim_snapshot_range(int first, int n) {
  if (first < 0 || n <= 0 || first + n > VOICE_MAX) return SKB_ERR_ARG;
  engine();
  skb_shim_flush();
  skb_voice_state *st = (skb_voice_state *)malloc((size_t)n * sizeof(*st));
  int r = skb_snapshot(g_engine, first, n, st);
  if (r == SKB_OK) {
    for (int k = 0; k < n; k++) {
      const int v = first + k;
      /* a voice-sharded engine holds the evolving state of ITS voices only: the mirror of a voice another GPU owns is
       * left as it is (what the setters last wrote) instead of being overwritten with the zeros skb_snapshot returns */
      if (g_cfg_world > 1 && !skb_owns_voice(g_engine, v)) continue;
      voice_phase[v] = st[k].phase;
      voice_finished[v] = st[k].finished;
      voice_sample[v] = st[k].sample;
      voice_sample_hold[v] = st[k].sh_hold;
      voice_sample_hold_count[v] = st[k].sh_count;
      voice_filter[v].x1 = st[k].x1; voice_filter[v].x2 = st[k].x2;
      voice_filter[v].y1 = st[k].y1; voice_filter[v].y2 = st[k].y2;
      voice_amp_envelope[v].is_active = st[k].env_active;
      voice_amp_envelope[v].velocity = st[k].env_velocity;
      voice_amp_envelope[v].sample_start = st[k].env_start;
      voice_amp_envelope[v].sample_release = st[k].env_release;
      voice_smoother_gain[v] = st[k].smoother_gain;
      voice_pan_left[v] = st[k].pan_left;
      voice_pan_right[v] = st[k].pan_right;
    }
  }
  free(st);
  return r;
}

void skb_shim_snapshot(void) { skb_shim_snapshot_range(0, VOICE_MAX); }

/* The inverse: host arrays -> device (checkpoint resume; a harness that wrote voice_phase[]
 * / voice_finished[] directly, the way wire.c writes other arrays). */
int skb_shim_restore_range(int first, int n) {
  if (first < 0 || n <= 0 || first + n > VOICE_MAX) return SKB_ERR_ARG;
  engine();
  skb_shim_flush();                       /* pending ops first: the arrays already include their effect */
  skb_voice_state *st = (skb_voice_state *)calloc((size_t)n, sizeof(*st));
  for (int k = 0; k < n; k++) {
    const int v = first + k;
    st[k].phase = voice_phase[v];
    st[k].finished = voice_finished[v];
    st[k].sample = voice_sample[v];
    st[k].sh_hold = voice_sample_hold[v];
    st[k].sh_count = voice_sample_hold_count[v];
    st[k].x1 = voice_filter[v].x1; st[k].x2 = voice_filter[v].x2;
    st[k].y1 = voice_filter[v].y1; st[k].y2 = voice_filter[v].y2;
    st[k].env_active = voice_amp_envelope[v].is_active;
    st[k].env_velocity = voice_amp_envelope[v].velocity;
    st[k].env_start = voice_amp_envelope[v].sample_start;
    st[k].env_release = voice_amp_envelope[v].sample_release;
    st[k].smoother_gain = voice_smoother_gain[v];
    st[k].pan_left = voice_pan_left[v];
    st[k].pan_right = voice_pan_right[v];
  }
  int r = skb_restore(g_engine, first, n, st);
  free(st);
  return r;
}

static int64_t ts_ns(const struct timespec *a, const struct timespec *b) {
  return ((int64_t)b->tv_sec - a->tv_sec) * 1000000000LL + ((int64_t)b->tv_nsec - a->tv_nsec);
}

#define APPEND(...) do { ptr += sprintf(ptr, __VA_ARGS__); } while (0)

char *voice_format(int v, char *out, int verbose) {
  /* Same text as synth.c:663-808 so `?`, `\` and patch re-emission round-trip. */
  if (out == NULL) return "(NULL)";
  if (bad_voice(v)) { out[0] = '\0'; return out; }
  if (verbose && g_engine) skb_shim_snapshot_range(v, 1);
  char *ptr = out;
  APPEND("v%d w%d f%g a%g", v, voice_wave_table_index[v], voice_freq[v], voice_user_amp[v]);
  if (verbose || voice_midi_transpose[v]) APPEND(" N%g", voice_midi_transpose[v]);
  if (verbose || voice_link_midi_a[v] >= 0 || voice_link_midi_b[v] >= 0) APPEND(" G%g,%g", voice_link_midi_a[v], voice_link_midi_b[v]);
  if (verbose || voice_link_velo_a[v] >= 0 || voice_link_velo_b[v] >= 0) APPEND(" H%g,%g", voice_link_velo_a[v], voice_link_velo_b[v]);
  if (verbose || voice_link_trig[v] >= 0) APPEND(" L%g", voice_link_trig[v]);
  if (verbose || voice_direction[v]) APPEND(" b%d", voice_direction[v]);
  if (verbose || voice_loop_enabled[v]) APPEND(" B%d", voice_loop_enabled[v]);
  if (verbose || voice_pan[v]) APPEND(" p%g", voice_pan[v]);
  if (verbose || voice_note[v]) APPEND(" n%g", voice_note[v]);
  if (verbose || voice_filter_mode[v]) APPEND(" J%d K%g Q%g", voice_filter_mode[v], voice_filter_freq[v], voice_filter_res[v]);
  if (verbose || voice_cz_mode[v]) APPEND(" c%d,%g", voice_cz_mode[v], voice_cz_distortion[v]);
  if (verbose || voice_quantize[v]) APPEND(" q%d", voice_quantize[v]);
  if (verbose || voice_sample_hold_max[v]) APPEND(" h%d", voice_sample_hold_max[v]);
  if (verbose || (voice_amp_mod_osc[v] >= 0 && voice_amp_mod_depth[v] > 0)) APPEND(" A%d,%g", voice_amp_mod_osc[v], voice_amp_mod_depth[v]);
  if (verbose || (voice_cz_mod_osc[v] >= 0 && voice_cz_mod_depth[v] > 0)) APPEND(" C%d,%g", voice_cz_mod_osc[v], voice_cz_mod_depth[v]);
  if (verbose || (voice_freq_mod_osc[v] >= 0 && voice_freq_mod_depth[v] > 0)) APPEND(" F%d,%g", voice_freq_mod_osc[v], voice_freq_mod_depth[v]);
  if (verbose || (voice_pan_mod_osc[v] >= 0 && voice_pan_mod_depth[v] > 0)) APPEND(" P%d,%g", voice_pan_mod_osc[v], voice_pan_mod_depth[v]);
  if (verbose || voice_disconnect[v]) APPEND(" m%d", voice_disconnect[v]);
  if (verbose || voice_record[v]) APPEND(" r%d", voice_record[v]);
  if ((verbose || voice_smoother_enable[v]) && voice_smoother_smoothing[v] != SHIM_SMOOTH_DEFAULT) APPEND(" s%g", voice_smoother_smoothing[v]);
  if (verbose || voice_glissando_enable[v]) APPEND(" g%g", voice_glissando_speed[v]);
  if (verbose || !envelope_is_flat(v))
    APPEND(" t%g,%g,%g,%g", voice_amp_envelope[v].a, voice_amp_envelope[v].d, voice_amp_envelope[v].s, voice_amp_envelope[v].r);
  if (verbose) {
    APPEND("\n#");
    APPEND(" freq_scale:%g", voice_freq_scale[v]);
    APPEND(" finished:%d one_shot:%d", voice_finished[v], voice_one_shot[v]);
    APPEND(" sample:%g", voice_sample[v]);
    APPEND(" smoother:%g", voice_smoother_gain[v]);
    APPEND(" phase:%g phase_inc:%g", voice_phase[v], voice_phase_inc[v]);
    APPEND(" offset_hz:%g", voice_offset_hz[v]);
    APPEND(" latency:%gms", (double)ts_ns(&voice_mark_a[v], &voice_mark_b[v]) / 1000000.0);
  }
  return out;
}

void voice_show(int v, char c, int verbose) {
  char s[1024];
  voice_format(v, s, verbose);
  if (strlen(s)) printf("; %s%s\n", s, (c != ' ') ? " # *" : "");
}

int voice_show_all(int voice, int verbose) {
  for (int i = 0; i < VOICE_MAX; i++)
    if (voice_amp[i] != 0) voice_show(i, (i == voice) ? '*' : ' ', verbose);
  return 0;
}

/* ======================================================================== */
/* wave tables (synth.c:1171-1307)                                           */
/* ======================================================================== */
#include "retro/korg.h"
#include "amysamples.h"

#define SHIM_BUILTIN_SIZE (4096)

void normalize_preserve_zero(float *data, int length) {          /* synth.c:1175-1197 */
  if (length == 0) return;
  float peak = 0.0f;
  for (int i = 0; i < length; i++) { float a = fabsf(data[i]); if (a > peak) peak = a; }
  if (peak == 0.0) return;
  const float k = 1.0f / peak;
  for (int i = 0; i < length; i++) data[i] *= k;
}

static void slot_plain(int w, float *t, int size) {
  wave_table_data[w] = t;
  wave_size[w] = size;
  wave_rate[w] = MAIN_SAMPLE_RATE;
  wave_one_shot[w] = 0;
  wave_loop_start[w] = 0;
  wave_loop_end[w] = size - 1;
}

void wave_table_init(void) {
  for (int i = 0; i < WAVE_TABLE_MAX; i++) { wave_table_data[i] = NULL; wave_size[i] = 0; wave_is_miniwav[i] = 0; }

  /* w0..w6: closed forms sampled with an ACCUMULATED float phase (synth.c:1231-1248) */
  uint64_t noise_rng;
  audio_rng_init(&noise_rng, 1);
  for (int w = WAVE_TABLE_SINE; w <= WAVE_TABLE_NOISE_ALT; w++) {
    float *t = (float *)malloc(SHIM_BUILTIN_SIZE * sizeof(float));
    slot_plain(w, t, SHIM_BUILTIN_SIZE);
    const float step = 1.0f / (float)SHIM_BUILTIN_SIZE;
    int k = 0;
    for (float x = 0; x < 1.0f; x += step) {
      float s = sinf(2.0f * (float)M_PI * x);
      float y = 0;
      switch (w) {
        case WAVE_TABLE_SINE: y = s; break;
        case WAVE_TABLE_SQR: y = (x < 0.5) ? 1.0f : -1.0f; break;
        case WAVE_TABLE_SAW_DOWN: y = 2.0f * x - 1.0f; break;
        case WAVE_TABLE_SAW_UP: y = 1.0f - 2.0f * x; break;
        case WAVE_TABLE_TRI: y = (x < 0.5f) ? (4.0f * x - 1.0f) : (3.0f - 4.0f * x); break;
        case WAVE_TABLE_NOISE:
        case WAVE_TABLE_NOISE_ALT: y = audio_rng_float(&noise_rng); break;
      }
      t[k++] = y;
    }
  }

  /* w32..w62: 31 Korg single-cycle tables, int16 / 32767 (synth.c:1255-1268) */
  korg_init();
  for (int w = WAVE_TABLE_KRG1; w < WAVE_TABLE_KRG32; w++) {
    const int k = w - WAVE_TABLE_KRG1, n = kwave_size[k];
    float *t = (float *)malloc((size_t)n * sizeof(float));
    for (int j = 0; j < n; j++) t[j] = (float)kwave[k][j] / (float)32767;
    slot_plain(w, t, n);
  }

  /* w100..: AMY PCM one-shots at 22,050 Hz, peak-normalised (synth.c:1270-1292) */
  for (int i = 0; i < PCM_SAMPLES; i++) {
    const int w = i + AMY_SAMPLE_00;
    if (w > AMY_SAMPLE_99 - 1) break;
    const int n = (int)pcm_map[i].length;
    float *t = (float *)malloc((size_t)n * sizeof(float));
    for (int k = 0; k < n; k++) t[k] = (float)pcm[pcm_map[i].offset + k] / 32767.0f;
    normalize_preserve_zero(t, n);
    wave_table_data[w] = t;
    wave_size[w] = n;
    wave_rate[w] = PCM_AMY_SAMPLE_RATE;
    wave_one_shot[w] = 1;
    wave_loop_enabled[w] = 0;
    wave_loop_start[w] = (int)pcm_map[i].loopstart;
    wave_loop_end[w] = (int)pcm_map[i].loopend;
    wave_midi_note[w] = (int)pcm_map[i].midinote;
    wave_offset_hz[w] = midi2hz((float)pcm_map[i].midinote);
  }
}

void wave_free(void) {
  for (int i = 0; i < WAVE_TABLE_MAX; i++) {
    if (!wave_table_data[i]) continue;
    if (wave_is_miniwav[i]) mw_free(wave_table_data[i]); else free(wave_table_data[i]);
    wave_table_data[i] = NULL;
    wave_size[i] = 0;
  }
}

void synth_init(void) {
  extern int debug;
  if (debug) {
#define ARRAY(type, name, size, init) printf("%s : %d\n", #name, name##__len__);
#include "synth.def"
#undef ARRAY
  }
}

void synth_free(void) {
  if (g_engine) { skb_destroy(g_engine); g_engine = NULL; }
}

/* ======================================================================== */
/* callback statistics (synth.c:435-500)                                     */
/* ======================================================================== */
#define SHIM_BENCH_SLOTS 16
static struct { struct timespec a, b; int state, frames; int64_t order; } g_bench[SHIM_BENCH_SLOTS];
static int64_t g_bench_n = 0;
static char g_stats_text[4096];

char *synth_stats(void) {
  char *ptr = g_stats_text;
  *ptr = '\0';
  for (int i = 0; i < SHIM_BENCH_SLOTS; i++) {
    if (g_bench[i].state != 2) continue;
    double budget_ms = (double)g_bench[i].frames / (double)MAIN_SAMPLE_RATE * 1000.0;
    double took_ms = ts_ns(&g_bench[i].a, &g_bench[i].b) / 1e6;
    APPEND("# %d %d %gms %gms\n", (int)g_bench[i].order, g_bench[i].frames, took_ms, budget_ms);
    g_bench[i].state = 0;
  }
  return g_stats_text;
}

static int g_any_mark = 0;

void synth_voice_bench(int voice) {
  g_any_mark = 1;
  voice_mark_b[voice].tv_sec = 0;
  voice_mark_b[voice].tv_nsec = 0;
  clock_gettime(CLOCK_MONOTONIC, &voice_mark_a[voice]);
  voice_mark_go[voice] = 1;
}

/* ======================================================================== */
/* the render entry point                                                    */
/* ======================================================================== */
static void ensure_traces(int n) {
  if (n <= g_trace_cap) return;
  g_trace_cap = n;
  g_gain = (float *)realloc(g_gain, (size_t)n * sizeof(float));
  g_noise = (float *)realloc(g_noise, (size_t)n * sizeof(float));
}

/* Advance the two voice-independent recurrences of the frame loop for
 * `n` frames: the shared noise draw (synth.c:525) and the master-volume
 * one-pole (synth.c:616-617).  Returns non-zero if any voice needs noise. */
static int g_gain_fill = 0;   /* split form: gain frames rendered but not yet finished */

static int step_traces(int n, int append) {
  const int base = append ? g_gain_fill : 0;
  ensure_traces(base + n);
  if (!g_rng_seeded) { audio_rng_init(&g_rng, 1); g_rng_seeded = 1; }      /* synth.c:508 */
  if (g_noise_voices > 0) {
    for (int i = 0; i < n; i++) g_noise[i] = audio_rng_float(&g_rng);
  } else {
    /* nobody listens: the reference still draws once per frame (synth.c:525), so the generator
     * must end up n steps further — x -> a^n x + c (a^n - 1)/(a - 1) (mod 2^64) by squaring */
    uint64_t am = 6364136223846793005ULL, cm = 1442695040888963407ULL, an = 1, cn = 0;
    for (unsigned k = (unsigned)n; k; k >>= 1) {
      if (k & 1) { an *= am; cn = cn * am + cm; }
      cm = (am + 1) * cm;
      am *= am;
    }
    g_rng = an * g_rng + cn;
  }
  float g = volume_smoother_gain;
  const float k = volume_smoother_smoothing, target = volume_final;
  /* synth.c:616-620.  Once a step no longer moves g (the one-pole has reached its float fixed point) no later step
   * does either — same operands, same result — so the rest of the block is that constant: the 12-cycle dependent chain
   * per frame (0.03 ms per 8,192 frames on the calling thread) is only walked while the volume is actually moving. */
  int i = 0;
  for (; i < n; i++) {
    const float gn = g + k * (target - g);
    if (gn == g && !signbit(gn) == !signbit(g)) break;
    g = gn;
    g_gain[base + i] = g;
  }
  for (; i < n; i++) g_gain[base + i] = g;
  volume_smoother_gain = g;
  g_gain_fill = base + n;
  return 1;
}

static void mark_latency(void) {
  if (!g_any_mark) return;
  g_any_mark = 0;
  for (int v = 0; v < VOICE_MAX; v++)
    if (voice_mark_go[v]) { clock_gettime(CLOCK_MONOTONIC, &voice_mark_b[v]); voice_mark_go[v] = 0; }
}

/* ---- timestamped event queue ------------------------------------------------
 * The binary twin of the reference's work_queue[] + seq() (seq.c:164-178,
 * 241-257): events carry an absolute sample time; after every callback-sized
 * sub-block ending at count c they fire when `when <= c + frames`
 * by calling the ordinary setters, so they reach the device as parameter
 * records / ordered ops at that block boundary.  Unlike work_queue (1,024
 * strings, silently dropping when full) the queue is unbounded. */
static skb_event *g_evq = NULL;
static size_t g_evq_head = 0, g_evq_n = 0, g_evq_cap = 0;
static uint64_t g_events_fired = 0;

int skb_shim_queue_events(const skb_event *ev, int n) {
  if (n < 0 || (n > 0 && !ev)) return SKB_ERR_ARG;
  if (g_evq_head > 0 && g_evq_head == g_evq_n) g_evq_head = g_evq_n = 0;
  if (g_evq_n + (size_t)n > g_evq_cap) {
    g_evq_cap = (g_evq_n + (size_t)n) * 2 + 1024;
    g_evq = (skb_event *)realloc(g_evq, g_evq_cap * sizeof(skb_event));
  }
  /* keep the queue ordered by `when`, first-come first-served among equals
   * (seq() scans work_queue in slot order; equal-time items fire in queue order) */
  for (int i = 0; i < n; i++) {
    size_t j = g_evq_n;
    while (j > g_evq_head && g_evq[j - 1].when > ev[i].when) { g_evq[j] = g_evq[j - 1]; j--; }
    g_evq[j] = ev[i];
    g_evq_n++;
  }
  return SKB_OK;
}

int skb_shim_pending_events(void) { return (int)(g_evq_n - g_evq_head); }

static void fire_event(const skb_event *e) {
  const int v = e->voice;
  if (v < 0 || v >= VOICE_MAX) return;
  switch (e->code) {
    case SKB_EV_TRIGGER:     voice_trigger(v); if (voice_link_trig[v] > 0) voice_trigger(voice_link_trig[v]); break;  /* wire.c:710-714 */
    case SKB_EV_VELOCITY:    envelope_velocity(v, e->a0);                                                     /* wire.c:674-679 */
                             if (voice_link_velo_a[v] >= 0) envelope_velocity(voice_link_velo_a[v], e->a0);
                             if (voice_link_velo_b[v] >= 0) envelope_velocity(voice_link_velo_b[v], e->a0); break;
    case SKB_EV_FREQ:        freq_set(v, e->a0); break;
    case SKB_EV_MIDI:        freq_midi(v, e->a0);                                                             /* wire.c:682-687 */
                             if (voice_link_midi_a[v] >= 0) freq_midi(voice_link_midi_a[v], e->a0);
                             if (voice_link_midi_b[v] >= 0) freq_midi(voice_link_midi_b[v], e->a0); break;
    case SKB_EV_AMP:         amp_set(v, e->a0); break;
    case SKB_EV_PAN:         pan_set(v, e->a0); break;
    case SKB_EV_WAVE:        wave_set(v, (int)e->a0); break;
    case SKB_EV_CZ:          cz_set(v, (int)e->a0, e->a1); break;
    case SKB_EV_FILTER_FREQ: mmf_set_freq(v, e->a0); break;
    case SKB_EV_FILTER_RES:  mmf_set_res(v, e->a0); break;
    case SKB_EV_MUTE:        wave_mute(v, (int)e->a0); break;
    default: break;
  }
}

/* seq()'s firing rule for the callback that just ended (seq.c:171-178). */
static void fire_due(int frame_count) {
  const uint64_t horizon = synth_sample_count + (uint64_t)frame_count;
  while (g_evq_head < g_evq_n && g_evq[g_evq_head].when <= horizon) {
    fire_event(&g_evq[g_evq_head++]);
    g_events_fired++;
  }
}

/* First sub-block boundary (relative to the call start, in (done, num_frames])
 * after which the head of the queue fires; num_frames + 1 if none does. */
static int next_firing_boundary(int done, int num_frames, uint64_t start_count) {
  if (g_evq_head >= g_evq_n) return num_frames + 1;
  const int EB = SYNTH_FRAMES_PER_CALLBACK;
  const uint64_t w = g_evq[g_evq_head].when;
  for (int b = (done / EB + 1) * EB;; b += EB) {
    const int end = b < num_frames ? b : num_frames;
    const int len = end - (b - EB);
    if (w <= start_count + (uint64_t)end + (uint64_t)len) return end;
    if (end == num_frames) break;
  }
  return num_frames + 1;
}

/* host-side time of synth(), seconds, accumulated: [0] flush + trace stepping, [1] skb_render_mix (queueing),
 * [2] firing the due events (setters -> ops), [3] skb_finish (launch, kernels, D2H, sync) + tap readback */
static double g_shim_time[4];
static double shim_now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
void skb_shim_timing(double *out4, int reset) {
  for (int i = 0; i < 4; i++) { if (out4) out4[i] = g_shim_time[i]; if (reset) g_shim_time[i] = 0.0; }
}

/* ---- selective per-voice tap ----------------------------------------------------------------------
 * The reference writes one_skred_frame[frame][voice][L, R] for EVERY voice every frame (synth.c:533-611) although only
 * the voices with voice_record[v] set are ever written to disk (wire.c:94-185, 698): 8 bytes per voice-sample, all of
 * it over PCIe here.  With skb_shim_tap_selective(1) (or SKB_TAP_SELECTIVE=1) the device keeps the full tap but only the
 * columns of the recorded voices come back; the other voices' entries of `user` stay 0 except the FIRST unrecorded
 * voice of frame 0 of a call, which carries (min(0, x), max(0, x)) over all the samples that did not come back:
 * save_wav scales by the extremes of every voice of the recording (wire.c:150-166) and skips unrecorded voices when it
 * writes, so the WAV file is the reference's byte for byte (tests). */
static int g_tap_selective = -1;                  /* -1 = read SKB_TAP_SELECTIVE on first use */
static int *g_tap_sel = NULL, g_tap_nsel = 0, g_tap_carrier = -1;
static unsigned char *g_tap_was = NULL;           /* voice_record[] as of the last call */
static float *g_tap_compact = NULL; static size_t g_tap_compact_cap = 0;
static int g_tap_frames_seen = 0;
void skb_shim_tap_selective(int on) { g_tap_selective = on ? 1 : 0; }

static void tap_sync_selection(float *tap_user) {
  if (!g_tap_was) { g_tap_was = (unsigned char *)calloc(VOICE_MAX, 1); g_tap_sel = (int *)malloc(sizeof(int) * VOICE_MAX); g_tap_nsel = -1; }
  int changed = g_tap_nsel < 0;
  for (int v = 0; v < VOICE_MAX && !changed; v++) changed = (voice_record[v] != 0) != (g_tap_was[v] != 0);
  if (!changed) return;
  g_tap_nsel = 0; g_tap_carrier = -1;
  for (int v = 0; v < VOICE_MAX; v++) {
    const int on = voice_record[v] != 0;
    if (on) g_tap_sel[g_tap_nsel++] = v; else if (g_tap_carrier < 0) g_tap_carrier = v;
    if (!on && g_tap_was[v])                       /* no longer recorded: its column reads 0 again */
      for (int f = 0; f < g_tap_frames_seen; f++) tap_user[((size_t)f * VOICE_MAX + v) * 2] = tap_user[((size_t)f * VOICE_MAX + v) * 2 + 1] = 0.0f;
    g_tap_was[v] = (unsigned char)on;
  }
  /* an empty selection would mean "all" to the engine: nothing recorded = a selection of one dummy column (the carrier) */
  if (skb_set_tap_voices(g_engine, g_tap_sel, g_tap_nsel) != SKB_OK) shim_die("skb_set_tap_voices");
}

static void tap_read_selected(float *tap_user, int chunk0, int frames) {
  if (g_tap_nsel == VOICE_MAX) {                   /* every voice is recorded: the plain copy */
    if (skb_read_tap(g_engine, 0, frames, tap_user + (size_t)chunk0 * VOICE_MAX * 2) != SKB_OK) shim_die("skb_read_tap");
    return;
  }
  float ext[2] = {0.0f, 0.0f};
  if (g_tap_nsel > 0) {
    const size_t need = (size_t)frames * g_tap_nsel * 2;
    if (need > g_tap_compact_cap) { g_tap_compact_cap = need * 2; g_tap_compact = (float *)realloc(g_tap_compact, g_tap_compact_cap * sizeof(float)); }
    if (skb_read_tap_selected(g_engine, 0, frames, g_tap_compact, ext) != SKB_OK) shim_die("skb_read_tap_selected");
    for (int f = 0; f < frames; f++) {
      float *dst = tap_user + (size_t)(chunk0 + f) * VOICE_MAX * 2;
      const float *src = g_tap_compact + (size_t)f * g_tap_nsel * 2;
      for (int i = 0; i < g_tap_nsel; i++) { dst[2 * g_tap_sel[i]] = src[2 * i]; dst[2 * g_tap_sel[i] + 1] = src[2 * i + 1]; }
    }
  }
  /* nothing selected: nothing is written to disk either (save_wav returns early, wire.c:106-109): no read-back at all */
  if (g_tap_carrier >= 0 && g_tap_nsel > 0) {
    float *c = tap_user + ((size_t)chunk0 * VOICE_MAX + g_tap_carrier) * 2;
    c[0] = ext[0]; c[1] = ext[1];
  }
  if (chunk0 + frames > g_tap_frames_seen) g_tap_frames_seen = chunk0 + frames;
}

/* skb_shim_synth_between: the sequencer timeline walked AHEAD of the audio.  The reference's audio callback
 * is `synth(512 frames); seq(512);` (skred.c:116-119): seq() fires deferred wire strings and pattern steps, which call
 * the setters, which take effect from the next callback on.  Nothing in seq() depends on the rendered audio, so a batch
 * host can run it for callback k + 1, k + 2, ... before callback k has left the GPU: this entry point renders
 * num_frames as callbacks of SYNTH_FRAMES_PER_CALLBACK frames, calls `between(frame_count)` where skred.c calls seq()
 * after each of them, and finishes ONCE — the engine turns the callbacks into one launch as long as `between` only
 * produced state edits (triggers, envelope gates, pans), and into a few launches where it changed parameters.
 * Same audio and state as the callback loop, bit for bit (tests: the shipped pattern patches). */
static void (*g_between)(int frame_count) = NULL;
void synth(float *buffer, float *input, int num_frames, int num_channels, void *user);
void skb_shim_synth_between(float *buffer, int num_frames, int num_channels, void *user, void (*between)(int frame_count)) {
  g_between = between;
  synth(buffer, NULL, num_frames, num_channels, user);
  g_between = NULL;
}

void synth(float *buffer, float *input, int num_frames, int num_channels, void *user) {
  (void)input;
  /* `user`: the per-voice tap one_skred_frame[frame][voice][L,R], latched on the FIRST call only like the
   * reference does (synth.c:503-511).  NULL = no tap (the reference would fault; skred.c always passes one). */
  static int first = 1;
  static float *tap_user = NULL;
  if (first) {
    synth_frames_per_callback = num_frames;
    tap_user = (float *)user;
    first = 0;
    if (tap_user && skb_set_tap(engine(), 1) != SKB_OK) shim_die("skb_set_tap");
    if (g_tap_selective < 0) { const char *s = getenv("SKB_TAP_SELECTIVE"); g_tap_selective = (s && atoi(s)) ? 1 : 0; }
  }
  if (tap_user && g_tap_selective == 1) { engine(); tap_sync_selection(tap_user); }
  const int slot = (int)(g_bench_n % SHIM_BENCH_SLOTS);
  clock_gettime(CLOCK_MONOTONIC, &g_bench[slot].a);
  g_bench[slot].frames = num_frames; g_bench[slot].order = g_bench_n; g_bench[slot].state = 1;

  engine();
  mark_latency();
  const int EB = SYNTH_FRAMES_PER_CALLBACK;
  const uint64_t start_count = synth_sample_count;
  float *d_mix = skb_mix_buffer(g_engine);
  int done = 0;
  while (done < num_frames) {
    /* one CHUNK of <= max_frames: its event-free segments are queued on the device back to
     * back (the host fires the events of segment k+1 while the GPU renders segment k) and
     * the chunk is finished — master volume, D2H, one synchronisation — at its end */
    const int chunk0 = done;
    const int chunk1 = (num_frames - done > g_cfg_max_frames) ? done + g_cfg_max_frames : num_frames;
    g_gain_fill = 0;
    int early = 0;
    while (done < chunk1) {
      /* render up to the next boundary at which a queued event fires (everything in between
       * is event-free, so one long launch equals many callbacks) */
      const double t_a = shim_now();
      int end = next_firing_boundary(done, num_frames, start_count);
      if (g_between) {                          /* the host's own seq() runs after every callback: every boundary counts */
        const int nb = (done / EB + 1) * EB;
        end = nb < num_frames ? nb : num_frames;
      }
      const int fires = end <= chunk1;
      if (!fires) end = chunk1;
      if (skb_shim_flush() != SKB_OK) shim_die("flush");
      const int n = end - done;
      step_traces(n, 1);
      const double t_b = shim_now();
      if (skb_render_mix(g_engine, n, synth_sample_count, g_noise, d_mix + (size_t)(done - chunk0) * 2, NULL) != SKB_OK)
        shim_die("skb_render_mix");
      const double t_c = shim_now();
      synth_sample_count += (uint64_t)n;
      if (fires && n > 0) {
        const int sub = (end % EB) ? (end % EB) : EB;     /* length of the sub-block that just ended */
        fire_due(sub);
        if (g_between) g_between(sub);          /* skred.c:119: seq(frame_count) right after the callback's synth() */
      }
      done = end;
      /* option: hand the GPU its first callbacks now, so that it renders them while the host fires and queues
       * the events of the rest (one deferred launch leaves it idle for ~95 us of host work per 4,096-frame
       * call); off by default, see g_early_flush_frames */
      if (!early && g_early_flush_frames > 0 && done - chunk0 >= g_early_flush_frames && chunk1 - done >= g_early_flush_rest) {
        if (skb_flush(g_engine) != SKB_OK) shim_die("skb_flush");
        early = 1;
      }
      g_shim_time[0] += t_b - t_a; g_shim_time[1] += t_c - t_b; g_shim_time[2] += shim_now() - t_c;
    }
    const double t_f = shim_now();
    if (skb_finish(g_engine, d_mix, chunk1 - chunk0, g_gain, buffer + (size_t)chunk0 * num_channels, num_channels, NULL) != SKB_OK)
      shim_die("skb_finish");
    if (tap_user && g_tap_selective == 1) tap_read_selected(tap_user, chunk0, chunk1 - chunk0);
    else if (tap_user && skb_read_tap(g_engine, 0, chunk1 - chunk0, tap_user + (size_t)chunk0 * VOICE_MAX * 2) != SKB_OK)
      shim_die("skb_read_tap");                    /* synth.c:533-611 */
    g_gain_fill = 0;
    g_shim_time[3] += shim_now() - t_f;
  }
  clock_gettime(CLOCK_MONOTONIC, &g_bench[slot].b);
  g_bench[slot].state = 2;
  g_bench_n++;
}

/* Split form of synth() for multi-GPU hosts: flush + render this engine's
 * voices into DEVICE memory d_mix[num_frames][2]; after the caller reduced the
 * partial mixes over NVLink, skb_shim_finish applies the master volume on the
 * root.  Both advance / consume the host traces exactly like synth().  Several
 * render_mix calls (one per 512-frame callback) may precede one finish over
 * their concatenated frames (batch mode reduces once per multi-block chunk). */
int skb_shim_render_mix(int num_frames, float *d_mix, void *stream) {
  engine();
  if (num_frames > g_cfg_max_frames) return SKB_ERR_ARG;
  if (skb_shim_flush() != SKB_OK) return skb_last_error(g_engine);
  step_traces(num_frames, 1);
  int r = skb_render_mix(g_engine, num_frames, synth_sample_count, g_noise, d_mix, stream);
  synth_sample_count += (uint64_t)num_frames;
  fire_due(num_frames);        /* seq()'s rule for this callback */
  return r;
}

/* `ncalls` consecutive callbacks of `frames_per_call` frames into d_mix (frames_per_call * ncalls * 2 floats),
 * then launch: what a host loop over skb_shim_render_mix + skb_shim_flush_render does, in one call. */
int skb_shim_render_calls(int frames_per_call, int ncalls, float *d_mix, void *stream) {
  for (int k = 0; k < ncalls; k++) {
    const int r = skb_shim_render_mix(frames_per_call, d_mix + (size_t)k * frames_per_call * 2, stream);
    if (r != SKB_OK) return r;
  }
  return skb_flush(engine());
}

void skb_shim_discard_gain(void) { g_gain_fill = 0; }

/* launch whatever skb_shim_render_mix deferred (before the caller consumes d_mix on its stream) */
int skb_shim_flush_render(void) { return skb_flush(engine()); }

int skb_shim_finish(const float *d_mix, int num_frames, float *out, int num_channels, void *stream) {
  /* consumes the gain trace of every skb_shim_render_mix since the last finish */
  if (num_frames != g_gain_fill) return SKB_ERR_STATE;
  g_gain_fill = 0;
  return skb_finish(engine(), d_mix, num_frames, g_gain, out, num_channels, stream);
}
