/* oracle/ref_harness.c — TEST INFRASTRUCTURE (see oracle/README.md).
 *
 * Our glue, linked together with the UNMODIFIED reference translation units
 * (synth.c wire.c seq.c skode.c amysamples.c, compiled where they lie under
 * /root/reference by oracle/build_oracle.py) into
 * oracle/_ref/libskred_ref_v<VOICE_MAX>.so.
 *
 * It provides
 *   1. the host symbols the reference path imports from skred.c
 *      (skred.c:38-56, 84-90; SURVEY §8b "Symbols the path imports"),
 *   2. an offline render loop in the reference's own callback order
 *      `synth(); seq();` (skred.c:116-119) with block = 512,
 *   3. small accessors so Python (ctypes) can drive it.
 *
 * The same file, compiled with -DSKB_DROPIN, glues the reference's wire.c /
 * seq.c / skode.c on top of the PRODUCT's synth.h implementation
 * (skred_b200/csrc/synth_shim.c) — the drop-in proof used by tests.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "skred.h"
#include "synth-types.h"
#include "synth.h"
#include "wire.h"
#include "seq.h"
#include "miniwav.h"
#include "scope-shared.h"

/* ---- 1. host symbols normally defined by skred.c ---------------------- */
int scope_enable = 0;
scope_buffer_t scope_safety;
scope_buffer_t *scope = &scope_safety;
float tempo_time_per_step = 60.0f;      /* skred.c:47 */
float tempo_bpm = 120.0f / 4.0f;        /* skred.c:48 */
float tempo_base = 0.0f;                /* skred.c:49 */
int debug = 0;
int trace = 0;
int console_voice = 0;
int main_running = 1;
int rec_state = 0;                      /* skred.c:84-90 */
long rec_ptr = 0;
float rec_sec = (float)REC_IN_SEC;
long rec_max = 0;
float *recording = NULL;

void util_set_thread_name(char *s) { (void)s; }
char *udp_info(void) { return ""; }
int udp_port = 0;

/* WAV loading, `:wN,slot[,ch]` (wire.c:406-441 -> mw_get, miniwav.c:103-147, which decodes through
 * miniaudio).  miniaudio (95 k lines) is not compiled into the harness; this is OUR reader for what it would
 * return for the files the reference ships — RIFF/WAVE integer PCM 16-bit, decoded to interleaved float32 with
 * miniaudio's scaling x * 2^-15 (miniaudio.h:45560) — followed by mw_get's own channel handling restated as
 * it is written, quirks included: with ch == -1 (the default of `:w`) nothing is stored, so the table is the
 * first `frames` floats of the INTERLEAVED data (miniwav.c:126-137); ch > channels reads one past the frame.
 * The same glue serves the compiled reference and the drop-in builds, so it cancels in the parity tests. */
static uint32_t rd_u32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t rd_u16(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
float *mw_get(char *name, int *frames_out, wav_t *w, int ch) {
  if (frames_out) *frames_out = 0;
  FILE *f = fopen(name, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  unsigned char *b = (unsigned char *)malloc(sz > 0 ? (size_t)sz : 1);
  if (sz < 44 || fread(b, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(b); return NULL; }
  fclose(f);
  if (memcmp(b, "RIFF", 4) || memcmp(b + 8, "WAVE", 4)) { free(b); return NULL; }
  int channels = 0, rate = 0, bits = 0, fmt = 0;
  const unsigned char *data = NULL;
  uint32_t data_len = 0;
  for (long at = 12; at + 8 <= sz;) {
    const uint32_t len = rd_u32(b + at + 4);
    if (!memcmp(b + at, "fmt ", 4) && len >= 16) {
      fmt = (int)rd_u16(b + at + 8); channels = (int)rd_u16(b + at + 10); rate = (int)rd_u32(b + at + 12); bits = (int)rd_u16(b + at + 22);
    } else if (!memcmp(b + at, "data", 4)) {
      data = b + at + 8;
      data_len = (at + 8 + (long)len <= sz) ? len : (uint32_t)(sz - at - 8);
      break;
    }
    at += 8 + (long)len + (len & 1);
  }
  if (!data || fmt != 1 || bits != 16 || channels < 1) { free(b); return NULL; }
  const long frameCount = (long)(data_len / 2) / channels;
  float *pSamples = (float *)malloc((size_t)(frameCount * channels + 1) * sizeof(float));
  for (long i = 0; i < frameCount * channels; i++) {
    const int16_t v = (int16_t)rd_u16(data + 2 * i);
    pSamples[i] = (float)v * 0.000030517578125f;
  }
  pSamples[frameCount * channels] = 0.0f;
  free(b);
  /* miniwav.c:126-137 */
  int j = 0;
  if (ch > channels) ch = channels;
  for (long i = 0; i < frameCount * channels; i += channels) {
    if (ch == -1) {
      float a = 0;
      for (int k = 0; k < ch; k++) a += pSamples[i + k];
      a /= (float)ch;
    } else {
      pSamples[j] = pSamples[i + ch];
    }
    j++;
  }
  w->SamplesRate = rate;
  w->Channels = channels;
  *frames_out = (int)frameCount;
  return pSamples;
}
float *mw_free(float *f) { free(f); return NULL; }

/* ---- 2. offline driver ------------------------------------------------- */
static float *g_tap = NULL;
static int g_tap_frames = 0;
static wire_t g_wire = WIRE();

int ref_voice_max(void) { return VOICE_MAX; }

/* Same order as main(): skred.c:232-236. */
void ref_init(void) {
  static int once = 0;
  if (once) return;
  once = 1;
  /* the per-voice tap `user` must hold num_frames*VOICE_MAX*2 floats
   * (synth.c:533-611); it is latched on the first synth() call. */
#ifndef SKB_DROPIN
  /* synth.c writes it unconditionally; the drop-in treats NULL as "no tap" (ref_enable_tap) */
  g_tap_frames = SYNTH_FRAMES_PER_CALLBACK;
  g_tap = (float *)calloc((size_t)g_tap_frames * VOICE_MAX * 2, sizeof(float));
#endif
  synth_init();
  wave_table_init();
  voice_init();
  seq_init();
  g_wire.printf = null_printf;
  g_wire.puts = null_puts;
}

float *ref_tap(void) { return g_tap; }
int ref_tap_frames(void) { return g_tap_frames; }

/* Size the per-voice tap for callbacks of up to `frames` frames.  Call before the first
 * ref_render: synth() latches the pointer on its first call (synth.c:503-511). */
float *ref_enable_tap(int frames) {
  free(g_tap);
  g_tap_frames = frames;
  g_tap = (float *)calloc((size_t)frames * VOICE_MAX * 2, sizeof(float));
  return g_tap;
}

/* Feed one line of skode to the reference parser (wire.c:924). */
int ref_wire(const char *line) {
  char buf[4096];
  strncpy(buf, line, sizeof(buf) - 1);
  buf[sizeof(buf) - 1] = '\0';
  return wire(buf, &g_wire);
}

/* Load "<dir>/<n>.sk" through the reference's own sk_load (wire.c:342). */
int ref_sk_load(const char *dir, int n) {
  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd))) return -1;
  if (chdir(dir) != 0) return -2;
  int r = sk_load(&g_wire, 0, n, 0);
  if (chdir(cwd) != 0) return -3;
  return r;
}

/* Render `nframes` frames as ceil(nframes/block) callbacks; returns seconds
 * of wall time spent inside the loop (CPU baseline).  out: nframes*2 f32. */
double ref_render(float *out, long nframes, int block, int run_seq) {
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  long done = 0;
  while (done < nframes) {
    int n = (nframes - done) < block ? (int)(nframes - done) : block;
    synth(out + done * 2, NULL, n, 2, g_tap);
    if (run_seq) seq(n);
    done += n;
  }
  clock_gettime(CLOCK_MONOTONIC, &b);
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}

uint64_t ref_sample_count(void) { return synth_sample_count; }

/* The same job as ref_render(out, nframes, 512, 1) handed to the drop-in as ONE call with seq() as the per-callback hook
 * (skb_shim_synth_between): the sequencer runs ahead of the audio, the engine batches the callbacks.  The reference
 * build has no such entry point and runs the plain loop. */
#ifdef SKB_DROPIN
#include "skred_b200_shim.h"
#endif
static void between_seq(int frame_count) { seq(frame_count); }
double ref_render_batched(float *out, long nframes, int call_frames) {
#ifdef SKB_DROPIN
  long done = 0;
  while (done < nframes) {
    int n = (nframes - done) < call_frames ? (int)(nframes - done) : call_frames;
    skb_shim_synth_between(out + done * 2, n, 2, g_tap, between_seq);
    done += n;
  }
  return 0.0;
#else
  (void)call_frames;
  return ref_render(out, nframes, SYNTH_FRAMES_PER_CALLBACK, 1);
#endif
}

/* ---- recording (skred.c:84-131, wire.c `<sec` / `*`): what synth_callback does with the per-voice tap ---- */
/* synth_callback_init(max_sec), skred.c:93-100, with a test-sized buffer; call before the first ref_render*.
 * The tap is sized for callbacks of `block` frames (synth() latches it on its first call). */
void ref_record_init(float max_sec, int block) {
  free(recording);
  rec_sec = max_sec;
  rec_max = (long)(max_sec * (float)(MAIN_SAMPLE_RATE * AUDIO_CHANNELS * VOICE_MAX));
  recording = (float *)malloc((size_t)rec_max * sizeof(float));
  if (!g_tap || g_tap_frames < block) ref_enable_tap(block);
}

/* ref_render with the body of synth_callback (skred.c:116-131): synth(); seq(); then, while rec_state is set, the tap of
 * EVERY voice of the callback is appended to the recording. */
double ref_render_recording(float *out, long nframes, int block, int run_seq) {
  long done = 0;
  while (done < nframes) {
    int n = (nframes - done) < block ? (int)(nframes - done) : block;
    synth(out + done * 2, NULL, n, 2, g_tap);
    if (run_seq) seq(n);
    if (rec_state) {
      float *f = g_tap;
      for (long i = 0; i < (long)n * 2 * VOICE_MAX; i += 2) {
        if (rec_ptr < rec_max) {
          recording[rec_ptr++] = f[i];
          recording[rec_ptr++] = f[i + 1];
        } else {
          rec_state = 0;
          break;
        }
      }
    }
    done += n;
  }
  return 0.0;
}
long ref_rec_ptr(void) { return rec_ptr; }

/* Install a caller-owned float table into a wave slot the way data_load does
 * (wire.c:374-404) but with explicit loop/one-shot fields, so the dead
 * notamy LUTs can be exercised.  The table is copied. */
int ref_install_table(int slot, const float *data, int len, float rate,
                      int one_shot, int loop_start, int loop_end,
                      float midi_note, float offset_hz) {
  if (slot < 0 || slot >= WAVE_TABLE_MAX || len <= 0) return 100;
  float *t = (float *)malloc((size_t)len * sizeof(float));
  memcpy(t, data, (size_t)len * sizeof(float));
  wave_table_data[slot] = t;
  wave_size[slot] = len;
  wave_rate[slot] = rate;
  wave_one_shot[slot] = one_shot;
  wave_loop_enabled[slot] = 0;
  wave_loop_start[slot] = loop_start;
  wave_loop_end[slot] = loop_end;
  wave_midi_note[slot] = midi_note;
  wave_offset_hz[slot] = offset_hz;
  wave_is_miniwav[slot] = 0;
  return 0;
}

/* ---- 3. raw state access (struct fields ctypes cannot reach by symbol) -- */
void ref_get_filter(int v, float *out9) {
  mmf_t *f = &voice_filter[v];
  out9[0] = f->x1; out9[1] = f->x2; out9[2] = f->y1; out9[3] = f->y2;
  out9[4] = f->b0; out9[5] = f->b1; out9[6] = f->b2; out9[7] = f->a1; out9[8] = f->a2;
}
void ref_get_envelope(int v, float *f9, uint64_t *u2, int *active) {
  envelope_t *e = &voice_amp_envelope[v];
  f9[0] = e->a; f9[1] = e->d; f9[2] = e->s; f9[3] = e->r;
  f9[4] = e->attack_time; f9[5] = e->decay_time; f9[6] = e->sustain_level; f9[7] = e->release_time;
  f9[8] = e->velocity;
  u2[0] = e->sample_start; u2[1] = e->sample_release;
  active[0] = e->is_active;
}

/* Bring the public arrays up to date before Python reads them: a no-op for
 * the reference (the arrays ARE the state), a device->host snapshot for the
 * product drop-in. */
#ifdef SKB_DROPIN
#include "skred_b200_shim.h"
void ref_sync_state(void) { skb_shim_snapshot(); }
int ref_is_dropin(void) { return 1; }
#else
void ref_sync_state(void) {}
int ref_is_dropin(void) { return 0; }
#endif
