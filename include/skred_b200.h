/* skred_b200.h — C-ABI of the B200 voice-render engine (libskred_b200.so).
 *
 * This is the device boundary that replaces the body of the reference's
 *
 *     void synth(float *buffer, float *input, int num_frames,
 *                int num_channels, void *user);          (synth.h:8, synth.c:502-630)
 *
 * Plain C, plain pointers and sizes, no torch / C++ types.  The host side of
 * the drop-in (skred_b200/csrc/synth_shim.c) implements the reference's whole
 * synth.h / synth.def API in C and talks to the GPU only through the calls
 * below; INTEGRATION.md shows the binding a skred maintainer adds.
 *
 * Model.  The reference keeps ~65 parallel `voice_*[VOICE_MAX]` arrays
 * (synth.def:12-89) that the audio callback re-streams every frame.  Here a
 * voice is split into
 *   - PARAMETERS (skb_voice_params): everything the render loop only READS
 *     (SURVEY App. C "hot parameters").  The host owns them; it re-sends a
 *     voice's record whenever a setter changed it.  All transcendental math
 *     (sinf/cosf/powf for biquad coefficients, phase increments, midi->Hz;
 *     synth.c:125-136, 929-1008, 1056-1059) stays on the host so that the
 *     device never has to reproduce glibc's libm.
 *   - EVOLVING STATE (skb_voice_state): what the loop WRITES (SURVEY §8a row
 *     11).  It lives in HBM between blocks; the host changes it only through
 *     ordered device OPS (skb_op) applied at the next block boundary — the
 *     reference's own event granularity (seq() fires once per
 *     callback, seq.c:170-178).
 *
 * Two implementations export exactly these symbols:
 *   - skred_b200/csrc/engine.cu   -> libskred_b200.so      (the product, CUDA sm_100a)
 *   - oracle/skred_port.c         -> oracle/_build/...     (CPU restatement, TEST ONLY)
 */
#ifndef SKRED_B200_H
#define SKRED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_ABI_VERSION 1

/* ---- error codes (synth() itself returns void in the reference; errors are
 *      sticky on the engine and surfaced through skb_last_error) ---------- */
enum {
  SKB_OK = 0,
  SKB_ERR_ARG = -1,          /* bad argument (voice / table / frame count out of range) */
  SKB_ERR_CUDA = -2,         /* a CUDA call failed; skb_error_string has the text */
  SKB_ERR_NO_DEVICE = -3,    /* no usable sm_100 device — there is NO CPU fallback */
  SKB_ERR_CAPACITY = -4,     /* table arena / op buffer / group size exceeded */
  SKB_ERR_STATE = -5,        /* call sequence error */
};

/* ---- voice parameter record ------------------------------------------- */
/* flag bits: the reference's 0/1 int arrays packed into one word */
#define SKB_F_ONE_SHOT     (1u << 0)  /* voice_one_shot        synth.c:221 */
#define SKB_F_LOOP_ENABLED (1u << 1)  /* voice_loop_enabled    synth.c:222 */
#define SKB_F_LOOP_VALID   (1u << 2)  /* voice_loop_valid      synth.c:235 */
#define SKB_F_REVERSE      (1u << 3)  /* voice_direction       synth.c:224 */
#define SKB_F_USE_ENV      (1u << 4)  /* voice_use_amp_envelope synth.c:582 */
#define SKB_F_SMOOTHER     (1u << 5)  /* voice_smoother_enable synth.c:589 */
#define SKB_F_DISCONNECT   (1u << 6)  /* voice_disconnect      synth.c:595 */
#define SKB_F_NOISE        (1u << 7)  /* wave index == WAVE_TABLE_NOISE_ALT, synth.c:543 */

typedef struct skb_voice_params {
  float    amp;              /* voice_amp                         synth.c:537,580 */
  float    phase_inc;        /* voice_phase_inc (host: synth.c:125-136) */
  float    freq_scale;       /* voice_freq_scale                  synth.c:554 */
  float    freq_mod_depth;   /* voice_freq_mod_depth              synth.c:553 */
  int32_t  freq_mod_osc;     /* voice_freq_mod_osc, <0 = off      synth.c:548 */
  int32_t  table_id;         /* handle from skb_table_upload (voice_table), <0 = none */
  int32_t  table_size;       /* voice_table_size                  synth.c:220 */
  float    loop_start_f;     /* voice_loop_start_f                synth.c:236 */
  float    loop_end_f;       /* voice_loop_end_f                  synth.c:238 */
  uint32_t flags;            /* SKB_F_* */
  int32_t  cz_mode;          /* voice_cz_mode                     synth.c:262 */
  float    cz_distortion;    /* voice_cz_distortion               synth.c:266 */
  int32_t  cz_mod_osc;       /* voice_cz_mod_osc                  synth.c:263 */
  float    cz_mod_depth;     /* voice_cz_mod_depth                synth.c:264 */
  int32_t  sample_hold_max;  /* voice_sample_hold_max             synth.c:560 */
  int32_t  quantize;         /* voice_quantize                    synth.c:574 */
  int32_t  filter_mode;      /* voice_filter_mode (0 = off)       synth.c:577 */
  float    b0, b1, b2, a1, a2; /* voice_filter[].b0..a2 (host: synth.c:929-1008) */
  float    env_attack;       /* voice_amp_envelope[].attack_time  (samples) synth.c:404 */
  float    env_decay;        /* .decay_time                       synth.c:410 */
  float    env_sustain;      /* .sustain_level                    synth.c:413 */
  float    env_release;      /* .release_time                     synth.c:423 */
  int32_t  amp_mod_osc;      /* voice_amp_mod_osc, <0 = off       synth.c:584 */
  float    amp_mod_depth;    /* voice_amp_mod_depth               synth.c:586 */
  float    smoother_k;       /* voice_smoother_smoothing          synth.c:590 */
  int32_t  pan_mod_osc;      /* voice_pan_mod_osc, <0 = off       synth.c:597 */
  float    pan_mod_depth;    /* voice_pan_mod_depth               synth.c:599 */
} skb_voice_params;          /* 32 words */

/* ---- evolving per-voice state ---------------------- */
typedef struct skb_voice_state {
  float    phase;            /* voice_phase                       synth.c:226,258 */
  int32_t  finished;         /* voice_finished                    synth.c:245,252 */
  float    sample;           /* voice_sample (what modulators read) synth.c:593 */
  float    sh_hold;          /* voice_sample_hold                 synth.c:562 */
  int32_t  sh_count;         /* voice_sample_hold_count           synth.c:565 */
  float    x1, x2, y1, y2;   /* voice_filter[] delay line         synth.c:358-361 */
  int32_t  env_active;       /* voice_amp_envelope[].is_active    synth.c:399,429 */
  float    env_velocity;     /* .velocity                         synth.c:582 */
  float    smoother_gain;    /* voice_smoother_gain               synth.c:590 */
  float    pan_left;         /* voice_pan_left  (pan-mod rewrites it, synth.c:600) */
  float    pan_right;        /* voice_pan_right                   synth.c:601 */
  uint64_t env_start;        /* .sample_start                     synth.c:401 */
  uint64_t env_release;      /* .sample_release                   synth.c:417,422 */
} skb_voice_state;

/* ---- device ops: ordered edits of evolving state at a block boundary ---- */
enum {
  SKB_OP_TRIGGER = 1,   /* osc_trigger (synth.c:316-339): phase = f0, finished = 0 */
  SKB_OP_SET_FINISHED,  /* osc_set_wave_table_index (synth.c:281-282): finished = i0 */
  SKB_OP_ENV_ON,        /* amp_envelope_trigger (synth.c:383-388): start = u0, release = 0, velocity = f0, active = 1 */
  SKB_OP_ENV_OFF,       /* amp_envelope_release (synth.c:391-395): if (active) release = u0 */
  SKB_OP_ENV_RESET,     /* envelope_init (synth.c:377-379): start = release = 0, active = 0 */
  SKB_OP_FILTER_CLEAR,  /* mmf_init (synth.c:1017-1018): x1 = x2 = y1 = y2 = 0 */
  SKB_OP_VOICE_CLEAR,   /* voice_reset (synth.c:1094,1124): sample = 0, smoother_gain = 0 */
  SKB_OP_SET_PAN,       /* pan_set / voice_reset (synth.c:841-842,1098-1099): pan_left = f0, pan_right = f1 */
  SKB_OP_SET_SH,        /* voice_copy (synth.c:1045-1046): sh_count = i0, sh_hold = f0 */
  SKB_OP_SET_PHASE,     /* checkpoint restore / tests: phase = f0 */
  SKB_OP__COUNT
};

typedef struct skb_op {
  int32_t  voice;
  int32_t  code;        /* SKB_OP_* */
  int32_t  i0;
  float    f0;
  float    f1;
  int32_t  _pad;
  uint64_t u0;
} skb_op;               /* 32 bytes */

/* ---- engine ------------------------------------------------------------ */
typedef struct skb_engine skb_engine;

typedef struct skb_config {
  int32_t  abi_version;   /* SKB_ABI_VERSION */
  int32_t  device;        /* CUDA ordinal */
  int32_t  n_voices;      /* the host's VOICE_MAX (skred.h:9) */
  int32_t  max_frames;    /* largest nframes of one skb_render* call (>= 512) */
  int32_t  rank;          /* voice shard of this engine ... */
  int32_t  world;         /* ... out of `world` engines (1 = everything) */
  uint32_t flags;         /* SKB_CFG_* */
  int32_t  _reserved;
} skb_config;

#define SKB_CFG_DEFAULT 0u
/* testing aid: render every free voice through the generic per-frame code path
 * instead of the warp-specialised one (both must produce identical bits) */
#define SKB_CFG_FORCE_GENERIC 1u
/* testing aid: one launch per skb_render_mix call (no batching of consecutive callbacks) */
#define SKB_CFG_NO_BATCH 2u
/* opt-in: render launches of >= 4 windows TIME-SPLIT (free_kernel.cuh: pass A advances every voice through
 * the launch with the light recurrences and snapshots its state per window, pass B renders the windows in
 * parallel, pass C runs the biquads).  Bit-identical state, same mix within the regrouping of the sum; measured
 * on B200 it does not pay (profiles/r01_time_split.txt), so the default is the sequential kernel. */
#define SKB_CFG_WIDE 4u
/* render every launch with k_render_rows (one CTA per 32-voice row, the warps are pipeline stages: row_kernel.cuh), the
 * kernel built for launches that leave the SMs nearly empty (one job cut over several GPUs).  Bit-identical state (the
 * whole parity suite runs through it with SKB_ROWS=1); measured no faster than k_render_free, so it is opt-in.
 * Environment: SKB_ROWS=0 / 1 / 2 (never / always / launches of at most SKB_ROWS_MAX rows). */
#define SKB_CFG_ROWS 16u
/* measuring aid: deal rows to CTAs by cost only (plain LPT) instead of class-affine (engine.cu: replan) */
#define SKB_CFG_NO_AFFINE 8u

int  skb_create(skb_engine **out, const skb_config *cfg);
void skb_destroy(skb_engine *e);

/* Sticky error of the engine (SKB_OK if none) and its text. */
int         skb_last_error(const skb_engine *e);
const char *skb_error_string(const skb_engine *e);

/* Copy a wave table (wave_table_data[slot], wave_size[slot]; synth.c:1199-1294,
 * wire.c:374-441) into the device arena.  Returns a table id >= 0 that
 * skb_voice_params.table_id refers to, or a negative SKB_ERR_*.  Tables are
 * immutable once uploaded (a replaced wave slot gets a new id, the old one
 * keeps serving voices that still hold it — the reference's graveyard rule,
 * wire.c:370-390). */
int  skb_table_upload(skb_engine *e, const float *data, int size);

/* Replace the parameter record of one voice; takes effect at the next render. */
int  skb_set_params(skb_engine *e, int voice, const skb_voice_params *p);

/* Queue ordered device ops; all are applied, in order, before the next render. */
int  skb_push_ops(skb_engine *e, const skb_op *ops, int n);

/* Render `nframes` frames of all voices this engine owns.
 *   ssc_before : synth_sample_count BEFORE the first frame (the loop
 *                pre-increments it, synth.c:521)
 *   gain       : host, nframes floats — the master-volume smoother trace
 *                (synth.c:616-620), voice independent, computed by the host
 *   noise      : host, nframes floats — the per-frame shared noise draw
 *                (synth.c:525), or NULL when no voice uses it
 *   out        : HOST buffer, interleaved, stride num_channels; channels 0 and
 *                1 are written (synth.c:623-624)
 * Blocking: returns after `out` is complete.  Returns SKB_OK or an error. */
int  skb_render(skb_engine *e, int nframes, uint64_t ssc_before,
                const float *gain, const float *noise,
                float *out, int num_channels);

/* Split form for multi-GPU / batch use.  skb_render_mix leaves the raw
 * (pre-master-volume) stereo sum of this engine's voices in DEVICE memory
 * d_mix[nframes][2] on `stream` (a cudaStream_t, NULL = engine stream) without
 * synchronising; the caller may reduce d_mix across engines (NCCL) and then
 * call skb_finish on the root to apply `gain` and copy to the host. */
int  skb_render_mix(skb_engine *e, int nframes, uint64_t ssc_before,
                    const float *noise, float *d_mix, void *stream);
/* skb_render_mix may DEFER its launch: consecutive calls that continue each other (next frames,
 * next d_mix address, same stream) and are separated only by state edits (skb_push_ops) are
 * rendered by ONE kernel launch that applies those edits itself at the callback boundaries.
 * skb_flush launches what is pending (no host synchronisation); skb_finish / skb_sync /
 * skb_snapshot do so implicitly.  Call it before consuming d_mix on the stream yourself
 * (e.g. before an NCCL reduce). */
int  skb_flush(skb_engine *e);
/* skb_finish applies the master-volume trace gain[nframes] (host memory, synth.c:619-624) to d_mix and returns the frames in
 * the HOST buffer `out`; blocking.  The finishing kernel writes the engine's pinned host buffer directly (mapped into the
 * device's address space) and raises a flag the call polls: no device-to-host copy, no stream synchronisation.  Environment:
 * SKB_FINISH_COPY=1 at skb_create selects the staged form (device buffer, cudaMemcpyAsync, cudaStreamSynchronize) — the same
 * bits (tests: test_direct_finish_equals_staged_finish). */
int  skb_finish(skb_engine *e, const float *d_mix, int nframes, const float *gain,
                float *out, int num_channels, void *stream);
/* The engine's own raw-mix buffer (DEVICE memory, max_frames x 2 floats) for callers of the
 * split form that do not bring one: synth() renders the event-free segments of one call into
 * it back to back and finishes once (one synchronisation per call, not per segment). */
float *skb_mix_buffer(skb_engine *e);
/* Per-voice stereo tap: the `user` buffer of synth() (synth.c:503-511, 533-611; skred.c:120-131 records
 * from it).  While enabled, every render also leaves tap[frame][voice][L, R] — each voice's contribution to
 * the mix before the master volume, zeros for skipped (finished, amp == 0), disconnected and not-owned
 * voices — in device memory, frames counted from the last skb_finish (at most max_frames of them).
 * skb_read_tap copies frames [frame0, frame0 + nframes) to out[nframes][n_voices][2] on the host; call it
 * after skb_finish.  The tap moves 8 bytes per voice-sample (256 MiB per 512-frame block at
 * 65,536 voices), so it is off until asked for; time-split launches are not used while it is on. */
int  skb_set_tap(skb_engine *e, int enable);
int  skb_read_tap(skb_engine *e, int frame0, int nframes, float *out);
/* Selective read-back (skred.c:120-131 and wire.c:94-185 only ever WRITE the voices with voice_record[v] set,
 * wire.c:698).  skb_set_tap_voices names the voices whose tap the host wants (n = 0: every voice); skb_read_tap_selected
 * copies only their columns, out[nframes][n_selected][2] in the order given, over PCIe — 8 bytes per SELECTED
 * voice-sample instead of 8 bytes per voice-sample — and reports in extremes[2] = {min(0, x), max(0, x)} over the samples
 * of the voices NOT selected: save_wav scales by the extremes of all voices of the recording (wire.c:150-166), so a host
 * that wants the reference's file bit for bit needs those two numbers and nothing else of the unselected voices. */
int  skb_set_tap_voices(skb_engine *e, const int32_t *voices, int n);
int  skb_read_tap_selected(skb_engine *e, int frame0, int nframes, float *out, float *extremes);
/* Wait for everything queued on `stream` (NULL = engine stream). */
int  skb_sync(skb_engine *e, void *stream);

/* Device -> host copy of evolving state for voices [first, first+n)
 * (so voice_format / `?` / `\` and voice_copy keep working, synth.c:663-808). */
int  skb_snapshot(skb_engine *e, int first, int n, skb_voice_state *out);
/* Host -> device overwrite of evolving state (checkpoint restore, tests). */
int  skb_restore(skb_engine *e, int first, int n, const skb_voice_state *in);

/* ---- the exchange step of the voice-sharded path (SURVEY 8e, 8b last row; north_star: "per-GPU stereo
 *      partial mixes are summed with an NCCL reduce over NVLink") ------------------------------------
 * The reference is one process with no device and no communication: these calls replace
 * nothing, they are what lets a plain C host (skred.c:107-119 calls synth() from the audio callback)
 * drive N B200s without Python.  NCCL (libnccl.so.2) is opened lazily by the first skb_comm_* call; an
 * engine that never communicates does not need the library.
 *
 *   one process per GPU:   rank 0 calls skb_comm_unique_id and ships the SKB_COMM_ID_BYTES bytes to the
 *                          other ranks over any transport; every rank then calls skb_comm_init_rank
 *                          (ncclCommInitRank) on its engine (created with cfg.rank / cfg.world);
 *   one process, N GPUs:   skb_comm_init_all(engines, n) — ncclCommInitAll over the engines' devices,
 *                          engines[r] was created with cfg.rank = r, cfg.world = n.
 *
 * skb_reduce_mix launches what is pending (skb_flush) and sums d_mix[nframes][2] of all ranks into rank
 * 0's d_mix on `stream` (NULL = engine stream), without host synchronisation; rank 0 then calls skb_finish.
 * Cross-rank order of the sum:
 *   SKB_COMM_NCCL_REDUCE  ncclReduce(sum, root 0): NCCL's tree / NVLS order — a deterministic function of
 *                         (ranks, frames, NCCL version), run-to-run reproducible, not rank order;
 *   SKB_COMM_ORDERED      partial mixes gathered on rank 0 (grouped ncclSend / ncclRecv) and added there
 *                         in rank order 0, 1, ... n-1 by k_sum_ranks: the order SURVEY 5 asks for,
 *                         independent of topology and NCCL version.
 * Returns SKB_OK, SKB_ERR_STATE (no communicator / NCCL missing) or SKB_ERR_CUDA (NCCL error text in
 * skb_error_string). */
#define SKB_COMM_ID_BYTES 128
#define SKB_COMM_NCCL_REDUCE 0
#define SKB_COMM_ORDERED 1
int  skb_comm_unique_id(void *id_out);
int  skb_comm_init_rank(skb_engine *e, const void *id, int rank, int nranks);
int  skb_comm_init_all(skb_engine *const *engines, int n);
int  skb_comm_set_mode(skb_engine *e, int mode);
int  skb_comm_size(const skb_engine *e);          /* ranks of the engine's communicator, 0 = none */
int  skb_reduce_mix(skb_engine *e, float *d_mix, int nframes, void *stream);
/* the same for ONE host thread that drives all n engines (skb_comm_init_all): every engine's own mix buffer
 * (skb_mix_buffer) on its own stream, the n calls inside one NCCL group */
int  skb_reduce_mix_all(skb_engine *const *engines, int n, int nframes);
int  skb_comm_destroy(skb_engine *e);

/* Does voice `v` belong to this engine's shard? (world > 1) */
int  skb_owns_voice(skb_engine *e, int voice);

/* Counters for bench/tests. */
typedef struct skb_stats {
  uint64_t kernel_launches;   /* our kernels launched since create */
  uint64_t frames_rendered;
  uint64_t ops_applied;
  uint64_t params_uploaded;
  uint64_t replans;
  int32_t  n_free_voices;     /* current plan */
  int32_t  n_group_voices;
  int32_t  n_groups;
  int32_t  n_owned_voices;
  float    last_render_ms;    /* device time of the last skb_render_mix (CUDA events) */
  int32_t  _pad;
  uint64_t active_voice_frames; /* voice-frames actually rendered (not skipped by synth.c:531-542) as of the
                                   last skb_finish / skb_sync: the numerator of voice-samples/s */
  uint64_t class_rows[8];     /* diagnostics: live 32-voice rows seen by k_render_free per code-path class
                                 (0..5 = CZ none/piecewise/pow x filter off/on, 6 = mixed row, 7 = generic) */
  uint64_t phase_cycles[8];   /* diagnostics: SM clocks thread 0 of every CTA spent per phase of k_render_free
                                 (compaction, voice set-up, table cache, envelope pre-pass, render, wait for the
                                 slowest warp, row sum, final store), summed over CTAs and launches */
  uint64_t cta_batches;       /* ... and how many (CTA, batch) passes that covers */
  uint64_t wide_launches;     /* launches rendered time-split (pass A advance, pass B windows, pass C biquad) */
  uint64_t wide_errors;       /* rows a time-split pass had to drop (must stay 0) */
  float    last_wide_ms[3];   /* device time of passes A, B, C (+ row reduce) of the last time-split launch */
  int32_t  _pad2;
  double   host_us[4];        /* diagnostics, host microseconds accumulated: [0] launching a batch (sorting its ops, staging,
                                 H2D, kernel launches), [1] queueing skb_finish (wait for staging, gain H2D, k_finish, D2H),
                                 [2] waiting for the stream in skb_finish, [3] unused */
  uint64_t rows_launches;     /* launches rendered by k_render_rows */
  uint64_t migrated_voices;   /* voices whose state moved to another GPU at a re-plan (a modulation edge joined two shards) */
  uint64_t lo_launches;       /* launches of the few-voices build of k_render_free (csrc/free_lo.cu) */
  uint64_t h2d_bytes, d2h_bytes; /* bytes the render path copied host -> device (parameter records, ops, per-launch staging
                                    block, noise and gain traces) and device -> host (stereo block, counters, tap) */
} skb_stats;
int  skb_get_stats(skb_engine *e, skb_stats *out);

/* Diagnostics for tuning (tools/bench_probe.py): per-CTA phase clocks of the last launch of the
 * free-voice kernel and the planner's row lists; class rank of the voice in a slot. */
int  skb_debug_cta_phases(skb_engine *e, uint64_t *phases, int32_t *rows, int max_ctas, int *rows_cap);
int  skb_debug_slot_rank(skb_engine *e, int slot);
int  skb_debug_warp_clocks(skb_engine *e, uint64_t *out, int max_ctas);

/* "cuda-sm100a" for the product, "cpu-port" for the oracle build. */
const char *skb_backend_name(void);

#ifdef __cplusplus
}
#endif
#endif /* SKRED_B200_H */
