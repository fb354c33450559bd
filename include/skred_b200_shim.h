/* skred_b200_shim.h — the entry points the drop-in ADDS to skred's synth.h.
 *
 * The drop-in library (libskred_shim_v<VOICE_MAX>.so, built from
 * skred_b200/csrc/synth_shim.c against a skred source tree) defines every
 * array of synth.def:1-89 and every function of synth.h:8-85 with the
 * reference's names, arguments and return codes (100 = invalid voice / value,
 * 101 = frequency out of range; synth.c:661,834,860) — those are declared by
 * skred's own synth.h and are not repeated here.  `synth()` (synth.h:8) keeps
 * its signature and blocks until the block is in `buffer`.
 *
 * Additions (SURVEY §8b "What a C-ABI replacement adds"):
 */
#ifndef SKRED_B200_SHIM_H
#define SKRED_B200_SHIM_H

#include "skred_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Choose device / voice shard before the first setter or synth() call
 * (default: device $SKB_DEVICE or 0, rank 0 of 1, 8192 frames per launch).
 * Replaces nothing in the reference (single process, no device). */
int  skb_shim_configure(int device, int rank, int world, int max_frames);

/* The engine behind the shim (created on first use; aborts if no B200). */
skb_engine *skb_shim_engine(void);

/* Push pending parameter records and device ops to the engine now (synth()
 * does this itself at every block boundary — seq.c:170-178 granularity). */
int  skb_shim_flush(void);

/* Sticky engine error (the reference's synth() returns void). */
int  skb_shim_last_error(void);

/* Device -> host copy of the evolving state into the public arrays
 * (voice_phase, voice_finished, voice_sample, voice_sample_hold*,
 * voice_filter[].x1..y2, voice_amp_envelope[].is_active/sample_*,
 * voice_smoother_gain, voice_pan_left/right) so that voice_format / `?` / `\`
 * (synth.c:663-808) and voice_copy (synth.c:1044-1046) keep working. */
void skb_shim_snapshot(void);
int  skb_shim_snapshot_range(int first, int n);
/* Host arrays -> device, same fields (checkpoint resume, or a harness that wrote
 * voice_phase[] / voice_finished[] directly). */
int  skb_shim_restore_range(int first, int n);

/* wire.c writes some arrays directly, bypassing the setters (wire.c:639-708:
 * `h`, `s`, `J`, …).  With VOICE_MAX <= 4096 the shim diffs every voice
 * against the last record sent at each block; above that only voices touched
 * by a setter (or marked here) are re-packed.  skb_shim_scan_all overrides. */
void skb_shim_mark_dirty(int voice);
void skb_shim_scan_all(int on);

/* Timestamped event batch — the binary twin of the reference's work_queue[] +
 * seq() (seq.c:164-178, 241-257; wire.c:869-892 computes `when`).  `when` is an
 * absolute sample time; an event fires after the 512-frame sub-block ending at
 * count c when `when <= c + 512` by calling the setter its code
 * names, exactly as seq() would re-parse the deferred wire string.  synth()
 * with num_frames > 512 is cut only at boundaries where something fires. */
enum {
  SKB_EV_TRIGGER = 1,   /* T        voice_trigger (+ link_trig), wire.c:710-714 */
  SKB_EV_VELOCITY,      /* l<a0>    envelope_velocity (+ link_velo), wire.c:674-679 */
  SKB_EV_FREQ,          /* f<a0>    freq_set */
  SKB_EV_MIDI,          /* n<a0>    freq_midi (+ link_midi), wire.c:682-687 */
  SKB_EV_AMP,           /* a<a0>    amp_set */
  SKB_EV_PAN,           /* p<a0>    pan_set */
  SKB_EV_WAVE,          /* w<a0>    wave_set */
  SKB_EV_CZ,            /* c<a0>,<a1> cz_set */
  SKB_EV_FILTER_FREQ,   /* K<a0>    mmf_set_freq */
  SKB_EV_FILTER_RES,    /* Q<a0>    mmf_set_res */
  SKB_EV_MUTE,          /* m<a0>    wave_mute */
};
typedef struct skb_event {
  uint64_t when;
  int32_t  voice;
  int32_t  code;
  float    a0, a1;
} skb_event;            /* 24 bytes */
/* Order.  Due events fire in TIME order, first come first served among equal times.  The reference's seq() scans its
 * 1,024 work_queue slots in SLOT order (seq.c:171-178; a slot is whatever queue_item found free, seq.c:243-257), so two
 * of its items that become due at the same callback with different times fire in an order that depends on which slots
 * happened to be free.  This queue does not reproduce that accident; hosts that need it keep using seq.c's own
 * work_queue on top of the drop-in (wire.c / seq.c run unmodified over the shim: tests/test_gpu_parity.py). */
int  skb_shim_queue_events(const skb_event *ev, int n);
int  skb_shim_pending_events(void);

/* A wave slot edited in place (wave_table_dynamic_expand, wire.c:553-586)
 * must be re-uploaded at its next `w`. */
void skb_shim_wave_touch(int wave);

/* Split form of synth() for multi-GPU hosts: render this engine's voice shard
 * into DEVICE memory d_mix[num_frames][2] (raw stereo sum, before the master
 * volume of synth.c:616-620) on `stream`; after the partial mixes were summed
 * across GPUs (ncclReduce over NVLink) the root calls skb_shim_finish. */
int  skb_shim_render_mix(int num_frames, float *d_mix, void *stream);
/* The engine batches consecutive skb_shim_render_mix calls into one launch (skb_flush in
 * skred_b200.h): launch what is pending before touching d_mix on the stream yourself. */
int  skb_shim_flush_render(void);
int  skb_shim_finish(const float *d_mix, int num_frames, float *out, int num_channels, void *stream);
/* A non-root rank (or a caller that keeps the raw mix on the device) drops the
 * master-volume trace accumulated by its skb_shim_render_mix calls. */
void skb_shim_discard_gain(void);

/* `ncalls` consecutive callbacks of `frames_per_call` frames rendered into d_mix back to back, then launched
 * (skb_shim_render_mix x ncalls + skb_shim_flush_render). */
int skb_shim_render_calls(int frames_per_call, int ncalls, float *d_mix, void *stream);

/* synth() over `num_frames` frames = num_frames / 512 callbacks with the host's per-callback work in between: `between`
 * is called where skred.c calls seq(frame_count) (skred.c:119), i.e. the sequencer timeline (seq.c:164-213: deferred
 * wire strings, pattern steps) is walked ahead of the audio and its setter calls are applied at the callback boundaries
 * of ONE batched launch.  Bit-identical to the loop `synth(512); seq(512);`. */
void skb_shim_synth_between(float *buffer, int num_frames, int num_channels, void *user, void (*between)(int frame_count));

/* Per-voice tap `user` of synth(): read back only the voices being recorded (voice_record[], wire.c:698) instead of every
 * voice (8 bytes per voice-sample over PCIe).  The other voices' entries stay 0, except one carrier pair per call that
 * holds the extremes save_wav's scale depends on, so the written WAV equals the reference's (synth_shim.c, tests).
 * Also switched on by SKB_TAP_SELECTIVE=1.  Call before the first synth(). */
void skb_shim_tap_selective(int on);

/* Diagnostics: host-side seconds spent inside synth() since the last reset: [0] flush + master-volume / noise
 * trace stepping, [1] skb_render_mix (queueing the segments), [2] firing due events (setters -> device ops),
 * [3] skb_finish (launch, kernels, D2H, synchronisation) and the tap readback. */
void skb_shim_timing(double *out4, int reset);

#ifdef __cplusplus
}
#endif
#endif
