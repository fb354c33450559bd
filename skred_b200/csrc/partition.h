/* partition.h — voice dependency graph, connected components, shard owner map.
 *
 * Plain C (header-only, no CUDA) so that the engine, the host-logic tests and
 * the CPU port agree on ownership by construction.
 *
 * Reference semantics being honoured: inside synth() a voice
 * may read `voice_sample[m]` of another voice through four edges —
 *   FM   voice_freq_mod_osc  (synth.c:548-555, only when mod != n)
 *   AM   voice_amp_mod_osc   (synth.c:584-587)
 *   PAN  voice_pan_mod_osc   (synth.c:597-602)
 *   CZ   voice_cz_mod_osc    (synth.c:263-266, only when cz_mode != 0)
 * Voices joined by such edges form a connected component that must be
 * rendered frame-lock-step on ONE GPU.  The default CZ modulator is
 * "voice 0, depth 0" for every voice (voice_reset never touches it,
 * synth.c:1090-1132), which multiplies the modulator by 0.0f: value-
 * independent, so a CZ edge counts only when its depth is non-zero
 * (SURVEY App. A-4).
 */
#ifndef SKB_PARTITION_H
#define SKB_PARTITION_H

#include <stdint.h>
#include <stdlib.h>
#include "skred_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Which modulator slots of voice n are live reads of voice_sample[]?
 * Fills m[4] = {fm, cz, am, pan} with the voice index or -1. */
static inline void skb_live_mods(const skb_voice_params *p, int n, int n_voices, int m[4]) {
  m[0] = (p->freq_mod_osc >= 0 && p->freq_mod_osc != n && !(p->flags & SKB_F_NOISE)) ? p->freq_mod_osc : -1;
  m[1] = (p->cz_mode != 0 && p->cz_mod_osc >= 0 && p->cz_mod_depth != 0.0f && !(p->flags & SKB_F_NOISE)) ? p->cz_mod_osc : -1;
  m[2] = (p->amp_mod_osc >= 0) ? p->amp_mod_osc : -1;
  m[3] = (p->pan_mod_osc >= 0 && !(p->flags & SKB_F_DISCONNECT)) ? p->pan_mod_osc : -1;
  for (int k = 0; k < 4; k++)
    if (m[k] >= n_voices) m[k] = -1;   /* out of range: reads as silence (reference: UB) */
}

static inline int skb_uf_find(int32_t *parent, int x) {
  while (parent[x] != x) {
    parent[x] = parent[parent[x]];
    x = parent[x];
  }
  return x;
}

/* comp[v] = smallest voice index of v's component.  A voice that is skipped
 * by the loop for the whole block (amp == 0, synth.c:537) still WRITES
 * voice_sample = 0 and can be read, so amplitude does not cut edges. */
static inline void skb_components(const skb_voice_params *params, int n_voices, int32_t *comp) {
  for (int v = 0; v < n_voices; v++) comp[v] = v;
  for (int v = 0; v < n_voices; v++) {
    int m[4];
    skb_live_mods(&params[v], v, n_voices, m);
    for (int k = 0; k < 4; k++) {
      if (m[k] < 0 || m[k] == v) continue;
      int a = skb_uf_find(comp, v), b = skb_uf_find(comp, m[k]);
      if (a != b) { if (a < b) comp[b] = a; else comp[a] = b; }
    }
  }
  for (int v = 0; v < n_voices; v++) comp[v] = skb_uf_find(comp, v);
}

/* owner[v] in [0, world): components in order of their smallest voice index
 * go to the currently least-loaded rank (ties -> lowest rank); load = voices.
 * Deterministic, so every rank computes the same map without communication. */
static inline void skb_partition(const int32_t *comp, int n_voices, int world, int32_t *owner) {
  if (world <= 1) { for (int v = 0; v < n_voices; v++) owner[v] = 0; return; }
  int64_t *load = (int64_t *)calloc((size_t)world, sizeof(int64_t));
  int32_t *csize = (int32_t *)calloc((size_t)n_voices, sizeof(int32_t));
  for (int v = 0; v < n_voices; v++) csize[comp[v]]++;
  for (int v = 0; v < n_voices; v++) {
    if (comp[v] != v) continue;           /* v is a component root (its minimum) */
    int best = 0;
    for (int r = 1; r < world; r++) if (load[r] < load[best]) best = r;
    owner[v] = best;
    load[best] += csize[v];
  }
  for (int v = 0; v < n_voices; v++) owner[v] = owner[comp[v]];
  free(load);
  free(csize);
}

/* The same with a memory: `prev` (or NULL) is the owner map of the previous plan.  A component whose voices all had
 * one owner keeps it; a component that MERGED voices of several owners goes to the rank that held most of them (ties ->
 * lowest rank), so only the smaller side has to move; a component that split keeps its voices where they were.
 * Modulation routing can therefore change after rendering has started without reshuffling unrelated voices (the greedy
 * map above re-deals every later component when an early one changes size).  Deterministic on every rank. */
static inline void skb_partition_stable(const int32_t *comp, int n_voices, int world, const int32_t *prev, int32_t *owner) {
  if (!prev || world <= 1) { skb_partition(comp, n_voices, world, owner); return; }
  int32_t *votes = (int32_t *)calloc((size_t)n_voices, sizeof(int32_t));   /* votes of the component's best rank so far */
  int32_t *best = (int32_t *)malloc((size_t)n_voices * sizeof(int32_t));
  int32_t *cnt = (int32_t *)calloc((size_t)world, sizeof(int32_t));
  for (int v = 0; v < n_voices; v++) best[v] = -1;
  /* per component: count its voices per previous owner; components are visited root by root */
  int32_t *next = (int32_t *)malloc((size_t)n_voices * sizeof(int32_t));   /* linked list of a component's voices */
  int32_t *head = (int32_t *)malloc((size_t)n_voices * sizeof(int32_t));
  for (int v = 0; v < n_voices; v++) head[v] = -1;
  for (int v = n_voices - 1; v >= 0; v--) { next[v] = head[comp[v]]; head[comp[v]] = v; }
  for (int r = 0; r < n_voices; r++) {
    if (comp[r] != r) continue;
    for (int w = 0; w < world; w++) cnt[w] = 0;
    for (int v = head[r]; v >= 0; v = next[v]) {
      const int o = prev[v];
      if (o >= 0 && o < world) cnt[o]++;
    }
    int b = 0;
    for (int w = 1; w < world; w++) if (cnt[w] > cnt[b]) b = w;
    best[r] = b;
  }
  for (int v = 0; v < n_voices; v++) owner[v] = best[comp[v]];
  free(votes); free(best); free(cnt); free(next); free(head);
}

#ifdef __cplusplus
}
#endif
#endif
