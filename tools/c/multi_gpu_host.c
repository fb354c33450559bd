/* multi_gpu_host.c — a plain C host driving N B200s through include/skred_b200.h, no Python, no torch.
 *
 *   gcc -O2 -Iinclude tools/c/multi_gpu_host.c -Lskred_b200 -lskred_b200 -Wl,-rpath,$PWD/skred_b200 -lm -o multi_gpu_host
 *   ./multi_gpu_host [n_gpus] [voices] [blocks] [ordered 0|1]
 *
 * One process, one engine per GPU (skb_create with rank r of n), one communicator over them (skb_comm_init_all =
 * ncclCommInitAll), the same parameter records and ops to every engine (each keeps what its shard owns), then per
 * block: skb_render_mix on every engine, skb_reduce_mix_all (ncclReduce of the stereo partial mixes to rank 0 over
 * NVLink), skb_finish on rank 0.  The result is compared with ONE engine rendering all voices: the per-voice
 * arithmetic is identical, only the grouping of the cross-voice sum differs (<= 1e-6 here).  Exit code 0 = equal.
 * (What skred.c:107-119 would do around synth() on an 8-GPU box.) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "skred_b200.h"

#define CHECK(x) do { int _r = (x); if (_r != SKB_OK) { fprintf(stderr, "%s failed: %d\n", #x, _r); return 2; } } while (0)

static void voice(skb_voice_params *p, int v, int V, int table, int tsize) {
  memset(p, 0, sizeof(*p));
  const float hz = 55.0f * powf(2.0f, (float)(v % 48) / 12.0f);
  p->amp = 40.0f / (float)V;
  p->phase_inc = (hz * (float)tsize) / 44100.0f;      /* osc_get_phase_inc, synth.c:125-131 (rate == 44100) */
  p->freq_mod_osc = p->amp_mod_osc = p->pan_mod_osc = -1;
  p->table_id = table; p->table_size = tsize;
  p->loop_end_f = (float)(tsize - 1);
  p->flags = SKB_F_SMOOTHER | SKB_F_LOOP_VALID;
  p->smoother_k = 0.02f;
  if (v % 3 == 1) { p->cz_mode = 1 + v % 5; p->cz_distortion = 0.1f + 0.07f * (float)(v % 11); }
  if (v % 4 == 2) { p->filter_mode = 1; p->b0 = 0.02f; p->b1 = 0.04f; p->b2 = 0.02f; p->a1 = -1.6f; p->a2 = 0.68f; }
  if (v % 64 == 5) { p->freq_mod_osc = v + 1; p->freq_mod_depth = 2.0f; p->freq_scale = 1.0f; }   /* FM pairs stay on one GPU */
}

static int render(int n, int V, int blocks, int ordered, float *out) {
  skb_engine *eng[8];
  float *sine = (float *)malloc(4096 * sizeof(float));
  for (int i = 0; i < 4096; i++) sine[i] = sinf(2.0f * (float)M_PI * (float)i / 4096.0f);
  for (int r = 0; r < n; r++) {
    skb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = SKB_ABI_VERSION; cfg.device = r; cfg.n_voices = V; cfg.max_frames = 512; cfg.rank = r; cfg.world = n;
    CHECK(skb_create(&eng[r], &cfg));
    const int tid = skb_table_upload(eng[r], sine, 4096);
    for (int v = 0; v < V; v++) {
      skb_voice_params p;
      voice(&p, v, V, tid, 4096);
      CHECK(skb_set_params(eng[r], v, &p));
      skb_op pan = { v, SKB_OP_SET_PAN, 0, 0.5f - 0.4f * (float)(v % 5 - 2) / 2.0f, 0.5f + 0.4f * (float)(v % 5 - 2) / 2.0f, 0, 0 };
      CHECK(skb_push_ops(eng[r], &pan, 1));
    }
  }
  if (n > 1) CHECK(skb_comm_init_all(eng, n));
  if (n > 1 && ordered) for (int r = 0; r < n; r++) CHECK(skb_comm_set_mode(eng[r], SKB_COMM_ORDERED));   /* fixed rank order of the sum */
  float gain[512];
  float g = 0.0f;
  for (int b = 0; b < blocks; b++) {
    for (int i = 0; i < 512; i++) { g += 0.002f * (0.025f - g); gain[i] = g; }       /* master volume trace, synth.c:616-620 */
    for (int r = 0; r < n; r++) CHECK(skb_render_mix(eng[r], 512, (uint64_t)b * 512, NULL, skb_mix_buffer(eng[r]), NULL));
    if (n > 1) CHECK(skb_reduce_mix_all(eng, n, 512));
    CHECK(skb_finish(eng[0], skb_mix_buffer(eng[0]), 512, gain, out + (size_t)b * 1024, 2, NULL));
    for (int r = 1; r < n; r++) CHECK(skb_sync(eng[r], NULL));
  }
  int owned = 0;
  for (int r = 0; r < n; r++) { skb_stats st; skb_get_stats(eng[r], &st); owned += st.n_owned_voices; }
  for (int r = 0; r < n; r++) { skb_comm_destroy(eng[r]); skb_destroy(eng[r]); }
  free(sine);
  return owned == V ? 0 : 3;
}

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2, V = argc > 2 ? atoi(argv[2]) : 4096, blocks = argc > 3 ? atoi(argv[3]) : 16;
  const int ordered = argc > 4 ? atoi(argv[4]) : 0;
  float *a = (float *)calloc((size_t)blocks * 1024, sizeof(float)), *b = (float *)calloc((size_t)blocks * 1024, sizeof(float));
  int r = render(1, V, blocks, 0, a);
  if (r) return r;
  r = render(n, V, blocks, ordered, b);
  if (r) return r;
  double d = 0.0, peak = 0.0;
  for (int i = 0; i < blocks * 1024; i++) { d = fmax(d, fabs((double)a[i] - (double)b[i])); peak = fmax(peak, fabs((double)a[i])); }
  printf("C host, %d GPUs in one process (skb_comm_init_all + skb_reduce_mix_all, %s): %d voices x %d frames, max|diff| vs one GPU %.3g, peak %.3g -> %s\n",
         n, ordered ? "ordered sum" : "ncclReduce", V, blocks * 512, d, peak, (d <= 1e-6 && peak > 1e-3) ? "OK" : "FAIL");
  return (d <= 1e-6 && peak > 1e-3) ? 0 : 1;
}
