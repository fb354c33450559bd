/* free_lo.cu — k_render_free compiled a second time for FEW VOICES PER GPU (BASELINE configs[4] cut 4 or 8 ways, VOICE_MAX 64).
 *
 * When a GPU holds at most a few rows per SM, a launch lasts as long as its slowest WARP running alone on its scheduler
 * (profiles/r02_probe_shards.txt: one pow + filter row alone on an SM, 392 us per 8,192 frames, 92 cycles per frame, while
 * every other SM has long finished).  A lone warp is paced by its own instruction-level parallelism, and the default build
 * caps that at 128 registers per thread (14 warps per CTA) and 4 frames per pipeline stage.  This translation unit is the
 * same source — free_kernel.cuh, every arithmetic op identical, bit-identical state — with 8 frames per stage and 8 warps
 * per CTA (224 registers): the lone pow + filter warp takes 325 us, the launch 0.39 instead of 0.47 ms; at 65,536 voices per
 * GPU it is 40 % SLOWER (half the warps per SM), so the engine launches it only when no CTA holds more than 8 rows
 * (engine.cu: lo_ok).  Everything lives in namespace skb_lo; the engine reaches it through the three C functions below. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math_constants.h>
#include "skred_b200.h"

#define SKB_LO_VARIANT 1
#define SKB_SUB 8
#define SKB_CTA_WARPS 8
namespace skb_lo {
#include "voice_kernels.cuh"
}

extern "C" size_t skb_lo_free_args_bytes(void) { return sizeof(skb_lo::FreeArgs); }
extern "C" size_t skb_lo_free_smem_bytes(int env_warp_floats) { return skb_lo::skb_free_smem_bytes(env_warp_floats); }
extern "C" int skb_lo_free_threads(void) { return SKB_CTA_THREADS; }
extern "C" int skb_lo_free_rows_per_cta(void) { return SKB_CTA_WARPS; }
extern "C" int skb_lo_free_init(void) {
  return cudaFuncSetAttribute(skb_lo::k_render_free, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)skb_lo::skb_free_smem_bytes(SKB_ENV_WARP_FLOATS_LARGE)) == cudaSuccess ? 0 : -1;
}
/* `args` is the engine's FreeArgs (the same struct, compiled in the other translation unit) */
extern "C" void skb_lo_free_launch(const void *args, int ctas, size_t smem, void *stream) {
  skb_lo::FreeArgs a;
  memcpy(&a, args, sizeof(a));
  skb_lo::k_render_free<<<ctas, SKB_CTA_THREADS, smem, (cudaStream_t)stream>>>(a);
}
