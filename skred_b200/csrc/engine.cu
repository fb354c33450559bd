/* engine.cu — the B200 voice-render engine behind include/skred_b200.h.
 *
 * Product code.  Built ONLY for sm_100a with the parity flags
 *   -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true
 *   -prec-sqrt=true -ftz=false -lineinfo
 * (SURVEY H1/F5).  There is no CPU path in this file: skb_create fails with
 * SKB_ERR_NO_DEVICE when no CUDA device of compute capability 10.x is present.
 *
 * Host side = a planner + uploader around the kernels of voice_kernels.cuh:
 *   plan     voice graph -> connected components (partition.h) -> shard owner
 *            -> SLOTS: free voices sorted by feature key (warp-homogeneous
 *            branches), then modulation BINS (frame-lock-step CTAs)
 *   upload   changed parameter records -> k_scatter_params (float4 SoA)
 *   ops      ordered state edits -> k_apply_ops (block-boundary events, F8)
 *   render   k_render_free + k_render_bins -> partial rows -> k_reduce_rows
 *   finish   k_finish (master volume) -> pinned host -> caller's buffer
 * Reference mapping: synth() synth.c:502-630; see DESIGN.md.
 */
#include <cuda_runtime.h>
#include <atomic>
#include <dlfcn.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <queue>
#include <vector>

#include "skred_b200.h"
#include "partition.h"
#include "voice_kernels.cuh"

#define SKB_N_COUNTERS 20      /* rendered voice-frames (free, bins), live rows per class, phase clocks, CTA batches */
#define SKB_BIN_WARP 32       /* components of <= 32 voices are packed into bins of <= 32: one warp per bin */
#define SKB_BIN_MAX 1024      /* one CTA per bin: hard upper bound of a component */

struct TableDesc { size_t off; int size; };

struct skb_engine {
  skb_config cfg;
  int n = 0, cap = 0;
  int err = SKB_OK;
  char errtxt[512] = {0};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_h2d = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_comm = nullptr;
  bool timing_pending = false, wide_timing_pending = false;

  /* host mirror of what the host sent */
  std::vector<skb_voice_params> par;
  std::vector<uint8_t> dirty;
  std::vector<int32_t> dirty_list;
  bool need_plan = true, planned = false;
  bool any_noise = false;
  std::vector<uint8_t> noise_flag;
  int noise_count = 0;
  cudaStream_t last_stream = nullptr;

  /* plan */
  std::vector<int32_t> comp, owner, slot_of_voice, voice_of_slot, level;
  /* levelled modulated voices (level_kernel.cuh): components whose modulation graph is a DAG */
  std::vector<int32_t> lev_of, trace_of;            /* per voice: level (-1 = not levelled), trace row (-1 = nobody reads it) */
  std::vector<std::vector<int32_t>> lev_deps;       /* per voice: the levelled voices that read it */
  std::vector<std::pair<int, int>> lev_rows;        /* per level: (first row of 32 slots, rows) */
  int n_lev_pad = 0, n_lev_rows = 0, n_traces = 0;  /* slots [n_free_pad, n_lev_pad) hold the levelled voices */
  float *d_trace = nullptr; size_t trace_cap = 0;   /* [n_traces][max_frames + 1] */
  std::vector<int32_t> bin_of_voice;                /* -1 = free */
  std::vector<skb_bin_desc> bins;
  std::vector<uint64_t> edge_sig;                   /* per voice: hash of its live edges (re-plan trigger) */
  int n_free = 0, n_free_pad = 0, n_slots = 0, n_free_rows = 0, max_bin_threads = 0;
  int n_small_bins = 0;              /* bins [0, n_small_bins) hold <= 32 voices each (k_render_bins_warp), the rest one big component each */
  int n_big_end = 0;                 /* bins [n_small_bins, n_big_end): one component of 33 ... 1,024 voices (k_render_bins); [n_big_end, size()):
                                        cyclic components above that (k_render_bins_huge) */
  float *d_binx = nullptr; size_t binx_cap = 0;      /* voice_sample exchange of the huge bins: [3][cap] floats */
  /* How many envelopes are in a transient, roughly: envelope triggers / releases seen, halved every 0.5 s of audio.  Only
   * picks the shared-memory size of the free-voice kernels (free_kernel.cuh: SKB_ENV_WARP_FLOATS_*), never a result. */
  /* few rows per SM: the launch lasts as long as its slowest lone warp, and free_lo.cu's build of the kernel (more
   * registers and frames per pipeline stage, fewer warps) runs a lone warp ~17 % faster.  lo_ok = the current plan gives
   * no CTA more rows than that kernel has warps; lo_mode: $SKB_LO (0 never, 1 when lo_ok [default]). */
  bool lo_ok = false; int lo_mode = 1; uint64_t lo_launches = 0;
  double env_activity = 0.0; long env_ops_pending = 0;
  int env_floats_forced = 0;         /* $SKB_ENV_FLOATS: testing aid */
  int rows_cap = 0;                  /* entries per CTA in d_ctarows */
  int *d_ctarows = nullptr; size_t ctarows_cap = 0;
  std::vector<int> h_ctarows;
  unsigned long long *d_ctaphase = nullptr; size_t ctaphase_cap = 0;   /* diagnostics: per-CTA phase clocks of the last launch */
  int n_sm = 148, free_ctas = 0, free_groups = 0, n_prows = 0;   /* partial rows: one per (CTA, batch) of k_render_free, one per bin */

  /* time-split ("wide") launches: free_kernel.cuh, passes A / B / C */
  struct RowList { int off = 0, ctas = 0, rows_cap = 0; };   /* rows at d_lists + off: [ctas][rows_cap] */
  RowList list_a_wide, list_c;
  std::vector<int> cta_of_row, cta_of_row_wide;   /* pass-A CTA of every free row (the in-kernel ops are bucketed by it) */
  std::vector<RowList> list_b;       /* indexed by the number of windows of the launch */
  int *d_lists = nullptr; size_t lists_cap = 0;
  int n_wide_rows = 0, n_xrows = 0, snap_nwin = 0;
  int *d_xrow = nullptr; size_t xrow_cap = 0;
  float4 *d_snap = nullptr; size_t snap_cap = 0;
  float *d_xs = nullptr; size_t xs_cap = 0;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;

  /* per-voice tap (synth.c:533-611) */
  bool tap_on = false, vos_dirty = true;
  int tap_cursor = 0;                               /* frames rendered since the last skb_finish */
  float2 *d_tap = nullptr; size_t tap_cap = 0;      /* [max_frames][n] */
  int *d_vos = nullptr; size_t vos_cap = 0;          /* voice of slot */
  float2 *h_tap = nullptr; size_t h_tap_cap = 0;
  std::vector<int32_t> tap_sel;                     /* voices whose tap is read back (empty = all) */
  int *d_tapcol = nullptr; size_t tapcol_cap = 0;    /* per voice: column in the compact copy, -1 = not selected */
  float2 *d_tapsel = nullptr; size_t tapsel_cap = 0; /* compact copy [frames][selected] */
  unsigned int *d_tapext = nullptr;                  /* extremes of the voices not selected (bit patterns) */
  bool tapcol_dirty = true;

  /* device */
  float4 *d_pq = nullptr, *d_sq[2] = {nullptr, nullptr};
  int cur = 0;
  float *d_tables = nullptr;
  size_t tables_used = 0, tables_cap = 0;
  std::vector<TableDesc> tables;
  skb_bin_desc *d_bins = nullptr; int d_bins_cap = 0;
  float2 *d_partials = nullptr; size_t partials_cap = 0;
  float2 *d_part2 = nullptr;
  unsigned int *d_tickets = nullptr;
  unsigned long long *d_counters = nullptr, *h_counters = nullptr;
  float2 *d_mix = nullptr, *d_out = nullptr;
  float *d_gain = nullptr, *d_noise = nullptr;
  int *d_idx = nullptr; size_t d_idx_cap = 0;       /* slots / old_slot scratch */
  float4 *d_recs = nullptr; size_t d_recs_cap = 0;
  skb_op *d_ops = nullptr; size_t d_ops_cap = 0;
  int2 *d_runs = nullptr; size_t d_runs_cap = 0;
  skb_voice_state *d_vsnap = nullptr; size_t d_vsnap_cap = 0;

  /* pinned staging */
  float *h_gain = nullptr, *h_noise = nullptr, *h_out = nullptr;
  /* skb_finish without copies: k_finish_host writes the frames and the counters into mapped host memory and raises h_flag
   * (SKB_FINISH_COPY=1 selects the staged path: k_finish -> d_out -> two device-to-host copies -> stream synchronise) */
  volatile unsigned int *h_flag = nullptr;
  unsigned int *d_fin_ticket = nullptr;
  unsigned int fin_seq = 0;
  bool finish_direct = false;
  float2 *dv_out = nullptr; unsigned long long *dv_counters = nullptr; unsigned int *dv_flag = nullptr;   /* device views */
  int *h_idx = nullptr; size_t h_idx_cap = 0;
  float4 *h_recs = nullptr; size_t h_recs_cap = 0;
  skb_op *h_ops = nullptr; size_t h_ops_cap = 0;
  int2 *h_runs = nullptr; size_t h_runs_cap = 0;
  skb_voice_state *h_snap = nullptr; size_t h_snap_cap = 0;

  std::vector<skb_op> ops;
  skb_stats stats;

  /* k_render_rows (row_kernel.cuh): 0 = never, 1 = every launch, 2 = launches of at most rows_auto_max free rows */
  int rows_mode = 2, rows_auto_max = 0;

  /* exchange step (skb_comm_*): NCCL communicator over the ranks of the voice-sharded render */
  ncclComm_t comm = nullptr;
  int comm_n = 0, comm_rank = 0, comm_mode = SKB_COMM_NCCL_REDUCE;
  float2 *d_gather = nullptr; size_t gather_cap = 0;      /* rank 0, ordered mode: [rank][max_frames] partial mixes */
  skb_voice_state *d_mig = nullptr; size_t mig_cap = 0;   /* records of voices that change GPUs at a re-plan */
  int *d_migidx = nullptr; size_t migidx_cap = 0;

  /* pending batch of consecutive callbacks: rendered by ONE launch of k_render_free */
  struct {
    bool open = false;
    cudaStream_t st = nullptr;
    float *mix = nullptr;               /* device, batch frames x 2 */
    int frames = 0;
    uint64_t ssc0 = 0;
    std::vector<int> win_frames;        /* each <= SKB_ENV_WIN */
    std::vector<int> win_ob;            /* ops applied before window w start at ops[win_ob[w]] */
    std::vector<skb_op> ops;            /* op.voice = slot */
    std::vector<uint32_t> wake;         /* bit per slot touched by an op of the batch */
    std::vector<int> wake_words;        /* ... and which words are non-zero */
    int tap_frame0 = 0;                 /* first frame of the batch in the tap buffer (= its offset in the caller's mix) */
  } batch;
  /* per-launch staging block (window list, op CSR, ops, wake bits), double buffered, uploaded on copy_stream */
  char *h_stage[2] = {nullptr, nullptr}, *d_stage[2] = {nullptr, nullptr};
  size_t h_stage_cap[2] = {0, 0}, d_stage_cap[2] = {0, 0};
  int stage_idx = 0;
  std::vector<int> sort_cursor;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_stage_copied[2] = {nullptr, nullptr}, ev_stage_done[2] = {nullptr, nullptr};
};

/* free_lo.cu: k_render_free compiled for few voices per GPU (8 frames per pipeline stage, 8 warps per CTA) */
extern "C" size_t skb_lo_free_args_bytes(void);
extern "C" size_t skb_lo_free_smem_bytes(int env_warp_floats);
extern "C" int skb_lo_free_rows_per_cta(void);
extern "C" int skb_lo_free_init(void);
extern "C" void skb_lo_free_launch(const void *args, int ctas, size_t smem, void *stream);

#if SKB_FAST_MODE
const char *skb_backend_name(void) { return "cuda-sm100a-fast-nonparity"; }
#elif SKB_CANARY
const char *skb_backend_name(void) { return "cuda-sm100a-canary"; }      /* the exchange self-check build (voice_kernels.cuh) */
#else
const char *skb_backend_name(void) { return "cuda-sm100a"; }
#endif

static double host_now_us() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return 1e6 * (double)t.tv_sec + 1e-3 * (double)t.tv_nsec; }

static int batch_launch(skb_engine *e);
/* can the pending ops ride inside a batched launch?  Free voices always; modulation bins when every bin is a warp bin */
static inline bool bins_batch_ok(const skb_engine *e) { return e->bins.empty() || (int)e->bins.size() == e->n_small_bins; }

static int fail(skb_engine *e, int code, const char *what, const char *detail = nullptr) {
  if (e && e->err == SKB_OK) {
    e->err = code;
    snprintf(e->errtxt, sizeof(e->errtxt), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
  }
  return code;
}

#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t _r = (call);                                                       \
    if (_r != cudaSuccess) return fail(e, SKB_ERR_CUDA, #call, cudaGetErrorString(_r)); \
  } while (0)

/* ---- exchange step: NCCL reduce of the partial mixes ------------------------------- */
/* libnccl.so.2 is opened on first use: a single-GPU host never needs it.  In a process that already
 * holds an NCCL (torch's bundled copy) the loader hands back that one. */
struct NcclApi {
  void *h = nullptr;
  bool tried = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load() {
  if (g_nccl.tried) return g_nccl.h != nullptr;
  g_nccl.tried = true;
  const char *names[] = {getenv("SKB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (int i = 0; i < 3 && !h; i++) if (names[i] && names[i][0]) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
#define SKB_NCCL_SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { dlclose(h); return false; }
  SKB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") SKB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  SKB_NCCL_SYM(CommInitAll, "ncclCommInitAll") SKB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  SKB_NCCL_SYM(Reduce, "ncclReduce") SKB_NCCL_SYM(Broadcast, "ncclBroadcast")
  SKB_NCCL_SYM(Send, "ncclSend") SKB_NCCL_SYM(Recv, "ncclRecv")
  SKB_NCCL_SYM(GroupStart, "ncclGroupStart") SKB_NCCL_SYM(GroupEnd, "ncclGroupEnd") SKB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SKB_NCCL_SYM
  g_nccl.h = h;
  return true;
}

#define NK(call)                                                                                   \
  do {                                                                                             \
    ncclResult_t _r = (call);                                                                      \
    if (_r != ncclSuccess) return fail(e, SKB_ERR_CUDA, #call, g_nccl.GetErrorString(_r));          \
  } while (0)


template <class T>
static cudaError_t grow_dev(T **p, size_t *cap, size_t need) {
  if (need <= *cap) return cudaSuccess;
  size_t ncap = std::max(need, *cap * 2 + 64);
  if (*p) cudaFree(*p);
  *p = nullptr;
  cudaError_t r = cudaMalloc((void **)p, ncap * sizeof(T));
  *cap = (r == cudaSuccess) ? ncap : 0;
  return r;
}
template <class T>
static cudaError_t grow_pin(T **p, size_t *cap, size_t need) {
  if (need <= *cap) return cudaSuccess;
  size_t ncap = std::max(need, *cap * 2 + 64);
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  cudaError_t r = cudaMallocHost((void **)p, ncap * sizeof(T));
  *cap = (r == cudaSuccess) ? ncap : 0;
  return r;
}

int skb_create(skb_engine **out, const skb_config *cfg) {
  if (!out) return SKB_ERR_ARG;
  *out = nullptr;
  if (!cfg || cfg->abi_version != SKB_ABI_VERSION || cfg->n_voices <= 0 || cfg->world < 1 ||
      cfg->rank < 0 || cfg->rank >= cfg->world)
    return SKB_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return SKB_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) {
    /* the fatbin holds sm_100a SASS only */
    cudaGetLastError();
    return SKB_ERR_NO_DEVICE;
  }
  skb_engine *e = new skb_engine();
  e->cfg = *cfg;
  if (e->cfg.max_frames < 512) e->cfg.max_frames = 512;
  e->n = cfg->n_voices;
  e->n_sm = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
  /* few voices on the GPU (one job cut over several GPUs): one CTA per row, the warps are pipeline stages */
  /* measured on B200 (profiles/r02_rows_kernel.txt): 0.40-0.44 ms against 0.445 ms of k_render_free for an 8,192-frame
   * launch of 8,192 voices, slower from 16,384 voices up — not enough to switch by default: opt-in */
  e->rows_mode = (cfg->flags & SKB_CFG_ROWS) ? 1 : 0;
  e->rows_auto_max = 2 * e->n_sm;
  { const char *s = getenv("SKB_LO"); if (s && s[0]) e->lo_mode = atoi(s); }
  { const char *s = getenv("SKB_ENV_FLOATS"); if (s && s[0]) { const int v = atoi(s); if (v >= SKB_ENV_WARP_FLOATS_SMALL && v <= SKB_ENV_WARP_FLOATS_LARGE) e->env_floats_forced = v & ~3; } }
  { const char *s = getenv("SKB_ROWS"); if (s && s[0]) e->rows_mode = atoi(s);
    s = getenv("SKB_ROWS_MAX"); if (s && s[0]) e->rows_auto_max = atoi(s); }
  memset(&e->stats, 0, sizeof(e->stats));
  const int n = e->n, mf = e->cfg.max_frames;
  e->cap = ((n + 31) / 32) * 32 + 64 + 32 * 8;        /* + class padding of the free range */
  e->par.resize(n);
  for (int v = 0; v < n; v++) {
    memset(&e->par[v], 0, sizeof(skb_voice_params));
    e->par[v].table_id = -1;
    e->par[v].freq_mod_osc = e->par[v].amp_mod_osc = e->par[v].pan_mod_osc = -1;
  }
  e->dirty.assign(n, 0);
  e->noise_flag.assign(n, 0);
  e->comp.assign(n, 0); e->owner.assign(n, 0); e->slot_of_voice.assign(n, -1);
  e->level.assign(n, 0); e->bin_of_voice.assign(n, -1); e->edge_sig.assign(n, 0);
  e->lev_of.assign(n, -1); e->trace_of.assign(n, -1); e->lev_deps.assign(n, std::vector<int32_t>());
  bool ok = cudaSetDevice(cfg->device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_h2d, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreate(&e->ev_t0) == cudaSuccess && cudaEventCreate(&e->ev_t1) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc((void **)&e->d_pq, (size_t)SKB_NPQ * e->cap * sizeof(float4)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_sq[0], (size_t)SKB_NSQ * e->cap * sizeof(float4)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_sq[1], (size_t)SKB_NSQ * e->cap * sizeof(float4)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_mix, (size_t)mf * sizeof(float2)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_out, (size_t)mf * sizeof(float2)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_gain, (size_t)mf * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_noise, (size_t)mf * sizeof(float)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_part2, (size_t)SKB_RED_CHUNKS * mf * sizeof(float2)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_tickets, (size_t)(mf / SKB_RED_X + 1) * sizeof(unsigned int)) == cudaSuccess &&
            cudaMalloc((void **)&e->d_counters, SKB_N_COUNTERS * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMallocHost((void **)&e->h_counters, SKB_N_COUNTERS * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMemset(e->d_tickets, 0, (size_t)(mf / SKB_RED_X + 1) * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(e->d_counters, 0, SKB_N_COUNTERS * sizeof(unsigned long long)) == cudaSuccess &&
            cudaEventCreate(&e->ev_a) == cudaSuccess && cudaEventCreate(&e->ev_b) == cudaSuccess &&
            cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_stage_copied[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_stage_copied[1], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_stage_done[0], cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_stage_done[1], cudaEventDisableTiming) == cudaSuccess &&
            cudaFuncSetAttribute(k_render_free, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)skb_free_smem_bytes(SKB_ENV_WARP_FLOATS_LARGE)) == cudaSuccess &&
            cudaFuncSetAttribute(k_render_free_tap, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)skb_free_smem_bytes(SKB_ENV_WARP_FLOATS_LARGE)) == cudaSuccess &&
            cudaFuncSetAttribute(k_render_window, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)skb_free_smem_bytes(SKB_ENV_WARP_FLOATS_LARGE)) == cudaSuccess &&
            cudaFuncSetAttribute(k_render_biquad, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)skb_free_smem_bytes(SKB_ENV_WARP_FLOATS_LARGE)) == cudaSuccess &&
            skb_lo_free_init() == 0 && skb_lo_free_args_bytes() == sizeof(FreeArgs) &&
            cudaFuncSetAttribute(k_render_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)skb_rows_smem_bytes()) == cudaSuccess &&
            cudaFuncSetAttribute(k_render_levels, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)skb_levels_smem_bytes()) == cudaSuccess &&
            cudaMallocHost((void **)&e->h_gain, (size_t)mf * sizeof(float)) == cudaSuccess &&
            cudaMallocHost((void **)&e->h_noise, (size_t)mf * sizeof(float)) == cudaSuccess &&
            cudaMallocHost((void **)&e->h_out, (size_t)mf * sizeof(float2)) == cudaSuccess &&
            cudaMemset(e->d_pq, 0, (size_t)SKB_NPQ * e->cap * sizeof(float4)) == cudaSuccess &&
            cudaMemset(e->d_sq[0], 0, (size_t)SKB_NSQ * e->cap * sizeof(float4)) == cudaSuccess &&
            cudaMemset(e->d_sq[1], 0, (size_t)SKB_NSQ * e->cap * sizeof(float4)) == cudaSuccess;
  if (ok) {
    /* direct finish: needs the pinned buffers in the device's address space (UVA: they are) */
    const char *fc = getenv("SKB_FINISH_COPY");
    unsigned int *hf = nullptr;
    if (!(fc && atoi(fc) != 0) &&
        cudaHostAlloc((void **)&hf, 64, cudaHostAllocMapped) == cudaSuccess &&
        cudaMalloc((void **)&e->d_fin_ticket, sizeof(unsigned int)) == cudaSuccess &&
        cudaMemset(e->d_fin_ticket, 0, sizeof(unsigned int)) == cudaSuccess &&
        cudaHostGetDevicePointer((void **)&e->dv_out, e->h_out, 0) == cudaSuccess &&
        cudaHostGetDevicePointer((void **)&e->dv_counters, e->h_counters, 0) == cudaSuccess &&
        cudaHostGetDevicePointer((void **)&e->dv_flag, hf, 0) == cudaSuccess) {
      *hf = 0u;
      e->h_flag = hf;
      e->finish_direct = true;
    } else {
      cudaGetLastError();
      if (hf) cudaFreeHost(hf);
    }
  }
  if (ok) {
    /* the bin kernel may need more than 48 KB? no: 3*nt floats + 2*nwarps float2 <= 12.8 KB */
    e->tables_cap = (size_t)4 << 20;
    ok = cudaMalloc((void **)&e->d_tables, e->tables_cap * sizeof(float)) == cudaSuccess;
  }
  if (!ok) {
    fprintf(stderr, "skred_b200: skb_create: %s\n", cudaGetErrorString(cudaGetLastError()));
    skb_destroy(e);
    return SKB_ERR_CUDA;
  }
  *out = e;
  return SKB_OK;
}

void skb_destroy(skb_engine *e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  batch_launch(e);
  if (e->stream) cudaStreamSynchronize(e->stream);
  skb_comm_destroy(e);
  cudaFree(e->d_gather); cudaFree(e->d_mig); cudaFree(e->d_migidx); cudaFree(e->d_trace);
  cudaFree(e->d_pq); cudaFree(e->d_sq[0]); cudaFree(e->d_sq[1]); cudaFree(e->d_tables);
  cudaFree(e->d_bins); cudaFree(e->d_binx); cudaFree(e->d_partials); cudaFree(e->d_mix); cudaFree(e->d_out);
  cudaFree(e->d_gain); cudaFree(e->d_noise); cudaFree(e->d_idx); cudaFree(e->d_recs);
  cudaFree(e->d_ops); cudaFree(e->d_runs); cudaFree(e->d_vsnap);
  cudaFree(e->d_ctarows); cudaFree(e->d_ctaphase);
  for (int i = 0; i < 2; i++) {
    cudaFree(e->d_stage[i]); cudaFreeHost(e->h_stage[i]);
    if (e->ev_stage_copied[i]) cudaEventDestroy(e->ev_stage_copied[i]);
    if (e->ev_stage_done[i]) cudaEventDestroy(e->ev_stage_done[i]);
  }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  cudaFree(e->d_lists); cudaFree(e->d_xrow); cudaFree(e->d_snap); cudaFree(e->d_xs);
  cudaFree(e->d_tap); cudaFree(e->d_vos); cudaFreeHost(e->h_tap);
  cudaFree(e->d_tapcol); cudaFree(e->d_tapsel); cudaFree(e->d_tapext);
  if (e->ev_a) cudaEventDestroy(e->ev_a);
  if (e->ev_b) cudaEventDestroy(e->ev_b);
  cudaFree(e->d_part2); cudaFree(e->d_tickets); cudaFree(e->d_counters);
  cudaFreeHost(e->h_counters);
  if (e->h_flag) cudaFreeHost((void *)e->h_flag);
  cudaFree(e->d_fin_ticket);
  cudaFreeHost(e->h_gain); cudaFreeHost(e->h_noise); cudaFreeHost(e->h_out); cudaFreeHost(e->h_idx);
  cudaFreeHost(e->h_recs); cudaFreeHost(e->h_ops); cudaFreeHost(e->h_runs); cudaFreeHost(e->h_snap);
  if (e->ev_h2d) cudaEventDestroy(e->ev_h2d);
  if (e->ev_t0) cudaEventDestroy(e->ev_t0);
  if (e->ev_t1) cudaEventDestroy(e->ev_t1);
  if (e->ev_comm) cudaEventDestroy(e->ev_comm);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int skb_last_error(const skb_engine *e) { return e ? e->err : SKB_ERR_ARG; }
const char *skb_error_string(const skb_engine *e) { return e ? e->errtxt : "null engine"; }

int skb_table_upload(skb_engine *e, const float *data, int size) {
  if (!e || !data || size <= 0) return fail(e, SKB_ERR_ARG, "table_upload: bad argument");
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return e->err;
  const size_t need = e->tables_used + (size_t)((size + 31) & ~31);   /* 128-byte aligned starts */
  if (need + SKB_TBL_CHUNK > e->tables_cap) {               /* slack: the table cache copies whole chunks */
    size_t ncap = std::max(need + SKB_TBL_CHUNK, e->tables_cap * 2);
    float *nt = nullptr;
    CK(cudaMalloc((void **)&nt, ncap * sizeof(float)));
    CK(cudaStreamSynchronize(e->last_stream ? e->last_stream : e->stream));
    CK(cudaMemcpy(nt, e->d_tables, e->tables_used * sizeof(float), cudaMemcpyDeviceToDevice));
    cudaFree(e->d_tables);
    e->d_tables = nt;
    e->tables_cap = ncap;
  }
  /* synchronous copy from pageable memory: table loads are cold (wire.c:406-441) */
  CK(cudaMemcpy(e->d_tables + e->tables_used, data, (size_t)size * sizeof(float), cudaMemcpyHostToDevice));
  TableDesc t = {e->tables_used, size};
  e->tables.push_back(t);
  e->tables_used = need;
  return (int)e->tables.size() - 1;
}

int skb_set_params(skb_engine *e, int voice, const skb_voice_params *p) {
  if (!e || !p || voice < 0 || voice >= e->n) return fail(e, SKB_ERR_ARG, "set_params: bad voice");
  e->par[voice] = *p;
  if (!e->dirty[voice]) { e->dirty[voice] = 1; e->dirty_list.push_back(voice); }
  e->stats.params_uploaded++;
  return SKB_OK;
}

int skb_push_ops(skb_engine *e, const skb_op *ops, int n) {
  if (!e || n < 0 || (n > 0 && !ops)) return fail(e, SKB_ERR_ARG, "push_ops: bad argument");
  e->ops.insert(e->ops.end(), ops, ops + n);
  for (int i = 0; i < n; i++) e->env_ops_pending += (ops[i].code == SKB_OP_ENV_ON || ops[i].code == SKB_OP_ENV_OFF);
  return SKB_OK;
}

/* ---- planner ------------------------------------------------------------ */
static uint64_t edge_signature(const skb_voice_params *p, int v, int n) {
  int m[4];
  skb_live_mods(p, v, n, m);
  uint64_t h = 1469598103934665603ull;
  for (int k = 0; k < 4; k++) {
    int x = (m[k] == v) ? -2 : m[k];          /* self reads are not edges */
    h = (h ^ (uint64_t)(uint32_t)x) * 1099511628211ull;
  }
  return h;
}

/* Sort key of a free voice: voices that take the same branches share warps. */
/* Sort key of a free voice.  Primary: which body of k_render_free it needs, in ascending
 * order of cost — rows of one class are packed together by the kernel, and later (costlier)
 * rows get the higher warp ids, which the SM's issue arbiter prefers: the long warps run
 * at full speed from the start and the short ones fill the gaps.  Then: same branches,
 * same wave table (neighbouring rows go to the same CTA and share its table cache). */
#define SKB_KEY_CLASS_SHIFT 40
static uint64_t feature_key(const skb_voice_params *p) {
  uint64_t k = 0;
  const bool silent = (p->amp == 0.0f);
  const bool generic = (p->flags & (SKB_F_NOISE | SKB_F_REVERSE | SKB_F_DISCONNECT)) || !(p->flags & SKB_F_SMOOTHER) ||
                       p->sample_hold_max != 0 || p->quantize != 0 || p->amp_mod_osc >= 0 || p->pan_mod_osc >= 0 ||
                       ((p->flags & SKB_F_LOOP_ENABLED) && (p->flags & SKB_F_LOOP_VALID) && p->loop_start_f != 0.0f) ||
                       p->cz_mode < 0 || p->cz_mode > 7;
  const int czv = (p->cz_mode == 0) ? 0 : (p->cz_mode >= 6 ? 2 : 1);
  static const int rank_of[6] = {0, 2, 1, 3, 4, 5};      /* (czv, filter): none, none+f, pw, pw+f, pow, pow+f */
  int rank = rank_of[czv * 2 + (p->filter_mode ? 1 : 0)] + 1;
  if (generic) rank = 7;
  if (silent) rank = 0;                                  /* skipped by the loop: never rendered */
  k |= (uint64_t)rank << SKB_KEY_CLASS_SHIFT;
  k |= (uint64_t)((p->flags & SKB_F_ONE_SHOT) ? 1 : 0) << 39;
  k |= (uint64_t)((p->flags & SKB_F_USE_ENV) ? 1 : 0) << 38;
  k |= (uint64_t)(p->cz_mode & 7) << 35;
  k |= (uint64_t)((p->flags & SKB_F_NOISE) ? 1 : 0) << 34;
  k |= (uint64_t)(p->quantize ? 1 : 0) << 33;
  k |= (uint64_t)(p->sample_hold_max ? 1 : 0) << 32;
  k |= (uint64_t)((uint32_t)(p->table_id + 1) & 0x7fffffffu);
  return k;
}

/* Translate a modulator voice index into the device reference of `v`. */
static int make_ref(const skb_engine *e, int v, int m, bool none_if_negative) {
  if (m < 0) return none_if_negative ? SKB_REF_NONE : SKB_REF_ZERO;
  if (m >= e->n) return SKB_REF_ZERO;
  if (m == v) return SKB_REF_SELF;
  if (e->lev_of[v] >= 0) {
    /* a levelled voice reads its modulator's trace: current frame if the modulator comes first in the voice loop */
    if (e->lev_of[m] < 0 || e->trace_of[m] < 0) return SKB_REF_ZERO;    /* cannot happen for a live edge */
    return SKB_REF_TRACE | e->trace_of[m] | (m < v ? SKB_REF_CUR : 0);
  }
  const int b = e->bin_of_voice[v];
  if (b < 0 || e->bin_of_voice[m] != b) return SKB_REF_ZERO;   /* cannot happen for a live edge */
  const int l = e->slot_of_voice[m] - e->bins[b].slot0;
  return l | (m < v ? SKB_REF_CUR : 0);
}

static void pack_record(const skb_engine *e, int v, float4 *r) {
  const skb_voice_params &p = e->par[v];
  int live[4];
  skb_live_mods(&p, v, e->n, live);
  const bool noise = (p.flags & SKB_F_NOISE) != 0;
  /* FM: only a live edge modulates (mod != n, synth.c:549) */
  const int fm_ref = (live[0] >= 0) ? make_ref(e, v, live[0], true) : SKB_REF_NONE;
  /* CZ: negative osc -> the literal 1.0f (synth.c:264); depth 0 / out of range -> 0.0f */
  int cz_ref;
  if (p.cz_mod_osc < 0) cz_ref = SKB_REF_NONE;
  else if (p.cz_mod_osc == v) cz_ref = (p.cz_mod_depth != 0.0f) ? SKB_REF_SELF : SKB_REF_ZERO;   /* sample * 0.0f (App. A-4) */
  else if (live[1] >= 0) cz_ref = make_ref(e, v, live[1], true);
  else cz_ref = SKB_REF_ZERO;
  const int am_ref = make_ref(e, v, p.amp_mod_osc, true);
  int pm_ref = make_ref(e, v, p.pan_mod_osc, true);
  if (p.flags & SKB_F_DISCONNECT) pm_ref = SKB_REF_NONE;        /* never evaluated, synth.c:595 */
  int toff = -1, tsize = p.table_size;
  if (!noise && p.table_id >= 0 && p.table_id < (int)e->tables.size()) {
    toff = (int)e->tables[p.table_id].off;
    if (tsize > e->tables[p.table_id].size) tsize = e->tables[p.table_id].size;   /* never read past the upload */
  }
  auto fi = [](int x) { float f; memcpy(&f, &x, 4); return f; };
  auto fu = [](uint32_t x) { float f; memcpy(&f, &x, 4); return f; };
  r[0] = make_float4(p.amp, p.phase_inc, p.freq_scale, p.freq_mod_depth);
  r[1] = make_float4(fi(fm_ref), fi(toff), fi(tsize), fu(p.flags));
  r[2] = make_float4(p.loop_start_f, p.loop_end_f, p.cz_distortion, p.cz_mod_depth);
  r[3] = make_float4(fi(p.cz_mode), fi(cz_ref), fi(p.sample_hold_max), fi(p.quantize));
  r[4] = make_float4(p.b0, p.b1, p.b2, p.a1);
  r[5] = make_float4(p.a2, p.env_attack, p.env_decay, p.env_sustain);
  r[6] = make_float4(p.env_release, p.amp_mod_depth, p.smoother_k, p.pan_mod_depth);
  r[7] = make_float4(fi(p.filter_mode), fi(am_ref), fi(pm_ref), fi(e->level[v]));
  /* levelled voice: its FM modulator's phase increment (synth.c:553 reads voice_phase_inc[mod] at render time) and the
   * trace row its own voice_sample goes to (+ 1: a cleared record means none) */
  const float fm_minc = (e->lev_of[v] >= 0 && live[0] >= 0) ? e->par[live[0]].phase_inc : 0.0f;
  r[8] = make_float4(fm_minc, fi(e->lev_of[v] >= 0 ? e->trace_of[v] + 1 : 0), 0.0f, 0.0f);
}

static int wait_staging(skb_engine *e) {
  CK(cudaEventSynchronize(e->ev_h2d));
  return SKB_OK;
}

/* deal `items` = (cost, row entry) to at most max_ctas CTAs; returns [ctas][rows_cap], -1 padded.
 * rank != nullptr: CLASS-AFFINE dealing.  Every distinct loop body a CTA's warps run is code its SM must keep streaming
 * (each body is 3-6 KB of SASS; one more body per SM measured -20 %, profiles/r01_s4_ab.txt), so the rows of a costly
 * class (CZ and / or filter) go to a subset of the CTAs sized by the class's share of the costly work — a CTA then
 * renders ONE costly class — and the light rows (plain, one-shot) fill every CTA up by LPT. */
static void deal_rows(bool tbl_affine, std::vector<std::pair<int, int>> items, int max_ctas, std::vector<int> &out, int *ctas_out,
                      int *cap_out, const std::vector<int> *rank, int heavy_cost = 29) {
      const int n = (int)items.size();
      const int ctas = std::max(1, std::min(max_ctas, n));
      const int nb = n ? ((n + ctas - 1) / ctas + SKB_CTA_WARPS - 1) / SKB_CTA_WARPS : 0;
      const int rcap = nb * SKB_CTA_WARPS;
      std::stable_sort(items.begin(), items.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) { return x.first > y.first; });
      std::vector<std::vector<std::pair<int, int>>> mine((size_t)ctas);
      std::vector<long long> load((size_t)ctas, 0);
      /* LPT of `sub` over CTAs [c0, c1) on top of the current loads */
      auto lpt = [&](const std::vector<std::pair<int, int>> &sub, int c0, int c1) {
        typedef std::pair<long long, int> Load;                              /* (load, cta) */
        std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pqd;
        for (int c = c0; c < c1; c++) if ((int)mine[c].size() < rcap) pqd.push(Load(load[c], c));
        std::vector<std::pair<int, int>> left;
        for (size_t i = 0; i < sub.size(); i++) {
          if (pqd.empty()) { left.push_back(sub[i]); continue; }
          Load l = pqd.top(); pqd.pop();
          mine[l.second].push_back(sub[i]);
          l.first += sub[i].first; load[l.second] = l.first;
          if ((int)mine[l.second].size() < rcap) pqd.push(l);                /* a full CTA leaves the heap */
        }
        return left;
      };
      const int HEAVY = heavy_cost;                                          /* cost_of_rank of anything with CZ or a filter */
      std::vector<std::pair<int, int>> light;
      if (rank && ctas >= 8) {
        std::vector<std::vector<std::pair<int, int>>> byc(8);
        long long hsum = 0, csum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i < items.size(); i++) {
          const int rk = (*rank)[items[i].second & ~SKB_ROW_WIDE] & 7;
          if (items[i].first >= HEAVY) { byc[rk].push_back(items[i]); csum[rk] += items[i].first; hsum += items[i].first; }
          else light.push_back(items[i]);
        }
        int c0 = 0, nheavy = 0, seen = 0;
        for (int rk = 7; rk >= 0; rk--) nheavy += byc[rk].empty() ? 0 : 1;
        for (int rk = 7; rk >= 0; rk--) {
          if (byc[rk].empty()) continue;
          seen++;
          int k = (seen == nheavy) ? ctas - c0 : (int)((double)ctas * (double)csum[rk] / (double)hsum + 0.5);
          k = std::max(1, std::min(k, ctas - c0 - (nheavy - seen)));
          std::vector<std::pair<int, int>> left;
          if (tbl_affine) {
            /* TABLE-AFFINE: consecutive rows of a class hold voices sorted by wave table (feature_key), so a CTA that
             * renders a CONTIGUOUS run of them sees 1-3 distinct tables and can stage all of them in shared memory
             * (free_kernel.cuh, SKB_TMA_TABLES); LPT would deal neighbouring rows to different CTAs.  Rows of a class
             * cost the same (one-shot rows a quarter), so cutting the row-ordered list at equal cost is as balanced. */
            std::vector<std::pair<int, int>> byrow = byc[rk];
            std::stable_sort(byrow.begin(), byrow.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) {
              return (x.second & ~SKB_ROW_WIDE) < (y.second & ~SKB_ROW_WIDE); });
            long long acc = 0;
            for (size_t i = 0; i < byrow.size(); i++) {
              int c = c0 + (int)std::min<long long>((long long)k - 1, (acc * k) / std::max<long long>(csum[rk], 1));
              while (c < c0 + k && (int)mine[c].size() >= rcap) c++;
              if (c >= c0 + k) { left.push_back(byrow[i]); continue; }
              mine[c].push_back(byrow[i]); load[c] += byrow[i].first;
              acc += byrow[i].first;
            }
          } else {
            left = lpt(byc[rk], c0, c0 + k);
          }
          light.insert(light.end(), left.begin(), left.end());              /* (did not fit: anywhere) */
          c0 += k;
        }
        std::stable_sort(light.begin(), light.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) { return x.first > y.first; });
      } else {
        light = items;
      }
      lpt(light, 0, ctas);
      out.assign((size_t)std::max(ctas * rcap, 1), -1);
      for (int c = 0; c < ctas; c++) {
        std::stable_sort(mine[c].begin(), mine[c].end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) {
          return x.first != y.first ? x.first < y.first : (x.second & ~SKB_ROW_WIDE) < (y.second & ~SKB_ROW_WIDE); });   /* cheapest first: costliest = highest warp id */
        for (size_t k = 0; k < mine[c].size(); k++) out[(size_t)c * rcap + k] = mine[c][k].second;
      }
      *ctas_out = n ? ctas : 0; *cap_out = rcap;
    }

/* Can the stage pipeline of k_render_levels render this voice (the parameter-only part of lane_needs_generic with
 * levelled = true, free_kernel.cuh)?  What it cannot — S&H, quantize, noise, reverse, smoother off, odd loop windows —
 * keeps the voice's whole component in the frame-lock-step bins. */
static bool levelable_voice(const skb_engine *e, int v) {
  const skb_voice_params &p = e->par[v];
  if ((p.flags & (SKB_F_NOISE | SKB_F_REVERSE)) || !(p.flags & SKB_F_SMOOTHER) || p.sample_hold_max != 0 || p.quantize != 0) return false;
  if (p.table_id < 0 || p.table_id >= (int)e->tables.size() || p.table_size <= 0 || p.table_size > 8388608) return false;
  if (p.cz_mode < 0 || p.cz_mode > 7) return false;
  const bool use_loop = (p.flags & SKB_F_LOOP_ENABLED) && (p.flags & SKB_F_LOOP_VALID);
  const float lo = use_loop ? p.loop_start_f : 0.0f, hi = use_loop ? p.loop_end_f : (float)p.table_size;
  if (!(lo >= 0.0f) || !(hi <= (float)p.table_size) || !(hi > lo)) return false;
  int live[4];
  skb_live_mods(&p, v, e->n, live);
  const bool fm = live[0] >= 0;
  if (!fm && (!(lo == 0.0f) || !(p.phase_inc >= 0.0f && p.phase_inc < hi))) return false;
  /* self-references read the voice's own value of this or the last frame: lock-step */
  if (p.cz_mode != 0 && p.cz_mod_osc == v && p.cz_mod_depth != 0.0f) return false;
  if (p.amp_mod_osc == v || (p.pan_mod_osc == v && !(p.flags & SKB_F_DISCONNECT))) return false;
  if (p.cz_mode != 0 && live[1] < 0 && (p.table_size & (p.table_size - 1)) != 0) return false;   /* (untraced CZ needs a 2^k table) */
  return true;
}

static int replan(skb_engine *e, cudaStream_t st) {
  const int n = e->n, rank = e->cfg.rank, world = e->cfg.world;
  std::vector<int32_t> old_owner;
  if (e->planned && world > 1) old_owner = e->owner;
  skb_components(e->par.data(), n, e->comp.data());
  /* ownership is sticky: a re-plan keeps every voice on its rank unless its component merged with one that lives elsewhere */
  skb_partition_stable(e->comp.data(), n, world, old_owner.empty() ? nullptr : old_owner.data(), e->owner.data());
  std::vector<int32_t> migrating;                     /* voices whose evolving state has to change GPUs, ascending */
  if (e->planned && world > 1 && e->stats.frames_rendered > 0) {
    for (int v = 0; v < n; v++)
      if (old_owner[v] != e->owner[v]) migrating.push_back(v);
    if (!migrating.empty() && !e->comm) {
      e->owner = old_owner;
      return fail(e, SKB_ERR_STATE, "a modulation edge joined voices of two shards: their state has to move, which needs the engines' communicator (skb_comm_init_rank / skb_comm_init_all)");
    }
  }
  if (!migrating.empty()) {
    /* every rank computes the same list (the host mirrors are identical).  The old owner of each voice gathers its record
     * from its (old) slot; one broadcast per source rank carries the records to everybody; the new owners scatter the
     * ones they now own once the new slots exist (below).  Order on `st`: after every launch that wrote the state. */
    const size_t nm = migrating.size();
    cudaError_t rr;
    if ((rr = grow_dev(&e->d_mig, &e->mig_cap, nm)) != cudaSuccess || (rr = grow_dev(&e->d_migidx, &e->migidx_cap, nm)) != cudaSuccess)
      return fail(e, SKB_ERR_CUDA, "migration alloc", cudaGetErrorString(rr));
    std::stable_sort(migrating.begin(), migrating.end(), [&old_owner](int32_t a, int32_t b) { return old_owner[a] < old_owner[b]; });
    std::vector<int> idx(nm);
    for (size_t i = 0; i < nm; i++) idx[i] = (old_owner[migrating[i]] == rank) ? e->slot_of_voice[migrating[i]] : -1;
    CK(cudaMemcpyAsync(e->d_migidx, idx.data(), nm * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                    /* pageable source */
    k_gather_state<<<(int)((nm + 127) / 128), 128, 0, st>>>(e->d_sq[e->cur], e->cap, e->d_migidx, (int)nm, e->d_mig);
    e->stats.kernel_launches++;
    for (size_t i = 0; i < nm;) {
      size_t j = i;
      const int src = old_owner[migrating[i]];
      while (j < nm && old_owner[migrating[j]] == src) j++;
      NK(g_nccl.Broadcast(e->d_mig + i, e->d_mig + i, (j - i) * sizeof(skb_voice_state), ncclChar, src, e->comm, st));
      i = j;
    }
  }
  std::vector<int32_t> csize(n, 0);
  for (int v = 0; v < n; v++) csize[e->comp[v]]++;
  /* free voices */
  std::vector<std::pair<uint64_t, int32_t>> fr;
  std::vector<int32_t> roots;
  for (int v = 0; v < n; v++) {
    if (e->owner[v] != rank) continue;
    if (csize[e->comp[v]] == 1) fr.push_back(std::make_pair(feature_key(&e->par[v]), v));
    else if (e->comp[v] == v) roots.push_back(v);
  }
  std::sort(fr.begin(), fr.end());
  const std::vector<int32_t> old_slot_of_voice = e->slot_of_voice;
  std::fill(e->slot_of_voice.begin(), e->slot_of_voice.end(), -1);
  std::fill(e->bin_of_voice.begin(), e->bin_of_voice.end(), -1);
  std::fill(e->level.begin(), e->level.end(), 0);
  e->voice_of_slot.assign(e->cap, -1);
  e->n_free = (int)fr.size();
  {
    /* a row of 32 slots never mixes classes: pad to the next row where the class changes */
    int sl = 0;
    for (int i = 0; i < e->n_free; i++) {
      if (i > 0 && (fr[i].first >> SKB_KEY_CLASS_SHIFT) != (fr[i - 1].first >> SKB_KEY_CLASS_SHIFT)) sl = (sl + 31) & ~31;
      e->slot_of_voice[fr[i].second] = sl;
      e->voice_of_slot[sl] = fr[i].second;
      sl++;
    }
    e->n_free_pad = (sl + 31) & ~31;
  }
  /* bins: components in order of their smallest voice, first-fit in order */
  std::vector<std::vector<int32_t>> members(roots.size());
  {
    std::vector<int32_t> ridx(n, -1);
    for (size_t i = 0; i < roots.size(); i++) ridx[roots[i]] = (int32_t)i;
    for (int v = 0; v < n; v++)
      if (e->owner[v] == rank && csize[e->comp[v]] > 1) members[ridx[e->comp[v]]].push_back(v);
  }
  e->bins.clear();
  /* LEVELS (level_kernel.cuh).  A component whose live edges form a DAG, all of whose voices the stage pipeline can
   * render, is not stepped frame by frame: its voices are rendered level by level (a voice after every voice it reads),
   * each level by one launch of k_render_levels, the modulators' samples handed on as per-frame traces.  No size limit. */
  std::fill(e->lev_of.begin(), e->lev_of.end(), -1);
  std::fill(e->trace_of.begin(), e->trace_of.end(), -1);
  for (int v = 0; v < n; v++) e->lev_deps[v].clear();
  std::vector<int32_t> lev_voices;
  int max_level = -1;
  const bool levels_on = !e->tap_on && !(e->cfg.flags & SKB_CFG_FORCE_GENERIC) && !(getenv("SKB_LEVELS") && atoi(getenv("SKB_LEVELS")) == 0);
  for (size_t i = 0; i < roots.size() && levels_on; i++) {
    const std::vector<int32_t> &mem = members[i];
    bool ok = true;
    for (size_t k = 0; k < mem.size() && ok; k++) ok = levelable_voice(e, mem[k]);
    if (!ok) continue;
    /* longest-path levels by relaxation (Kahn): indegree = live modulators inside the component */
    std::vector<int> indeg(mem.size(), 0), lvl(mem.size(), 0);
    std::vector<int32_t> local(n > 0 ? 0 : 0);
    std::vector<std::vector<int>> outs(mem.size());
    auto idx_of = [&mem](int v) { return (int)(std::lower_bound(mem.begin(), mem.end(), v) - mem.begin()); };
    for (size_t k = 0; k < mem.size(); k++) {
      int live[4];
      skb_live_mods(&e->par[mem[k]], mem[k], n, live);
      for (int q = 0; q < 4; q++) {
        if (live[q] < 0 || live[q] == mem[k]) continue;
        bool dup = false;
        for (int r = 0; r < q; r++) dup = dup || live[r] == live[q];
        if (dup) continue;
        outs[idx_of(live[q])].push_back((int)k);
        indeg[k]++;
      }
    }
    std::vector<int> queue;
    for (size_t k = 0; k < mem.size(); k++) if (indeg[k] == 0) queue.push_back((int)k);
    size_t done = 0;
    while (done < queue.size()) {
      const int k = queue[done++];
      for (size_t j = 0; j < outs[k].size(); j++) {
        const int d = outs[k][j];
        lvl[d] = std::max(lvl[d], lvl[k] + 1);
        if (--indeg[d] == 0) queue.push_back(d);
      }
    }
    if (done != mem.size()) continue;                       /* a cycle: lock-step bins */
    for (size_t k = 0; k < mem.size(); k++) {
      e->lev_of[mem[k]] = lvl[k];
      max_level = std::max(max_level, lvl[k]);
      lev_voices.push_back(mem[k]);
      for (size_t j = 0; j < outs[k].size(); j++) e->lev_deps[mem[k]].push_back(mem[outs[k][j]]);
    }
    members[i].clear();                                     /* not a bin */
  }
  /* slots of the levelled voices: after the free range, level by level (rows never mix levels), voices in index order */
  e->lev_rows.assign((size_t)(max_level + 1), std::make_pair(0, 0));
  e->n_traces = 0;
  {
    /* inside a level: by feature key like the free voices (a row's lanes then share CZ mode, filter, table: the general
     * CZ form of the gather stage switches on the mode per lane), then by index; the index rule lives in the references */
    std::vector<uint64_t> fkey((size_t)n, 0);
    for (size_t i = 0; i < lev_voices.size(); i++) fkey[lev_voices[i]] = feature_key(&e->par[lev_voices[i]]);
    std::stable_sort(lev_voices.begin(), lev_voices.end(), [e, &fkey](int32_t a2, int32_t b2) {
      if (e->lev_of[a2] != e->lev_of[b2]) return e->lev_of[a2] < e->lev_of[b2];
      if (fkey[a2] != fkey[b2]) return fkey[a2] < fkey[b2];
      return a2 < b2; });
    int sl = e->n_free_pad;
    for (size_t i = 0; i < lev_voices.size(); i++) {
      const int v = lev_voices[i];
      if (i > 0 && e->lev_of[v] != e->lev_of[lev_voices[i - 1]]) sl = (sl + 31) & ~31;
      if (i == 0 || e->lev_of[v] != e->lev_of[lev_voices[i - 1]]) e->lev_rows[e->lev_of[v]].first = sl / 32;
      e->slot_of_voice[v] = sl;
      e->voice_of_slot[sl] = v;
      if (!e->lev_deps[v].empty()) e->trace_of[v] = e->n_traces++;
      sl++;
      e->lev_rows[e->lev_of[v]].second = (sl + 31) / 32 - e->lev_rows[e->lev_of[v]].first;
    }
    e->n_lev_pad = (sl + 31) & ~31;
    e->n_lev_rows = (e->n_lev_pad - e->n_free_pad) / 32;
  }
  /* components of <= 32 voices are packed into bins of <= 32 and rendered one WARP per bin (k_render_bins_warp, batched
   * launches with in-kernel boundary ops); a larger component gets a CTA of its own (k_render_bins) */
  std::vector<std::vector<int32_t>> binv, binv_big, binv_huge;
  for (size_t i = 0; i < roots.size(); i++) {
    const int sz = (int)members[i].size();
    if (sz == 0) continue;                                  /* levelled */
    if (sz > SKB_BIN_MAX) { binv_huge.push_back(members[i]); continue; }       /* legal in the reference: rendered, slowly */
    if (sz > SKB_BIN_WARP) { binv_big.push_back(members[i]); continue; }
    if (binv.empty() || (int)binv.back().size() + sz > SKB_BIN_WARP) binv.push_back(std::vector<int32_t>());
    binv.back().insert(binv.back().end(), members[i].begin(), members[i].end());
  }
  e->n_small_bins = (int)binv.size();
  binv.insert(binv.end(), binv_big.begin(), binv_big.end());
  e->n_big_end = (int)binv.size();
  binv.insert(binv.end(), binv_huge.begin(), binv_huge.end());
  int slot = e->n_lev_pad;
  e->n_free_rows = e->n_free_pad / 32;
  /* partial-row groups: one per (CTA, batch) of k_render_free, then the bins 16 to a group */
  std::vector<int> ctarows;
  std::vector<int> lists;                    /* row lists of the time-split passes, uploaded as one array */
  {
    /* Rows -> CTAs, balanced by estimated cost (LPT: costliest row first, to the least loaded
     * CTA).  One CTA per SM; a CTA renders its rows in batches of SKB_CTA_WARPS, all rows of a
     * batch concurrently, so its time is roughly (sum of row costs) / issue rate. */
#ifndef SKB_COST_R4
#define SKB_COST_R4 36
#endif
#ifndef SKB_COST_R6
#define SKB_COST_R6 46
#endif
#ifndef SKB_COST_ONESHOT_DIV
#define SKB_COST_ONESHOT_DIV 4
#endif
    static const int cost_of_rank[8] = {0, 20, 24, 29, SKB_COST_R4, 36, SKB_COST_R6, 160};   /* instructions per voice-frame, measured */
    const int nrows = e->n_free_rows;
    std::vector<int> row_rank((size_t)nrows, 0), row_cost((size_t)nrows, 0);
    std::vector<uint8_t> row_wide((size_t)nrows, 0), row_filt((size_t)nrows, 0);
    e->n_wide_rows = 0;
    std::vector<int> xrow((size_t)std::max(nrows, 1), -1);
    e->n_xrows = 0;
    for (int r = 0; r < nrows; r++) {
      int cost = 0, rank = 0;
      bool oneshot_any = false, any = false;
      for (int l = 0; l < 32; l++) {
        const int v = e->voice_of_slot[r * 32 + l];
        if (v < 0) continue;
        any = true;
        if (e->par[v].flags & SKB_F_ONE_SHOT) oneshot_any = true;
        if (!cost) {
          const uint64_t key = feature_key(&e->par[v]);
          rank = (int)((key >> SKB_KEY_CLASS_SHIFT) & 7);
          cost = cost_of_rank[rank];
          if (e->par[v].flags & SKB_F_ONE_SHOT) cost = (cost + SKB_COST_ONESHOT_DIV - 1) / SKB_COST_ONESHOT_DIV;      /* one-shots are mostly over; the kernel packs the rest */
        }
      }
      row_rank[r] = rank; row_cost[r] = cost;
      /* time-split eligibility: a pipelined class (rows are class-pure), and a row with a filter holds no
       * one-shot voice (pass C carries the biquad across windows without the one-shot end logic) */
      const bool filt = rank == 3 || rank == 4 || rank == 6;
      row_filt[r] = filt ? 1 : 0;
      if (any && rank >= 1 && rank <= 6 && !(filt && oneshot_any) && (e->cfg.flags & SKB_CFG_WIDE) && !(e->cfg.flags & SKB_CFG_FORCE_GENERIC)) {
        row_wide[r] = 1;
        e->n_wide_rows++;
        if (filt) xrow[r] = e->n_xrows++;
      }
    }
    /* deal `items` = (cost, row entry) to at most max_ctas CTAs; returns [ctas][rows_cap], -1 padded.
     * rank != nullptr: CLASS-AFFINE dealing.  Every distinct loop body a CTA's warps run is code its SM
     * must keep streaming (each body is 3-6 KB of SASS; one more body per SM measured -20 %,
     * profiles/r01_s4_ab.txt), so the rows of a costly class (CZ and / or filter) go to a subset of the
     * CTAs sized by the class's share of the costly work — a CTA then renders ONE costly class — and the
     * light rows (plain, one-shot) fill every CTA up by LPT as before. */
    const bool tbl_affine = !(getenv("SKB_TBL_AFFINE") && atoi(getenv("SKB_TBL_AFFINE")) == 0);
    auto deal = [tbl_affine](std::vector<std::pair<int, int>> items, int max_ctas, std::vector<int> &out, int *ctas_out, int *cap_out,
                   const std::vector<int> *rank) { deal_rows(tbl_affine, items, max_ctas, out, ctas_out, cap_out, rank); };
    const std::vector<int> *affine = (e->cfg.flags & SKB_CFG_NO_AFFINE) ? nullptr : &row_rank;
    std::vector<std::pair<int, int>> items((size_t)nrows);
    for (int r = 0; r < nrows; r++) items[r] = std::make_pair(row_cost[r], r);
    deal(items, e->n_sm, ctarows, &e->free_ctas, &e->rows_cap, affine);
    auto invert = [nrows](const std::vector<int> &l, int ctas, int rcap, std::vector<int> &cta_of) {
      cta_of.assign((size_t)std::max(nrows, 1), 0);
      for (int c = 0; c < ctas; c++)
        for (int k = 0; k < rcap; k++) {
          const int rr = l[(size_t)c * rcap + k];
          if (rr >= 0) cta_of[rr & ~SKB_ROW_WIDE] = c;
        }
    };
    invert(ctarows, e->free_ctas, e->rows_cap, e->cta_of_row);
    e->free_groups = e->free_ctas * (e->rows_cap / SKB_CTA_WARPS);
    {
      /* free_lo.cu's kernel takes the first skb_lo_free_rows_per_cta() entries of a CTA's list in ONE pass: usable when the
       * lists are one batch long and nothing lies beyond those entries (a GPU with few voices) */
      const int lo_rows = skb_lo_free_rows_per_cta();
      bool ok = nrows > 0 && e->rows_cap == SKB_CTA_WARPS;
      for (int c = 0; ok && c < e->free_ctas; c++)
        for (int k = lo_rows; k < e->rows_cap; k++)
          if (ctarows[(size_t)c * e->rows_cap + k] >= 0) { ok = false; break; }
      e->lo_ok = ok;
    }
    /* pass A of a time-split launch: the same rows, the wide ones flagged and costed as the light body */
    auto append = [&lists](const std::vector<int> &l, int ctas, int rcap) {
      skb_engine::RowList rl; rl.off = (int)lists.size(); rl.ctas = ctas; rl.rows_cap = rcap;
      lists.insert(lists.end(), l.begin(), l.end());
      return rl;
    };
    e->list_b.clear();
    e->list_a_wide = skb_engine::RowList(); e->list_c = skb_engine::RowList();
    if (e->n_wide_rows > 0) {
      std::vector<int> l; int ctas = 0, rcap = 0;
      for (int r = 0; r < nrows; r++) items[r] = row_wide[r] ? std::make_pair(5, r | SKB_ROW_WIDE) : std::make_pair(row_cost[r], r);
      deal(items, e->n_sm, l, &ctas, &rcap, affine);
      if (ctas != e->free_ctas || rcap != e->rows_cap) return fail(e, SKB_ERR_STATE, "planner: pass A list shape");
      e->list_a_wide = append(l, ctas, rcap);
      invert(l, ctas, rcap, e->cta_of_row_wide);
      /* pass C: rows with a filter, the biquad / gain / mix part */
      items.clear();
      for (int r = 0; r < nrows; r++) if (row_wide[r] && row_filt[r]) items.push_back(std::make_pair(16, r));
      if (!items.empty()) { deal(items, e->n_sm, l, &ctas, &rcap, nullptr); e->list_c = append(l, ctas, rcap); }
      /* pass B, per number of windows W of a launch: floor(#SM / W) CTAs per window */
      items.clear();
      for (int r = 0; r < nrows; r++)
        if (row_wide[r]) items.push_back(std::make_pair(row_filt[r] ? std::max(row_cost[r] - 14, 4) : row_cost[r], r));
      const int wmax = (e->cfg.max_frames + SKB_ENV_WIN - 1) / SKB_ENV_WIN;
      e->list_b.assign((size_t)wmax + 1, skb_engine::RowList());
      for (int W = SKB_WIDE_MIN_WIN; W <= wmax && W <= e->n_sm; W++) {
        deal(items, std::max(1, e->n_sm / W), l, &ctas, &rcap, nullptr);
        e->list_b[W] = append(l, ctas, rcap);
      }
      e->snap_nwin = wmax;
    }
    {
      cudaError_t rr = grow_dev(&e->d_xrow, &e->xrow_cap, xrow.size());
      if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "x row map alloc", cudaGetErrorString(rr));
      CK(cudaMemcpyAsync(e->d_xrow, xrow.data(), xrow.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      e->vos_dirty = true;                       /* (the bins' slots are assigned further down) */
      if (!lists.empty()) {
        rr = grow_dev(&e->d_lists, &e->lists_cap, lists.size());
        if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "row list alloc", cudaGetErrorString(rr));
        CK(cudaMemcpyAsync(e->d_lists, lists.data(), lists.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      }
      CK(cudaStreamSynchronize(st));          /* pageable sources */
      if (e->n_traces > 0) {
        rr = grow_dev(&e->d_trace, &e->trace_cap, (size_t)e->n_traces * (e->cfg.max_frames + 1));
        if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "modulator trace alloc", cudaGetErrorString(rr));
      }
      if (e->n_wide_rows > 0) {
        rr = grow_dev(&e->d_snap, &e->snap_cap, (size_t)SKB_NSQ * e->snap_nwin * e->cap);
        if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "window snapshot alloc", cudaGetErrorString(rr));
      }
      if (e->n_xrows > 0) {
        rr = grow_dev(&e->d_xs, &e->xs_cap, (size_t)e->n_xrows * e->cfg.max_frames * 32);
        if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "x scratch alloc", cudaGetErrorString(rr));
      }
    }
  }
  e->n_prows = e->free_groups + (int)binv.size() + e->n_lev_rows;      /* (CTA, batch) rows, bins, levelled rows */
  e->max_bin_threads = 0;
  for (size_t b = 0; b < binv.size(); b++) {
    std::sort(binv[b].begin(), binv[b].end());
    skb_bin_desc d;
    d.slot0 = slot; d.size = (int)binv[b].size(); d.nlevels = 1; d.row = e->free_groups + (int)b;
    for (int i = 0; i < d.size; i++) {
      const int v = binv[b][i];
      e->slot_of_voice[v] = slot + i;
      e->voice_of_slot[slot + i] = v;
      e->bin_of_voice[v] = (int)b;
      int live[4], lv = 0;
      skb_live_mods(&e->par[v], v, n, live);
      for (int k = 0; k < 4; k++)
        if (live[k] >= 0 && live[k] < v) lv = std::max(lv, e->level[live[k]] + 1);
      e->level[v] = lv;
      d.nlevels = std::max(d.nlevels, lv + 1);
    }
    e->bins.push_back(d);
    if ((int)b >= e->n_small_bins && (int)b < e->n_big_end) e->max_bin_threads = std::max(e->max_bin_threads, (d.size + 31) & ~31);
    slot += d.size;
  }
  e->n_slots = slot;
  if (e->n_slots > e->cap) return fail(e, SKB_ERR_CAPACITY, "slot capacity exceeded");
  for (int v = 0; v < n; v++) e->edge_sig[v] = edge_signature(&e->par[v], v, n);

  cudaError_t r;
  /* carry evolving state over to the new slots */
  if (wait_staging(e)) return e->err;
  if ((r = grow_pin(&e->h_idx, &e->h_idx_cap, (size_t)e->cap)) != cudaSuccess ||
      (r = grow_dev(&e->d_idx, &e->d_idx_cap, (size_t)e->cap)) != cudaSuccess)
    return fail(e, SKB_ERR_CUDA, "replan alloc", cudaGetErrorString(r));
  for (int s = 0; s < e->cap; s++) {
    const int v = e->voice_of_slot[s];
    e->h_idx[s] = (v >= 0) ? old_slot_of_voice[v] : -1;
  }
  CK(cudaMemcpyAsync(e->d_idx, e->h_idx, (size_t)e->cap * sizeof(int), cudaMemcpyHostToDevice, st));
  {
    const int total = e->cap * SKB_NSQ;
    k_permute_state<<<(total + 255) / 256, 256, 0, st>>>(e->d_sq[e->cur], e->d_sq[e->cur ^ 1], e->cap, e->d_idx, e->cap);
    e->stats.kernel_launches++;
    e->cur ^= 1;
  }
  if (!migrating.empty()) {
    const size_t nm = migrating.size();
    std::vector<int> idx(nm);
    for (size_t i = 0; i < nm; i++) idx[i] = (e->owner[migrating[i]] == rank) ? e->slot_of_voice[migrating[i]] : -1;
    CK(cudaMemcpyAsync(e->d_migidx, idx.data(), nm * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    k_scatter_state<<<(int)((nm + 127) / 128), 128, 0, st>>>(e->d_sq[e->cur], e->cap, e->d_migidx, (int)nm, e->d_mig);
    e->stats.kernel_launches++;
    e->stats.migrated_voices += nm;
  }
  CK(cudaMemsetAsync(e->d_pq, 0, (size_t)SKB_NPQ * e->cap * sizeof(float4), st));
  CK(cudaEventRecord(e->ev_h2d, st));
  if (!e->bins.empty()) {
    if ((int)e->bins.size() > e->d_bins_cap) {
      cudaFree(e->d_bins);
      e->d_bins_cap = (int)e->bins.size() * 2;
      CK(cudaMalloc((void **)&e->d_bins, (size_t)e->d_bins_cap * sizeof(skb_bin_desc)));
    }
    /* pageable source: synchronous with respect to the host, ordered on st */
    CK(cudaMemcpyAsync(e->d_bins, e->bins.data(), e->bins.size() * sizeof(skb_bin_desc), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    if ((int)e->bins.size() > e->n_big_end) {
      cudaError_t rx = grow_dev(&e->d_binx, &e->binx_cap, (size_t)(SKB_CANARY ? 5 : 3) * e->cap);
      if (rx != cudaSuccess) return fail(e, SKB_ERR_CUDA, "huge-bin exchange alloc", cudaGetErrorString(rx));
    }
  }
  {
    cudaError_t rr;
    rr = grow_dev(&e->d_ctarows, &e->ctarows_cap, ctarows.size());
    if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "row list alloc", cudaGetErrorString(rr));
    CK(cudaMemcpyAsync(e->d_ctarows, ctarows.data(), ctarows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    e->h_ctarows = ctarows;
    rr = grow_dev(&e->d_ctaphase, &e->ctaphase_cap, (size_t)std::max(e->n_sm, 1) * (8 + SKB_CTA_WARPS));
    if (rr != cudaSuccess) return fail(e, SKB_ERR_CUDA, "phase clock alloc", cudaGetErrorString(rr));
  }
  /* every owned voice gets a fresh record */
  std::fill(e->noise_flag.begin(), e->noise_flag.end(), 0);
  e->noise_count = 0;
  for (size_t i = 0; i < e->dirty_list.size(); i++) e->dirty[e->dirty_list[i]] = 0;
  e->dirty_list.clear();
  for (int v = 0; v < n; v++)
    if (e->slot_of_voice[v] >= 0) { e->dirty[v] = 1; e->dirty_list.push_back(v); }
  e->planned = true;
  e->need_plan = false;
  e->stats.replans++;
  e->stats.n_free_voices = e->n_free;
  {
    /* voices rendered by the modulation kernels (levelled rows are padded to 32 per level: count voices, not slots) */
    int ng = 0;
    for (int sl = e->n_free_pad; sl < e->n_slots; sl++) ng += e->voice_of_slot[(size_t)sl] >= 0;
    e->stats.n_group_voices = ng;
  }
  e->stats.n_groups = (int)e->bins.size();
  e->stats.n_owned_voices = e->n_free + e->stats.n_group_voices;
  return SKB_OK;
}

/* Work is ordered by stream; when the caller switches streams (its own NCCL
 * stream vs the engine's), drain the previous one first. */
static int use_stream(skb_engine *e, cudaStream_t st) {
  if (e->last_stream && e->last_stream != st) CK(cudaStreamSynchronize(e->last_stream));
  e->last_stream = st;
  return SKB_OK;
}

/* Bring the device up to date with everything the host sent: plan, parameter
 * records, ordered ops. */
/* defer_ops: leave the queued ops for the kernel to apply at its first boundary when it can (no bins). */
static int sync_inputs(skb_engine *e, cudaStream_t st, bool defer_ops = false) {
  if (e->err) return e->err;
  if (!e->planned) e->need_plan = true;
  if (!e->need_plan)
    for (size_t i = 0; i < e->dirty_list.size(); i++) {
      const int v = e->dirty_list[i];
      if (edge_signature(&e->par[v], v, e->n) != e->edge_sig[v]) { e->need_plan = true; break; }
      if (e->lev_of[v] >= 0 && !levelable_voice(e, v)) { e->need_plan = true; break; }   /* e.g. `h` / `q` / `b` on a levelled voice */
    }
  if (!e->need_plan) {
    /* the readers of a changed modulator carry its phase increment in their own records */
    const size_t nd = e->dirty_list.size();
    for (size_t i = 0; i < nd; i++) {
      const int v = e->dirty_list[i];
      if (e->lev_of[v] < 0) continue;
      for (size_t k = 0; k < e->lev_deps[v].size(); k++) {
        const int d = e->lev_deps[v][k];
        if (!e->dirty[d]) { e->dirty[d] = 1; e->dirty_list.push_back(d); }
      }
    }
  }
  if (e->need_plan && replan(e, st)) return e->err;
  cudaError_t r;
  if (!e->dirty_list.empty()) {
    size_t cnt = 0;
    for (size_t i = 0; i < e->dirty_list.size(); i++) cnt += e->slot_of_voice[e->dirty_list[i]] >= 0;
    if (wait_staging(e)) return e->err;
    if ((r = grow_pin(&e->h_idx, &e->h_idx_cap, cnt)) != cudaSuccess ||
        (r = grow_dev(&e->d_idx, &e->d_idx_cap, cnt)) != cudaSuccess ||
        (r = grow_pin(&e->h_recs, &e->h_recs_cap, cnt * SKB_NPQ)) != cudaSuccess ||
        (r = grow_dev(&e->d_recs, &e->d_recs_cap, cnt * SKB_NPQ)) != cudaSuccess)
      return fail(e, SKB_ERR_CUDA, "param staging alloc", cudaGetErrorString(r));
    size_t j = 0;
    for (size_t i = 0; i < e->dirty_list.size(); i++) {
      const int v = e->dirty_list[i];
      e->dirty[v] = 0;
      if (e->slot_of_voice[v] < 0) continue;
      e->h_idx[j] = e->slot_of_voice[v];
      pack_record(e, v, e->h_recs + j * SKB_NPQ);
      const uint8_t nz = (e->par[v].flags & SKB_F_NOISE) ? 1 : 0;
      e->noise_count += (int)nz - (int)e->noise_flag[v];
      e->noise_flag[v] = nz;
      j++;
    }
    e->dirty_list.clear();
    if (cnt) {
      CK(cudaMemcpyAsync(e->d_idx, e->h_idx, cnt * sizeof(int), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(e->d_recs, e->h_recs, cnt * SKB_NPQ * sizeof(float4), cudaMemcpyHostToDevice, st));
      e->stats.h2d_bytes += (uint64_t)(cnt * (SKB_NPQ * sizeof(float4) + sizeof(int)));
      const int total = (int)cnt * SKB_NPQ;
      k_scatter_params<<<(total + 255) / 256, 256, 0, st>>>(e->d_pq, e->cap, e->d_idx, e->d_recs, (int)cnt);
      e->stats.kernel_launches++;
      CK(cudaEventRecord(e->ev_h2d, st));
    }
    e->any_noise = e->noise_count > 0;
  }
  if (!e->ops.empty() && !(defer_ops && bins_batch_ok(e) && !(e->cfg.flags & SKB_CFG_NO_BATCH))) {
    /* keep per-voice order: stable sort by slot, then one run per slot */
    std::vector<skb_op> &ops = e->ops;
    size_t cnt = 0;
    for (size_t i = 0; i < ops.size(); i++) {
      const int v = ops[i].voice;
      if (v < 0 || v >= e->n || e->slot_of_voice[v] < 0) continue;
      ops[cnt] = ops[i];
      ops[cnt].voice = e->slot_of_voice[v];
      cnt++;
    }
    e->stats.ops_applied += ops.size();
    ops.resize(cnt);
    if (cnt) {
      std::stable_sort(ops.begin(), ops.end(), [](const skb_op &a, const skb_op &b) { return a.voice < b.voice; });
      if (wait_staging(e)) return e->err;
      if ((r = grow_pin(&e->h_ops, &e->h_ops_cap, cnt)) != cudaSuccess ||
          (r = grow_pin(&e->h_runs, &e->h_runs_cap, cnt)) != cudaSuccess ||
          (r = grow_dev(&e->d_ops, &e->d_ops_cap, cnt)) != cudaSuccess ||
          (r = grow_dev(&e->d_runs, &e->d_runs_cap, cnt)) != cudaSuccess)
        return fail(e, SKB_ERR_CUDA, "op staging alloc", cudaGetErrorString(r));
      memcpy(e->h_ops, ops.data(), cnt * sizeof(skb_op));
      int nruns = 0;
      for (size_t i = 0; i < cnt;) {
        size_t j = i + 1;
        while (j < cnt && ops[j].voice == ops[i].voice) j++;
        e->h_runs[nruns++] = make_int2((int)i, (int)j);
        i = j;
      }
      CK(cudaMemcpyAsync(e->d_ops, e->h_ops, cnt * sizeof(skb_op), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(e->d_runs, e->h_runs, (size_t)nruns * sizeof(int2), cudaMemcpyHostToDevice, st));
      e->stats.h2d_bytes += (uint64_t)(cnt * sizeof(skb_op) + (size_t)nruns * sizeof(int2));
      k_apply_ops<<<(nruns + 127) / 128, 128, 0, st>>>(e->d_sq[e->cur], e->cap, e->d_ops, e->d_runs, nruns);
      e->stats.kernel_launches++;
      CK(cudaEventRecord(e->ev_h2d, st));
    }
    ops.clear();
  }
  return SKB_OK;
}

/* ---- exchange step: the skb_comm_* entry points (NCCL loader and NK() are defined near the top) ---- */
int skb_comm_unique_id(void *id_out) {
  if (!id_out) return SKB_ERR_ARG;
  if (!nccl_load()) return SKB_ERR_STATE;
  static_assert(SKB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return SKB_ERR_CUDA;
  memcpy(id_out, &id, SKB_COMM_ID_BYTES);
  return SKB_OK;
}

int skb_comm_init_rank(skb_engine *e, const void *id, int rank, int nranks) {
  if (!e || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(e, SKB_ERR_ARG, "comm_init_rank: bad argument");
  /* a sharded engine joins as its shard; an engine that holds ALL its voices (world = 1) may join any communicator
   * (N independent full renders whose mixes are summed: the weak-scaling job of bench.py) */
  if (e->cfg.world != 1 && (rank != e->cfg.rank || nranks != e->cfg.world))
    return fail(e, SKB_ERR_ARG, "comm_init_rank: rank / nranks differ from the engine's voice shard (skb_config.rank / world)");
  if (e->comm) return fail(e, SKB_ERR_STATE, "comm_init_rank: the engine already has a communicator");
  if (!nccl_load()) return fail(e, SKB_ERR_STATE, "comm_init_rank: libnccl.so.2 not found (set SKB_NCCL_LIB)");
  cudaSetDevice(e->cfg.device);
  ncclUniqueId nid;
  memcpy(&nid, id, SKB_COMM_ID_BYTES);
  NK(g_nccl.CommInitRank(&e->comm, nranks, nid, rank));
  e->comm_n = nranks; e->comm_rank = rank;
  return e->err;
}

int skb_comm_init_all(skb_engine *const *engines, int n) {
  if (!engines || n < 1) return SKB_ERR_ARG;
  skb_engine *e = engines[0];
  if (!nccl_load()) return fail(e, SKB_ERR_STATE, "comm_init_all: libnccl.so.2 not found (set SKB_NCCL_LIB)");
  std::vector<int> devs((size_t)n);
  for (int r = 0; r < n; r++) {
    if (!engines[r] || engines[r]->cfg.rank != r || engines[r]->cfg.world != n || engines[r]->comm)
      return fail(e, SKB_ERR_ARG, "comm_init_all: engines[r] must be created with rank = r, world = n and hold no communicator");
    devs[r] = engines[r]->cfg.device;
  }
  std::vector<ncclComm_t> comms((size_t)n);
  NK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
  for (int r = 0; r < n; r++) { engines[r]->comm = comms[r]; engines[r]->comm_n = n; engines[r]->comm_rank = r; }
  return SKB_OK;
}

int skb_comm_set_mode(skb_engine *e, int mode) {
  if (!e || (mode != SKB_COMM_NCCL_REDUCE && mode != SKB_COMM_ORDERED)) return fail(e, SKB_ERR_ARG, "comm_set_mode: bad mode");
  e->comm_mode = mode;
  return e->err;
}

int skb_comm_size(const skb_engine *e) { return (e && e->comm) ? e->comm_n : 0; }

int skb_comm_destroy(skb_engine *e) {
  if (!e) return SKB_ERR_ARG;
  if (e->comm) {
    cudaSetDevice(e->cfg.device);
    if (e->last_stream) cudaStreamSynchronize(e->last_stream);
    g_nccl.CommDestroy(e->comm);
    e->comm = nullptr; e->comm_n = 0;
  }
  return SKB_OK;
}

/* rank 0, ordered mode: mix[f] = part[0][f] + part[1][f] + ... + part[n-1][f], left to right */
__global__ void k_sum_ranks(const float2 *__restrict__ parts, int nranks, int stride, float2 *__restrict__ mix, int nframes) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nframes) return;
  float2 a = parts[f];
  for (int r = 1; r < nranks; r++) { const float2 b = parts[(size_t)r * stride + f]; a.x += b.x; a.y += b.y; }
  mix[f] = a;
}

/* The exchange step in three parts, so that one host thread can put the NCCL calls of all its engines in one group. */
static int reduce_prepare(skb_engine *e, float *d_mix, int nframes, cudaStream_t st) {
  if (!d_mix || nframes < 0 || nframes > e->cfg.max_frames) return fail(e, SKB_ERR_ARG, "reduce_mix: bad argument");
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return e->err;
  if (!e->comm) {
    if (e->cfg.world == 1) return e->err;                     /* one shard, nobody to add: the partial mix is the mix */
    return fail(e, SKB_ERR_STATE, "reduce_mix: world > 1 and no communicator (skb_comm_init_rank / skb_comm_init_all)");
  }
  if (e->comm_n == 1 || nframes == 0) return e->err;
  /* A caller may run the exchange on a stream of its own so that it overlaps the next render: the reduce is then
   * ordered after everything queued on the render stream so far by an event (no host synchronisation, and the
   * engine keeps launching on its render stream); making the NEXT writer of d_mix wait for `stream` is the caller's. */
  if (e->last_stream && e->last_stream != st) {
    CK(cudaEventRecord(e->ev_comm, e->last_stream));
    CK(cudaStreamWaitEvent(st, e->ev_comm, 0));
  } else if (!e->last_stream) {
    e->last_stream = st;
  }
  if (e->comm_mode == SKB_COMM_ORDERED && e->comm_rank == 0 && (size_t)e->comm_n * e->cfg.max_frames > e->gather_cap) {
    CK(cudaStreamSynchronize(st));
    cudaError_t r = grow_dev(&e->d_gather, &e->gather_cap, (size_t)e->comm_n * e->cfg.max_frames);
    if (r != cudaSuccess) return fail(e, SKB_ERR_CUDA, "gather alloc", cudaGetErrorString(r));
  }
  return e->err;
}

static int reduce_enqueue(skb_engine *e, float *d_mix, int nframes, cudaStream_t st) {
  if (!e->comm || e->comm_n == 1 || nframes == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  if (e->comm_mode == SKB_COMM_NCCL_REDUCE) {
    NK(g_nccl.Reduce(d_mix, d_mix, (size_t)nframes * 2, ncclFloat32, ncclSum, 0, e->comm, st));
  } else if (e->comm_rank == 0) {
    const int mf = e->cfg.max_frames;
    NK(g_nccl.GroupStart());
    for (int r = 1; r < e->comm_n; r++)
      NK(g_nccl.Recv(e->d_gather + (size_t)r * mf, (size_t)nframes * 2, ncclFloat32, r, e->comm, st));
    NK(g_nccl.GroupEnd());
  } else {
    NK(g_nccl.Send(d_mix, (size_t)nframes * 2, ncclFloat32, 0, e->comm, st));
  }
  return e->err;
}

static int reduce_complete(skb_engine *e, float *d_mix, int nframes, cudaStream_t st) {
  if (!e->comm || e->comm_n == 1 || nframes == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  if (e->comm_mode == SKB_COMM_ORDERED && e->comm_rank == 0) {
    /* rank 0's own partial sum is row 0; then rank 1, 2, ... in that order (k_sum_ranks) */
    const int mf = e->cfg.max_frames;
    CK(cudaMemcpyAsync(e->d_gather, d_mix, (size_t)nframes * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    k_sum_ranks<<<(nframes + 255) / 256, 256, 0, st>>>(e->d_gather, e->comm_n, mf, (float2 *)d_mix, nframes);
    e->stats.kernel_launches++;
  }
  CK(cudaGetLastError());
  return e->err;
}

int skb_reduce_mix(skb_engine *e, float *d_mix, int nframes, void *stream) {
  if (!e) return SKB_ERR_ARG;
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (reduce_prepare(e, d_mix, nframes, st)) return e->err;
  if (reduce_enqueue(e, d_mix, nframes, st)) return e->err;
  return reduce_complete(e, d_mix, nframes, st);
}

/* One host thread, several engines (skb_comm_init_all): NCCL wants the ranks' calls of one collective inside a group.
 * Every engine reduces its own mix buffer (skb_mix_buffer) on its own stream. */
int skb_reduce_mix_all(skb_engine *const *engines, int n, int nframes) {
  if (!engines || n < 1 || !engines[0]) return SKB_ERR_ARG;
  skb_engine *e = engines[0];
  for (int r = 0; r < n; r++) {
    if (!engines[r]) return fail(e, SKB_ERR_ARG, "reduce_mix_all: null engine");
    if (n > 1 && !engines[r]->comm) return fail(e, SKB_ERR_STATE, "reduce_mix_all: engine without a communicator");
    if (reduce_prepare(engines[r], (float *)engines[r]->d_mix, nframes, engines[r]->stream)) return engines[r]->err;
  }
  if (n == 1) return e->err;
  NK(g_nccl.GroupStart());
  int rc = SKB_OK;
  for (int r = 0; r < n && rc == SKB_OK; r++) rc = reduce_enqueue(engines[r], (float *)engines[r]->d_mix, nframes, engines[r]->stream);
  NK(g_nccl.GroupEnd());
  for (int r = 0; r < n && rc == SKB_OK; r++) rc = reduce_complete(engines[r], (float *)engines[r]->d_mix, nframes, engines[r]->stream);
  return rc;
}

int skb_owns_voice(skb_engine *e, int voice) {
  if (!e || voice < 0 || voice >= e->n) return 0;
  if (e->cfg.world == 1) return 1;
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return 0;
  if (use_stream(e, e->stream) || sync_inputs(e, e->stream)) return 0;
  return e->owner[voice] == e->cfg.rank;
}

/* Launch the pending batch (no host synchronisation). */
static int batch_launch(skb_engine *e) {
  if (!e->batch.open) return e->err;
  e->batch.open = false;
  const double t_bl0 = host_now_us();
  const int nframes = e->batch.frames;
  if (nframes == 0) return e->err;
  cudaStream_t st = e->batch.st;
  const int nwin = (int)e->batch.win_frames.size();
  const size_t nops = e->batch.ops.size();
  e->batch.win_ob.push_back((int)nops);                       /* CSR sentinel */
  /* time-split launch?  Needs enough windows to split, eligible rows, and no modulation bins */
  const bool wide = e->n_wide_rows > 0 && nwin >= SKB_WIDE_MIN_WIN && nwin < (int)e->list_b.size() &&
                    e->list_b[nwin].ctas > 0 && e->bins.empty() && (e->cfg.flags & SKB_CFG_WIDE) && !e->tap_on;
  /* per boundary: by (CTA that renders the voice, slot), queue order within a voice; the kernel gets one CSR
   * row per (boundary, CTA), so a CTA touches only its own ops — and most boundaries hold none for it */
  /* few rows: one CTA per row (k_render_rows); its op buckets are the rows */
  const bool rows = !wide && !e->tap_on && e->n_free_rows > 0 &&
                    (e->rows_mode == 1 || (e->rows_mode == 2 && e->n_free_rows <= e->rows_auto_max));
  const std::vector<int> &cta_of = wide ? e->cta_of_row_wide : e->cta_of_row;
  const int ncta = rows ? e->n_free_rows : std::max(e->free_ctas, 1);
  const int n_free_pad = e->n_free_pad;
  /* op buckets: the CTAs (rows) of the free kernel, then the bins (one bucket), then the levelled rows (one each) */
  const int nbk = ncta + 1 + e->n_lev_rows;
  const int n_lev_pad = e->n_lev_pad;
  auto cta_of_slot = [&cta_of, n_free_pad, n_lev_pad, rows, ncta](int slot) {
    if (slot >= 0 && slot < n_free_pad) return rows ? (slot >> 5) : cta_of[(size_t)(slot >> 5)];
    if (slot >= n_free_pad && slot < n_lev_pad) return ncta + 1 + ((slot - n_free_pad) >> 5);
    return ncta; };
  for (size_t i = 0; i < nops; i++) e->batch.ops[i]._pad = cta_of_slot(e->batch.ops[i].voice);   /* bucket key; the kernel ignores it */
  cudaError_t r;
  if (wait_staging(e)) return e->err;
  /* ONE staging block per launch — [window frames | per-(boundary, CTA) CSR | ops | wake bits] — copied by one
   * H2D on a COPY stream, double buffered: while launch k renders, the host stages and uploads launch k + 1,
   * and the render stream only waits on an event that has long fired (three dependent copies on the render
   * stream cost ~15-20 us of a 0.4 ms step). */
  const size_t nwi = (size_t)nwin + (size_t)nwin * (nbk + 1);
  const size_t nw = nops ? ((size_t)e->cap + 31) / 32 : 0;
  const size_t off_ops = (nwi * sizeof(int) + 15) & ~(size_t)15;
  const size_t off_wake = off_ops + nops * sizeof(skb_op);
  const size_t stage_bytes = off_wake + nw * sizeof(uint32_t);
  const int sb = e->stage_idx;
  e->stage_idx ^= 1;
  CK(cudaEventSynchronize(e->ev_stage_done[sb]));          /* the launch that last read this block is over (two launches ago) */
  if ((r = grow_pin(&e->h_stage[sb], &e->h_stage_cap[sb], stage_bytes)) != cudaSuccess ||
      (r = grow_dev(&e->d_stage[sb], &e->d_stage_cap[sb], stage_bytes)) != cudaSuccess)
    return fail(e, SKB_ERR_CUDA, "launch staging alloc", cudaGetErrorString(r));
  char *hs = e->h_stage[sb];
  char *ds = e->d_stage[sb];
  memcpy(hs, e->batch.win_frames.data(), (size_t)nwin * sizeof(int));
  {
    /* counting sort of every boundary's ops by CTA straight into the staging block (stable: queue order within
     * a voice survives), then by slot inside the few-op buckets; the bucket starts are the CSR rows */
    int *csr = (int *)hs + nwin;                              /* [nwin][nbk + 1] */
    skb_op *outp = (skb_op *)(hs + off_ops);
    e->sort_cursor.assign((size_t)nbk + 1, 0);
    int *cur = e->sort_cursor.data();
    for (int w = 0; w < nwin; w++) {
      int *row = csr + (size_t)w * (nbk + 1);
      const int b = e->batch.win_ob[w], en = e->batch.win_ob[w + 1];
      for (int c = 0; c <= nbk; c++) cur[c] = 0;
      for (int i = b; i < en; i++) cur[e->batch.ops[i]._pad]++;
      int pos = b;
      for (int c = 0; c < nbk; c++) { row[c] = pos; pos += cur[c]; cur[c] = row[c]; }
      row[nbk] = en;
      for (int i = b; i < en; i++) outp[cur[e->batch.ops[i]._pad]++] = e->batch.ops[i];
      for (int c = 0; c < nbk; c++) {
        const int s0 = row[c], s1 = row[c + 1];
        for (int i = s0 + 1; i < s1; i++) {                   /* insertion sort by slot, stable */
          const skb_op key = outp[i];
          int j = i;
          while (j > s0 && outp[j - 1].voice > key.voice) { outp[j] = outp[j - 1]; j--; }
          outp[j] = key;
        }
      }
    }
  }
  const int *d_winp = (const int *)ds;
  const skb_op *d_bopsp = (const skb_op *)(ds + off_ops);
  const unsigned *d_wake = nullptr;
  if (nops) {
    memcpy(hs + off_wake, e->batch.wake.data(), nw * sizeof(uint32_t));
    d_wake = (const unsigned *)(ds + off_wake);
    for (size_t i = 0; i < e->batch.wake_words.size(); i++) e->batch.wake[e->batch.wake_words[i]] = 0u;
    e->batch.wake_words.clear();
  }
  CK(cudaMemcpyAsync(ds, hs, stage_bytes, cudaMemcpyHostToDevice, e->copy_stream));
  e->stats.h2d_bytes += (uint64_t)(stage_bytes);
  CK(cudaEventRecord(e->ev_stage_copied[sb], e->copy_stream));
  CK(cudaStreamWaitEvent(st, e->ev_stage_copied[sb], 0));
  if (e->any_noise) {
    CK(cudaMemcpyAsync(e->d_noise, e->h_noise, (size_t)nframes * sizeof(float), cudaMemcpyHostToDevice, st));
    e->stats.h2d_bytes += (uint64_t)((size_t)nframes * sizeof(float));
  }
  CK(cudaEventRecord(e->ev_h2d, st));
  const skb_engine::RowList lb = wide ? e->list_b[nwin] : skb_engine::RowList();
  const int groups_b = wide ? lb.ctas * (lb.rows_cap / SKB_CTA_WARPS) : 0;
  const int groups_c = (wide && e->list_c.ctas > 0) ? e->list_c.ctas * (e->list_c.rows_cap / SKB_CTA_WARPS) : 0;
  /* rows mode: the bins keep their partial rows [free_groups, n_prows), the free rows follow at n_prows */
  const int n_prows_now = wide ? e->free_groups + groups_b + groups_c : (rows ? e->n_prows + e->n_free_rows : e->n_prows);
  const size_t n_prow = (size_t)std::max(n_prows_now, 1);
  if (n_prow * (size_t)e->cfg.max_frames > e->partials_cap) {
    CK(cudaStreamSynchronize(st));
    r = grow_dev(&e->d_partials, &e->partials_cap, n_prow * (size_t)e->cfg.max_frames);
    if (r != cudaSuccess) return fail(e, SKB_ERR_CUDA, "partials alloc", cudaGetErrorString(r));
  }
  if (e->tap_on) {
    if (e->vos_dirty) {
      r = grow_dev(&e->d_vos, &e->vos_cap, (size_t)e->cap);
      if (r != cudaSuccess) return fail(e, SKB_ERR_CUDA, "tap slot map alloc", cudaGetErrorString(r));
      CK(cudaMemcpyAsync(e->d_vos, e->voice_of_slot.data(), (size_t)e->cap * sizeof(int), cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));            /* pageable source */
      e->vos_dirty = false;
    }
    /* skipped, disconnected and not-owned voices read 0 (synth.c:534-541, 609-611) */
    CK(cudaMemsetAsync(e->d_tap + (size_t)e->batch.tap_frame0 * e->n, 0, (size_t)nframes * e->n * sizeof(float2), st));
  }
  CK(cudaEventRecord(e->ev_t0, st));
  if (e->n_free_rows > 0) {
    FreeArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.pq = e->d_pq; fa.sq = e->d_sq[e->cur]; fa.cap = e->cap; fa.n_rows = e->n_free_rows; fa.n_free = e->n_free_pad;
    fa.cta_rowlist = wide ? e->d_lists + e->list_a_wide.off : e->d_ctarows; fa.rows_cap = e->rows_cap;
    fa.tables = e->d_tables; fa.noise = e->d_noise;
    fa.nframes = nframes; fa.ssc_before = (unsigned long long)e->batch.ssc0;
    fa.win_frames = d_winp; fa.win_ob = d_winp + nwin; fa.nwin = nwin; fa.ob_stride = nbk + 1;
    fa.bops = d_bopsp; fa.wake = d_wake;
    fa.ctarows = e->d_partials; fa.row_stride = nframes;
    fa.envbuf = nullptr; fa.counters = e->d_counters; fa.cta_phase = e->d_ctaphase;
    {
      /* envelope triggers / releases of this launch join the running count; a quarter of the voices in a transient asks
       * for the long envelope slices (more shared memory, less L1) */
      e->env_activity = e->env_activity * exp2(-(double)nframes / 22050.0) + (double)e->env_ops_pending;
      e->env_ops_pending = 0;
      const bool many = e->env_activity > 0.25 * (double)std::max(e->n_free, 1);
      fa.env_warp_floats = e->env_floats_forced ? e->env_floats_forced : (many ? SKB_ENV_WARP_FLOATS_LARGE : SKB_ENV_WARP_FLOATS_SMALL);
    }
    const size_t free_smem = skb_free_smem_bytes(fa.env_warp_floats);
    fa.force_generic = (e->cfg.flags & SKB_CFG_FORCE_GENERIC) ? 1 : 0;
    fa.wide_on = wide ? 1 : 0;
    fa.snap = e->d_snap; fa.snap_nwin = e->snap_nwin;
    fa.xs = e->d_xs; fa.xs_frames = e->cfg.max_frames; fa.xrow_of = e->d_xrow;
    fa.cpw = 0; fa.group0 = 0;
    fa.tap = e->tap_on ? e->d_tap + (size_t)e->batch.tap_frame0 * e->n : nullptr; fa.voice_of_slot = e->d_vos; fa.tap_n = e->tap_on ? e->n : 0;
    { const char *pp = getenv("SKB_PHASE_PASS"); fa.phase_pass = pp ? atoi(pp) : 0; }
    fa.warp_diag = e->d_ctaphase ? e->d_ctaphase + (size_t)e->n_sm * 8 : nullptr;
    if (wide)       /* "no snapshot" = finished: group 0 of every window's snapshot */
      CK(cudaMemsetAsync(e->d_snap, 0x01, (size_t)nwin * e->cap * sizeof(float4), st));
    if (rows) {
      fa.group0 = e->n_prows;
      k_render_rows<<<e->n_free_rows, RP_THREADS, skb_rows_smem_bytes(), st>>>(fa);
      e->stats.rows_launches++;
    } else if (e->tap_on) k_render_free_tap<<<e->free_ctas, SKB_CTA_THREADS, free_smem, st>>>(fa);
    else if (e->lo_ok && e->lo_mode) { skb_lo_free_launch(&fa, e->free_ctas, skb_lo_free_smem_bytes(fa.env_warp_floats), st); e->lo_launches++; e->stats.lo_launches++; }
    else k_render_free<<<e->free_ctas, SKB_CTA_THREADS, free_smem, st>>>(fa);
    e->stats.kernel_launches++;
    if (wide) {
      CK(cudaEventRecord(e->ev_a, st));
      FreeArgs fb = fa;
      fb.wide_on = 0; fb.bops = nullptr; fb.wake = nullptr;
      fb.cta_rowlist = e->d_lists + lb.off; fb.rows_cap = lb.rows_cap; fb.cpw = lb.ctas; fb.group0 = e->free_groups;
      k_render_window<<<lb.ctas * nwin, SKB_CTA_THREADS, free_smem, st>>>(fb);
      e->stats.kernel_launches++;
      CK(cudaEventRecord(e->ev_b, st));
      if (groups_c > 0) {
        FreeArgs fc = fb;
        fc.cta_rowlist = e->d_lists + e->list_c.off; fc.rows_cap = e->list_c.rows_cap; fc.cpw = 0;
        fc.group0 = e->free_groups + groups_b;
        k_render_biquad<<<e->list_c.ctas, SKB_CTA_THREADS, free_smem, st>>>(fc);
        e->stats.kernel_launches++;
      }
      e->stats.wide_launches++;
      e->wide_timing_pending = true;
    }
  }
  if (e->n_small_bins > 0) {                                  /* one warp per bin, boundary ops applied in-kernel */
    BinWarpArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.pq = e->d_pq; ba.sq = e->d_sq[e->cur]; ba.cap = e->cap; ba.bins = e->d_bins; ba.nbins = e->n_small_bins;
    ba.tables = e->d_tables; ba.noise = e->d_noise; ba.nframes = nframes; ba.ssc_before = (unsigned long long)e->batch.ssc0;
    ba.win_frames = d_winp; ba.win_ob = d_winp + nwin; ba.nwin = nwin; ba.ob_stride = nbk + 1; ba.ob_bucket = ncta;
    ba.bops = d_bopsp; ba.partials = e->d_partials; ba.row_stride = nframes; ba.counter = e->d_counters + 1;
    ba.tap = e->tap_on ? e->d_tap + (size_t)e->batch.tap_frame0 * e->n : nullptr; ba.voice_of_slot = e->d_vos; ba.tap_n = e->tap_on ? e->n : 0;
    k_render_bins_warp<<<(e->n_small_bins + SKB_BINW_WARPS - 1) / SKB_BINW_WARPS, SKB_BINW_WARPS * 32, 0, st>>>(ba);
    e->stats.kernel_launches++;
  }
  if ((int)e->bins.size() > e->n_big_end) {                   /* cyclic components above 1,024 voices: the slow, correct fallback */
    k_render_bins_huge<<<(int)e->bins.size() - e->n_big_end, 1024, 0, st>>>(e->d_pq, e->d_sq[e->cur], e->cap, e->d_bins + e->n_big_end, e->d_tables,
                                                         e->d_noise, nframes, (unsigned long long)e->batch.ssc0,
                                                         e->d_partials, nframes, e->d_counters + 1,
                                                         e->tap_on ? e->d_tap + (size_t)e->batch.tap_frame0 * e->n : nullptr, e->d_vos,
                                                         e->tap_on ? e->n : 0, e->d_binx);
    e->stats.kernel_launches++;
  }
  if (e->n_big_end > e->n_small_bins) {                       /* big components: one CTA each, one launch per callback, ops by k_apply_ops */
    const int nt = e->max_bin_threads;
    const size_t smem = (size_t)3 * nt * sizeof(float) + (size_t)2 * (nt / 32) * sizeof(float2) + (SKB_CANARY ? (size_t)2 * nt * sizeof(int) : 0);
    k_render_bins<<<e->n_big_end - e->n_small_bins, nt, smem, st>>>(e->d_pq, e->d_sq[e->cur], e->cap, e->d_bins + e->n_small_bins, e->d_tables,
                                                         e->d_noise, nframes, (unsigned long long)e->batch.ssc0,
                                                         e->d_partials, nframes, e->d_counters + 1,
                                                         e->tap_on ? e->d_tap + (size_t)e->batch.tap_frame0 * e->n : nullptr, e->d_vos,
                                                         e->tap_on ? e->n : 0);
    e->stats.kernel_launches++;
  }
  if (e->n_lev_rows > 0) {
    /* one launch per level, in level order: a level reads the traces the levels before it wrote */
    FreeArgs fl;
    memset(&fl, 0, sizeof(fl));
    fl.pq = e->d_pq; fl.sq = e->d_sq[e->cur]; fl.cap = e->cap; fl.n_rows = e->n_lev_rows; fl.n_free = e->n_lev_pad;
    fl.tables = e->d_tables; fl.noise = e->d_noise;
    fl.nframes = nframes; fl.ssc_before = (unsigned long long)e->batch.ssc0;
    fl.win_frames = d_winp; fl.win_ob = d_winp + nwin; fl.nwin = nwin; fl.ob_stride = nbk + 1;
    fl.bops = d_bopsp; fl.wake = d_wake;
    fl.ctarows = e->d_partials; fl.row_stride = nframes;
    fl.counters = e->d_counters; fl.force_generic = 0;
    fl.trace = e->d_trace; fl.trace_stride = e->cfg.max_frames + 1;
    const int first_row = e->n_free_pad / 32;
    for (size_t L = 0; L < e->lev_rows.size(); L++) {
      if (e->lev_rows[L].second <= 0) continue;
      fl.row0 = e->lev_rows[L].first;
      fl.group0 = e->free_groups + (int)e->bins.size() + (fl.row0 - first_row);
      fl.ob_bucket0 = ncta + 1 + (fl.row0 - first_row);
      k_render_levels<<<e->lev_rows[L].second, RP_THREADS, skb_levels_smem_bytes(), st>>>(fl);
      e->stats.kernel_launches++;
    }
  }
  if (n_prows_now > 0) {
    dim3 blk(SKB_RED_X, SKB_RED_Y);
    dim3 grd((nframes + SKB_RED_X - 1) / SKB_RED_X, std::max(1, std::min(SKB_RED_CHUNKS, n_prows_now / 16)));
    const int skip = rows ? e->free_groups : 0;                /* rows mode: k_render_free's rows are not written */
    k_reduce_rows<<<grd, blk, 0, st>>>(e->d_partials + (size_t)skip * nframes, n_prows_now - skip, nframes, nframes, e->d_part2,
                                       e->d_tickets, (float2 *)e->batch.mix);
    e->stats.kernel_launches++;
  } else {
    CK(cudaMemsetAsync(e->batch.mix, 0, (size_t)nframes * sizeof(float2), st));
  }
  CK(cudaEventRecord(e->ev_t1, st));
  CK(cudaEventRecord(e->ev_stage_done[sb], st));
  e->stats.host_us[0] += host_now_us() - t_bl0;
  e->timing_pending = true;
  CK(cudaGetLastError());
  e->batch.frames = 0;
  e->batch.win_frames.clear(); e->batch.win_ob.clear(); e->batch.ops.clear();
  return e->err;
}

int skb_flush(skb_engine *e) {
  if (!e) return SKB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  return batch_launch(e);
}

int skb_render_mix(skb_engine *e, int nframes, uint64_t ssc_before, const float *noise,
                   float *d_mix, void *stream) {
  if (!e) return SKB_ERR_ARG;
  if (nframes < 0 || nframes > e->cfg.max_frames || !d_mix) return fail(e, SKB_ERR_ARG, "render_mix: bad argument");
  if (e->tap_on && e->tap_cursor + nframes > e->cfg.max_frames)
    return fail(e, SKB_ERR_ARG, "render_mix: tap on and more than max_frames rendered since the last skb_finish");
  cudaSetDevice(e->cfg.device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (e->err) return e->err;
  /* Can this segment ride in the pending launch?  Yes if it continues it (same stream, next
   * frames, next mix address), fits, and everything the host queued since is a pure state edit
   * (ops): those are applied by the kernel itself at the window boundary.  Parameter changes,
   * re-plans and modulation bins end the batch. */
  bool extend = e->batch.open && e->batch.st == st && e->batch.mix + (size_t)e->batch.frames * 2 == d_mix &&
                e->batch.ssc0 + (uint64_t)e->batch.frames == ssc_before &&
                e->batch.frames + nframes <= e->cfg.max_frames && bins_batch_ok(e) && e->planned && !e->need_plan &&
                e->dirty_list.empty() && !(e->cfg.flags & SKB_CFG_NO_BATCH);
  if (extend && !e->ops.empty() && e->batch.frames % SKB_ENV_WIN != 0) extend = false;   /* (boundaries are window starts) */
  if (extend && e->n_lev_rows > 0)
    /* voice_reset clears voice_sample between two callbacks: a modulator's trace cannot say "0 from here on, but the old
     * value for the frame before" — such a boundary starts a new launch (trace[0] then holds the cleared value) */
    for (size_t i = 0; i < e->ops.size() && extend; i++) if (e->ops[i].code == SKB_OP_VOICE_CLEAR) extend = false;
  if (!extend) {
    if (batch_launch(e)) return e->err;
    if (use_stream(e, st)) return e->err;
    if (sync_inputs(e, st, true)) return e->err;  /* plan, parameter records; ops -> k_apply_ops only if the kernel cannot take them */
    if (nframes == 0) return e->err;
    e->batch.open = true; e->batch.st = st; e->batch.mix = d_mix; e->batch.frames = 0; e->batch.ssc0 = ssc_before;
    e->batch.tap_frame0 = e->tap_cursor;
    e->batch.win_frames.clear(); e->batch.win_ob.clear(); e->batch.ops.clear();
    if (e->batch.wake.size() < ((size_t)e->cap + 31) / 32) e->batch.wake.assign(((size_t)e->cap + 31) / 32, 0u);
  } else if (nframes == 0) {
    return e->err;
  }
  if (e->any_noise) {
    if (!noise) return fail(e, SKB_ERR_ARG, "render_mix: a voice uses the shared noise source but noise == NULL");
    if (e->batch.frames == 0 && wait_staging(e)) return e->err;
    memcpy(e->h_noise + e->batch.frames, noise, (size_t)nframes * sizeof(float));
  }
  /* this segment's windows; the queued ops belong to the boundary before its first window */
  e->batch.win_ob.push_back((int)e->batch.ops.size());
  if (!e->ops.empty()) {          /* (a new batch: the boundary before its first window) */
    for (size_t i = 0; i < e->ops.size(); i++) {
      skb_op op = e->ops[i];
      const int v = op.voice;
      if (v < 0 || v >= e->n || e->slot_of_voice[v] < 0) continue;
      op.voice = e->slot_of_voice[v];
      e->batch.ops.push_back(op);
      const int w = op.voice >> 5;
      if (!e->batch.wake[w]) e->batch.wake_words.push_back(w);
      e->batch.wake[w] |= 1u << (op.voice & 31);
    }
    e->stats.ops_applied += e->ops.size();
    e->ops.clear();
  }
  for (int done = 0; done < nframes; done += SKB_ENV_WIN) {
    if (done) e->batch.win_ob.push_back((int)e->batch.ops.size());
    e->batch.win_frames.push_back(std::min(SKB_ENV_WIN, nframes - done));
  }
  e->batch.frames += nframes;
  if (e->tap_on) e->tap_cursor += nframes;
  e->stats.frames_rendered += (uint64_t)nframes;
  if (!bins_batch_ok(e) || (e->cfg.flags & SKB_CFG_NO_BATCH)) return batch_launch(e);
  return e->err;
}

int skb_finish(skb_engine *e, const float *d_mix, int nframes, const float *gain, float *out,
               int num_channels, void *stream) {
  if (!e) return SKB_ERR_ARG;
  if (!d_mix || !gain || !out || num_channels < 2 || nframes < 0 || nframes > e->cfg.max_frames)
    return fail(e, SKB_ERR_ARG, "finish: bad argument");
  if (nframes == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (batch_launch(e)) return e->err;
  const double t_f0 = host_now_us();
  if (use_stream(e, st)) return e->err;
  if (wait_staging(e)) return e->err;
  /* the master-volume trace: one value when the one-pole sits on its fixed point (the usual case) */
  bool gain_flat = true;
  for (int i = 1; i < nframes && gain_flat; i++) gain_flat = gain[i] == gain[0];
  if (!(e->finish_direct && gain_flat)) {
    memcpy(e->h_gain, gain, (size_t)nframes * sizeof(float));
    CK(cudaMemcpyAsync(e->d_gain, e->h_gain, (size_t)nframes * sizeof(float), cudaMemcpyHostToDevice, st));
    e->stats.h2d_bytes += (uint64_t)((size_t)nframes * sizeof(float));
    CK(cudaEventRecord(e->ev_h2d, st));
  }
  double t_f1, t_f2;
  if (e->finish_direct) {
    const unsigned int seq = ++e->fin_seq ? e->fin_seq : ++e->fin_seq;       /* (never 0: the flag's idle value) */
    k_finish_host<<<(nframes + 255) / 256, 256, 0, st>>>((const float2 *)d_mix, gain_flat ? nullptr : e->d_gain, gain[0], e->dv_out,
                                                         nframes, e->d_counters, e->dv_counters, SKB_N_COUNTERS, e->d_fin_ticket,
                                                         e->dv_flag, seq);
    e->stats.kernel_launches++;
    e->stats.d2h_bytes += (uint64_t)nframes * sizeof(float2) + SKB_N_COUNTERS * sizeof(unsigned long long);
    t_f1 = host_now_us();
    /* poll the flag; every few thousand reads ask the stream, so that a failed launch ends the wait with its error */
    unsigned int spins = 0;
    cudaError_t qs = cudaErrorNotReady;
    while (*e->h_flag != seq) {
      if ((++spins & 0xfffu) == 0u) {
        qs = cudaStreamQuery(st);
        if (qs != cudaErrorNotReady) break;
      }
    }
    if (*e->h_flag != seq) {                       /* the stream ended (or failed) before we saw the flag */
      CK(qs == cudaErrorNotReady ? cudaSuccess : qs);
      CK(cudaStreamSynchronize(st));
      if (*e->h_flag != seq) return fail(e, SKB_ERR_CUDA, "finish: the completion flag was not written");
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    t_f2 = host_now_us();
  } else {
    k_finish<<<(nframes + 255) / 256, 256, 0, st>>>((const float2 *)d_mix, e->d_gain, e->d_out, nframes);
    e->stats.kernel_launches++;
    CK(cudaMemcpyAsync(e->h_out, e->d_out, (size_t)nframes * sizeof(float2), cudaMemcpyDeviceToHost, st));
    e->stats.d2h_bytes += (uint64_t)nframes * sizeof(float2) + SKB_N_COUNTERS * sizeof(unsigned long long);
    CK(cudaMemcpyAsync(e->h_counters, e->d_counters, SKB_N_COUNTERS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    t_f1 = host_now_us();
    CK(cudaStreamSynchronize(st));
    t_f2 = host_now_us();
  }
  e->stats.host_us[1] += t_f1 - t_f0; e->stats.host_us[2] += t_f2 - t_f1;
  e->stats.active_voice_frames = e->h_counters[0] + e->h_counters[1];
  for (int i = 0; i < 8; i++) e->stats.class_rows[i] = e->h_counters[2 + i];
  for (int i = 0; i < 8; i++) e->stats.phase_cycles[i] = e->h_counters[10 + i];
  e->stats.cta_batches = e->h_counters[18];
  e->stats.wide_errors = e->h_counters[19];
  e->tap_cursor = 0;
  if (num_channels == 2) {
    memcpy(out, e->h_out, (size_t)nframes * sizeof(float2));
  } else {
    const float *h = e->h_out;
    for (int i = 0; i < nframes; i++) {           /* only channels 0 and 1 are written, synth.c:623-624 */
      out[(size_t)i * num_channels + 0] = h[2 * i + 0];
      out[(size_t)i * num_channels + 1] = h[2 * i + 1];
    }
  }
  if (e->timing_pending) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1) == cudaSuccess) e->stats.last_render_ms = ms;
    else cudaGetLastError();
    e->stats.last_wide_ms[0] = e->stats.last_wide_ms[1] = e->stats.last_wide_ms[2] = 0.0f;
    if (e->wide_timing_pending) {
      if (cudaEventElapsedTime(&ms, e->ev_t0, e->ev_a) == cudaSuccess) e->stats.last_wide_ms[0] = ms; else cudaGetLastError();
      if (cudaEventElapsedTime(&ms, e->ev_a, e->ev_b) == cudaSuccess) e->stats.last_wide_ms[1] = ms; else cudaGetLastError();
      if (cudaEventElapsedTime(&ms, e->ev_b, e->ev_t1) == cudaSuccess) e->stats.last_wide_ms[2] = ms; else cudaGetLastError();
      e->wide_timing_pending = false;
    }
    e->timing_pending = false;
  }
  return e->err;
}

int skb_render(skb_engine *e, int nframes, uint64_t ssc_before, const float *gain, const float *noise,
               float *out, int num_channels) {
  if (!e) return SKB_ERR_ARG;
  if (nframes < 0 || !gain || !out) return fail(e, SKB_ERR_ARG, "render: bad argument");
  int done = 0;
  while (done < nframes) {
    const int n = std::min(nframes - done, e->cfg.max_frames);
    int r = skb_render_mix(e, n, ssc_before + (uint64_t)done, noise ? noise + done : nullptr, (float *)e->d_mix, nullptr);
    if (r) return r;
    r = skb_finish(e, (const float *)e->d_mix, n, gain + done, out + (size_t)done * num_channels, num_channels, nullptr);
    if (r) return r;
    done += n;
  }
  return e->err;
}

float *skb_mix_buffer(skb_engine *e) { return e ? (float *)e->d_mix : nullptr; }

int skb_set_tap(skb_engine *e, int enable) {
  if (!e) return SKB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return e->err;
  if (enable && !e->d_tap) {
    const size_t need = (size_t)e->cfg.max_frames * e->n;
    if (need * sizeof(float2) > ((size_t)16 << 30))
      return fail(e, SKB_ERR_CAPACITY, "tap: max_frames x voices x 8 bytes exceeds 16 GiB; create the engine with a smaller max_frames");
    cudaError_t r = grow_dev(&e->d_tap, &e->tap_cap, need);
    if (r != cudaSuccess) return fail(e, SKB_ERR_CUDA, "tap alloc", cudaGetErrorString(r));
  }
  if (e->tap_on != (enable != 0)) e->need_plan = true;     /* (levelled voices have no tap: with the tap on they go to the bins) */
  e->tap_on = enable != 0;
  e->tap_cursor = 0;
  return e->err;
}

int skb_read_tap(skb_engine *e, int frame0, int nframes, float *out) {
  if (!e || !out || frame0 < 0 || nframes < 0 || frame0 + nframes > e->cfg.max_frames) return fail(e, SKB_ERR_ARG, "read_tap: bad argument");
  if (!e->tap_on || !e->d_tap) return fail(e, SKB_ERR_STATE, "read_tap: the tap is off");
  if (nframes == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return e->err;
  cudaStream_t st = e->last_stream ? e->last_stream : e->stream;
  /* straight into the caller's (pageable) buffer: the copy is as large as the tap itself */
  CK(cudaMemcpyAsync(out, e->d_tap + (size_t)frame0 * e->n, (size_t)nframes * e->n * sizeof(float2), cudaMemcpyDeviceToHost, st));
  e->stats.d2h_bytes += (uint64_t)nframes * e->n * sizeof(float2);
  CK(cudaStreamSynchronize(st));
  return e->err;
}

int skb_set_tap_voices(skb_engine *e, const int32_t *voices, int n) {
  if (!e || n < 0 || (n > 0 && !voices)) return fail(e, SKB_ERR_ARG, "set_tap_voices: bad argument");
  for (int i = 0; i < n; i++) if (voices[i] < 0 || voices[i] >= e->n) return fail(e, SKB_ERR_ARG, "set_tap_voices: bad voice");
  e->tap_sel.assign(voices, voices + n);
  e->tapcol_dirty = true;
  return e->err;
}

int skb_read_tap_selected(skb_engine *e, int frame0, int nframes, float *out, float *extremes) {
  if (!e || !out || frame0 < 0 || nframes < 0 || frame0 + nframes > e->cfg.max_frames) return fail(e, SKB_ERR_ARG, "read_tap_selected: bad argument");
  if (!e->tap_on || !e->d_tap) return fail(e, SKB_ERR_STATE, "read_tap_selected: the tap is off");
  if (e->tap_sel.empty()) {                                  /* everything selected: the plain copy, no voice left out */
    if (extremes) extremes[0] = extremes[1] = 0.0f;
    return skb_read_tap(e, frame0, nframes, out);
  }
  if (nframes == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return e->err;
  cudaStream_t st = e->last_stream ? e->last_stream : e->stream;
  const int nsel = (int)e->tap_sel.size();
  cudaError_t r;
  if (e->tapcol_dirty) {
    std::vector<int> col((size_t)e->n, -1);
    for (int i = 0; i < nsel; i++) col[e->tap_sel[i]] = i;
    if ((r = grow_dev(&e->d_tapcol, &e->tapcol_cap, (size_t)e->n)) != cudaSuccess) return fail(e, SKB_ERR_CUDA, "tap column map alloc", cudaGetErrorString(r));
    CK(cudaMemcpyAsync(e->d_tapcol, col.data(), (size_t)e->n * sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                           /* pageable source */
    e->tapcol_dirty = false;
  }
  if (!e->d_tapext) CK(cudaMalloc((void **)&e->d_tapext, 2 * sizeof(unsigned int)));
  const size_t need = (size_t)nframes * nsel;
  if ((r = grow_dev(&e->d_tapsel, &e->tapsel_cap, need)) != cudaSuccess || (r = grow_pin(&e->h_tap, &e->h_tap_cap, need + 1)) != cudaSuccess)
    return fail(e, SKB_ERR_CUDA, "tap selection alloc", cudaGetErrorString(r));
  CK(cudaMemsetAsync(e->d_tapext, 0, 2 * sizeof(unsigned int), st));
  const size_t items = (size_t)e->n * nframes;
  k_tap_select<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(e->d_tap + (size_t)frame0 * e->n, e->n, nframes, e->d_tapcol, nsel,
                                                               e->d_tapsel, e->d_tapext);
  e->stats.kernel_launches++;
  CK(cudaMemcpyAsync(e->h_tap, e->d_tapsel, need * sizeof(float2), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(e->h_tap + need, e->d_tapext, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  e->stats.d2h_bytes += (uint64_t)(need * sizeof(float2) + 8);
  CK(cudaStreamSynchronize(st));
  memcpy(out, e->h_tap, need * sizeof(float2));
  if (extremes) memcpy(extremes, e->h_tap + need, 2 * sizeof(float));
  return e->err;
}

int skb_sync(skb_engine *e, void *stream) {
  if (!e) return SKB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  if (batch_launch(e)) return e->err;
  CK(cudaMemcpyAsync(e->h_counters, e->d_counters, SKB_N_COUNTERS * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                     stream ? (cudaStream_t)stream : e->stream));
  CK(cudaStreamSynchronize(stream ? (cudaStream_t)stream : e->stream));
  e->stats.active_voice_frames = e->h_counters[0] + e->h_counters[1];
  for (int i = 0; i < 8; i++) e->stats.class_rows[i] = e->h_counters[2 + i];
  for (int i = 0; i < 8; i++) e->stats.phase_cycles[i] = e->h_counters[10 + i];
  e->stats.cta_batches = e->h_counters[18];
  e->stats.wide_errors = e->h_counters[19];
  if (e->timing_pending) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, e->ev_t0, e->ev_t1) == cudaSuccess) e->stats.last_render_ms = ms;
    else cudaGetLastError();
    e->stats.last_wide_ms[0] = e->stats.last_wide_ms[1] = e->stats.last_wide_ms[2] = 0.0f;
    if (e->wide_timing_pending) {
      if (cudaEventElapsedTime(&ms, e->ev_t0, e->ev_a) == cudaSuccess) e->stats.last_wide_ms[0] = ms; else cudaGetLastError();
      if (cudaEventElapsedTime(&ms, e->ev_a, e->ev_b) == cudaSuccess) e->stats.last_wide_ms[1] = ms; else cudaGetLastError();
      if (cudaEventElapsedTime(&ms, e->ev_b, e->ev_t1) == cudaSuccess) e->stats.last_wide_ms[2] = ms; else cudaGetLastError();
      e->wide_timing_pending = false;
    }
    e->timing_pending = false;
  }
  return e->err;
}

static int stage_slots(skb_engine *e, int first, int n) {
  cudaError_t r;
  if (wait_staging(e)) return e->err;
  if ((r = grow_pin(&e->h_idx, &e->h_idx_cap, (size_t)n)) != cudaSuccess ||
      (r = grow_dev(&e->d_idx, &e->d_idx_cap, (size_t)n)) != cudaSuccess ||
      (r = grow_pin(&e->h_snap, &e->h_snap_cap, (size_t)n)) != cudaSuccess ||
      (r = grow_dev(&e->d_vsnap, &e->d_vsnap_cap, (size_t)n)) != cudaSuccess)
    return fail(e, SKB_ERR_CUDA, "snapshot alloc", cudaGetErrorString(r));
  for (int i = 0; i < n; i++) e->h_idx[i] = e->slot_of_voice[first + i];
  return SKB_OK;
}

int skb_snapshot(skb_engine *e, int first, int n, skb_voice_state *out) {
  if (!e || first < 0 || n < 0 || first + n > e->n || !out) return fail(e, SKB_ERR_ARG, "snapshot: bad range");
  if (n == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  cudaStream_t st = e->stream;
  if (batch_launch(e)) return e->err;
  if (use_stream(e, st)) return e->err;
  if (sync_inputs(e, st)) return e->err;
  if (stage_slots(e, first, n)) return e->err;
  CK(cudaMemcpyAsync(e->d_idx, e->h_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  k_gather_state<<<(n + 127) / 128, 128, 0, st>>>(e->d_sq[e->cur], e->cap, e->d_idx, n, e->d_vsnap);
  e->stats.kernel_launches++;
  CK(cudaMemcpyAsync(e->h_snap, e->d_vsnap, (size_t)n * sizeof(skb_voice_state), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(out, e->h_snap, (size_t)n * sizeof(skb_voice_state));
  return e->err;
}

int skb_restore(skb_engine *e, int first, int n, const skb_voice_state *in) {
  if (!e || first < 0 || n < 0 || first + n > e->n || !in) return fail(e, SKB_ERR_ARG, "restore: bad range");
  if (n == 0) return e->err;
  cudaSetDevice(e->cfg.device);
  cudaStream_t st = e->stream;
  if (batch_launch(e)) return e->err;
  if (use_stream(e, st)) return e->err;
  if (sync_inputs(e, st)) return e->err;
  if (stage_slots(e, first, n)) return e->err;
  memcpy(e->h_snap, in, (size_t)n * sizeof(skb_voice_state));
  CK(cudaMemcpyAsync(e->d_idx, e->h_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->d_vsnap, e->h_snap, (size_t)n * sizeof(skb_voice_state), cudaMemcpyHostToDevice, st));
  k_scatter_state<<<(n + 127) / 128, 128, 0, st>>>(e->d_sq[e->cur], e->cap, e->d_idx, n, e->d_vsnap);
  e->stats.kernel_launches++;
  CK(cudaStreamSynchronize(st));
  return e->err;
}

/* Diagnostics: phase clocks (8 per CTA, SM cycles) of the last k_render_free launch and the
 * planner's row list (rows_cap per CTA, -1 = none; row r = slots 32r..32r+31).  Returns the
 * number of CTAs written (<= max_ctas). */
int skb_debug_cta_phases(skb_engine *e, uint64_t *phases, int32_t *rows, int max_ctas, int *rows_cap) {
  if (!e || !phases || max_ctas <= 0) return SKB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  const int n = std::min(max_ctas, e->free_ctas);
  if (n <= 0 || !e->d_ctaphase) return 0;
  if (batch_launch(e)) return e->err;
  CK(cudaStreamSynchronize(e->last_stream ? e->last_stream : e->stream));
  CK(cudaMemcpy(phases, e->d_ctaphase, (size_t)n * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (rows_cap) *rows_cap = e->rows_cap;
  if (rows) memcpy(rows, e->h_ctarows.data(), (size_t)n * e->rows_cap * sizeof(int32_t));
  return n;
}

/* Diagnostics: per physical warp of the first batch of every CTA of the last launch: render cycles (bits 0-46),
 * dyn body (bit 47), live lanes (48-55), class (56-63).  Returns the number of CTAs written. */
int skb_debug_warp_clocks(skb_engine *e, uint64_t *out, int max_ctas) {
  if (!e || !out || max_ctas <= 0) return SKB_ERR_ARG;
  cudaSetDevice(e->cfg.device);
  const int n = std::min(max_ctas, e->n_sm);
  if (n <= 0 || !e->d_ctaphase) return 0;
  if (batch_launch(e)) return e->err;
  CK(cudaStreamSynchronize(e->last_stream ? e->last_stream : e->stream));
  CK(cudaMemcpy(out, e->d_ctaphase + (size_t)e->n_sm * 8, (size_t)n * SKB_CTA_WARPS * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return n;
}

/* class rank (0 silent, 1..6 pipelined bodies by cost, 7 generic) of the voice in a slot, -1 = empty */
int skb_debug_slot_rank(skb_engine *e, int slot) {
  if (!e || slot < 0 || slot >= (int)e->voice_of_slot.size()) return -1;
  const int v = e->voice_of_slot[slot];
  if (v < 0) return -1;
  return (int)((feature_key(&e->par[v]) >> SKB_KEY_CLASS_SHIFT) & 7) | ((e->par[v].flags & SKB_F_ONE_SHOT) ? 8 : 0);
}

int skb_get_stats(skb_engine *e, skb_stats *out) {
  if (!e || !out) return SKB_ERR_ARG;
  *out = e->stats;
  return SKB_OK;
}
