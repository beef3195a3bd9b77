// fealess_b200 internal declarations shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/fealess_b200.h"

#define FL_MAX_DEVICES 64        // devices of one process (per-device kernel attribute tables)
#define FL_LM_PAD 4096          // zero bytes after each label's linear memory (flat-addressing over-read)
#define FL_MAX_T 16
#define FL_SKIP 0xFFFFFFFFu     // packed-feature offset of a feature that falls outside the image

// packed feature for one pyramid level and one frame geometry: lm_off is the byte offset of the feature's first
// response inside the (level, modality) linear-memory block = label*label_stride + row*cells + (y/T)*W' + x/T
// (accessLinearMemory, reference linemod/linemod.cpp:1094-1117); x,y kept for the in-image test of similarityLocal.
struct __align__(8) fl_pfeat { uint32_t lm_off; int16_t x, y; };

struct fl_level_geom {
  int W, H, T, Wd, Hd, cells;      // image size at the level, sampling step, decimated size
  size_t label_stride;             // bytes per label (T*T*cells + pad, 32-aligned: the 4-bit copy of a label starts 16-byte aligned)
  size_t mod_stride;               // bytes per modality = 8 * label_stride
};

struct fl_sort_key { unsigned long long hi, lo; };

void fl_set_error(const char* fmt, ...);
int fl_entries_per_template(const fl_handle* h);   // n_levels * n_modalities of the handle (api.cu)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may have its CTAs resident while the previous kernel
// of the stream drains: they run their prologue and block in fl_grid_dep_wait() until that kernel has completed and its writes
// are visible; fl_grid_dep_launch() tells the scheduler that the NEXT kernel may start being placed.  Rule: a kernel touches no
// mutable global buffer before fl_grid_dep_wait().  Launched without the attribute both calls are no-ops.  Only the staged
// similarity kernel is launched this way (behind the front end); for every other kernel of the frame the attribute was measured
// to give nothing or to cost, so they use the plain fl_launch below.
#ifdef __CUDACC__
__device__ __forceinline__ void fl_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void fl_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t fl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define FL_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      fl_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return FL_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

// ---- front end (frontend.cu) -------------------------------------------------------------------
int fl_launch_tables_init();   // uploads the two LUTs to __constant__ memory of the current device
void fl_launch_pyrdown_bgr(const uint8_t* src, int W, int H, uint8_t* dst, cudaStream_t s);
void fl_launch_resize_nn_half(const uint8_t* src, int W, int H, uint8_t* dst, cudaStream_t s);
void fl_launch_apply_mask(const uint8_t* q, const uint8_t* mask, int n, uint8_t* out, cudaStream_t s);
void fl_launch_spread_lm(const uint8_t* q, fl_level_geom g, uint8_t* lm_mod, uint8_t* spread_or_null, uint8_t* lm4_mod_or_null, cudaStream_t s);

// one launch = several independent front-end jobs (see k_front_end_wave)
enum { FL_JOB_PYRDOWN = 2, FL_JOB_RESIZE = 3, FL_JOB_SPREAD = 4, FL_JOB_COLOR2 = 5, FL_JOB_DEPTH2 = 6, FL_JOB_PREFETCH = 7 };
struct fl_fe_job {
  int kind, cta_begin, gx, W, H, p0, p1;
  float thr_sq;
  const uint8_t* src; uint8_t* dst; uint8_t* dst2; uint8_t* dst3;   // spread job: dst = linear memories, dst2 = spread image (or NULL), dst3 = 4-bit linear memories (or NULL)
  fl_level_geom g;
  // in-grid dependencies (single-launch front end): a CTA of this job first waits until counters[wait_slot] has reached
  // wait_target (all CTAs of the producing job have finished), and bumps counters[signal_slot] when it is done; -1 = none
  int wait_slot, signal_slot;
  unsigned wait_target;
  // finer grain for the spread jobs: a producing tile job also counts its finished tiles PER TILE ROW (counters[row_base + by]);
  // a spread CTA that owns grid row gy waits only for the tile rows that hold its label rows [gy T, gy T + 2T - 2] (row y of
  // its level lives in producer row (y << wait_row_shift) / 16), so it starts while later rows are still being quantised
  int row_base, wait_row_base, wait_row_shift, wait_row_count;
  unsigned wait_row_target;
};
#define FL_FE_MAX_TILE_ROWS 128
// NN-downsampled copies of the level-0 depth labels written by the depth job itself: entry i = pyramid level i + 1
#define FL_FE_MAX_PYR 7
struct fl_depth_pyr { int n; int W[FL_FE_MAX_PYR], H[FL_FE_MAX_PYR]; uint8_t* dst[FL_FE_MAX_PYR]; };
#define FL_FE_MAX_JOBS 16
#define FL_FE_MAX_DPYR 2
struct fl_fe_wave { int n_jobs, n_ctas; size_t smem; int* zero_me; fl_fe_job job[FL_FE_MAX_JOBS]; int n_pyr; fl_depth_pyr pyr[FL_FE_MAX_DPYR];
                    unsigned* counters; int* dep_error; int* dep_error_host;   // dep_error_host: the same flag in mapped host memory (written only on a time-out)
                    unsigned long long* trace; };   // trace (FL_OPT_TRACE): per job {first CTA start, last CTA end} in globaltimer ns   // zero_me: int reset by CTA 0 (candidate counter), or NULL; counters: FL_FE_MAX_JOBS monotonic CTA counters
// word-parallel quantisers (frontend_v2.cuh); pyr (nullable, at most FL_FE_MAX_DPYR per wave): also write the NN pyramid of the labels
void fl_fe_add_color_v2(fl_fe_wave* w, const uint8_t* bgr, int W, int H, float thr_sq, uint8_t* q);
bool fl_fe_add_depth_v2(fl_fe_wave* w, const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* q, const fl_depth_pyr* pyr);
void fl_fe_add_pyrdown(fl_fe_wave* w, const uint8_t* src, int W, int H, uint8_t* dst, bool src_static);
// pull `bytes` at src towards L2 (one 128-byte line per thread): the similarity kernel's per-template feature lists, touched
// in the shadow of the front end so that its prologue does not start with DRAM round trips
void fl_fe_add_prefetch(fl_fe_wave* w, const void* src, size_t bytes);
void fl_fe_add_resize(fl_fe_wave* w, const uint8_t* src, int W, int H, uint8_t* dst);
void fl_fe_add_spread(fl_fe_wave* w, const uint8_t* q, fl_level_geom g, uint8_t* lm_mod, uint8_t* spread_or_null, uint8_t* lm4_mod_or_null);
cudaError_t fl_launch_fe_wave(const fl_fe_wave& w, cudaStream_t s);

// ---- template training, per-pixel half (train.cu) -----------------------------------------------
void fl_launch_train_magnitude(const uint8_t* bgr, int W, int H, float* mag, cudaStream_t s);
void fl_launch_train_erode3(const uint8_t* src, int W, int H, uint8_t* dst, cudaStream_t s);
void fl_launch_train_color_score(const uint8_t* angle, const float* mag, const uint8_t* mask, const uint8_t* eroded, int W, int H, float threshold_sq,
                                 float* score, cudaStream_t s);
void fl_launch_train_depth_score(const uint8_t* normal, const uint8_t* local_mask, uint16_t* hd, int* any_zero8, int W, int H, int extract_threshold,
                                 float* score, int* label_counts8, cudaStream_t s);

// ---- similarity / refinement / sort (similarity.cu) ---------------------------------------------
struct fl_tdb {                     // device template database
  int n_templates, L, M, n_classes;
  const fl_template_hdr_t* hdr;     // [n_templates*L*M]
  const fl_feature_t* feat;         // [n_features]
  const int32_t* class_of;          // [n_templates]
  const int32_t* class_first;       // [n_classes]
  const int32_t* tid_of;            // [n_templates] per-class template_id reported in matches (global id for shards)
  const uint8_t* class_enabled;     // [n_classes] (class filter of the current call)
  fl_pfeat* pfeat;                  // [n_features] packed for the current geometry
  // refinement records (levels 0 .. L-2), one fixed-size block per (template, level): {width, height, n_features, 0, fc[0..3]}
  // (32 B) + M x 64 feature slots of 8 B.  A candidate's template index alone gives the address of everything its refinement
  // reads, so header and features load concurrently (no dependent pointer chase) and can be prefetched into L2 by the kernel
  // that emits the candidate.  Slot = fl_pfeat; a feature outside the image at offset 0 carries lm_off = FL_RSKIP | label.
  uint8_t* rrec; int rrec_bytes;
};
#define FL_RSKIP 0xFFFFFFF0u
__host__ __device__ inline int fl_rrec_bytes(int M) { return (32 + M * 64 * 8 + 127) & ~127; }
void fl_launch_pack_refine_records(fl_tdb db, const fl_level_geom* d_geom, cudaStream_t s);
// plan of the shared-memory-staged global similarity kernel (similarity_staged.cu) for one frame geometry
struct fl_staged_plan {
  int phase_rows, n_rowblocks, n_phases;   // a phase = phase_rows linear-memory rows of one (modality, label) ...
  int labels_per_phase;                    // ... or, when whole labels fit, this many consecutive labels of one modality (1, 2 or 4)
  int halo_bytes, buf_bytes, n_buf;        // bytes staged past the last row; size and number of the smem ring buffers
  int n_words, nw_template, tpw;           // 32-bit words per similarity map; template parameters of the instantiation used
  int tpc, n_cta, block_threads;           // templates per CTA, grid, block
  int smem_bytes, pre_stride;              // dynamic shared memory; bytes per gpre row
  int stage_off;                           // emission staging area (one similarity map, nw_template x 32 words)
  uint32_t* gfeat;                         // [n_templates][64] feature words sorted by phase: word offset in the phase buffer << 5 | 8 * byte misalignment
  uint8_t* gpre;                           // [n_templates][pre_stride] prefix counts per phase
  int4* gmeta;                             // [n_templates] {template_positions, n_features, class, 0}
  unsigned long long* trace;               // developer timeline (FL_OPT_TRACE): 8 words per CTA, or NULL
};
bool fl_plan_staged(const fl_level_geom& g, int M, int n_templates, int max_positions, int n_sm, fl_staged_plan* plan);
void fl_launch_pack_staged(fl_tdb db, fl_level_geom g, fl_staged_plan plan, cudaStream_t s);
// every level's geometry and linear memories, for the refinement that shares a launch with the sort (k_refine_sort)
struct fl_refine_args {
  int n_levels;
  fl_level_geom g[FL_MAX_LEVELS];
  const uint8_t* lm[FL_MAX_LEVELS];
};
int fl_launch_similarity_staged(fl_tdb db, fl_level_geom g, const uint8_t* lm4_level, float threshold, fl_match_t* cand, int cap,
                                int* d_count, fl_staged_plan plan, cudaStream_t s);
void fl_launch_pack_features(fl_tdb db, const fl_level_geom* d_geom, int n_features_total, cudaStream_t s);
void fl_launch_similarity_global(fl_tdb db, fl_level_geom g, const uint8_t* lm_level, float threshold,
                                 fl_match_t* cand, int cap, int* d_count, cudaStream_t s);
void fl_launch_similarity_debug(fl_tdb db, fl_level_geom g, const uint8_t* lm_level, int t, uint16_t* out, cudaStream_t s);
void fl_launch_refine_level(fl_tdb db, fl_level_geom g, int level, const uint8_t* lm_level, float threshold,
                            fl_match_t* cand, int cap, const int* d_count, cudaStream_t s);
// candidate lists handed to sort + unique: list l holds min(n_in[l * n_in_stride], list_cap) records starting at
// in + l * list_stride.  (list_stride = list_cap, n_in_stride = 1 for plain arrays; the all-gather block layout
// [header record with the count in .x | list_cap records] has list_stride = list_cap + 1, n_in_stride = 5 * (list_cap + 1).)
struct fl_lists { const fl_match_t* in; int n_lists, list_cap, list_stride; const int* n_in; int n_in_stride; };
// Direct peer exchange fused into the head of the sort kernel (template-sharded matching, one process per GPU): every rank's
// exchange buffer = [FL_XCHG_SIGNALS u32 signal words][2 parities x world blocks of (cap + 1) records] is mapped into every
// process (peer[p] = base of rank p's buffer as seen from THIS process).  The kernel pushes this rank's block
// [count | live records] into slot `rank` of every peer's buffer over NVLink, publishes `epoch` in the peer's signal word
// `rank` (release, system scope), waits until all of its own signal words show `epoch` (acquire), and then sorts the
// union straight out of its own buffer.  world == 0: no exchange.  The parity (epoch & 1) double-buffers the blocks: a rank
// can be at most one frame ahead of a peer, because it cannot pass frame k + 1's wait before the peer has published
// epoch k + 1, which the peer's stream does only after its frame-k sort kernel has finished.
#define FL_XCHG_MAX_WORLD 8
#define FL_XCHG_SIGNALS 64
struct fl_xchg { int world, rank, cap; unsigned epoch; const fl_match_t* local_block; uint8_t* peer[FL_XCHG_MAX_WORLD]; };

// sort + unique: result in d_out/d_out_count, summary in d_hdr[16] = {unique count, live records, flag_big, raw counts...}.
// One launch; when flag_big comes back set the host runs fl_launch_sort_unique_big (key workspace: next_pow2(n_upper) keys
// + 1 int).  Both return the number of launches.  h_hdr / h_first: mapped pinned host copies of the summary and of the
// first h_first_cap matches (nullable).
int fl_launch_sort_unique(fl_lists L, fl_xchg X, int key_cap, fl_match_t* d_out, int out_cap, int* d_out_count, int* d_hdr,
                          int* h_hdr, fl_match_t* h_first, int h_first_cap, cudaStream_t s);
// refinement of every level + sort + unique in one launch (the last CTA to finish sorts); done_ctr: zero-initialised device int
int fl_launch_refine_sort(fl_tdb db, const fl_refine_args& ra, float threshold, fl_match_t* cand, int cap, const int* d_count, int* done_ctr, int n_sm,
                          fl_lists L, fl_xchg X, int key_cap, fl_match_t* d_out, int out_cap, int* d_out_count, int* d_hdr,
                          int* h_hdr, fl_match_t* h_first, int h_first_cap, cudaStream_t s);
int fl_launch_sort_unique_big(fl_lists L, fl_sort_key* keys, int key_cap, int n_upper, fl_match_t* d_out, int out_cap, int* d_out_count, cudaStream_t s);

// ---- ICP (icp.cu) --------------------------------------------------------------------------------
struct fl_icp_hyp {                 // one hypothesis of a batch, device-visible
  const uint16_t* model_depth;      // device, dense crop rows of rect_model.width (NULL for cloud mode)
  fl_rect_t rect_model, rect_ref;
  float r_match[9], t_match[3];
  int status;
};
struct fl_icp_ws {                  // per-batch workspace, all device pointers; max_pts is a multiple of 4
  int n_hyp, max_pts;
  float* pts_ref;                   // [n_hyp][max_pts][3]
  float* pts_mod;                   // [n_hyp][max_pts][3]  (pts_model_tmp)
  float* cor_m; float* cor_r;       // [n_hyp][max_pts][3]  ordered correspondences
  float* dist;                      // [n_hyp][max_pts]     inlier distances of the index pairs (0 for the others)
  float* nn_d2; int* nn_slot;       // [n_hyp][max_pts]     nearest reference point of every model point: squared distance, slot in the grid
  float4* grid_pts;                 // [n_hyp][max_pts]     ref points sorted by cell (x,y,z,index) - only for clouds that do not fit shared memory
  int* n_ref; int* n_mod;           // [n_hyp]              cloud mode only
};
#define FL_ICP_TRACE_WORDS 20      // developer timeline of the fused ICP launch: words per hypothesis (fl_debug_icp_trace)
#define FL_ICP_GRID 64
#define FL_ICP_CELLS (FL_ICP_GRID * FL_ICP_GRID)
// cv::resize INTER_LINEAR (resize.cu): device tables built by the host for one (source size -> destination size) pair
struct fl_resize_tables {
  int* xofs; int* yofs;             // [dW], [dH] source index of the left / upper tap (rows unclipped: the kernel clips)
  short2* ialpha; short2* ibeta;    // 8U: 11-bit fixed-point weight pairs
  float2* alpha; float2* beta;      // 16U: float weight pairs
};
void fl_launch_resize_linear(const void* src, int sW, int sH, int type, void* dst, int dW, int dH, const fl_resize_tables& t, cudaStream_t s);
void fl_launch_depth_to_3d(const uint16_t* depth, int W, int H, fl_intrinsics_t K, float* out3, cudaStream_t s);
// the whole of detection() / icpCloudToCloud_Ex for a batch in ONE persistent launch (icp.cu); hyps == NULL: cloud mode.
// ticket: device int (reset by the call); trace: NULL or FL_ICP_TRACE_WORDS words per hypothesis (SM cycles per phase).  Returns the number of launches, -1 on a launch error.
int fl_launch_icp(fl_icp_ws ws, fl_icp_params_t p, const fl_icp_hyp* hyps_or_null, const uint16_t* ref_depth, int W, int H, fl_intrinsics_t K_ref,
                  fl_icp_result_t* results, int* ticket, unsigned long long* trace, int n_sm, cudaStream_t s);
void fl_launch_nms(const float* t3, const int32_t* n_model, const float* icp_dist, int n, float th, int32_t* out_idx,
                   int32_t* out_count, cudaStream_t s);
