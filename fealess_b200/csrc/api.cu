// C ABI of fealess_b200 (include/fealess_b200.h): handle lifecycle, device workspaces, and the orchestration of the
// kernels in frontend.cu / similarity.cu / icp.cu on the handle's stream.  CUDA only - there is no CPU path here.
#include "fl_internal.cuh"
#include "resize_tables.h"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <new>
#include <vector>


static thread_local char g_err[512] = "";
void fl_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
extern "C" const char* fl_last_error(void) { return g_err; }
// Programmatic dependent launch for every kernel of the pipeline was measured neutral (126.7 vs 125.8 us per frame: the boundaries
// cost ~1.5 us each and griddepcontrol.wait still has to see the previous grid drain); only the similarity kernel, whose prologue is
// long enough to matter, is launched that way (similarity_staged.cu).
extern "C" const char* fl_version(void) { return "fealess_b200 0.1 (sm_100a)"; }

#define FETCH_FIRST 1024   // matches copied back together with the count in the common case

struct fl_handle {
  fl_params_t p;
  cudaStream_t stream;
  int64_t launches;
  // template database
  int n_templates, n_features, n_classes;
  fl_template_hdr_t* d_hdr; fl_feature_t* d_feat; int32_t* d_class_of; int32_t* d_class_first; uint8_t* d_class_enabled;
  fl_pfeat* d_pfeat; int32_t* d_tid_of; uint8_t* d_rrec;
  std::vector<int32_t> tid_of_h;
  std::vector<int32_t> coarse_wh;                          // (width, height) of every template at the coarsest level
  bool staged_eligible, use_staged; int n_sm; int force_baseline;   // staged global-similarity kernel (similarity_staged.cu)
  fl_staged_plan plan;
  std::vector<int32_t> class_first_h, class_of_h;
  std::vector<float> pose13_h;
  std::vector<uint8_t> class_enabled_h; bool class_enabled_valid;   // host mirror of d_class_enabled
  // geometry of the last frame
  int gW, gH; bool packed;
  fl_level_geom geom[FL_MAX_LEVELS]; fl_level_geom* d_geom;
  // frame-sized buffers
  uint8_t* d_in_bgr; uint16_t* d_in_depth;                 // staging for host-input calls
  uint8_t* d_bgr[FL_MAX_LEVELS];                           // colour pyramid, levels >= 1
  uint8_t* d_q[FL_MAX_LEVELS][FL_MAX_MODALITIES];          // quantised images (unmasked)
  uint8_t* d_qm[FL_MAX_LEVELS][FL_MAX_MODALITIES];         // masked copies (allocated on first use)
  uint8_t* d_mask[FL_MAX_LEVELS][FL_MAX_MODALITIES];
  uint8_t* d_spread[FL_MAX_LEVELS][FL_MAX_MODALITIES];     // debug only
  uint8_t* d_lm[FL_MAX_LEVELS]; size_t lm_bytes[FL_MAX_LEVELS]; uint8_t* d_lm4;   // d_lm4: 4-bit copy of the coarsest level's linear memories
  bool used_mask[FL_MAX_MODALITIES]; bool keep_spread;
  // state between the enqueue half (fl_match_device_async, ..._async) and fl_match_wait
  bool pend_sort, pend_match, pend_own, pend_masks_valid, pend_small_fused; fl_lists pend_lists; fl_match_t* pend_out; int pend_out_cap; int* pend_out_count;
  const void* pend_bgr; const void* pend_depth; int pend_W, pend_H; float pend_threshold; const void* pend_masks[FL_MAX_MODALITIES]; std::vector<int32_t> pend_filter;
  unsigned long long* d_fe_trace; int fe_trace_jobs, fe_trace_kind[FL_FE_MAX_JOBS], fe_trace_ctas[FL_FE_MAX_JOBS];
  bool opt_fe_waves, opt_split_refine, opt_trace, opt_dep_test;   // fl_debug_option
  int* h_fe_err; int* d_fe_err_host; bool fe_force_waves;   // in-grid dependency time-out: mapped host flag (+ its device alias); true = this handle launches one kernel per wave from now on
  unsigned* d_fe_counters; unsigned fe_counter_base[FL_FE_MAX_JOBS]; unsigned fe_row_base[FL_FE_MAX_JOBS];   // in-grid dependency counters of the single-launch front end (+ 1 error word)
  // candidates / matches
  fl_match_t* d_cand; int* d_count; fl_sort_key* d_keys; int key_cap; uint8_t* d_outblk; fl_match_t* d_out; int* d_out_count;   // d_outblk = [16-int summary][matches]
  // pinned host staging
  uint8_t* h_bgr; uint16_t* h_depth; uint8_t* h_mask; uint8_t* h_outblk; int* h_small; fl_match_t* h_first; uint8_t* h_class_enabled;   // h_small/h_first point into h_outblk
  bool have_result, overflow;
  // profiling
  bool profile; cudaEvent_t ev[5]; float stage_ms[4]; float icp_ms;
  // ICP workspace (grown on demand)
  int icp_hyp_cap, icp_pts_cap;
  fl_icp_ws icp; fl_icp_hyp* d_hyps; fl_icp_result_t* d_results; uint16_t* d_model_crops; uint16_t* d_ref_depth; int* d_icp_ticket; unsigned long long* d_icp_trace; int icp_trace_cap, icp_trace_n;
  size_t ref_depth_cap;
  uint16_t* h_model_crops; fl_icp_hyp* h_hyps; fl_icp_result_t* h_results; size_t h_crop_cap;
  // rendered template depth crops kept on the device (fl_upload_model_depths) and the geometry of the depth frame that the
  // last host-input fl_match left in d_in_depth (0 x 0 = none)
  uint16_t* d_resident; std::vector<size_t> res_off; std::vector<fl_rect_t> res_rect; int res_W, res_H;
  int in_depth_W, in_depth_H;
  int32_t* d_nms; int nms_cap;   // fl_nms workspace (hypotheses)
  bool blocking_wait; cudaEvent_t ev_block;   // fl_set_blocking_wait
  uint8_t* d_train;                           // fl_add_template workspace (allocated at the first call)
  // input rescale: device tables of the current (source -> destination) geometry, source-frame staging
  fl_resize_tables rz; int rz_sW, rz_sH, rz_dW, rz_dH;
  uint8_t* d_src_bgr; uint16_t* d_src_depth; size_t src_cap;
};

// The host waits for the handle's stream.  Default: cudaStreamSynchronize (the runtime spins - lowest latency, one busy core per
// waiting thread).  fl_set_blocking_wait(h, 1): an event created with cudaEventBlockingSync, so the thread sleeps until the GPU's
// interrupt - for hosts that keep more frames in flight (threads) than they have cores to spin on.
static cudaError_t wait_stream(fl_handle* h);
template <typename T> static int dalloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  FL_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return FL_OK;
}
template <typename T> static int halloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  FL_CUDA(cudaMallocHost((void**)p, n * sizeof(T)));
  return FL_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__ != FL_OK) return rc__; } while (0)
// temporary device buffers of one call (debug, conversion and training entry points - never the per-frame path): freed on every exit path
struct dev_buf {
  std::vector<void*> p;
  ~dev_buf() { for (void* q : p) cudaFree(q); }
  template <typename T> int get(T** out, size_t n) { int rc = dalloc(out, n); if (rc == FL_OK) p.push_back(*out); return rc; }
};

extern "C" void fl_default_params(fl_params_t* p) {
  memset(p, 0, sizeof *p);
  p->n_levels = 2; p->T[0] = 5; p->T[1] = 8;                                   // T_DEFAULTS, linemod.cpp:1820
  p->n_modalities = 2; p->modality_kind[0] = FL_MODALITY_COLOR_GRADIENT; p->modality_kind[1] = FL_MODALITY_DEPTH_NORMAL;
  p->weak_threshold = 10.0f; p->distance_threshold = 2000; p->difference_threshold = 50;
  p->max_width = 640; p->max_height = 480; p->max_candidates = 1 << 16; p->device = 0;
}

static size_t level_lm_bytes(const fl_params_t& p, int l, int W, int H, fl_level_geom* g) {
  fl_level_geom gg;
  gg.W = W >> l; gg.H = H >> l; gg.T = p.T[l];
  gg.Wd = gg.W / gg.T; gg.Hd = gg.H / gg.T; gg.cells = gg.Wd * gg.Hd;
  size_t ls = (size_t)gg.T * gg.T * gg.cells + FL_LM_PAD;
  gg.label_stride = (ls + 31) & ~(size_t)31;
  gg.mod_stride = gg.label_stride * 8;
  if (g) *g = gg;
  return gg.mod_stride * p.n_modalities;
}

extern "C" int fl_create(const fl_params_t* params, fl_handle** out) {
  if (!params || !out) return FL_ERR_ARG;
  *out = nullptr;
  const fl_params_t& p = *params;
  if (p.n_levels < 1 || p.n_levels > FL_MAX_LEVELS || p.n_modalities < 1 || p.n_modalities > FL_MAX_MODALITIES) return FL_ERR_ARG;
  for (int l = 0; l < p.n_levels; ++l) if (p.T[l] < 1 || p.T[l] > FL_MAX_T) return FL_ERR_ARG;
  for (int m = 0; m < p.n_modalities; ++m) if (p.modality_kind[m] != FL_MODALITY_COLOR_GRADIENT && p.modality_kind[m] != FL_MODALITY_DEPTH_NORMAL) return FL_ERR_ARG;
  if (p.max_width < 16 || p.max_height < 16 || p.max_candidates < 1) return FL_ERR_ARG;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || p.device < 0 || p.device >= ndev) {
    fl_set_error("no usable CUDA device (count=%d, requested %d): %s", ndev, p.device, cudaGetErrorString(e));
    return FL_ERR_CUDA;
  }
  FL_CUDA(cudaSetDevice(p.device));
  fl_handle* h = new (std::nothrow) fl_handle();
  if (!h) return FL_ERR_ARG;
  h->p = p; h->class_enabled_valid = false; h->launches = 0; h->gW = h->gH = 0; h->packed = false; h->have_result = false; h->profile = false; h->keep_spread = false;
  h->n_templates = h->n_features = h->n_classes = 0;
  h->d_hdr = nullptr; h->d_feat = nullptr; h->d_class_of = nullptr; h->d_class_first = nullptr; h->d_class_enabled = nullptr; h->d_pfeat = nullptr; h->d_rrec = nullptr;
  h->d_tid_of = nullptr;
  h->staged_eligible = false; h->use_staged = false; h->force_baseline = 0; memset(&h->plan, 0, sizeof h->plan);
  { cudaDeviceProp prop; FL_CUDA(cudaGetDeviceProperties(&prop, p.device)); h->n_sm = prop.multiProcessorCount; }
  h->icp_hyp_cap = h->icp_pts_cap = 0; memset(&h->icp, 0, sizeof h->icp);
  h->d_hyps = nullptr; h->d_results = nullptr; h->d_model_crops = nullptr; h->d_ref_depth = nullptr; h->ref_depth_cap = 0; h->d_icp_ticket = nullptr; h->d_icp_trace = nullptr; h->icp_trace_cap = h->icp_trace_n = 0;
  h->h_model_crops = nullptr; h->h_hyps = nullptr; h->h_results = nullptr; h->h_crop_cap = 0;
  h->d_resident = nullptr; h->res_W = h->res_H = 0; h->in_depth_W = h->in_depth_H = 0;
  memset(&h->rz, 0, sizeof h->rz); h->rz_sW = h->rz_sH = h->rz_dW = h->rz_dH = 0;
  h->d_src_bgr = nullptr; h->d_src_depth = nullptr; h->src_cap = 0;
  memset(h->d_bgr, 0, sizeof h->d_bgr); memset(h->d_q, 0, sizeof h->d_q); memset(h->d_qm, 0, sizeof h->d_qm);
  memset(h->d_mask, 0, sizeof h->d_mask); memset(h->d_spread, 0, sizeof h->d_spread); memset(h->d_lm, 0, sizeof h->d_lm); h->d_lm4 = nullptr; h->d_nms = nullptr; h->nms_cap = 0; h->blocking_wait = false; h->ev_block = nullptr; h->d_train = nullptr;
  memset(h->used_mask, 0, sizeof h->used_mask); memset(h->stage_ms, 0, sizeof h->stage_ms); h->icp_ms = 0.f;
  *out = h;
  FL_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 5; ++i) FL_CUDA(cudaEventCreate(&h->ev[i]));
  TRY(fl_launch_tables_init());
  const size_t npx = (size_t)p.max_width * p.max_height;
  TRY(dalloc(&h->d_in_bgr, npx * 3)); TRY(dalloc(&h->d_in_depth, npx));
  TRY(dalloc(&h->d_geom, FL_MAX_LEVELS));
  for (int l = 0; l < p.n_levels; ++l) {
    size_t n = (size_t)(p.max_width >> l) * (p.max_height >> l);
    if (l > 0) TRY(dalloc(&h->d_bgr[l], n * 3));
    for (int m = 0; m < p.n_modalities; ++m) TRY(dalloc(&h->d_q[l][m], n));
    // linear memories: T need not divide the maximum size, so bound the cell count from above
    int T = p.T[l];
    size_t cells = (size_t)(((p.max_width >> l) + T - 1) / T) * (((p.max_height >> l) + T - 1) / T);
    size_t ls = (((size_t)T * T * cells + FL_LM_PAD) + 31) & ~(size_t)31;
    h->lm_bytes[l] = ls * 8 * p.n_modalities + cells + 256;   // slack: windowed reads may run one map past the last label
    TRY(dalloc(&h->d_lm[l], h->lm_bytes[l]));
    if (l == p.n_levels - 1) {
      // 4-bit copy of the coarsest level's linear memories (two cells per byte), read by the staged similarity kernel
      TRY(dalloc(&h->d_lm4, h->lm_bytes[l] / 2 + 256));
      FL_CUDA(cudaMemset(h->d_lm4, 0, h->lm_bytes[l] / 2 + 256));
    }
  }
  TRY(dalloc(&h->d_cand, (size_t)p.max_candidates)); TRY(dalloc(&h->d_count, 4));
  FL_CUDA(cudaMemset(h->d_count, 0, 4 * sizeof(int)));
  // [FL_FE_MAX_JOBS job counters][1 error word][FL_FE_MAX_JOBS x FL_FE_MAX_TILE_ROWS tile-row counters]
  TRY(dalloc(&h->d_fe_counters, FL_FE_MAX_JOBS + 1 + FL_FE_MAX_JOBS * FL_FE_MAX_TILE_ROWS));
  FL_CUDA(cudaMemset(h->d_fe_counters, 0, (FL_FE_MAX_JOBS + 1 + FL_FE_MAX_JOBS * FL_FE_MAX_TILE_ROWS) * sizeof(unsigned)));
  memset(h->fe_counter_base, 0, sizeof h->fe_counter_base); memset(h->fe_row_base, 0, sizeof h->fe_row_base);
  FL_CUDA(cudaHostAlloc((void**)&h->h_fe_err, 64, cudaHostAllocMapped)); *h->h_fe_err = 0; h->fe_force_waves = false;
  h->opt_fe_waves = h->opt_split_refine = h->opt_trace = h->opt_dep_test = false;
  FL_CUDA(cudaHostGetDevicePointer((void**)&h->d_fe_err_host, h->h_fe_err, 0));
  h->pend_sort = h->pend_match = false; h->pend_small_fused = false; h->d_fe_trace = nullptr; h->fe_trace_jobs = 0;                             // [0] candidate count, [1] CTA ticket counter of k_refine_sort
  int kc = 2; while (kc < p.max_candidates) kc <<= 1;
  h->key_cap = kc;
  TRY(dalloc(&h->d_keys, (size_t)kc + 1));      // + room for the two scratch ints behind the keys
  TRY(dalloc(&h->d_outblk, 64 + (size_t)p.max_candidates * sizeof(fl_match_t)));
  FL_CUDA(cudaMemset(h->d_outblk, 0, 64));                                         // summary words, incl. the fused-tail overflow flag
  h->d_out_count = reinterpret_cast<int*>(h->d_outblk); h->d_out = reinterpret_cast<fl_match_t*>(h->d_outblk + 64);
  TRY(halloc(&h->h_bgr, npx * 3)); TRY(halloc(&h->h_depth, npx)); TRY(halloc(&h->h_mask, npx * p.n_modalities));
  TRY(halloc(&h->h_outblk, 64 + FETCH_FIRST * sizeof(fl_match_t))); TRY(halloc(&h->h_class_enabled, 4096));
  h->h_small = reinterpret_cast<int*>(h->h_outblk); h->h_first = reinterpret_cast<fl_match_t*>(h->h_outblk + 64);
  return FL_OK;
}

static void icp_free(fl_handle* h) {
  cudaFree(h->d_icp_trace); h->d_icp_trace = nullptr; h->icp_trace_cap = h->icp_trace_n = 0;
  cudaFree(h->icp.pts_ref); cudaFree(h->icp.pts_mod); cudaFree(h->icp.cor_m); cudaFree(h->icp.cor_r); cudaFree(h->icp.dist);
  cudaFree(h->icp.grid_pts); cudaFree(h->icp.nn_d2); cudaFree(h->icp.nn_slot); cudaFree(h->icp.n_ref); cudaFree(h->icp.n_mod);
  cudaFree(h->d_hyps); cudaFree(h->d_results); cudaFree(h->d_model_crops); cudaFree(h->d_icp_ticket);
  cudaFreeHost(h->h_model_crops); cudaFreeHost(h->h_hyps); cudaFreeHost(h->h_results);
  memset(&h->icp, 0, sizeof h->icp);
  h->d_hyps = nullptr; h->d_results = nullptr; h->d_model_crops = nullptr; h->d_icp_ticket = nullptr;
  h->h_model_crops = nullptr; h->h_hyps = nullptr; h->h_results = nullptr; h->h_crop_cap = 0;
  h->icp_hyp_cap = h->icp_pts_cap = 0;
}

static void free_templates(fl_handle* h) {
  cudaFree(h->d_hdr); cudaFree(h->d_feat); cudaFree(h->d_class_of); cudaFree(h->d_class_first); cudaFree(h->d_class_enabled); cudaFree(h->d_pfeat); cudaFree(h->d_rrec);
  cudaFree(h->d_tid_of); cudaFree(h->plan.gfeat); cudaFree(h->plan.gpre); cudaFree(h->plan.gmeta); cudaFree(h->plan.trace);
  h->plan.gfeat = nullptr; h->plan.gpre = nullptr; h->plan.gmeta = nullptr; h->plan.trace = nullptr;
  h->use_staged = false; h->staged_eligible = false;
  h->d_hdr = nullptr; h->d_feat = nullptr; h->d_class_of = nullptr; h->d_class_first = nullptr; h->d_class_enabled = nullptr; h->d_pfeat = nullptr; h->d_rrec = nullptr;
  h->d_tid_of = nullptr;
  h->n_templates = h->n_features = h->n_classes = 0; h->packed = false;
}

extern "C" int fl_destroy(fl_handle* h) {
  if (!h) return FL_OK;
  cudaSetDevice(h->p.device);
  cudaStreamSynchronize(h->stream);
  free_templates(h); icp_free(h);
  cudaFree(h->d_ref_depth); cudaFree(h->d_resident); cudaFree(h->d_nms); cudaFree(h->d_train); if (h->ev_block) cudaEventDestroy(h->ev_block);
  cudaFree(h->rz.xofs); cudaFree(h->rz.yofs); cudaFree(h->rz.ialpha); cudaFree(h->rz.ibeta); cudaFree(h->rz.alpha); cudaFree(h->rz.beta);
  cudaFree(h->d_src_bgr); cudaFree(h->d_src_depth);
  cudaFree(h->d_in_bgr); cudaFree(h->d_in_depth); cudaFree(h->d_geom);
  for (int l = 0; l < FL_MAX_LEVELS; ++l) {
    cudaFree(h->d_bgr[l]); cudaFree(h->d_lm[l]); if (l == 0) cudaFree(h->d_lm4);
    for (int m = 0; m < FL_MAX_MODALITIES; ++m) { cudaFree(h->d_q[l][m]); cudaFree(h->d_qm[l][m]); cudaFree(h->d_mask[l][m]); cudaFree(h->d_spread[l][m]); }
  }
  cudaFree(h->d_fe_counters); cudaFree(h->d_fe_trace); cudaFreeHost(h->h_fe_err);
  cudaFree(h->d_cand); cudaFree(h->d_count); cudaFree(h->d_keys); cudaFree(h->d_outblk);
  cudaFreeHost(h->h_bgr); cudaFreeHost(h->h_depth); cudaFreeHost(h->h_mask); cudaFreeHost(h->h_outblk); cudaFreeHost(h->h_class_enabled);
  for (int i = 0; i < 5; ++i) cudaEventDestroy(h->ev[i]);
  cudaStreamDestroy(h->stream);
  delete h;
  return FL_OK;
}

static cudaError_t wait_stream(fl_handle* h) {
  if (!h->blocking_wait) return cudaStreamSynchronize(h->stream);
  cudaError_t e = cudaEventRecord(h->ev_block, h->stream);
  return e == cudaSuccess ? cudaEventSynchronize(h->ev_block) : e;
}
extern "C" int fl_set_blocking_wait(fl_handle* h, int enable) {
  if (!h) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  if (enable && !h->ev_block) FL_CUDA(cudaEventCreateWithFlags(&h->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
  h->blocking_wait = enable != 0;
  return FL_OK;
}
extern "C" int fl_sync(fl_handle* h) { if (!h) return FL_ERR_ARG; FL_CUDA(cudaStreamSynchronize(h->stream)); return FL_OK; }
extern "C" void* fl_stream(fl_handle* h) { return h ? (void*)h->stream : nullptr; }
extern "C" int64_t fl_launch_count(fl_handle* h) { return h ? h->launches : 0; }
int fl_entries_per_template(const fl_handle* h) { return h ? h->p.n_levels * h->p.n_modalities : 0; }
extern "C" int fl_profile(fl_handle* h, int enable) { if (!h) return FL_ERR_ARG; h->profile = enable != 0; return FL_OK; }
extern "C" int fl_last_icp_ms(fl_handle* h, float* ms) { if (!h || !ms) return FL_ERR_ARG; *ms = h->icp_ms; return FL_OK; }
extern "C" int fl_debug_icp_trace(fl_handle* h, uint64_t* out, int32_t n_hyp) {
  if (!h || !out || n_hyp < 0) return FL_ERR_ARG;
  if (!h->d_icp_trace || n_hyp > h->icp_trace_n) { fl_set_error("no ICP phase clock: call fl_profile(h, 1) before the batch"); return FL_ERR_STATE; }
  FL_CUDA(cudaSetDevice(h->p.device));
  FL_CUDA(cudaMemcpy(out, h->d_icp_trace, (size_t)n_hyp * FL_ICP_TRACE_WORDS * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return FL_OK;
}
extern "C" int fl_last_stage_ms(fl_handle* h, float out4[4]) { if (!h || !out4) return FL_ERR_ARG; memcpy(out4, h->stage_ms, sizeof h->stage_ms); return FL_OK; }
// 1 = always use the baseline (L1/L2-fed) global similarity kernel; 0 = use the shared-memory-staged kernel when eligible
extern "C" int fl_debug_force_baseline(fl_handle* h, int enable) { if (!h) return FL_ERR_ARG; h->force_baseline = enable != 0; h->packed = false; return FL_OK; }
extern "C" int fl_debug_uses_staged(fl_handle* h) { return h ? (h->use_staged ? 1 : 0) : FL_ERR_ARG; }
// developer options of one handle (all off by default; each one is exercised by tests/test_gpu_match.py::test_debug_options)
extern "C" int fl_debug_option(fl_handle* h, int option, int value) {
  if (!h) return FL_ERR_ARG;
  switch (option) {
    case FL_OPT_FE_WAVES: h->opt_fe_waves = value != 0; return FL_OK;
    case FL_OPT_SPLIT_REFINE: h->opt_split_refine = value != 0; return FL_OK;
    case FL_OPT_TRACE: h->opt_trace = value != 0; h->packed = false; return FL_OK;     // (the staged plan allocates its timeline at planning time)
    case FL_OPT_FE_DEP_TIMEOUT_TEST: h->opt_dep_test = value != 0; return FL_OK;
    case FL_OPT_FE_FORCED_WAVES: return h->fe_force_waves ? 1 : 0;                      // query: has a dependency time-out switched this handle to wave launches?
  }
  return FL_ERR_ARG;
}
extern "C" int fl_debug_keep_spread(fl_handle* h, int enable) { if (!h) return FL_ERR_ARG; h->keep_spread = enable != 0; return FL_OK; }
extern "C" int fl_num_templates(fl_handle* h) { return h ? h->n_templates : FL_ERR_ARG; }

// ---------------------------------------------------------------------------------------------------
extern "C" int fl_upload_templates(fl_handle* h, int32_t n_templates, const fl_template_hdr_t* headers, const fl_feature_t* features,
                                   int32_t n_features, const int32_t* class_of, const float* pose13) {
  if (!h || n_templates < 0 || n_features < 0 || (n_templates > 0 && (!headers || !class_of)) || (n_features > 0 && !features)) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  const int L = h->p.n_levels, M = h->p.n_modalities;
  const size_t ne = (size_t)n_templates * L * M;
  int nc = 0;
  for (int t = 0; t < n_templates; ++t) { if (class_of[t] < 0) return FL_ERR_ARG; nc = std::max(nc, class_of[t] + 1); }
  if (nc > 4096) return FL_ERR_ARG;
  std::vector<int32_t> first(nc, -1);
  for (int t = 0; t < n_templates; ++t) {
    int c = class_of[t];
    if (first[c] < 0) first[c] = t; else if (class_of[t - 1] != c) { fl_set_error("templates of class %d are not contiguous", c); return FL_ERR_ARG; }
  }
  for (size_t e = 0; e < ne; ++e) {
    const fl_template_hdr_t& hd = headers[e];
    if (hd.feature_begin < 0 || hd.feature_count < 0 || (int64_t)hd.feature_begin + hd.feature_count > n_features) return FL_ERR_ARG;
    if (hd.feature_count > 63) return FL_ERR_FEATURES;                          // CV_Assert(features.size() <= 63)
    for (int k = 0; k < hd.feature_count; ++k) { int lab = features[hd.feature_begin + k].label; if (lab < 0 || lab > 7) return FL_ERR_ARG; }
  }
  // eligibility for the shared-memory-staged global similarity kernel: u8 totals cannot overflow, one template_positions
  // per template, and every coarsest-level feature inside its template's box (keeps flat over-reads within the staged halo)
  bool eligible = n_templates > 0;
  for (int t = 0; t < n_templates && eligible; ++t) {
    const fl_template_hdr_t* hd = headers + ((size_t)t * L + (L - 1)) * M;
    int nf = 0;
    for (int m = 0; m < M; ++m) {
      nf += hd[m].feature_count;
      if (hd[m].width != hd[0].width || hd[m].height != hd[0].height || hd[m].width < 1 || hd[m].height < 1) eligible = false;
      for (int k = 0; k < hd[m].feature_count; ++k) {
        const fl_feature_t& f = features[hd[m].feature_begin + k];
        if (f.x < 0 || f.y < 0 || f.x > hd[m].width || f.y > hd[m].height) eligible = false;
      }
    }
    if (nf > 63) eligible = false;
  }
  FL_CUDA(cudaStreamSynchronize(h->stream));
  free_templates(h);
  h->staged_eligible = eligible;
  h->coarse_wh.resize((size_t)n_templates * 2);
  for (int t = 0; t < n_templates; ++t) {
    const fl_template_hdr_t& hd = headers[((size_t)t * L + (L - 1)) * M];
    h->coarse_wh[2 * t] = hd.width; h->coarse_wh[2 * t + 1] = hd.height;
  }
  TRY(dalloc(&h->d_hdr, ne)); TRY(dalloc(&h->d_feat, (size_t)n_features)); TRY(dalloc(&h->d_class_of, (size_t)n_templates));
  TRY(dalloc(&h->d_class_first, (size_t)nc)); TRY(dalloc(&h->d_class_enabled, (size_t)std::max(nc, 1))); TRY(dalloc(&h->d_pfeat, (size_t)n_features));
  if (ne) FL_CUDA(cudaMemcpy(h->d_hdr, headers, ne * sizeof(fl_template_hdr_t), cudaMemcpyHostToDevice));
  if (n_features) FL_CUDA(cudaMemcpy(h->d_feat, features, (size_t)n_features * sizeof(fl_feature_t), cudaMemcpyHostToDevice));
  if (n_templates) FL_CUDA(cudaMemcpy(h->d_class_of, class_of, (size_t)n_templates * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (nc) FL_CUDA(cudaMemcpy(h->d_class_first, first.data(), (size_t)nc * sizeof(int32_t), cudaMemcpyHostToDevice));
  h->tid_of_h.resize((size_t)n_templates);
  for (int t = 0; t < n_templates; ++t) h->tid_of_h[t] = t - first[class_of[t]];                // matchClass numbering (:1458)
  TRY(dalloc(&h->d_tid_of, (size_t)n_templates));
  if (n_templates) FL_CUDA(cudaMemcpy(h->d_tid_of, h->tid_of_h.data(), (size_t)n_templates * sizeof(int32_t), cudaMemcpyHostToDevice));
  h->n_templates = n_templates; h->n_features = n_features; h->n_classes = nc;
  h->class_first_h = first; h->class_of_h.assign(class_of, class_of + n_templates);
  h->pose13_h.clear();
  if (pose13) h->pose13_h.assign(pose13, pose13 + (size_t)n_templates * 13);
  h->class_enabled_h.assign((size_t)std::max(nc, 1), 1); h->class_enabled_valid = false;
  h->packed = false;
  return FL_OK;
}

extern "C" int fl_set_template_ids(fl_handle* h, const int32_t* template_ids) {
  if (!h || !template_ids) return FL_ERR_ARG;
  if (h->n_templates == 0) return FL_OK;
  FL_CUDA(cudaSetDevice(h->p.device));
  FL_CUDA(cudaStreamSynchronize(h->stream));
  for (int t = 0; t < h->n_templates; ++t) if (template_ids[t] < 0) return FL_ERR_ARG;
  h->tid_of_h.assign(template_ids, template_ids + h->n_templates);
  FL_CUDA(cudaMemcpy(h->d_tid_of, template_ids, (size_t)h->n_templates * sizeof(int32_t), cudaMemcpyHostToDevice));
  return FL_OK;
}

extern "C" int fl_get_pose_info(fl_handle* h, int32_t class_idx, int32_t template_id, float out13[13]) {
  if (!h || !out13 || class_idx < 0 || class_idx >= h->n_classes || h->pose13_h.empty()) return FL_ERR_ARG;
  // the reference keeps ONE flat TemplatePoseInfo indexed by the per-class template_id (linemod.cpp:1624-1634), which is
  // wrong for more than one class (SURVEY A.6 iv); this ABI keys the pose by (class, template_id).
  int t = -1;
  for (int k = h->class_first_h[class_idx]; k >= 0 && k < h->n_templates && h->class_of_h[k] == class_idx; ++k)
    if (h->tid_of_h[k] == template_id) { t = k; break; }
  if (t < 0) return FL_ERR_ARG;
  memcpy(out13, &h->pose13_h[(size_t)t * 13], 13 * sizeof(float));
  return FL_OK;
}

static fl_tdb make_tdb(fl_handle* h) {
  fl_tdb db;
  db.n_templates = h->n_templates; db.L = h->p.n_levels; db.M = h->p.n_modalities; db.n_classes = h->n_classes;
  db.hdr = h->d_hdr; db.feat = h->d_feat; db.class_of = h->d_class_of; db.class_first = h->d_class_first;
  db.class_enabled = h->d_class_enabled; db.pfeat = h->d_pfeat; db.tid_of = h->d_tid_of;
  db.rrec = h->d_rrec; db.rrec_bytes = fl_rrec_bytes(h->p.n_modalities);
  return db;
}

// largest template_positions (linemod.cpp:1155) of the uploaded templates at the coarsest level of the current geometry
static int max_positions(const fl_handle* h) {
  const fl_level_geom& g = h->geom[h->p.n_levels - 1];
  int best = 0;
  for (size_t t = 0; t + 1 < h->coarse_wh.size(); t += 2) {
    const int wf = (h->coarse_wh[t] - 1) / g.T + 1, hf = (h->coarse_wh[t + 1] - 1) / g.T + 1;
    best = std::max(best, std::min((g.Hd - hf) * g.Wd + (g.Wd - wf) + 1, g.cells));
  }
  return best;
}

static int ensure_geometry(fl_handle* h, int W, int H) {
  const fl_params_t& p = h->p;
  if (W <= 0 || H <= 0 || W > p.max_width || H > p.max_height) { fl_set_error("frame %dx%d exceeds handle capacity %dx%d", W, H, p.max_width, p.max_height); return FL_ERR_SIZE; }
  for (int l = 0; l < p.n_levels; ++l) {                                        // CV_Asserts :981, :1062-1063
    int w = W >> l, hh = H >> l, T = p.T[l];
    if (w < 1 || hh < 1 || w % T || hh % T || (w * hh) % 16) return FL_ERR_GEOMETRY;
  }
  if (W != h->gW || H != h->gH) {
    for (int l = 0; l < p.n_levels; ++l) {
      size_t need = level_lm_bytes(p, l, W, H, &h->geom[l]);
      if (need > h->lm_bytes[l]) { fl_set_error("internal: LM capacity"); return FL_ERR_CAPACITY; }
      FL_CUDA(cudaMemsetAsync(h->d_lm[l], 0, h->lm_bytes[l], h->stream));       // pads must read as zero
      if (l == p.n_levels - 1) FL_CUDA(cudaMemsetAsync(h->d_lm4, 0, h->lm_bytes[l] / 2 + 256, h->stream));
    }
    FL_CUDA(cudaMemcpyAsync(h->d_geom, h->geom, sizeof(fl_level_geom) * p.n_levels, cudaMemcpyHostToDevice, h->stream));
    FL_CUDA(cudaStreamSynchronize(h->stream));                                  // h->geom is pageable
    h->gW = W; h->gH = H; h->packed = false;
  }
  if (!h->packed) {
    // the front end's tile-row counters are per job SLOT and advance by "tiles per row" every frame; a new geometry or template
    // set can change the job layout, so they restart from zero (stream-ordered with the launches that follow)
    FL_CUDA(cudaMemsetAsync(h->d_fe_counters + FL_FE_MAX_JOBS + 1, 0, (size_t)FL_FE_MAX_JOBS * FL_FE_MAX_TILE_ROWS * sizeof(unsigned), h->stream));
    memset(h->fe_row_base, 0, sizeof h->fe_row_base);
    if (h->n_templates == 0) h->packed = true;
  }
  if (!h->packed && h->n_templates > 0) {
    fl_launch_pack_features(make_tdb(h), h->d_geom, h->n_features, h->stream); ++h->launches;
    if (p.n_levels > 1) {
      if (!h->d_rrec) TRY(dalloc(&h->d_rrec, (size_t)h->n_templates * (p.n_levels - 1) * fl_rrec_bytes(p.n_modalities)));
      fl_launch_pack_refine_records(make_tdb(h), h->d_geom, h->stream); ++h->launches;
    }
    h->use_staged = false;
    fl_staged_plan plan;
    if (h->staged_eligible && !h->force_baseline && h->n_templates >= 256 &&
        fl_plan_staged(h->geom[p.n_levels - 1], p.n_modalities, h->n_templates, max_positions(h), h->n_sm, &plan)) {
      cudaFree(h->plan.gfeat); cudaFree(h->plan.gpre); cudaFree(h->plan.gmeta); cudaFree(h->plan.trace);
      h->plan = plan;
      TRY(dalloc(&h->plan.gfeat, (size_t)h->n_templates * 64));
      TRY(dalloc(&h->plan.gpre, (size_t)h->n_templates * plan.pre_stride));
      TRY(dalloc(&h->plan.gmeta, (size_t)h->n_templates));
      if (h->opt_trace) {                                                        // developer timeline of the staged kernel (FL_DBG_STAGED_TRACE)
        TRY(dalloc(&h->plan.trace, (size_t)plan.n_cta * 136 + 8));                 // + stamps of a 1-thread kernel before / after the launch
        FL_CUDA(cudaMemsetAsync(h->plan.trace, 0, ((size_t)plan.n_cta * 136 + 8) * sizeof(unsigned long long), h->stream));
      }
      fl_launch_pack_staged(make_tdb(h), h->geom[p.n_levels - 1], h->plan, h->stream); ++h->launches;
      h->use_staged = true;
    }
    h->packed = true;
  }
  return FL_OK;
}

// the 4-bit linear memories of (level, modality): only the coarsest level has them, and only grids of even width (two cells per byte)
static uint8_t* lm4_of(fl_handle* h, int l, int m) {
  return (l == h->p.n_levels - 1 && (h->geom[l].Wd & 1) == 0) ? h->d_lm4 + (((size_t)m * h->geom[l].mod_stride) >> 1) : nullptr;
}
static void fe_wave_init(fl_fe_wave* w) { memset(w, 0, sizeof *w); }   // an empty launch description (no jobs, no in-grid dependencies)

// front end + matchClass on the handle's templates; candidates land in (cand, count)
// defer_refine: leave the candidates at the coarsest level; the caller refines them in the same launch that sorts (k_refine_sort)
static int run_match_stages(fl_handle* h, const uint8_t* d_bgr, const uint16_t* d_depth, int W, int H, const void* const* d_masks,
                            float threshold, const int32_t* class_filter, int n_filter, fl_match_t* cand, int cap, int* d_count,
                            bool defer_refine = false) {
  const fl_params_t& p = h->p;
  FL_CUDA(cudaSetDevice(p.device));
  TRY(ensure_geometry(h, W, H));
  cudaStream_t s = h->stream;
  for (int m = 0; m < p.n_modalities; ++m) {
    if (p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT && !d_bgr) return FL_ERR_SIZE;
    if (p.modality_kind[m] == FL_MODALITY_DEPTH_NORMAL && !d_depth) return FL_ERR_SIZE;
  }
  // class filter (Detector::match :1418-1434): unknown ids are ignored, empty = all.  Uploaded only when it changes.
  if (h->n_classes > 0) {
    bool changed = false;
    for (int c = 0; c < h->n_classes; ++c) {
      uint8_t on = (n_filter > 0 && class_filter) ? 0 : 1;
      if (n_filter > 0 && class_filter) for (int k = 0; k < n_filter; ++k) if (class_filter[k] == c) on = 1;
      if (h->class_enabled_h[c] != on) { h->class_enabled_h[c] = on; changed = true; }
    }
    if (changed || !h->class_enabled_valid) {
      FL_CUDA(cudaStreamSynchronize(s));                                          // the pinned staging copy may still be in flight
      memcpy(h->h_class_enabled, h->class_enabled_h.data(), (size_t)h->n_classes);
      FL_CUDA(cudaMemcpyAsync(h->d_class_enabled, h->h_class_enabled, (size_t)h->n_classes, cudaMemcpyHostToDevice, s));
      h->class_enabled_valid = true;
    }
  }
  if (h->profile) cudaEventRecord(h->ev[0], s);
  const float thr_sq = p.weak_threshold * p.weak_threshold;
  bool any_mask = false;
  for (int m = 0; m < p.n_modalities; ++m) { h->used_mask[m] = d_masks && d_masks[m]; any_mask |= h->used_mask[m]; }
  int first_color = -1;
  for (int m = 0; m < p.n_modalities; ++m) if (p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT && first_color < 0) first_color = m;
  const bool zero_in_wave = !any_mask;                       // wave 0 resets the candidate counter: no memset node between the kernels
  if (!any_mask) {
    // wave schedule (frontend.cu): L + 1 launches for the whole front end; wave l produces the labels of level l and the
    // linear memories of level l - 1.  Jobs are added longest first: the kernel ends with its slowest CTA, and CTAs are
    // scheduled in job order.  (Measured alternative: slicing the level-0 colour quantisation over the waves and moving
    // every spread to the last wave - 16.6 + 10.8 + 9.2 us instead of 17.7 + 9.4 + 6.0 us, i.e. no gain: each wave has a
    // latency floor of one colour-tile CTA, about 8 us.)
    const int L = p.n_levels;
    fl_fe_wave w;
    auto wave_begin = [&](bool zero) { w.n_jobs = 0; w.n_ctas = 0; w.smem = 0; w.n_pyr = 0; w.zero_me = zero ? d_count : nullptr; w.counters = nullptr; w.dep_error = nullptr; w.dep_error_host = nullptr; w.trace = nullptr; };
    bool launch_ok = true;
    auto wave_flush = [&]() { if (w.n_jobs > 0) { launch_ok &= fl_launch_fe_wave(w, s) == cudaSuccess; ++h->launches; } w.n_jobs = 0; w.n_ctas = 0; w.smem = 0; w.n_pyr = 0; w.zero_me = nullptr; };
    // word-parallel quantisers (frontend_v2.cuh).  The depth job also writes the NN-downsampled label pyramid when every level
    // halves exactly (then dst_l(y,x) = src(2^l y, 2^l x)).
    bool depth_pyr_fused[FL_MAX_MODALITIES] = {false, false, false, false};
    bool halves = true;
    for (int l = 0; l + 1 < L; ++l) halves &= (h->geom[l].W % 2 == 0) && (h->geom[l].H % 2 == 0);
    auto wave_room = [&]() { if (w.n_jobs == FL_FE_MAX_JOBS) wave_flush(); };     // jobs of one wave are independent: splitting is always safe
    // ---- single-launch front end: every job of the frame in ONE grid, producers before consumers, dependencies resolved by
    // per-job CTA counters inside the kernel (k_front_end_wave).  Removes L launches + their dependency gaps and lets the
    // level-0 work fill the machine while the short critical chain pyrDown -> colour L(top) -> spread L(top) runs.
    const bool fe_waves = h->opt_fe_waves;                                       // FL_OPT_FE_WAVES: one launch per wave (also the fallback after a dependency time-out)
    int n_color = 0, n_depth = 0;
    for (int m = 0; m < p.n_modalities; ++m) { n_color += p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT; n_depth += p.modality_kind[m] == FL_MODALITY_DEPTH_NORMAL; }
    const int n_jobs_single = (n_color ? L - 1 : 0) + n_color * L + n_depth + p.n_modalities * L;
    const bool single = !fe_waves && !h->fe_force_waves && n_jobs_single <= FL_FE_MAX_JOBS && n_depth <= FL_FE_MAX_DPYR && (L == 1 || (halves && L - 1 <= FL_FE_MAX_PYR));
    if (single) {
      wave_begin(zero_in_wave);
      w.counters = h->d_fe_counters; w.dep_error = reinterpret_cast<int*>(h->d_fe_counters + FL_FE_MAX_JOBS); w.dep_error_host = h->d_fe_err_host;
      int slot_pyr[FL_MAX_LEVELS], slot_color[FL_MAX_LEVELS][FL_MAX_MODALITIES], slot_depth[FL_MAX_MODALITIES];
      auto last_job = [&]() -> fl_fe_job& { return w.job[w.n_jobs - 1]; };
      auto job_ctas = [&](int idx) { return (idx + 1 < w.n_jobs ? w.job[idx + 1].cta_begin : w.n_ctas) - w.job[idx].cta_begin; };
      auto produce = [&]() { last_job().signal_slot = w.n_jobs - 1; return w.n_jobs - 1; };
      // tile jobs (32 x 16 tiles) also count per tile row; returns false when the image has more tile rows than counters
      auto produce_rows = [&](int slot) {
        fl_fe_job& jb = w.job[slot];
        const int rows = (jb.H + 15) / 16;
        if (rows > FL_FE_MAX_TILE_ROWS) return;                 // (then the consumers wait for the whole job)
        jb.row_base = FL_FE_MAX_JOBS + 1 + slot * FL_FE_MAX_TILE_ROWS;
      };
      auto consume_rows = [&](int slot, int shift) {      // spread job: wait per tile row of producer `slot` (level shift for the depth pyramid)
        const fl_fe_job& pj = w.job[slot];
        if (pj.row_base < 0) return;
        fl_fe_job& jb = last_job();
        jb.wait_row_base = pj.row_base; jb.wait_row_shift = shift; jb.wait_row_count = (pj.H + 15) / 16;
        jb.wait_row_target = h->fe_row_base[slot] + (unsigned)abs(pj.gx);
      };
      // FL_OPT_FE_DEP_TIMEOUT_TEST: the next frame waits for one CTA more than its producer has, i.e. the wait times out (test of the fallback)
      const unsigned dep_test = h->opt_dep_test ? 1u : 0u;
      h->opt_dep_test = false;
      auto consume = [&](int slot) { last_job().wait_slot = slot; last_job().wait_target = h->fe_counter_base[slot] + (unsigned)job_ctas(slot) + dep_test; };
      // 0. the staged similarity kernel's per-template feature lists -> L2 (they would otherwise cost its prologue DRAM round trips)
      if (h->use_staged && h->n_templates > 0 && n_jobs_single + 3 <= FL_FE_MAX_JOBS) {
        fl_fe_add_prefetch(&w, h->plan.gfeat, (size_t)h->n_templates * 64 * sizeof(uint32_t));
        fl_fe_add_prefetch(&w, h->plan.gpre, (size_t)h->n_templates * h->plan.pre_stride);
        fl_fe_add_prefetch(&w, h->plan.gmeta, (size_t)h->n_templates * sizeof(int4));
        if (n_jobs_single + 6 <= FL_FE_MAX_JOBS) {                               // the prologue's dependent lookups: template -> class -> enabled
          fl_fe_add_prefetch(&w, h->d_class_of, (size_t)h->n_templates * sizeof(int32_t));
          fl_fe_add_prefetch(&w, h->d_class_enabled, (size_t)std::max(h->n_classes, 1));
          fl_fe_add_prefetch(&w, h->d_tid_of, (size_t)h->n_templates * sizeof(int32_t));
        }
      }
      // 1. colour pyramid chain (short jobs, on the critical path of the coarsest level)
      if (first_color >= 0)
        for (int l = 0; l + 1 < L; ++l) {
          fl_fe_add_pyrdown(&w, l == 0 ? d_bgr : h->d_bgr[l], h->geom[l].W, h->geom[l].H, h->d_bgr[l + 1], l == 0);
          slot_pyr[l + 1] = produce();
          if (l > 0) consume(slot_pyr[l]);
        }
      // 2. depth labels of every level (one job per depth modality)
      for (int m = 0; m < p.n_modalities; ++m)
        if (p.modality_kind[m] == FL_MODALITY_DEPTH_NORMAL) {
          fl_depth_pyr pyr; memset(&pyr, 0, sizeof pyr);
          pyr.n = L - 1;
          for (int k = 1; k < L; ++k) { pyr.W[k - 1] = h->geom[k].W; pyr.H[k - 1] = h->geom[k].H; pyr.dst[k - 1] = h->d_q[k][m]; }
          fl_fe_add_depth_v2(&w, d_depth, h->geom[0].W, h->geom[0].H, p.distance_threshold, p.difference_threshold, h->d_q[0][m], pyr.n > 0 ? &pyr : nullptr);
          slot_depth[m] = produce();
          produce_rows(slot_depth[m]);
        }
      // 3. colour labels, coarsest level first
      for (int l = L - 1; l >= 0; --l)
        for (int m = 0; m < p.n_modalities; ++m)
          if (p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT) {
            fl_fe_add_color_v2(&w, l == 0 ? d_bgr : h->d_bgr[l], h->geom[l].W, h->geom[l].H, thr_sq, h->d_q[l][m]);
            slot_color[l][m] = produce();
            produce_rows(slot_color[l][m]);
            if (l > 0) consume(slot_pyr[l]);
          }
      // 4. spread + response maps + linear memories, coarsest level first (the similarity kernel needs only that one)
      for (int l = L - 1; l >= 0; --l)
        for (int m = 0; m < p.n_modalities; ++m) {
          uint8_t* spread = nullptr;
          if (h->keep_spread) { if (!h->d_spread[l][m]) TRY(dalloc(&h->d_spread[l][m], (size_t)(p.max_width >> l) * (p.max_height >> l))); spread = h->d_spread[l][m]; }
          fl_fe_add_spread(&w, h->d_q[l][m], h->geom[l], h->d_lm[l] + (size_t)m * h->geom[l].mod_stride, spread, lm4_of(h, l, m));
          const bool is_color = p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT;
          consume(is_color ? slot_color[l][m] : slot_depth[m]);
          consume_rows(is_color ? slot_color[l][m] : slot_depth[m], is_color ? 0 : l);
        }
      if (h->opt_trace) {                                                        // developer timeline: per job first start / last end
        if (!h->d_fe_trace) TRY(dalloc(&h->d_fe_trace, 2 * FL_FE_MAX_JOBS + 2 * FL_FE_MAX_JOBS));
        unsigned long long init[2 * FL_FE_MAX_JOBS];
        for (int i = 0; i < FL_FE_MAX_JOBS; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0; }
        FL_CUDA(cudaMemcpyAsync(h->d_fe_trace, init, sizeof init, cudaMemcpyHostToDevice, s));
        FL_CUDA(cudaStreamSynchronize(s));
        w.trace = h->d_fe_trace;
        h->fe_trace_jobs = w.n_jobs;
        for (int i = 0; i < w.n_jobs; ++i) { h->fe_trace_kind[i] = w.job[i].kind; h->fe_trace_ctas[i] = job_ctas(i); }
      }
      // the monotonic counter bases advance only when the launch really went out (a failed launch bumps no counter)
      unsigned add_ctr[FL_FE_MAX_JOBS] = {0}, add_row[FL_FE_MAX_JOBS] = {0};
      for (int i = 0; i < w.n_jobs; ++i) { if (w.job[i].signal_slot >= 0) add_ctr[i] = (unsigned)job_ctas(i); if (w.job[i].row_base >= 0) add_row[i] = (unsigned)abs(w.job[i].gx); }
      wave_flush();
      if (!launch_ok) { fl_set_error("front-end launch failed: %s", cudaGetErrorString(cudaGetLastError())); return FL_ERR_CUDA; }
      for (int i = 0; i < FL_FE_MAX_JOBS; ++i) { h->fe_counter_base[i] += add_ctr[i]; h->fe_row_base[i] += add_row[i]; }
    } else
    for (int wv = 0; wv <= L; ++wv) {
      wave_begin(wv == 0 && zero_in_wave);                                        // the candidate counter is reset by the first wave
      const int l = wv;                                                             // level whose labels this wave produces
      if (l < L) {
        const fl_level_geom& g = h->geom[l];
        for (int m = 0; m < p.n_modalities; ++m)
          if (p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT) {
            wave_room();
            fl_fe_add_color_v2(&w, l == 0 ? d_bgr : h->d_bgr[l], g.W, g.H, thr_sq, h->d_q[l][m]);
          }
        for (int m = 0; m < p.n_modalities; ++m)
          if (p.modality_kind[m] == FL_MODALITY_DEPTH_NORMAL && l == 0) {
            wave_room();
            fl_depth_pyr pyr; memset(&pyr, 0, sizeof pyr);
            if (halves && L > 1 && L - 1 <= FL_FE_MAX_PYR) {
              pyr.n = L - 1;
              for (int k = 1; k < L; ++k) { pyr.W[k - 1] = h->geom[k].W; pyr.H[k - 1] = h->geom[k].H; pyr.dst[k - 1] = h->d_q[k][m]; }
            }
            if (pyr.n > 0 && fl_fe_add_depth_v2(&w, d_depth, g.W, g.H, p.distance_threshold, p.difference_threshold, h->d_q[0][m], &pyr)) depth_pyr_fused[m] = true;
            else fl_fe_add_depth_v2(&w, d_depth, g.W, g.H, p.distance_threshold, p.difference_threshold, h->d_q[0][m], nullptr);
          }
      }
      if (l >= 1) {
        const fl_level_geom& g = h->geom[l - 1];
        for (int m = 0; m < p.n_modalities; ++m) {
          uint8_t* spread = nullptr;
          if (h->keep_spread) { if (!h->d_spread[l - 1][m]) TRY(dalloc(&h->d_spread[l - 1][m], (size_t)(p.max_width >> (l - 1)) * (p.max_height >> (l - 1)))); spread = h->d_spread[l - 1][m]; }
          wave_room();
          fl_fe_add_spread(&w, h->d_q[l - 1][m], g, h->d_lm[l - 1] + (size_t)m * g.mod_stride, spread, lm4_of(h, l - 1, m));
        }
      }
      if (l < L) {
        const fl_level_geom& g = h->geom[l];
        if (first_color >= 0 && l + 1 < L) { wave_room(); fl_fe_add_pyrdown(&w, l == 0 ? d_bgr : h->d_bgr[l], g.W, g.H, h->d_bgr[l + 1], true); }   // shared by all colour modalities
        for (int m = 0; m < p.n_modalities; ++m)
          if (p.modality_kind[m] == FL_MODALITY_DEPTH_NORMAL && l > 0 && !depth_pyr_fused[m]) { wave_room(); fl_fe_add_resize(&w, h->d_q[l - 1][m], h->geom[l - 1].W, h->geom[l - 1].H, h->d_q[l][m]); }
      }
      wave_flush();
    }
  } else {
    // masked path (rare): one kernel per stage, masks NN-downsampled per level and applied before spreading
    bool color_pyr_done = false;
    for (int l = 0; l < p.n_levels; ++l) {
      const fl_level_geom& g = h->geom[l];
      const size_t npx = (size_t)g.W * g.H;
      for (int m = 0; m < p.n_modalities; ++m) {
        const bool has_mask = h->used_mask[m];
        if (p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT) {
          const uint8_t* src = d_bgr;
          if (l > 0) {
            if (!color_pyr_done) {
              const uint8_t* prev = l == 1 ? d_bgr : h->d_bgr[l - 1];
              fl_launch_pyrdown_bgr(prev, h->geom[l - 1].W, h->geom[l - 1].H, h->d_bgr[l], s); ++h->launches;
              color_pyr_done = true;
            }
            src = h->d_bgr[l];
          }
          { fl_fe_wave w1; fe_wave_init(&w1); fl_fe_add_color_v2(&w1, src, g.W, g.H, thr_sq, h->d_q[l][m]); fl_launch_fe_wave(w1, s); ++h->launches; }
        } else {
          if (l == 0) { fl_fe_wave w1; fe_wave_init(&w1); fl_fe_add_depth_v2(&w1, d_depth, g.W, g.H, p.distance_threshold, p.difference_threshold, h->d_q[0][m], nullptr); fl_launch_fe_wave(w1, s); ++h->launches; }
          else { fl_launch_resize_nn_half(h->d_q[l - 1][m], h->geom[l - 1].W, h->geom[l - 1].H, h->d_q[l][m], s); ++h->launches; }
        }
        const uint8_t* qsrc = h->d_q[l][m];
        if (has_mask) {
          if (!h->d_mask[l][m]) { size_t n = (size_t)(p.max_width >> l) * (p.max_height >> l); TRY(dalloc(&h->d_mask[l][m], n)); }
          if (!h->d_qm[l][m]) { size_t n = (size_t)(p.max_width >> l) * (p.max_height >> l); TRY(dalloc(&h->d_qm[l][m], n)); }
          const uint8_t* mk = (const uint8_t*)d_masks[m];
          if (l > 0) { fl_launch_resize_nn_half(l == 1 ? (const uint8_t*)d_masks[m] : h->d_mask[l - 1][m], h->geom[l - 1].W, h->geom[l - 1].H, h->d_mask[l][m], s); ++h->launches; mk = h->d_mask[l][m]; }
          fl_launch_apply_mask(h->d_q[l][m], mk, (int)npx, h->d_qm[l][m], s); ++h->launches;
          qsrc = h->d_qm[l][m];
        }
        uint8_t* spread = nullptr;
        if (h->keep_spread) { if (!h->d_spread[l][m]) TRY(dalloc(&h->d_spread[l][m], (size_t)(p.max_width >> l) * (p.max_height >> l))); spread = h->d_spread[l][m]; }
        fl_launch_spread_lm(qsrc, g, h->d_lm[l] + (size_t)m * g.mod_stride, spread, lm4_of(h, l, m), s); ++h->launches;
      }
      color_pyr_done = false;
    }
  }
  if (h->profile) cudaEventRecord(h->ev[1], s);
  if (!zero_in_wave) FL_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
  if (h->n_templates > 0) {
    fl_tdb db = make_tdb(h);
    const int lowest = p.n_levels - 1;
    if (h->use_staged) {
      if (fl_launch_similarity_staged(db, h->geom[lowest], h->d_lm4, threshold, cand, cap, d_count, h->plan, s) != 0) {
        fl_set_error("staged similarity kernel could not be configured"); return FL_ERR_CUDA;
      }
    } else {
      fl_launch_similarity_global(db, h->geom[lowest], h->d_lm[lowest], threshold, cand, cap, d_count, s);
    }
    ++h->launches;
    if (h->profile) cudaEventRecord(h->ev[2], s);
    if (!defer_refine)
      for (int l = p.n_levels - 2; l >= 0; --l) { fl_launch_refine_level(db, h->geom[l], l, h->d_lm[l], threshold, cand, cap, d_count, s); ++h->launches; }
  } else if (h->profile) cudaEventRecord(h->ev[2], s);
  if (h->profile) cudaEventRecord(h->ev[3], s);
  FL_CUDA(cudaGetLastError());
  return FL_OK;
}

// sort + unique of candidate lists into (d_out, d_out_count); handles the rare > 8,192-record path; leaves
// h_small = {count, n_live, flag, raw list counts...} and, for the handle's own output block, the first matches in h_first
struct fl_refine_req { float threshold; fl_match_t* cand; int cap; const int* d_count; };   // refine these candidates in the sort launch (k_refine_sort)
// first half of sort + unique (optionally with the refinement in the same launch): enqueue only
static int sort_launch(fl_handle* h, fl_lists L, fl_match_t* d_out, int out_cap, int* d_out_count, bool fetch_first, const fl_xchg* xchg = nullptr,
                       const fl_refine_req* refine = nullptr) {
  cudaStream_t s = h->stream;
  const int n_lists = L.n_lists, list_cap = L.list_cap;
  if ((int64_t)n_lists * list_cap > h->key_cap) { fl_set_error("sort capacity %d < %lld", h->key_cap, (long long)n_lists * list_cap); return FL_ERR_CAPACITY; }
  int* d_hdr = reinterpret_cast<int*>(h->d_outblk);
  const bool own = fetch_first && d_out == h->d_out;
  // the kernel itself posts the summary and (own output block) the first matches into the mapped pinned block h_outblk
  fl_xchg X;
  memset(&X, 0, sizeof X);
  if (xchg) X = *xchg;
  if (refine) {
    fl_refine_args ra;
    memset(&ra, 0, sizeof ra);
    ra.n_levels = h->p.n_levels;
    for (int l = 0; l < h->p.n_levels; ++l) { ra.g[l] = h->geom[l]; ra.lm[l] = h->d_lm[l]; }
    const int nl = fl_launch_refine_sort(make_tdb(h), ra, refine->threshold, refine->cand, refine->cap, refine->d_count, h->d_count + 1, h->n_sm, L, X, h->key_cap,
                                         d_out, out_cap, d_out_count, d_hdr, h->h_small, own ? h->h_first : nullptr, std::min(FETCH_FIRST, out_cap), s);
    if (nl < 0) { fl_set_error("refinement + sort launch failed: %s", cudaGetErrorString(cudaGetLastError())); return FL_ERR_CUDA; }
    h->launches += nl;
    h->pend_small_fused = true;
  } else {
    const int nl = fl_launch_sort_unique(L, X, h->key_cap, d_out, out_cap, d_out_count, d_hdr, h->h_small, own ? h->h_first : nullptr,
                                         std::min(FETCH_FIRST, out_cap), s);
    if (nl < 0) { fl_set_error("sort launch failed: %s", cudaGetErrorString(cudaGetLastError())); return FL_ERR_CUDA; }
    h->launches += nl;
  }
  if (h->profile) cudaEventRecord(h->ev[4], s);
  // what the second half (sort_finish) needs: it runs after the host has waited for the stream
  h->pend_lists = L; h->pend_out = d_out; h->pend_out_cap = out_cap; h->pend_out_count = d_out_count; h->pend_own = own; h->pend_sort = true;
  return FL_OK;
}

// second half of sort + unique: wait for the stream, read the posted summary, run the multi-kernel path if the one-CTA sort
// reported more records than it holds
static int sort_finish(fl_handle* h) {
  cudaStream_t s = h->stream;
  if (!h->pend_sort) return FL_ERR_STATE;
  h->pend_sort = false;
  const fl_lists L = h->pend_lists;
  fl_match_t* d_out = h->pend_out; const int out_cap = h->pend_out_cap; int* d_out_count = h->pend_out_count; const bool own = h->pend_own;
  const int n_lists = L.n_lists, list_cap = L.list_cap;
  FL_CUDA(wait_stream(h));
  h->overflow = false;
  int n_upper = 0;
  for (int i = 0; i < std::min(n_lists, 11); ++i) { if (h->h_small[3 + i] > list_cap) h->overflow = true; n_upper += std::min(std::max(h->h_small[3 + i], 0), list_cap); }
  if (n_lists > 11) n_upper = n_lists * list_cap;
  if (h->h_small[2] && h->pend_small_fused && n_upper <= 8192) {
    // the fused refinement + sort launch sorts up to 1,024 records in its last CTA; this frame has more (already refined in
    // place): run the stand-alone one-CTA sort (8,192 keys) on them.  What the first launch reported about the exchange
    // (a peer that never arrived) survives the second launch, which knows nothing about it.
    const int keep15 = h->h_small[15];
    const bool keep_overflow = h->overflow;
    h->pend_small_fused = false;
    TRY(sort_launch(h, L, d_out, out_cap, d_out_count, own));
    const int rc = sort_finish(h);
    if (keep15) h->h_small[15] = keep15;
    h->overflow |= keep_overflow;
    return rc;
  }
  h->pend_small_fused = false;
  if (h->h_small[2]) {                                                          // more records than the one-CTA sort holds: multi-kernel sort
    int rc = fl_launch_sort_unique_big(L, h->d_keys, h->key_cap, n_upper, d_out, out_cap, d_out_count, s);
    if (rc < 0) return FL_ERR_CAPACITY;
    h->launches += rc;
    FL_CUDA(cudaMemcpyAsync(h->h_small, d_out_count, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (own) FL_CUDA(cudaMemcpyAsync(h->h_first, d_out, sizeof(fl_match_t) * (size_t)std::min(FETCH_FIRST, out_cap), cudaMemcpyDeviceToHost, s));
    if (h->profile) cudaEventRecord(h->ev[4], s);
    FL_CUDA(wait_stream(h));
  }
  if (h->profile) for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&h->stage_ms[i], h->ev[i], h->ev[i + 1]);
  FL_CUDA(cudaGetLastError());
  return FL_OK;
}

static int run_sort_unique(fl_handle* h, fl_lists L, fl_match_t* d_out, int out_cap, int* d_out_count, bool fetch_first, const fl_xchg* xchg = nullptr,
                           const fl_refine_req* refine = nullptr) {
  TRY(sort_launch(h, L, d_out, out_cap, d_out_count, fetch_first, xchg, refine));
  return sort_finish(h);
}

extern "C" int fl_match_device_async(fl_handle* h, const void* d_bgr, const void* d_depth, int32_t W, int32_t H, const void* const* d_masks,
                                     float threshold, const int32_t* class_filter, int32_t n_filter) {
  if (!h) return FL_ERR_ARG;
  if (h->pend_sort) { fl_set_error("fl_match_wait has not been called for the previous frame"); return FL_ERR_STATE; }
  h->have_result = false; h->pend_match = false;
  // refinement and sort + unique share one launch (k_refine_sort: 256-thread CTAs, the last one to finish sorts up to 1,024
  // records); FL_OPT_SPLIT_REFINE keeps them as separate launches (what fl_match_shard_device needs anyway: its candidates must
  // be refined before an NCCL gather).  A variant with 1,024-thread CTAs and the 8,192-key sort inside was measured slower (those
  // CTAs cost more to launch than the launch they save) and removed.
  const bool defer = !h->opt_split_refine && h->n_templates > 0 && h->p.n_levels > 1;
  TRY(run_match_stages(h, (const uint8_t*)d_bgr, (const uint16_t*)d_depth, W, H, d_masks, threshold, class_filter, n_filter, h->d_cand,
                       h->p.max_candidates, h->d_count, defer));
  const fl_lists lists = {h->d_cand, 1, h->p.max_candidates, h->p.max_candidates, h->d_count, 1};
  const fl_refine_req req = {threshold, h->d_cand, h->p.max_candidates, h->d_count};
  TRY(sort_launch(h, lists, h->d_out, h->p.max_candidates, h->d_out_count, true, nullptr, defer ? &req : nullptr));
  // kept for fl_match_wait: a front-end dependency time-out re-runs the frame
  h->pend_match = true; h->pend_bgr = d_bgr; h->pend_depth = d_depth; h->pend_W = W; h->pend_H = H; h->pend_threshold = threshold;
  h->pend_masks_valid = d_masks != nullptr;
  for (int m = 0; m < FL_MAX_MODALITIES; ++m) h->pend_masks[m] = d_masks ? d_masks[m] : nullptr;
  h->pend_filter.assign(class_filter && n_filter > 0 ? class_filter : nullptr, class_filter && n_filter > 0 ? class_filter + n_filter : nullptr);
  return FL_OK;
}

extern "C" int fl_match_wait(fl_handle* h) {
  if (!h) return FL_ERR_ARG;
  if (!h->pend_sort) return FL_ERR_STATE;
  const bool was_match = h->pend_match;
  h->pend_match = false;
  TRY(sort_finish(h));
  if (*reinterpret_cast<volatile int*>(h->h_fe_err)) {
    // an in-grid dependency of the single-launch front end timed out (the producer CTAs were not dispatched before their consumers:
    // see k_front_end_wave): the frame's result is not trustworthy.  From now on this handle launches one kernel per wave; a frame
    // submitted through fl_match / fl_match_device* is re-run right away, the shard entry points report the state.
    const int job = *h->h_fe_err - 1;
    *h->h_fe_err = 0;
    FL_CUDA(cudaMemsetAsync(h->d_fe_counters + FL_FE_MAX_JOBS, 0, sizeof(unsigned), h->stream));
    h->fe_force_waves = true;
    if (!was_match) { fl_set_error("front end: in-grid dependency of job %d timed out; the handle now uses per-wave launches - resubmit the frame", job); return FL_ERR_STATE; }
    const fl_lists lists = {h->d_cand, 1, h->p.max_candidates, h->p.max_candidates, h->d_count, 1};
    int rc = run_match_stages(h, (const uint8_t*)h->pend_bgr, (const uint16_t*)h->pend_depth, h->pend_W, h->pend_H, h->pend_masks_valid ? h->pend_masks : nullptr,
                              h->pend_threshold, h->pend_filter.empty() ? nullptr : h->pend_filter.data(), (int)h->pend_filter.size(), h->d_cand,
                              h->p.max_candidates, h->d_count);
    if (rc == FL_OK) rc = run_sort_unique(h, lists, h->d_out, h->p.max_candidates, h->d_out_count, true);
    if (rc != FL_OK) return rc;
  }
  if (h->h_small[15]) { fl_set_error("peer exchange timed out waiting for rank %d", h->h_small[15] - 1); return FL_ERR_STATE; }
  h->have_result = true;
  return FL_OK;
}

extern "C" int fl_match_device(fl_handle* h, const void* d_bgr, const void* d_depth, int32_t W, int32_t H, const void* const* d_masks,
                               float threshold, const int32_t* class_filter, int32_t n_filter) {
  int rc = fl_match_device_async(h, d_bgr, d_depth, W, H, d_masks, threshold, class_filter, n_filter);
  if (rc != FL_OK) return rc;
  return fl_match_wait(h);
}

extern "C" int fl_match_fetch(fl_handle* h, fl_match_t* out, int32_t capacity, int32_t* count) {
  if (!h || !count || capacity < 0 || (capacity > 0 && !out)) return FL_ERR_ARG;
  if (!h->have_result) return FL_ERR_STATE;
  const int n = h->h_small[0];
  *count = n;
  const int ncopy = std::min(std::min(n, (int)capacity), h->p.max_candidates);
  const int nfirst = std::min(ncopy, FETCH_FIRST);
  if (nfirst > 0) memcpy(out, h->h_first, sizeof(fl_match_t) * nfirst);
  if (ncopy > nfirst) {
    FL_CUDA(cudaMemcpyAsync(out + nfirst, h->d_out + nfirst, sizeof(fl_match_t) * (size_t)(ncopy - nfirst), cudaMemcpyDeviceToHost, h->stream));
    FL_CUDA(wait_stream(h));
  }
  // overflow: matchClass produced more candidates than the device buffer holds (the list is then incomplete)
  return (n > capacity || n > h->p.max_candidates || h->overflow) ? FL_ERR_CAPACITY : FL_OK;
}

// Frame upload of the host-buffer entry points.  A caller buffer that is already page-locked (cudaHostAlloc / cudaHostRegister /
// torch pin_memory) is DMA'd directly, strided rows included, and has to stay valid until the frame has been waited for; pageable
// memory goes through the handle's pinned staging buffer first (copied before this returns).
static int stage_frame(fl_handle* h, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, int W, int H, const uint8_t* const* masks,
                       const uint8_t** d_bgr_out, const uint16_t** d_depth_out, const void** d_masks, bool* any_mask) {
  const fl_params_t& p = h->p;
  cudaStream_t s = h->stream;
  *d_bgr_out = nullptr; *d_depth_out = nullptr; *any_mask = false;
  auto is_pinned = [](const void* q) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  if (depth) {                                                                     // depth first: it is the smaller image
    if (depth_stride < (size_t)W * 2) return FL_ERR_SIZE;
    if (is_pinned(depth)) {
      // dense rows go as ONE linear DMA (a 2-D copy of the same bytes is split into row descriptors and runs at about half the rate)
      if (depth_stride == (size_t)W * 2) FL_CUDA(cudaMemcpyAsync(h->d_in_depth, depth, (size_t)W * H * 2, cudaMemcpyHostToDevice, s));
      else FL_CUDA(cudaMemcpy2DAsync(h->d_in_depth, (size_t)W * 2, depth, depth_stride, (size_t)W * 2, H, cudaMemcpyHostToDevice, s));
    } else {
      for (int y = 0; y < H; ++y) memcpy(h->h_depth + (size_t)y * W, (const uint8_t*)depth + (size_t)y * depth_stride, (size_t)W * 2);
      FL_CUDA(cudaMemcpyAsync(h->d_in_depth, h->h_depth, (size_t)W * H * 2, cudaMemcpyHostToDevice, s));
    }
    *d_depth_out = h->d_in_depth;
  }
  if (bgr) {
    if (bgr_stride < (size_t)W * 3) return FL_ERR_SIZE;
    if (is_pinned(bgr)) {
      if (bgr_stride == (size_t)W * 3) FL_CUDA(cudaMemcpyAsync(h->d_in_bgr, bgr, (size_t)W * H * 3, cudaMemcpyHostToDevice, s));
      else FL_CUDA(cudaMemcpy2DAsync(h->d_in_bgr, (size_t)W * 3, bgr, bgr_stride, (size_t)W * 3, H, cudaMemcpyHostToDevice, s));
    } else {
      // pageable source: stage in two halves so the second half's host copy overlaps the first half's DMA
      const int h0 = H / 2;
      for (int part = 0; part < 2; ++part) {
        const int y0 = part ? h0 : 0, y1 = part ? H : h0;
        for (int y = y0; y < y1; ++y) memcpy(h->h_bgr + (size_t)y * W * 3, bgr + (size_t)y * bgr_stride, (size_t)W * 3);
        if (y1 > y0) FL_CUDA(cudaMemcpyAsync(h->d_in_bgr + (size_t)y0 * W * 3, h->h_bgr + (size_t)y0 * W * 3, (size_t)(y1 - y0) * W * 3, cudaMemcpyHostToDevice, s));
      }
    }
    *d_bgr_out = h->d_in_bgr;
  }
  for (int m = 0; m < FL_MAX_MODALITIES; ++m) d_masks[m] = nullptr;
  if (masks) for (int m = 0; m < p.n_modalities; ++m) if (masks[m]) {
    if (!h->d_mask[0][m]) { size_t n = (size_t)p.max_width * p.max_height; TRY(dalloc(&h->d_mask[0][m], n)); TRY(dalloc(&h->d_qm[0][m], n)); }
    memcpy(h->h_mask + (size_t)m * W * H, masks[m], (size_t)W * H);
    FL_CUDA(cudaMemcpyAsync(h->d_mask[0][m], h->h_mask + (size_t)m * W * H, (size_t)W * H, cudaMemcpyHostToDevice, s));
    d_masks[m] = h->d_mask[0][m]; *any_mask = true;
  }
  return FL_OK;
}

// enqueue-only half of fl_match: upload + every kernel of the frame on the handle's stream; fl_match_wait / fl_match_fetch finish it
extern "C" int fl_match_async(fl_handle* h, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, int32_t W, int32_t H,
                              const uint8_t* const* masks, float threshold, const int32_t* class_filter, int32_t n_filter) {
  if (!h) return FL_ERR_ARG;
  const fl_params_t& p = h->p;
  // an asynchronous frame still in flight may be reading the staging buffers this call is about to overwrite
  if (h->pend_sort) { fl_set_error("fl_match_wait has not been called for the previous frame"); return FL_ERR_STATE; }
  h->in_depth_W = h->in_depth_H = 0;                                               // d_in_depth counts as "the matched frame" only once this call has succeeded
  if (W <= 0 || H <= 0 || W > p.max_width || H > p.max_height) return FL_ERR_SIZE;
  FL_CUDA(cudaSetDevice(p.device));
  const uint8_t* d_bgr; const uint16_t* d_depth;
  const void* d_masks[FL_MAX_MODALITIES];
  bool any_mask;
  TRY(stage_frame(h, bgr, bgr_stride, depth, depth_stride, W, H, masks, &d_bgr, &d_depth, d_masks, &any_mask));
  TRY(fl_match_device_async(h, d_bgr, d_depth, W, H, any_mask ? d_masks : nullptr, threshold, class_filter, n_filter));
  if (d_depth) { h->in_depth_W = W; h->in_depth_H = H; }
  return FL_OK;
}

extern "C" int fl_match(fl_handle* h, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, int32_t W, int32_t H,
                        const uint8_t* const* masks, float threshold, const int32_t* class_filter, int32_t n_filter, fl_match_t* out,
                        int32_t capacity, int32_t* count, uint8_t* const* quantized_out) {
  if (!h || !count) return FL_ERR_ARG;
  *count = 0;
  const fl_params_t& p = h->p;
  int rc = fl_match_async(h, bgr, bgr_stride, depth, depth_stride, W, H, masks, threshold, class_filter, n_filter);
  if (rc != FL_OK) return rc;
  rc = fl_match_wait(h);
  if (rc != FL_OK) { h->in_depth_W = h->in_depth_H = 0; return rc; }
  cudaStream_t s = h->stream;
  rc = fl_match_fetch(h, out, capacity, count);
  if (quantized_out) {
    for (int l = 0; l < p.n_levels; ++l) for (int m = 0; m < p.n_modalities; ++m) {
      uint8_t* dst = quantized_out[l * p.n_modalities + m];
      if (!dst) continue;
      const uint8_t* src = h->used_mask[m] ? h->d_qm[l][m] : h->d_q[l][m];
      FL_CUDA(cudaMemcpyAsync(dst, src, (size_t)h->geom[l].W * h->geom[l].H, cudaMemcpyDeviceToHost, s));
    }
    FL_CUDA(cudaStreamSynchronize(s));
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------------
// Detector::addTemplate (linemod.cpp:1579-1615)
// ---------------------------------------------------------------------------------------------------
extern "C" void fl_default_train_params(fl_train_params_t* p) {
  if (!p) return;
  for (int m = 0; m < FL_MAX_MODALITIES; ++m) p->num_features[m] = 63;
  p->strong_threshold = 55.0f; p->extract_threshold = 2;
}

namespace {
struct train_cand { int x, y, label; float score; };
// QuantizedPyramid::selectScatteredFeatures (linemod.cpp:134-163), candidates already sorted
void select_scattered(const std::vector<train_cand>& cands, std::vector<fl_feature_t>& out, size_t num_features, float distance) {
  out.clear();
  float distance_sq = distance * distance;
  int i = 0;
  while (out.size() < num_features) {
    const train_cand& c = cands[i];
    bool keep = true;
    for (int j = 0; j < (int)out.size() && keep; ++j)
      keep = (float)((c.x - out[j].x) * (c.x - out[j].x) + (c.y - out[j].y) * (c.y - out[j].y)) >= distance_sq;
    if (keep) { fl_feature_t f = {c.x, c.y, c.label}; out.push_back(f); }
    if (++i == (int)cands.size()) { i = 0; distance -= 1.0f; distance_sq = distance * distance; }   // start over with a relaxed distance
  }
}
}  // namespace

extern "C" int fl_add_template(fl_handle* h, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, const uint8_t* mask,
                               size_t mask_stride, int32_t W, int32_t H, const fl_train_params_t* params_or_null, fl_template_hdr_t* headers,
                               fl_feature_t* features, int32_t feature_capacity, int32_t* n_features, fl_rect_t* bounding_box) {
  if (!h || !headers || !n_features || feature_capacity < 0 || (feature_capacity > 0 && !features)) return FL_ERR_ARG;
  *n_features = 0;
  fl_train_params_t tp;
  if (params_or_null) tp = *params_or_null; else fl_default_train_params(&tp);
  const fl_params_t& p = h->p;
  const int L = p.n_levels, M = p.n_modalities;
  if (mask && mask_stride < (size_t)W) return FL_ERR_SIZE;
  // 1. the front end: quantised images of every level and modality (unmasked, as the pyramids' `angle` / `normal` members), the colour
  //    pyramid.  A threshold above 100 % keeps matchClass from emitting anything if templates are loaded.
  TRY(fl_match_async(h, bgr, bgr_stride, depth, depth_stride, W, H, nullptr, 200.f, nullptr, 0));
  TRY(fl_match_wait(h));
  cudaStream_t s = h->stream;
  const size_t npx = (size_t)W * H;
  // training workspace of the handle: one block sized for the largest frame, allocated at the first call (a database generator calls
  // this thousands of times; a malloc / free round per view cost more than the kernels)
  if (!h->d_train) {
    const size_t cap = (size_t)p.max_width * p.max_height;
    TRY(dalloc(&h->d_train, cap * (4 + 4 + 1 + 1 + 16 + 2) + 256));
  }
  uint8_t* base = h->d_train;
  const size_t cap_px = (size_t)p.max_width * p.max_height;
  float* d_mag = reinterpret_cast<float*>(base); base += cap_px * 4;
  float* d_score = reinterpret_cast<float*>(base); base += cap_px * 4;
  uint16_t* d_hd = reinterpret_cast<uint16_t*>(base); base += cap_px * 16;
  int* d_cnt = reinterpret_cast<int*>(base); base += 256;
  uint8_t* d_er1 = base; base += cap_px;
  uint8_t* d_er2 = base; base += cap_px;
  uint8_t* d_maskpyr = base;
  std::vector<const uint8_t*> d_mask(L, nullptr);
  if (mask) {
    FL_CUDA(cudaMemcpy2DAsync(d_maskpyr, W, mask, mask_stride, W, H, cudaMemcpyHostToDevice, s));
    d_mask[0] = d_maskpyr;
    size_t off = npx;
    for (int l = 1; l < L; ++l) {                           // resize(mask, next_mask, size / 2, INTER_NEAREST) (:447-450, :734-738)
      uint8_t* dst = d_maskpyr + off;
      fl_launch_resize_nn_half(d_mask[l - 1], W >> (l - 1), H >> (l - 1), dst, s); ++h->launches;
      d_mask[l] = dst; off += (size_t)(W >> l) * (H >> l);
    }
  }
  std::vector<std::vector<fl_feature_t> > feats((size_t)L * M);
  std::vector<float> h_score(npx);
  std::vector<uint8_t> h_q(npx), h_lmask(npx);
  for (int m = 0; m < M; ++m) {
    const bool is_color = p.modality_kind[m] == FL_MODALITY_COLOR_GRADIENT;
    for (int l = 0; l < L; ++l) {
      const int w = W >> l, hh = H >> l;
      const size_t n = (size_t)w * hh;
      const size_t num_features = (size_t)tp.num_features[m] >> l;              // num_features /= 2 per pyrDown (:429, :724)
      const uint8_t* d_q = h->d_q[l][m];
      std::vector<train_cand> cands;
      float distance;
      int counts[16] = {0};
      if (is_color) {
        const uint8_t* d_bgr_l = l == 0 ? h->d_in_bgr : h->d_bgr[l];
        fl_launch_train_magnitude(d_bgr_l, w, hh, d_mag, s); ++h->launches;
        if (d_mask[l]) { fl_launch_train_erode3(d_mask[l], w, hh, d_er1, s); ++h->launches; }
        fl_launch_train_color_score(d_q, d_mag, d_mask[l], d_mask[l] ? d_er1 : nullptr, w, hh, tp.strong_threshold * tp.strong_threshold, d_score, s); ++h->launches;
      } else {
        const uint8_t* d_local = nullptr;
        if (d_mask[l]) {                                     // erode(mask, local_mask, Mat(), Point(-1,-1), 2, BORDER_REPLICATE) (:753)
          fl_launch_train_erode3(d_mask[l], w, hh, d_er1, s); fl_launch_train_erode3(d_er1, w, hh, d_er2, s); h->launches += 2;
          d_local = d_er2;
          FL_CUDA(cudaMemcpyAsync(h_lmask.data(), d_er2, n, cudaMemcpyDeviceToHost, s));
        }
        fl_launch_train_depth_score(d_q, d_local, d_hd, d_cnt, w, hh, tp.extract_threshold >> l, d_score, d_cnt + 8, s); h->launches += 2;   // extract_threshold /= 2 per level (:725)
        FL_CUDA(cudaMemcpyAsync(counts, d_cnt, sizeof counts, cudaMemcpyDeviceToHost, s));
      }
      FL_CUDA(cudaMemcpyAsync(h_score.data(), d_score, n * sizeof(float), cudaMemcpyDeviceToHost, s));
      FL_CUDA(cudaMemcpyAsync(h_q.data(), d_q, n, cudaMemcpyDeviceToHost, s));
      FL_CUDA(wait_stream(h));
      for (int y = 0; y < hh; ++y)                            // candidates in row-major order: the order std::stable_sort keeps among ties
        for (int x = 0; x < w; ++x) {
          const float sc = h_score[(size_t)y * w + x];
          if (sc < 0.f) continue;
          const int q = h_q[(size_t)y * w + x];
          int label = 0; while (label < 7 && !(q & (1 << label))) ++label;      // getLabel: the set bit
          train_cand c = {x, y, label, sc};
          cands.push_back(c);
        }
      if (cands.size() < num_features) return FL_ERR_TRAIN;                       // (:500-501, :803-804)
      if (is_color) {
        distance = num_features ? static_cast<float>(cands.size() / num_features + 1) : 0.f;                         // (:506)
      } else {
        for (train_cand& c : cands) c.score /= (float)counts[8 + c.label];         // penalise labels with many candidates (:808-812)
        size_t area = n;
        if (d_mask[l]) { area = 0; for (size_t i = 0; i < n; ++i) area += h_lmask[i] != 0; }
        distance = num_features ? sqrtf((float)area) / sqrtf((float)num_features) + 1.5f : 0.f;                       // (:816-817)
      }
      std::stable_sort(cands.begin(), cands.end(), [](const train_cand& a, const train_cand& b) { return a.score > b.score; });   // Candidate::operator< (linemod.hpp:124-128)
      select_scattered(cands, feats[(size_t)l * M + m], num_features, distance);
    }
  }
  // cropTemplates (linemod.cpp:52-96)
  int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
  for (int l = 0; l < L; ++l) for (int m = 0; m < M; ++m) for (const fl_feature_t& f : feats[(size_t)l * M + m]) {
    const int x = f.x << l, y = f.y << l;
    min_x = std::min(min_x, x); min_y = std::min(min_y, y); max_x = std::max(max_x, x); max_y = std::max(max_y, y);
  }
  if (min_x % 2 == 1) --min_x;
  if (min_y % 2 == 1) --min_y;
  int nf = 0;
  for (int l = 0; l < L; ++l) for (int m = 0; m < M; ++m) {
    fl_template_hdr_t& hd = headers[(size_t)l * M + m];
    hd.width = (max_x - min_x) >> l; hd.height = (max_y - min_y) >> l; hd.offset_x = min_x >> l; hd.offset_y = min_y >> l; hd.pyramid_level = l;
    hd.feature_begin = nf; hd.feature_count = (int)feats[(size_t)l * M + m].size();
    for (const fl_feature_t& f : feats[(size_t)l * M + m]) {
      if (nf >= feature_capacity) return FL_ERR_CAPACITY;
      fl_feature_t g = {f.x - hd.offset_x, f.y - hd.offset_y, f.label};
      features[nf++] = g;
    }
  }
  *n_features = nf;
  if (bounding_box) { bounding_box->x = min_x; bounding_box->y = min_y; bounding_box->width = max_x - min_x; bounding_box->height = max_y - min_y; }
  return FL_OK;
}

extern "C" int fl_match_shard_device(fl_handle* h, const void* d_bgr, const void* d_depth, int32_t W, int32_t H, const void* const* d_masks,
                                     float threshold, const int32_t* class_filter, int32_t n_filter, fl_match_t* d_candidates,
                                     int32_t capacity, int32_t* d_count) {
  if (!h || !d_candidates || !d_count || capacity < 1) return FL_ERR_ARG;
  return run_match_stages(h, (const uint8_t*)d_bgr, (const uint16_t*)d_depth, W, H, d_masks, threshold, class_filter, n_filter, d_candidates,
                          capacity, d_count);
}

extern "C" int fl_sort_unique_device(fl_handle* h, const fl_match_t* d_in, int32_t n_lists, int32_t list_capacity, const int32_t* d_n_in,
                                     fl_match_t* d_out, int32_t out_capacity, int32_t* d_out_count) {
  if (!h || !d_in || !d_n_in || !d_out || !d_out_count || n_lists < 1 || list_capacity < 1 || out_capacity < 1) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  const fl_lists lists = {d_in, n_lists, list_capacity, list_capacity, d_n_in, 1};
  return run_sort_unique(h, lists, d_out, out_capacity, d_out_count, false);
}

extern "C" size_t fl_exchange_buffer_bytes(int32_t world, int32_t capacity) {
  if (world < 1 || capacity < 1) return 0;
  return FL_XCHG_SIGNALS * sizeof(unsigned) + (size_t)2 * world * ((size_t)capacity + 1) * sizeof(fl_match_t);
}

extern "C" int fl_exchange_sort_unique_device_async(fl_handle* h, int32_t rank, int32_t world, void* const* peer_buffers, int32_t capacity,
                                                    const fl_match_t* d_local_block, uint32_t epoch) {
  if (!h || !peer_buffers || !d_local_block || world < 1 || world > FL_XCHG_MAX_WORLD || rank < 0 || rank >= world || capacity < 1 || epoch == 0) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  if (h->pend_sort) { fl_set_error("fl_match_wait has not been called for the previous frame"); return FL_ERR_STATE; }
  h->have_result = false; h->pend_match = false;
  if (h->profile) for (int i = 0; i < 4; ++i) cudaEventRecord(h->ev[i], h->stream);
  fl_xchg X;
  memset(&X, 0, sizeof X);
  X.world = world; X.rank = rank; X.cap = capacity; X.epoch = epoch; X.local_block = d_local_block;
  for (int p = 0; p < world; ++p) { if (!peer_buffers[p]) return FL_ERR_ARG; X.peer[p] = static_cast<uint8_t*>(peer_buffers[p]); }
  // the lists the multi-kernel fallback would read: this rank's own buffer, parity of this epoch (filled by the kernel's exchange)
  const uint8_t* own = X.peer[rank] + FL_XCHG_SIGNALS * sizeof(unsigned) + (size_t)(epoch & 1u) * world * ((size_t)capacity + 1) * sizeof(fl_match_t);
  const fl_lists lists = {reinterpret_cast<const fl_match_t*>(own) + 1, world, capacity, capacity + 1, reinterpret_cast<const int*>(own), 5 * (capacity + 1)};
  return sort_launch(h, lists, h->d_out, h->p.max_candidates, h->d_out_count, true, &X);
}

// Template-sharded frame in one go (enqueue only): front end + matchClass on this handle's shard with the candidates written
// straight into the caller's block [count | records], then ONE launch that refines them, pushes the block to every peer,
// waits for the peers' blocks and sorts the union (k_refine_sort<1> with the exchange in its last CTA).
extern "C" int fl_match_shard_exchange_device_async(fl_handle* h, const void* d_bgr, const void* d_depth, int32_t W, int32_t H, float threshold,
                                                    const int32_t* class_filter, int32_t n_filter, int32_t rank, int32_t world,
                                                    void* const* peer_buffers, int32_t capacity, fl_match_t* d_local_block, uint32_t epoch) {
  if (!h || !peer_buffers || !d_local_block || world < 1 || world > FL_XCHG_MAX_WORLD || rank < 0 || rank >= world || capacity < 1 || epoch == 0) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  if (h->pend_sort) { fl_set_error("fl_match_wait has not been called for the previous frame"); return FL_ERR_STATE; }
  h->have_result = false; h->pend_match = false;
  const bool defer = !h->opt_split_refine && h->n_templates > 0 && h->p.n_levels > 1;
  fl_match_t* recs = d_local_block + 1;
  int* d_cnt = reinterpret_cast<int*>(d_local_block);
  TRY(run_match_stages(h, (const uint8_t*)d_bgr, (const uint16_t*)d_depth, W, H, nullptr, threshold, class_filter, n_filter, recs, capacity, d_cnt, defer));
  fl_xchg X;
  memset(&X, 0, sizeof X);
  X.world = world; X.rank = rank; X.cap = capacity; X.epoch = epoch; X.local_block = d_local_block;
  for (int p = 0; p < world; ++p) { if (!peer_buffers[p]) return FL_ERR_ARG; X.peer[p] = static_cast<uint8_t*>(peer_buffers[p]); }
  const uint8_t* own = X.peer[rank] + FL_XCHG_SIGNALS * sizeof(unsigned) + (size_t)(epoch & 1u) * world * ((size_t)capacity + 1) * sizeof(fl_match_t);
  const fl_lists lists = {reinterpret_cast<const fl_match_t*>(own) + 1, world, capacity, capacity + 1, reinterpret_cast<const int*>(own), 5 * (capacity + 1)};
  const fl_refine_req req = {threshold, recs, capacity, d_cnt};
  return sort_launch(h, lists, h->d_out, h->p.max_candidates, h->d_out_count, true, &X, defer ? &req : nullptr);
}

extern "C" int fl_match_shard_exchange_async(fl_handle* h, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, int32_t W, int32_t H,
                                             float threshold, const int32_t* class_filter, int32_t n_filter, int32_t rank, int32_t world,
                                             void* const* peer_buffers, int32_t capacity, fl_match_t* d_local_block, uint32_t epoch) {
  if (!h) return FL_ERR_ARG;
  if (h->pend_sort) { fl_set_error("fl_match_wait has not been called for the previous frame"); return FL_ERR_STATE; }
  h->in_depth_W = h->in_depth_H = 0;
  if (W <= 0 || H <= 0 || W > h->p.max_width || H > h->p.max_height) return FL_ERR_SIZE;
  FL_CUDA(cudaSetDevice(h->p.device));
  const uint8_t* d_bgr; const uint16_t* d_depth;
  const void* d_masks[FL_MAX_MODALITIES];
  bool any_mask;
  TRY(stage_frame(h, bgr, bgr_stride, depth, depth_stride, W, H, nullptr, &d_bgr, &d_depth, d_masks, &any_mask));
  TRY(fl_match_shard_exchange_device_async(h, d_bgr, d_depth, W, H, threshold, class_filter, n_filter, rank, world, peer_buffers, capacity, d_local_block, epoch));
  if (d_depth) { h->in_depth_W = W; h->in_depth_H = H; }
  return FL_OK;
}

extern "C" int fl_exchange_sort_unique_device(fl_handle* h, int32_t rank, int32_t world, void* const* peer_buffers, int32_t capacity,
                                              const fl_match_t* d_local_block, uint32_t epoch) {
  int rc = fl_exchange_sort_unique_device_async(h, rank, world, peer_buffers, capacity, d_local_block, epoch);
  if (rc != FL_OK) return rc;
  return fl_match_wait(h);
}

extern "C" int fl_sort_unique_blocks_device(fl_handle* h, const fl_match_t* d_blocks, int32_t n_blocks, int32_t capacity) {
  if (!h || !d_blocks || n_blocks < 1 || capacity < 1) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  h->have_result = false;
  if (h->profile) for (int i = 0; i < 4; ++i) cudaEventRecord(h->ev[i], h->stream);   // stage times are not defined for this entry point
  const fl_lists lists = {d_blocks + 1, n_blocks, capacity, capacity + 1, reinterpret_cast<const int*>(d_blocks), 5 * (capacity + 1)};
  TRY(run_sort_unique(h, lists, h->d_out, h->p.max_candidates, h->d_out_count, true));
  h->have_result = true;
  return FL_OK;
}

// ---------------------------------------------------------------------------------------------------
extern "C" int fl_debug_get(fl_handle* h, int what, int32_t a, int32_t b, int32_t c, void* host_out, size_t bytes) {
  if (!h || !host_out) return FL_ERR_ARG;
  if (h->gW == 0) return FL_ERR_STATE;
  const fl_params_t& p = h->p;
  FL_CUDA(cudaSetDevice(p.device));
  FL_CUDA(cudaStreamSynchronize(h->stream));
  if (what == FL_DBG_QUANTIZED || what == FL_DBG_SPREAD) {
    if (a < 0 || a >= p.n_levels || b < 0 || b >= p.n_modalities) return FL_ERR_ARG;
    size_t n = (size_t)h->geom[a].W * h->geom[a].H;
    if (bytes < n) return FL_ERR_CAPACITY;
    const uint8_t* src = what == FL_DBG_SPREAD ? h->d_spread[a][b] : (h->used_mask[b] ? h->d_qm[a][b] : h->d_q[a][b]);
    if (!src) return FL_ERR_STATE;
    FL_CUDA(cudaMemcpy(host_out, src, n, cudaMemcpyDeviceToHost));
    return FL_OK;
  }
  if (what == FL_DBG_LINEAR_MEMORY) {
    if (a < 0 || a >= p.n_levels || b < 0 || b >= p.n_modalities || c < 0 || c > 7) return FL_ERR_ARG;
    const fl_level_geom& g = h->geom[a];
    size_t n = (size_t)g.T * g.T * g.cells;
    if (bytes < n) return FL_ERR_CAPACITY;
    FL_CUDA(cudaMemcpy(host_out, h->d_lm[a] + (size_t)b * g.mod_stride + (size_t)c * g.label_stride, n, cudaMemcpyDeviceToHost));
    return FL_OK;
  }
  if (what == FL_DBG_SIMILARITY) {
    if (a < 0 || a >= h->n_templates) return FL_ERR_ARG;
    const fl_level_geom& g = h->geom[p.n_levels - 1];
    if (bytes < (size_t)g.cells * 2) return FL_ERR_CAPACITY;
    uint16_t* d_tmp = reinterpret_cast<uint16_t*>(h->d_keys);                   // scratch
    if ((size_t)g.cells * 2 > (size_t)h->key_cap * sizeof(fl_sort_key)) return FL_ERR_CAPACITY;
    fl_launch_similarity_debug(make_tdb(h), g, h->d_lm[p.n_levels - 1], a, d_tmp, h->stream); ++h->launches;
    FL_CUDA(cudaStreamSynchronize(h->stream));
    FL_CUDA(cudaMemcpy(host_out, d_tmp, (size_t)g.cells * 2, cudaMemcpyDeviceToHost));
    return FL_OK;
  }
  if (what == FL_DBG_LAST_COUNTS) {
    if (bytes < 16 * sizeof(int)) return FL_ERR_CAPACITY;
    memcpy(host_out, h->h_small, 16 * sizeof(int));
    return FL_OK;
  }
  if (what == FL_DBG_FE_TRACE) {                                                               // developer: front-end job timeline {kind, ctas, start, end} x jobs (FL_OPT_TRACE)
    if (!h->d_fe_trace || h->fe_trace_jobs <= 0) return FL_ERR_STATE;
    if (bytes < (size_t)h->fe_trace_jobs * 4 * sizeof(unsigned long long)) return FL_ERR_CAPACITY;
    unsigned long long raw[2 * FL_FE_MAX_JOBS];
    FL_CUDA(cudaMemcpy(raw, h->d_fe_trace, sizeof raw, cudaMemcpyDeviceToHost));
    unsigned long long* o = static_cast<unsigned long long*>(host_out);
    for (int i = 0; i < h->fe_trace_jobs; ++i) { o[4 * i] = (unsigned long long)h->fe_trace_kind[i]; o[4 * i + 1] = (unsigned long long)h->fe_trace_ctas[i]; o[4 * i + 2] = raw[2 * i]; o[4 * i + 3] = raw[2 * i + 1]; }
    return h->fe_trace_jobs;
  }
  if (what == FL_DBG_STAGED_TRACE) {
    if (!h->use_staged || !h->plan.trace) return FL_ERR_STATE;
    size_t n = ((size_t)h->plan.n_cta * 136 + 8) * sizeof(unsigned long long);
    if (bytes < n) return FL_ERR_CAPACITY;
    FL_CUDA(cudaMemcpy(host_out, h->plan.trace, n, cudaMemcpyDeviceToHost));
    return h->plan.n_cta;
  }
  return FL_ERR_ARG;
}

// ---------------------------------------------------------------------------------------------------
// input rescale: cv::resize INTER_LINEAR tables (OpenCV imgproc/src/resize.cpp, the loops that fill xofs / alpha / yofs / beta)
// ---------------------------------------------------------------------------------------------------
static int resize_tables(fl_handle* h, int sW, int sH, int dW, int dH) {
  if (h->rz.xofs && h->rz_sW == sW && h->rz_sH == sH && h->rz_dW == dW && h->rz_dH == dH) return FL_OK;
  FL_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(h->rz.xofs); cudaFree(h->rz.yofs); cudaFree(h->rz.ialpha); cudaFree(h->rz.ibeta); cudaFree(h->rz.alpha); cudaFree(h->rz.beta);
  memset(&h->rz, 0, sizeof h->rz);
  std::vector<int> xo, yo; std::vector<float> xa, yb;
  std::vector<short> ia, ib;
  fl_resize_axis(sW, dW, true, xo, xa, ia); fl_resize_axis(sH, dH, false, yo, yb, ib);
  TRY(dalloc(&h->rz.xofs, (size_t)dW)); TRY(dalloc(&h->rz.yofs, (size_t)dH));
  TRY(dalloc(&h->rz.ialpha, (size_t)dW)); TRY(dalloc(&h->rz.ibeta, (size_t)dH)); TRY(dalloc(&h->rz.alpha, (size_t)dW)); TRY(dalloc(&h->rz.beta, (size_t)dH));
  FL_CUDA(cudaMemcpy(h->rz.xofs, xo.data(), (size_t)dW * 4, cudaMemcpyHostToDevice)); FL_CUDA(cudaMemcpy(h->rz.yofs, yo.data(), (size_t)dH * 4, cudaMemcpyHostToDevice));
  FL_CUDA(cudaMemcpy(h->rz.ialpha, ia.data(), (size_t)dW * 4, cudaMemcpyHostToDevice)); FL_CUDA(cudaMemcpy(h->rz.ibeta, ib.data(), (size_t)dH * 4, cudaMemcpyHostToDevice));
  FL_CUDA(cudaMemcpy(h->rz.alpha, xa.data(), (size_t)dW * 8, cudaMemcpyHostToDevice)); FL_CUDA(cudaMemcpy(h->rz.beta, yb.data(), (size_t)dH * 8, cudaMemcpyHostToDevice));
  h->rz_sW = sW; h->rz_sH = sH; h->rz_dW = dW; h->rz_dH = dH;
  return FL_OK;
}
static int resize_staging(fl_handle* h, size_t npx) {
  if (npx <= h->src_cap) return FL_OK;
  FL_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_src_bgr); cudaFree(h->d_src_depth); h->d_src_bgr = nullptr; h->d_src_depth = nullptr; h->src_cap = 0;
  TRY(dalloc(&h->d_src_bgr, npx * 3)); TRY(dalloc(&h->d_src_depth, npx));
  h->src_cap = npx;
  return FL_OK;
}

extern "C" int fl_resize_linear(fl_handle* h, const void* src, size_t src_stride, int32_t sW, int32_t sH, int32_t type, void* dst, int32_t W, int32_t H) {
  if (!h || !src || !dst || sW <= 0 || sH <= 0 || W <= 0 || H <= 0 || (type != FL_IMG_8UC3 && type != FL_IMG_16UC1)) return FL_ERR_ARG;
  const size_t px = type == FL_IMG_8UC3 ? 3 : 2;
  if (src_stride < (size_t)sW * px) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  cudaStream_t s = h->stream;
  TRY(resize_tables(h, sW, sH, W, H));
  TRY(resize_staging(h, (size_t)sW * sH));
  void* d_src = type == FL_IMG_8UC3 ? (void*)h->d_src_bgr : (void*)h->d_src_depth;
  uint8_t* d_dst = nullptr;
  dev_buf tmp;
  TRY(tmp.get(&d_dst, (size_t)W * H * px));
  FL_CUDA(cudaMemcpy2DAsync(d_src, (size_t)sW * px, src, src_stride, (size_t)sW * px, sH, cudaMemcpyHostToDevice, s));
  fl_launch_resize_linear(d_src, sW, sH, type, d_dst, W, H, h->rz, s); ++h->launches;
  FL_CUDA(cudaMemcpyAsync(dst, d_dst, (size_t)W * H * px, cudaMemcpyDeviceToHost, s));
  FL_CUDA(cudaStreamSynchronize(s));
  FL_CUDA(cudaGetLastError());
  return FL_OK;
}

extern "C" int fl_match_rescaled(fl_handle* h, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, int32_t sW, int32_t sH,
                                 int32_t W, int32_t H, float threshold, const int32_t* class_filter, int32_t n_filter, fl_match_t* out, int32_t capacity,
                                 int32_t* count, uint16_t* rescaled_depth_out) {
  if (!h || !count || sW <= 0 || sH <= 0) return FL_ERR_ARG;
  *count = 0;
  const fl_params_t& p = h->p;
  if (h->pend_sort) { fl_set_error("fl_match_wait has not been called for the previous frame"); return FL_ERR_STATE; }
  h->in_depth_W = h->in_depth_H = 0;
  if (W <= 0 || H <= 0 || W > p.max_width || H > p.max_height) return FL_ERR_SIZE;
  if (sW == W && sH == H) {                                                      // TImage2Mat resizes only when the width differs (:42)
    int rc = fl_match(h, bgr, bgr_stride, depth, depth_stride, W, H, nullptr, threshold, class_filter, n_filter, out, capacity, count, nullptr);
    if (rescaled_depth_out && depth) for (int y = 0; y < H; ++y) memcpy(rescaled_depth_out + (size_t)y * W, (const uint8_t*)depth + (size_t)y * depth_stride, (size_t)W * 2);
    return rc;
  }
  if ((bgr && bgr_stride < (size_t)sW * 3) || (depth && depth_stride < (size_t)sW * 2)) return FL_ERR_SIZE;
  FL_CUDA(cudaSetDevice(p.device));
  cudaStream_t s = h->stream;
  TRY(resize_tables(h, sW, sH, W, H));
  TRY(resize_staging(h, (size_t)sW * sH));
  const uint8_t* d_bgr = nullptr; const uint16_t* d_depth = nullptr;
  if (depth) {
    FL_CUDA(cudaMemcpy2DAsync(h->d_src_depth, (size_t)sW * 2, depth, depth_stride, (size_t)sW * 2, sH, cudaMemcpyHostToDevice, s));
    fl_launch_resize_linear(h->d_src_depth, sW, sH, FL_IMG_16UC1, h->d_in_depth, W, H, h->rz, s); ++h->launches;
    d_depth = h->d_in_depth;
    if (rescaled_depth_out) FL_CUDA(cudaMemcpyAsync(rescaled_depth_out, h->d_in_depth, (size_t)W * H * 2, cudaMemcpyDeviceToHost, s));
  }
  if (bgr) {
    FL_CUDA(cudaMemcpy2DAsync(h->d_src_bgr, (size_t)sW * 3, bgr, bgr_stride, (size_t)sW * 3, sH, cudaMemcpyHostToDevice, s));
    fl_launch_resize_linear(h->d_src_bgr, sW, sH, FL_IMG_8UC3, h->d_in_bgr, W, H, h->rz, s); ++h->launches;
    d_bgr = h->d_in_bgr;
  }
  int rc = fl_match_device(h, d_bgr, d_depth, W, H, nullptr, threshold, class_filter, n_filter);
  if (rc != FL_OK) return rc;
  if (d_depth) { h->in_depth_W = W; h->in_depth_H = H; }
  return fl_match_fetch(h, out, capacity, count);
}

// ---------------------------------------------------------------------------------------------------
// ICP
// ---------------------------------------------------------------------------------------------------
static int icp_reserve(fl_handle* h, int n_hyp, int max_pts) {
  if (n_hyp <= h->icp_hyp_cap && max_pts <= h->icp_pts_cap) { h->icp.n_hyp = n_hyp; return FL_OK; }
  int nh = std::max(n_hyp, h->icp_hyp_cap), np = (std::max(max_pts, h->icp_pts_cap) + 3) & ~3;
  FL_CUDA(cudaStreamSynchronize(h->stream));
  icp_free(h);
  size_t tot = (size_t)nh * np;
  TRY(dalloc(&h->icp.pts_ref, tot * 3)); TRY(dalloc(&h->icp.pts_mod, tot * 3)); TRY(dalloc(&h->icp.cor_m, tot * 3)); TRY(dalloc(&h->icp.cor_r, tot * 3));
  TRY(dalloc(&h->icp.dist, tot)); TRY(dalloc(&h->icp.grid_pts, tot)); TRY(dalloc(&h->icp.nn_d2, tot)); TRY(dalloc(&h->icp.nn_slot, tot));
  TRY(dalloc(&h->d_icp_ticket, (size_t)1));
  TRY(dalloc(&h->icp.n_ref, (size_t)nh)); TRY(dalloc(&h->icp.n_mod, (size_t)nh));
  TRY(dalloc(&h->d_hyps, (size_t)nh)); TRY(dalloc(&h->d_results, (size_t)nh)); TRY(dalloc(&h->d_model_crops, tot));
  TRY(halloc(&h->h_model_crops, tot)); TRY(halloc(&h->h_hyps, (size_t)nh)); TRY(halloc(&h->h_results, (size_t)nh));
  h->icp_hyp_cap = nh; h->icp_pts_cap = np; h->icp.max_pts = np; h->icp.n_hyp = n_hyp;
  return FL_OK;
}

extern "C" int fl_depth_to_3d(fl_handle* h, const uint16_t* depth, size_t depth_stride, int32_t W, int32_t H, fl_intrinsics_t K, float* out3) {
  if (!h || !depth || !out3 || W <= 0 || H <= 0 || depth_stride < (size_t)W * 2) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  size_t n = (size_t)W * H;
  uint16_t* d_d = nullptr; float* d_o = nullptr;
  dev_buf tmp;
  TRY(tmp.get(&d_d, n)); TRY(tmp.get(&d_o, n * 3));
  FL_CUDA(cudaMemcpy2DAsync(d_d, (size_t)W * 2, depth, depth_stride, (size_t)W * 2, H, cudaMemcpyHostToDevice, h->stream));
  fl_launch_depth_to_3d(d_d, W, H, K, d_o, h->stream); ++h->launches;
  FL_CUDA(cudaMemcpyAsync(out3, d_o, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  FL_CUDA(cudaStreamSynchronize(h->stream));
  return FL_OK;
}

extern "C" int fl_icp_cloud_to_cloud_ex(fl_handle* h, const float* pts_ref, int32_t n_ref, const float* pts_model, int32_t n_model,
                                        fl_icp_params_t prm, fl_icp_result_t* out) {
  if (!h || !out || n_ref < 0 || n_model < 0 || (n_ref > 0 && !pts_ref) || (n_model > 0 && !pts_model)) return FL_ERR_ARG;
  if (n_ref >= 3 && n_model >= 3 && n_ref < n_model) { fl_set_error("icpCloudToCloud_Ex walks both clouds in lock step: n_ref (%d) must be >= n_model (%d)", n_ref, n_model); return FL_ERR_SIZE; }
  FL_CUDA(cudaSetDevice(h->p.device));
  cudaStream_t s = h->stream;
  TRY(icp_reserve(h, 1, std::max(std::max(n_ref, n_model), 16)));
  if (n_ref) FL_CUDA(cudaMemcpyAsync(h->icp.pts_ref, pts_ref, (size_t)n_ref * 12, cudaMemcpyHostToDevice, s));
  if (n_model) FL_CUDA(cudaMemcpyAsync(h->icp.pts_mod, pts_model, (size_t)n_model * 12, cudaMemcpyHostToDevice, s));
  FL_CUDA(cudaMemcpyAsync(h->icp.n_ref, &n_ref, 4, cudaMemcpyHostToDevice, s));
  FL_CUDA(cudaMemcpyAsync(h->icp.n_mod, &n_model, 4, cudaMemcpyHostToDevice, s));
  FL_CUDA(cudaStreamSynchronize(s));                                            // &n_ref / &n_model are stack addresses
  { const fl_intrinsics_t none = {0.f, 0.f, 0.f, 0.f};
    const int nl = fl_launch_icp(h->icp, prm, nullptr, nullptr, 0, 0, none, h->d_results, h->d_icp_ticket, nullptr, h->n_sm, s);
    if (nl < 0) { fl_set_error("ICP kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); return FL_ERR_CUDA; }
    h->launches += nl; }
  FL_CUDA(cudaMemcpyAsync(h->h_results, h->d_results, sizeof(fl_icp_result_t), cudaMemcpyDeviceToHost, s));
  FL_CUDA(cudaStreamSynchronize(s));
  FL_CUDA(cudaGetLastError());
  *out = h->h_results[0];
  return FL_OK;
}

// rects inside the W x H frame?  (cv::Mat ROI throw, detection.cpp:43-44)
static bool rect_inside(const fl_rect_t& a, int W, int H) {
  return a.x >= 0 && a.y >= 0 && a.width >= 0 && a.height >= 0 && a.x + a.width <= W && a.y + a.height <= H;
}
// the reference depth frame on the device: a host frame is uploaded into d_ref_depth; NULL selects the frame the last
// host-input fl_match left in d_in_depth (Recognition hands the same image to match and to detection, obj_reco_lmicp.cpp:101, 188)
static int icp_ref_frame(fl_handle* h, const uint16_t* ref_depth, size_t ref_stride, int W, int H, const uint16_t** d_ref) {
  if (!ref_depth) {
    if (h->in_depth_W != W || h->in_depth_H != H) { fl_set_error("ref_depth == NULL needs a preceding fl_match of a %d x %d depth frame on this handle (have %d x %d)", W, H, h->in_depth_W, h->in_depth_H); return FL_ERR_STATE; }
    *d_ref = h->d_in_depth;
    return FL_OK;
  }
  if (ref_stride < (size_t)W * 2) return FL_ERR_ARG;
  if ((size_t)W * H > h->ref_depth_cap) { cudaFree(h->d_ref_depth); TRY(dalloc(&h->d_ref_depth, (size_t)W * H)); h->ref_depth_cap = (size_t)W * H; }
  FL_CUDA(cudaMemcpy2DAsync(h->d_ref_depth, (size_t)W * 2, ref_depth, ref_stride, (size_t)W * 2, H, cudaMemcpyHostToDevice, h->stream));
  *d_ref = h->d_ref_depth;
  return FL_OK;
}
// h_hyps[0..n) are filled: upload them, run prepare + the ICP loop, fetch the results
static int icp_run_batch(fl_handle* h, const uint16_t* d_ref, int W, int H, fl_intrinsics_t K_ref, int n, fl_icp_params_t prm, fl_icp_result_t* out) {
  cudaStream_t s = h->stream;
  FL_CUDA(cudaMemcpyAsync(h->d_hyps, h->h_hyps, sizeof(fl_icp_hyp) * (size_t)n, cudaMemcpyHostToDevice, s));
  if (h->profile) cudaEventRecord(h->ev[0], s);
  unsigned long long* trace = nullptr;
  if (h->profile) {                                                             // per-hypothesis phase clock of the fused kernel (fl_debug_icp_trace)
    if (h->icp_trace_cap < n) { cudaFree(h->d_icp_trace); h->d_icp_trace = nullptr; TRY(dalloc(&h->d_icp_trace, (size_t)n * FL_ICP_TRACE_WORDS)); h->icp_trace_cap = n; }
    trace = h->d_icp_trace; h->icp_trace_n = n;
  }
  { const int nl = fl_launch_icp(h->icp, prm, h->d_hyps, d_ref, W, H, K_ref, h->d_results, h->d_icp_ticket, trace, h->n_sm, s);
    if (nl < 0) { fl_set_error("ICP kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); return FL_ERR_CUDA; }
    h->launches += nl; }
  if (h->profile) cudaEventRecord(h->ev[1], s);
  FL_CUDA(cudaMemcpyAsync(h->h_results, h->d_results, sizeof(fl_icp_result_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
  FL_CUDA(wait_stream(h));
  if (h->profile) cudaEventElapsedTime(&h->icp_ms, h->ev[0], h->ev[1]);
  FL_CUDA(cudaGetLastError());
  memcpy(out, h->h_results, sizeof(fl_icp_result_t) * (size_t)n);
  return FL_OK;
}
static void icp_fill_pose(fl_icp_hyp& hy, const float* r_match9, const float* t_match3, int i) {
  for (int k = 0; k < 9; ++k) hy.r_match[k] = r_match9 ? r_match9[9 * i + k] : (k % 4 == 0 ? 1.f : 0.f);
  for (int k = 0; k < 3; ++k) hy.t_match[k] = t_match3 ? t_match3[3 * i + k] : 0.f;
}

extern "C" int fl_detection_batch(fl_handle* h, const uint16_t* ref_depth, size_t ref_stride, int32_t W, int32_t H, fl_intrinsics_t K_ref,
                                  const uint16_t* const* model_depth, const size_t* model_stride, const fl_rect_t* rect_model,
                                  const fl_rect_t* rect_ref, const float* r_match9, const float* t_match3, const float* d_match, int32_t n,
                                  fl_icp_params_t prm, fl_icp_result_t* out) {
  (void)d_match;   // detection() receives d_match but its live branch (test_id == 2) never reads it (detection.cpp:147, 175-178)
  if (!h || !ref_depth || !model_depth || !model_stride || !rect_model || !rect_ref || !out || n < 0 || W <= 0 || H <= 0 || ref_stride < (size_t)W * 2) return FL_ERR_ARG;
  if (n == 0) return FL_OK;
  FL_CUDA(cudaSetDevice(h->p.device));
  cudaStream_t s = h->stream;
  int max_pts = 16;
  std::vector<int> st(n, FL_OK);
  for (int i = 0; i < n; ++i) {
    const fl_rect_t& a = rect_model[i]; const fl_rect_t& b = rect_ref[i];
    if (!(rect_inside(a, W, H) && rect_inside(b, W, H) && model_depth[i])) { st[i] = FL_ERR_ROI; continue; }
    max_pts = std::max(max_pts, std::max(a.width * a.height, b.width * b.height));
  }
  TRY(icp_reserve(h, n, max_pts));
  const int mp = h->icp.max_pts;
  const uint16_t* d_ref = nullptr;
  TRY(icp_ref_frame(h, ref_depth, ref_stride, W, H, &d_ref));
  for (int i = 0; i < n; ++i) {
    fl_icp_hyp& hy = h->h_hyps[i];
    hy.model_depth = h->d_model_crops + (size_t)i * mp;
    hy.rect_model = rect_model[i]; hy.rect_ref = rect_ref[i]; hy.status = st[i];
    icp_fill_pose(hy, r_match9, t_match3, i);
    if (st[i] != FL_OK) continue;
    const fl_rect_t& a = rect_model[i];
    uint16_t* dst = h->h_model_crops + (size_t)i * mp;
    for (int y = 0; y < a.height; ++y)
      memcpy(dst + (size_t)y * a.width, (const uint8_t*)model_depth[i] + (size_t)(a.y + y) * model_stride[i] + (size_t)a.x * 2, (size_t)a.width * 2);
  }
  // one copy for all crops (the staging block is contiguous): 256 separate copies cost more than the ICP itself
  FL_CUDA(cudaMemcpyAsync(h->d_model_crops, h->h_model_crops, (size_t)n * mp * 2, cudaMemcpyHostToDevice, s));
  return icp_run_batch(h, d_ref, W, H, K_ref, n, prm, out);
}

// The rendered template depth crops of the detector's templates, uploaded ONCE (AddObj time) and packed back to back.
extern "C" int fl_upload_model_depths(fl_handle* h, int32_t n_models, const uint16_t* const* model_depth, const size_t* model_stride,
                                      const fl_rect_t* rect_model, int32_t W, int32_t H) {
  if (!h || n_models < 0 || W <= 0 || H <= 0 || (n_models > 0 && (!model_depth || !model_stride || !rect_model))) return FL_ERR_ARG;
  FL_CUDA(cudaSetDevice(h->p.device));
  std::vector<size_t> off((size_t)n_models + 1, 0);
  for (int i = 0; i < n_models; ++i) {
    if (!model_depth[i] || model_stride[i] < (size_t)W * 2) return FL_ERR_ARG;
    if (!rect_inside(rect_model[i], W, H)) { fl_set_error("model rect %d lies outside the %d x %d image", i, W, H); return FL_ERR_ROI; }
    off[i + 1] = off[i] + (size_t)rect_model[i].width * rect_model[i].height;
  }
  std::vector<uint16_t> packed(std::max<size_t>(off[n_models], 1));
  for (int i = 0; i < n_models; ++i) {
    const fl_rect_t& a = rect_model[i];
    for (int y = 0; y < a.height; ++y)
      memcpy(&packed[off[i] + (size_t)y * a.width], (const uint8_t*)model_depth[i] + (size_t)(a.y + y) * model_stride[i] + (size_t)a.x * 2, (size_t)a.width * 2);
  }
  FL_CUDA(cudaStreamSynchronize(h->stream));                                    // an ICP batch that still reads the old set
  cudaFree(h->d_resident); h->d_resident = nullptr;
  TRY(dalloc(&h->d_resident, packed.size()));
  FL_CUDA(cudaMemcpy(h->d_resident, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  h->res_off.assign(off.begin(), off.end() - 1);
  h->res_rect.assign(rect_model, rect_model + n_models);
  h->res_W = W; h->res_H = H;
  // size the ICP workspace for the largest crop now (8 hypotheses; more grow it again): growing it later means cudaFree / cudaMalloc,
  // which synchronise the whole device - harmless for one handle, a stall for the other frames in flight on this GPU, and with a
  // peer-memory exchange in flight on another handle a cross-rank stall until that exchange times out
  int max_pts = 16;
  for (int i = 0; i < n_models; ++i) max_pts = std::max(max_pts, rect_model[i].width * rect_model[i].height);
  TRY(icp_reserve(h, std::max(h->icp_hyp_cap, 8), max_pts));
  return FL_OK;
}

static int detection_batch_resident(fl_handle* h, const uint16_t* ref_depth, size_t ref_stride, const uint16_t* d_ref_given, int32_t W, int32_t H,
                                    fl_intrinsics_t K_ref, const int32_t* model_index, const fl_rect_t* rect_ref, const float* r_match9,
                                    const float* t_match3, int32_t n, fl_icp_params_t prm, fl_icp_result_t* out) {
  if (!h || !model_index || !rect_ref || !out || n < 0 || W <= 0 || H <= 0) return FL_ERR_ARG;
  if (n == 0) return FL_OK;
  if (!h->d_resident || h->res_W != W || h->res_H != H) { fl_set_error("no model depth crops uploaded for %d x %d frames (fl_upload_model_depths)", W, H); return FL_ERR_STATE; }
  FL_CUDA(cudaSetDevice(h->p.device));
  int max_pts = 16;
  std::vector<int> st(n, FL_OK);
  for (int i = 0; i < n; ++i) {
    if (model_index[i] < 0 || (size_t)model_index[i] >= h->res_rect.size()) return FL_ERR_ARG;
    const fl_rect_t& a = h->res_rect[model_index[i]]; const fl_rect_t& b = rect_ref[i];
    if (!rect_inside(b, W, H)) { st[i] = FL_ERR_ROI; continue; }
    max_pts = std::max(max_pts, std::max(a.width * a.height, b.width * b.height));
  }
  TRY(icp_reserve(h, n, max_pts));
  const uint16_t* d_ref = d_ref_given;
  if (!d_ref) TRY(icp_ref_frame(h, ref_depth, ref_stride, W, H, &d_ref));
  for (int i = 0; i < n; ++i) {
    fl_icp_hyp& hy = h->h_hyps[i];
    hy.model_depth = h->d_resident + h->res_off[model_index[i]];
    hy.rect_model = h->res_rect[model_index[i]]; hy.rect_ref = rect_ref[i]; hy.status = st[i];
    icp_fill_pose(hy, r_match9, t_match3, i);
  }
  return icp_run_batch(h, d_ref, W, H, K_ref, n, prm, out);
}

extern "C" int fl_detection_batch_resident(fl_handle* h, const uint16_t* ref_depth, size_t ref_stride, int32_t W, int32_t H, fl_intrinsics_t K_ref,
                                           const int32_t* model_index, const fl_rect_t* rect_ref, const float* r_match9, const float* t_match3,
                                           int32_t n, fl_icp_params_t prm, fl_icp_result_t* out) {
  return detection_batch_resident(h, ref_depth, ref_stride, nullptr, W, H, K_ref, model_index, rect_ref, r_match9, t_match3, n, prm, out);
}
// the same with the reference depth frame ALREADY ON THE DEVICE (dense rows of W u16), e.g. the frame fl_match_device* was given
extern "C" int fl_detection_batch_resident_device(fl_handle* h, const void* d_ref_depth, int32_t W, int32_t H, fl_intrinsics_t K_ref,
                                                  const int32_t* model_index, const fl_rect_t* rect_ref, const float* r_match9,
                                                  const float* t_match3, int32_t n, fl_icp_params_t prm, fl_icp_result_t* out) {
  if (!d_ref_depth) return FL_ERR_ARG;
  return detection_batch_resident(h, nullptr, 0, static_cast<const uint16_t*>(d_ref_depth), W, H, K_ref, model_index, rect_ref, r_match9, t_match3, n, prm, out);
}

extern "C" int fl_detection(fl_handle* h, const uint16_t* model_depth, size_t model_stride, const uint16_t* ref_depth, size_t ref_stride,
                            int32_t W, int32_t H, fl_intrinsics_t K_ref, fl_rect_t rect_model, fl_rect_t rect_ref, int32_t icp_it_thr,
                            float dist_mean_thr, float dist_diff_thr, const float r_match[9], const float t_match[3], float d_match,
                            float T_final[3], float R_final[9]) {
  if (!T_final || !R_final) return FL_ERR_ARG;
  fl_icp_params_t prm = {icp_it_thr, dist_mean_thr, dist_diff_thr};
  fl_icp_result_t r;
  int rc = fl_detection_batch(h, ref_depth, ref_stride, W, H, K_ref, &model_depth, &model_stride, &rect_model, &rect_ref, r_match, t_match,
                              &d_match, 1, prm, &r);
  if (rc != FL_OK) return rc;
  if (r.status != FL_OK) return r.status;
  memcpy(T_final, r.T, 12); memcpy(R_final, r.R, 36);
  return FL_OK;
}

extern "C" int fl_nms_ex(fl_handle* h, const float* t3, const int32_t* n_model_pts, const float* icp_dist, int32_t n, float th, int32_t* out_idx,
                         uint8_t* absorbed) {
  if (!h || n < 0 || (n > 0 && (!t3 || !n_model_pts || !icp_dist || !out_idx))) return FL_ERR_ARG;
  if (n == 0) return 0;
  FL_CUDA(cudaSetDevice(h->p.device));
  cudaStream_t s = h->stream;
  // handle-owned workspace, grown on demand: a cudaMalloc / cudaFree pair per call would synchronise the whole device - with other
  // handles' frames in flight on the same GPU that serialises them, and a kernel of theirs that is waiting for a peer rank's
  // block turns it into a cross-rank stall until the exchange times out
  if (h->nms_cap < n) {
    cudaFree(h->d_nms); h->d_nms = nullptr; h->nms_cap = 0; h->blocking_wait = false; h->ev_block = nullptr; h->d_train = nullptr;
    const int cap = std::max(64, 2 * n);
    TRY(dalloc(&h->d_nms, (size_t)cap * 7 + 2));
    h->nms_cap = cap;
  }
  float* d_t = reinterpret_cast<float*>(h->d_nms);
  int32_t* d_n = h->d_nms + (size_t)h->nms_cap * 3;
  float* d_d = reinterpret_cast<float*>(h->d_nms + (size_t)h->nms_cap * 4);
  int32_t* d_o = h->d_nms + (size_t)h->nms_cap * 5;                        // [n out_idx][n absorbed bytes (n words reserved)][count]
  FL_CUDA(cudaMemcpyAsync(d_t, t3, (size_t)n * 12, cudaMemcpyHostToDevice, s));
  FL_CUDA(cudaMemcpyAsync(d_n, n_model_pts, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  FL_CUDA(cudaMemcpyAsync(d_d, icp_dist, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  int32_t* d_cnt = d_o + 2 * n + 1;
  fl_launch_nms(d_t, d_n, d_d, n, th, d_o, d_cnt, s); ++h->launches;       // the absorbed flags live right after out_idx (n bytes)
  int32_t cnt = 0;
  FL_CUDA(cudaMemcpyAsync(out_idx, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  if (absorbed) FL_CUDA(cudaMemcpyAsync(absorbed, d_o + n, (size_t)n, cudaMemcpyDeviceToHost, s));
  FL_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s));
  FL_CUDA(wait_stream(h));
  return cnt;
}

extern "C" int fl_nms(fl_handle* h, const float* t3, const int32_t* n_model_pts, const float* icp_dist, int32_t n, float th, int32_t* out_idx) {
  return fl_nms_ex(h, t3, n_model_pts, icp_dist, n, th, out_idx, nullptr);
}
