// fl_group: template-sharded matching over N handles of ONE process (one handle per GPU, or several on one GPU), for hosts that
// have no torch / NCCL / Python around - the reference's caller is a C++ class (CadReco/obj_reco_lmicp.cpp:86-204).
//
// The group owns N handles, N exchange buffers (one on each handle's device) and N candidate blocks.  Peer access is enabled
// between every pair of devices, so under unified addressing every handle's exchange buffer is reachable from every device
// with the pointer cudaMalloc returned: the kernel-side exchange (push into every peer's buffer, publish the epoch, wait for the
// peers' epochs, sort the union - the last CTA of k_refine_sort<1>) is the same one the one-process-per-GPU path uses with torch
// symmetric memory.  Templates are dealt round-robin by global index (gid % N) and carry their global per-class ids, so the
// merged list equals a single handle's.  One fl_group_match = the frame copied to every device from the caller's
// buffers (fl_match_shard_exchange_async), ONE enqueue per handle, one wait per handle, the merged list read from handle 0.
#include "fl_internal.cuh"
#include <algorithm>
#include <vector>

struct fl_group {
  int n, cap;
  std::vector<fl_handle*> h;
  std::vector<int> device;
  std::vector<uint8_t*> xbuf;          // exchange buffer of handle i (on device[i])
  std::vector<fl_match_t*> block;      // candidate block [count | cap records] of handle i
  std::vector<void*> peers;            // xbuf as void* (the same array for every rank: unified addressing)
  int max_w, max_h;
  uint32_t epoch;
};

#define GCUDA(call)                                                                    \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      fl_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return FL_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

extern "C" int fl_group_destroy(fl_group* g) {
  if (!g) return FL_ERR_ARG;
  for (int i = 0; i < g->n; ++i) {
    if (i < (int)g->device.size()) cudaSetDevice(g->device[i]);
    if (i < (int)g->h.size() && g->h[i]) { fl_sync(g->h[i]); }
    if (i < (int)g->xbuf.size()) cudaFree(g->xbuf[i]);
    if (i < (int)g->block.size()) cudaFree(g->block[i]);
    if (i < (int)g->h.size() && g->h[i]) fl_destroy(g->h[i]);
  }
  delete g;
  return FL_OK;
}

extern "C" int fl_group_create(const fl_params_t* params, const int32_t* devices, int32_t n, int32_t exchange_capacity, fl_group** out) {
  if (!params || !devices || !out || n < 1 || n > FL_XCHG_MAX_WORLD || exchange_capacity < 1) return FL_ERR_ARG;
  *out = nullptr;
  fl_group* g = new fl_group;
  g->n = n; g->cap = exchange_capacity; g->epoch = 0; g->max_w = params->max_width; g->max_h = params->max_height;
  g->h.assign(n, nullptr); g->device.assign(devices, devices + n); g->xbuf.assign(n, nullptr); g->block.assign(n, nullptr);
  g->peers.assign(n, nullptr);
  auto fail = [&](int rc) { fl_group_destroy(g); return rc; };
  // peer access between every pair of distinct devices (already enabled is fine)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      if (devices[i] == devices[j]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can) {
        fl_set_error("device %d cannot access device %d directly (no NVLink / PCIe peer path)", devices[i], devices[j]);
        return fail(FL_ERR_CUDA);
      }
      if (cudaSetDevice(devices[i]) != cudaSuccess) return fail(FL_ERR_CUDA);
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { fl_set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[i], devices[j], cudaGetErrorString(e)); return fail(FL_ERR_CUDA); }
      cudaGetLastError();
    }
  const size_t xbytes = fl_exchange_buffer_bytes(n, exchange_capacity);
  for (int i = 0; i < n; ++i) {
    fl_params_t p = *params;
    p.device = devices[i];
    p.max_candidates = std::max(p.max_candidates, n * (exchange_capacity + 1) + 16);
    int rc = fl_create(&p, &g->h[i]);
    if (rc != FL_OK) return fail(rc);
    if (cudaSetDevice(devices[i]) != cudaSuccess) return fail(FL_ERR_CUDA);
    if (cudaMalloc(&g->xbuf[i], xbytes) != cudaSuccess || cudaMemset(g->xbuf[i], 0, xbytes) != cudaSuccess ||
        cudaMalloc(&g->block[i], sizeof(fl_match_t) * ((size_t)exchange_capacity + 1)) != cudaSuccess ||
        cudaMemset(g->block[i], 0, sizeof(fl_match_t) * ((size_t)exchange_capacity + 1)) != cudaSuccess) {
      fl_set_error("fl_group_create: device allocation failed on device %d: %s", devices[i], cudaGetErrorString(cudaGetLastError()));
      return fail(FL_ERR_CUDA);
    }
    g->peers[i] = g->xbuf[i];
  }
  for (int i = 0; i < n; ++i) { cudaSetDevice(devices[i]); cudaDeviceSynchronize(); }
  *out = g;
  return FL_OK;
}

extern "C" fl_handle* fl_group_handle(fl_group* g, int32_t i) { return (g && i >= 0 && i < g->n) ? g->h[i] : nullptr; }
extern "C" int32_t fl_group_size(const fl_group* g) { return g ? g->n : 0; }

// the whole set (layout of fl_upload_templates): template gid goes to handle gid % n with its global per-class id
extern "C" int fl_group_upload_templates(fl_group* g, int32_t n_templates, const fl_template_hdr_t* headers, const fl_feature_t* features, int32_t n_features,
                                         const int32_t* class_of, const float* pose13) {
  if (!g || n_templates < 0 || (n_templates > 0 && (!headers || !features || !class_of))) return FL_ERR_ARG;
  const int LM = fl_entries_per_template(g->h[0]);               // n_levels * n_modalities headers per template
  if (LM <= 0) return FL_ERR_STATE;
  std::vector<int32_t> first;                                   // first global index of every class
  for (int t = 0; t < n_templates; ++t) {
    const int c = class_of[t];
    if (c < 0) return FL_ERR_ARG;
    if ((int)first.size() <= c) first.resize((size_t)c + 1, -1);
    if (first[c] < 0) first[c] = t;
  }
  for (int r = 0; r < g->n; ++r) {
    std::vector<fl_template_hdr_t> hs;
    std::vector<fl_feature_t> fs;
    std::vector<int32_t> co, ids;
    std::vector<float> pose;
    for (int t = r; t < n_templates; t += g->n) {
      for (int e = 0; e < LM; ++e) {
        fl_template_hdr_t hd = headers[(size_t)t * LM + e];
        if (hd.feature_begin < 0 || hd.feature_count < 0 || hd.feature_begin + hd.feature_count > n_features) return FL_ERR_ARG;
        const int32_t b = (int32_t)fs.size();
        fs.insert(fs.end(), features + hd.feature_begin, features + hd.feature_begin + hd.feature_count);
        hd.feature_begin = b;
        hs.push_back(hd);
      }
      co.push_back(class_of[t]);
      ids.push_back(t - first[class_of[t]]);
      if (pose13) pose.insert(pose.end(), pose13 + (size_t)t * 13, pose13 + (size_t)t * 13 + 13);
    }
    const int nt = (int)co.size();
    fl_feature_t dummy = {0, 0, 0};
    int rc = fl_upload_templates(g->h[r], nt, nt ? hs.data() : nullptr, fs.empty() ? &dummy : fs.data(), (int32_t)fs.size(), nt ? co.data() : nullptr,
                                 pose13 && nt ? pose.data() : nullptr);
    if (rc != FL_OK) return rc;
    if (nt) { rc = fl_set_template_ids(g->h[r], ids.data()); if (rc != FL_OK) return rc; }
  }
  return FL_OK;
}

// Detector::match over the whole (sharded) set.  Host buffers in, merged host list out; every handle ends up with the same list.
extern "C" int fl_group_match(fl_group* g, const uint8_t* bgr, size_t bgr_stride, const uint16_t* depth, size_t depth_stride, int32_t W, int32_t H,
                              float threshold, const int32_t* class_filter, int32_t n_filter, fl_match_t* out, int32_t capacity, int32_t* count) {
  if (!g || !count || (!bgr && !depth)) return FL_ERR_ARG;
  *count = 0;
  if (W <= 0 || H <= 0 || W > g->max_w || H > g->max_h) return FL_ERR_SIZE;
  if ((bgr && bgr_stride < (size_t)W * 3) || (depth && depth_stride < (size_t)W * 2)) return FL_ERR_SIZE;
  // (a second attempt only when a handle's single-launch front end timed out on an in-grid dependency during the first: that handle
  //  has switched itself to per-wave launches - fl_match_wait - and asks for the frame again; every rank has to take part in it)
  for (int attempt = 0; attempt < 2; ++attempt) {
    ++g->epoch;
    int first_err = FL_OK;
    bool switched = false;
    int launched = 0;
    for (int i = 0; i < g->n; ++i) {                            // enqueue on every handle; nothing waits until all are under way
      GCUDA(cudaSetDevice(g->device[i]));
      const int rc = fl_match_shard_exchange_async(g->h[i], bgr, bgr_stride, depth, depth_stride, W, H, threshold, class_filter, n_filter,
                                                   i, g->n, g->peers.data(), g->cap, g->block[i], g->epoch);
      if (rc != FL_OK) { first_err = rc; break; }                // (the ranks already launched report the missing peer instead of hanging)
      ++launched;
    }
    for (int i = 0; i < launched; ++i) {
      cudaSetDevice(g->device[i]);
      const int forced = fl_debug_option(g->h[i], FL_OPT_FE_FORCED_WAVES, 0);
      const int rc = fl_match_wait(g->h[i]);
      if (rc == FL_ERR_STATE && !forced && fl_debug_option(g->h[i], FL_OPT_FE_FORCED_WAVES, 0) == 1) switched = true;   // resubmit
      else if (rc != FL_OK && first_err == FL_OK) first_err = rc;                                                        // (incl. a peer that never arrived)
    }
    if (first_err != FL_OK) return first_err;
    if (!switched) break;
    if (attempt == 1) { fl_set_error("fl_group_match: the front end timed out twice"); return FL_ERR_STATE; }
  }
  GCUDA(cudaSetDevice(g->device[0]));
  return fl_match_fetch(g->h[0], out, capacity, count);
}
