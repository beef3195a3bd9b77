// fl_pipe: Detector::match over a STREAM of frames with several frames in flight on one GPU.
//
// One frame through a handle is three dependent launches (front end -> similarity -> refinement + sort) whose first and last have
// far fewer CTAs than the GPU has SM slots, followed by a host wait: a single handle leaves ~40 % of the SM time idle.
// The reference's caller loops over camera frames (CadReco/obj_reco_lmicp.cpp:86-204, test/linemod_acq.cpp:120-190) - the frames are
// independent, so the pipe owns `depth` handles on one device (each with the whole template set, its own stream, linear memories,
// candidate and result blocks), deals the submitted frames round-robin and hands the match lists back in submission order.  The GPU
// runs the front end of frame i+1 in the SM slots the similarity kernel of frame i leaves free and under its tail; the upload of
// frame i+1 and the result copy of frame i-1 ride on the copy engines meanwhile.  Nothing here synchronises more than one stream.
// fl_pipe_set_exchange turns the pipe into ONE RANK of a template-sharded detector: every slot then runs the fused local match +
// peer-memory exchange (fl_match_shard_exchange_*) with its own exchange buffers and epoch sequence, so a rank's whole per-frame host
// loop - upload, enqueue, wait, fetch - lives in fl_pipe_match_batch.
#include "fl_internal.cuh"
#include <vector>

struct fl_pipe {
  int depth, device;
  std::vector<fl_handle*> h;
  long long submitted, collected;
  // template-sharded mode (fl_pipe_set_exchange): this pipe is rank `rank` of `world`; slot i owns exchange buffers peers[i * world ..]
  // and the candidate block blocks[i]; its frames carry the epochs 1, 2, 3 ... of that slot
  int rank, world, xcap;
  std::vector<void*> peers;
  std::vector<fl_match_t*> blocks;
  std::vector<uint32_t> epoch;
};

extern "C" int fl_pipe_destroy(fl_pipe* p) {
  if (!p) return FL_ERR_ARG;
  for (fl_handle* h : p->h) if (h) { fl_sync(h); fl_destroy(h); }
  delete p;
  return FL_OK;
}

extern "C" int fl_pipe_create(const fl_params_t* params, int32_t depth, fl_pipe** out) {
  if (!params || !out || depth < 1 || depth > FL_PIPE_MAX_DEPTH) return FL_ERR_ARG;
  *out = nullptr;
  fl_pipe* p = new fl_pipe;
  p->depth = depth; p->device = params->device; p->submitted = p->collected = 0; p->rank = 0; p->world = 1; p->xcap = 0;
  p->h.assign(depth, nullptr);
  for (int i = 0; i < depth; ++i) {
    const int rc = fl_create(params, &p->h[i]);
    if (rc != FL_OK) { fl_pipe_destroy(p); return rc; }
  }
  *out = p;
  return FL_OK;
}

extern "C" int32_t fl_pipe_depth(const fl_pipe* p) { return p ? p->depth : 0; }
extern "C" int32_t fl_pipe_in_flight(const fl_pipe* p) { return p ? (int32_t)(p->submitted - p->collected) : 0; }
extern "C" fl_handle* fl_pipe_handle(fl_pipe* p, int32_t i) { return (p && i >= 0 && i < p->depth) ? p->h[i] : nullptr; }

extern "C" int fl_pipe_upload_templates(fl_pipe* p, int32_t n_templates, const fl_template_hdr_t* headers, const fl_feature_t* features, int32_t n_features,
                                        const int32_t* class_of, const float* pose13) {
  if (!p) return FL_ERR_ARG;
  if (p->submitted != p->collected) { fl_set_error("fl_pipe_upload_templates: %lld frame(s) still in flight", p->submitted - p->collected); return FL_ERR_STATE; }
  for (fl_handle* h : p->h) {
    const int rc = fl_upload_templates(h, n_templates, headers, features, n_features, class_of, pose13);
    if (rc != FL_OK) return rc;
  }
  return FL_OK;
}

// Template-sharded pipe: every rank (process or fl_pipe) of a `world`-way sharded detector calls this once, after uploading its shard
// (fl_pipe_upload_templates) and its global template ids (fl_set_template_ids on fl_pipe_handle(p, i)).  peer_buffers: depth x world
// pointers - entry [i * world + r] is rank r's exchange buffer OF SLOT i as mapped into this process (fl_exchange_buffer_bytes each,
// zeroed once); local_blocks: depth device blocks of exchange_capacity + 1 records.  From then on fl_pipe_submit runs
// fl_match_shard_exchange_(device_)async: all ranks have to submit the same sequence of frames.
extern "C" int fl_pipe_set_exchange(fl_pipe* p, int32_t rank, int32_t world, int32_t exchange_capacity, void* const* peer_buffers, fl_match_t* const* local_blocks) {
  if (!p || world < 1 || rank < 0 || rank >= world || exchange_capacity < 1 || !peer_buffers || !local_blocks) return FL_ERR_ARG;
  if (p->submitted != p->collected) { fl_set_error("fl_pipe_set_exchange: %lld frame(s) still in flight", p->submitted - p->collected); return FL_ERR_STATE; }
  for (int i = 0; i < p->depth; ++i) {
    if (!local_blocks[i]) return FL_ERR_ARG;
    for (int r = 0; r < world; ++r) if (!peer_buffers[(size_t)i * world + r]) return FL_ERR_ARG;
  }
  p->rank = rank; p->world = world; p->xcap = exchange_capacity;
  p->peers.assign(peer_buffers, peer_buffers + (size_t)p->depth * world);
  p->blocks.assign(local_blocks, local_blocks + p->depth);
  p->epoch.assign(p->depth, 0u);
  return FL_OK;
}

extern "C" int fl_pipe_submit(fl_pipe* p, const void* bgr, size_t bgr_stride, const void* depth, size_t depth_stride, int32_t W, int32_t H, float threshold,
                              const int32_t* class_filter, int32_t n_filter, int32_t on_device) {
  if (!p || (!bgr && !depth)) return FL_ERR_ARG;
  if (p->submitted - p->collected >= p->depth) { fl_set_error("fl_pipe_submit: %d frames in flight already - collect one first", p->depth); return FL_ERR_STATE; }
  const int slot = (int)(p->submitted % p->depth);
  fl_handle* h = p->h[slot];
  if (on_device && ((bgr && bgr_stride != (size_t)W * 3) || (depth && depth_stride != (size_t)W * 2))) { fl_set_error("fl_pipe_submit: device frames must have dense rows"); return FL_ERR_SIZE; }
  int rc;
  if (p->world > 1) {                                           // template-sharded: local match + peer-memory exchange + merge, one enqueue
    void* const* peers = p->peers.data() + (size_t)slot * p->world;
    const uint32_t e = p->epoch[slot] + 1;
    rc = on_device ? fl_match_shard_exchange_device_async(h, bgr, depth, W, H, threshold, class_filter, n_filter, p->rank, p->world, peers, p->xcap, p->blocks[slot], e)
                   : fl_match_shard_exchange_async(h, static_cast<const uint8_t*>(bgr), bgr_stride, static_cast<const uint16_t*>(depth), depth_stride, W, H, threshold,
                                                   class_filter, n_filter, p->rank, p->world, peers, p->xcap, p->blocks[slot], e);
    if (rc == FL_OK) p->epoch[slot] = e;
  } else if (on_device) {
    rc = fl_match_device_async(h, bgr, depth, W, H, nullptr, threshold, class_filter, n_filter);
  } else {
    rc = fl_match_async(h, static_cast<const uint8_t*>(bgr), bgr_stride, static_cast<const uint16_t*>(depth), depth_stride, W, H, nullptr, threshold, class_filter, n_filter);
  }
  if (rc != FL_OK) return rc;
  ++p->submitted;
  return FL_OK;
}

extern "C" int fl_pipe_collect(fl_pipe* p, fl_match_t* out, int32_t capacity, int32_t* count) {
  if (!p || !count) return FL_ERR_ARG;
  *count = 0;
  if (p->submitted == p->collected) { fl_set_error("fl_pipe_collect: no frame in flight"); return FL_ERR_STATE; }
  fl_handle* h = p->h[p->collected % p->depth];
  ++p->collected;                                               // the slot is free again whatever the frame's status
  const int rc = fl_match_wait(h);
  if (rc != FL_OK) return rc;
  return fl_match_fetch(h, out, capacity, count);
}

// n frames of one geometry through the pipe: frame f's list goes to out + f * capacity_per_frame, its length (before truncation)
// to counts[f].  Returns the first error; FL_ERR_CAPACITY if some list did not fit (the others are complete).
extern "C" int fl_pipe_match_batch(fl_pipe* p, int32_t n_frames, const void* const* bgr, size_t bgr_stride, const void* const* depth, size_t depth_stride,
                                   int32_t W, int32_t H, float threshold, const int32_t* class_filter, int32_t n_filter, int32_t on_device,
                                   fl_match_t* out, int32_t capacity_per_frame, int32_t* counts) {
  if (!p || n_frames < 0 || (!bgr && !depth) || !counts || capacity_per_frame < 0 || (capacity_per_frame > 0 && !out)) return FL_ERR_ARG;
  if (p->submitted != p->collected) { fl_set_error("fl_pipe_match_batch: %lld frame(s) still in flight", p->submitted - p->collected); return FL_ERR_STATE; }
  int status = FL_OK, next_out = 0;
  auto collect = [&]() -> int {
    const int rc = fl_pipe_collect(p, out + (size_t)next_out * capacity_per_frame, capacity_per_frame, counts + next_out);
    ++next_out;
    if (rc == FL_ERR_CAPACITY) { if (status == FL_OK) status = rc; return FL_OK; }
    return rc;
  };
  for (int f = 0; f < n_frames; ++f) {
    if (fl_pipe_in_flight(p) == p->depth) { const int rc = collect(); if (rc != FL_OK) { while (fl_pipe_in_flight(p)) { int32_t c; fl_pipe_collect(p, nullptr, 0, &c); } return rc; } }
    const int rc = fl_pipe_submit(p, bgr ? bgr[f] : nullptr, bgr_stride, depth ? depth[f] : nullptr, depth_stride, W, H, threshold, class_filter, n_filter, on_device);
    if (rc != FL_OK) { while (fl_pipe_in_flight(p)) { int32_t c; fl_pipe_collect(p, nullptr, 0, &c); } return rc; }
  }
  while (fl_pipe_in_flight(p)) { const int rc = collect(); if (rc != FL_OK) { while (fl_pipe_in_flight(p)) { int32_t c; fl_pipe_collect(p, nullptr, 0, &c); } return rc; } }
  return status;
}
