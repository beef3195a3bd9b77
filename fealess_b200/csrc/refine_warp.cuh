// Local refinement of one candidate by ONE WARP (similarityLocal + the per-level update of matchClass, reference
// linemod/linemod.cpp:1226-1300, 1509-1573), used by the fused tail of the staged similarity kernel: the warps of a CTA
// refine the candidates their CTA produced without another kernel launch in between.
//
// Lane = (patch row 0..15, half 0..1) owns 8 consecutive cells of the 16 x 16 patch: per feature 3 aligned word loads from
// the level's linear memories (L2 resident, written by the front end of the same frame) + 2 funnel shifts + 2 packed-u8
// adds.  Feature window origins are resolved 32 at a time (one lane per feature) and broadcast with shuffles; the loads
// of 8 features are in flight together.  Same arithmetic as k_refine_level (similarity.cu), which remains the path of the
// L1/L2-fed baseline kernel.
#pragma once

// (x, y) are the candidate's coordinates at level + 1 on entry and at `level` on return.  Returns false when the candidate
// is dropped (similarity < threshold, :1570-1572).  All 32 lanes must call it with identical arguments.
__device__ __forceinline__ bool fl_refine_candidate_warp(const fl_tdb& db, const fl_level_geom& g, int level, const uint8_t* __restrict__ lm_level,
                                                         float threshold, int t, int& x_io, int& y_io, float& sim_out) {
  const int lane = threadIdx.x & 31;
  const int T = g.T, border = 8 * T;
  const fl_template_hdr_t* hdr = db.hdr + ((size_t)t * db.L + level) * db.M;
  int x = x_io * 2 + 1, y = y_io * 2 + 1;                          // :1525-1534
  x = max(x, border); y = max(y, border);
  x = min(x, g.W - hdr[0].width - border);
  y = min(y, g.H - hdr[0].height - border);
  const int ox = (x / T - 8) * T, oy = (y / T - 8) * T;            // :1240-1241 (C division truncates towards zero)
  const int delta = (oy / T) * g.Wd + ox / T;
  const int row = lane >> 1, half = lane & 1;
  const uint32_t lane_off = (uint32_t)(row * g.Wd + half * 8);
  uint32_t tot[4] = {0u, 0u, 0u, 0u};                              // u16 pairs: cells (0,1) (2,3) (4,5) (6,7) of this lane
  int nf = 0;
  for (int m = 0; m < db.M; ++m) {
    const int fb = hdr[m].feature_begin, fc = hdr[m].feature_count;
    nf += fc;
    uint32_t a0 = 0, a1 = 0;                                       // packed u8 sums of one modality (<= 63 x 4)
    for (int k0 = 0; k0 < fc; k0 += 32) {
      const int k = k0 + lane;
      uint32_t o = FL_SKIP;
      if (k < fc) {
        const fl_pfeat p = db.pfeat[fb + k];
        const int fx = p.x + ox, fy = p.y + oy;
        if (fx >= 0 && fy >= 0 && fx < g.W && fy < g.H) {          // :1257
          if (p.lm_off == FL_SKIP)                                 // outside the image unshifted, inside when shifted
            o = (uint32_t)((size_t)db.feat[fb + k].label * g.label_stride +
                           (size_t)((p.y % T + T) % T * T + (p.x % T + T) % T) * g.cells + (size_t)(fy / T) * g.Wd + fx / T);
          else
            o = (uint32_t)((int)p.lm_off + delta);                 // shifting by a multiple of T moves only the cell index
          o += (uint32_t)((size_t)m * g.mod_stride);
        }
      }
      const int nk = min(32, fc - k0);
#pragma unroll 8
      for (int j = 0; j < nk; ++j) {
        const uint32_t oj = __shfl_sync(0xffffffffu, o, j);
        const bool valid = oj != FL_SKIP;
        const uint32_t a = (valid ? oj : 0u) + lane_off;           // skipped features read a valid dummy address and add 0
        const uint32_t* w = reinterpret_cast<const uint32_t*>(lm_level) + (a >> 2);
        const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
        const uint32_t sh = (a & 3u) * 8u;
        a0 += valid ? __funnelshift_r(w0, w1, sh) : 0u;
        a1 += valid ? __funnelshift_r(w1, w2, sh) : 0u;
      }
    }
    tot[0] += (a0 & 0xFFu) | ((a0 & 0xFF00u) << 8);
    tot[1] += ((a0 >> 16) & 0xFFu) | ((a0 >> 24) << 16);
    tot[2] += (a1 & 0xFFu) | ((a1 & 0xFF00u) << 8);
    tot[3] += ((a1 >> 16) & 0xFFu) | ((a1 >> 24) << 16);
  }
  // first maximum in row-major order: key = score << 8 | (255 - index); all-zero patch -> best 0 at (-1,-1) (:1547-1562)
  uint32_t key = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t sc = (tot[k >> 1] >> (16 * (k & 1))) & 0xFFFFu;
    const uint32_t idx = (uint32_t)(row * 16 + half * 8 + k);
    const uint32_t kk = (sc << 8) | (255u - idx);
    if (sc > 0 && kk > key) key = kk;
  }
  key = __reduce_max_sync(0xffffffffu, key);
  const int best = (int)(key >> 8);
  int br = -1, bc = -1;
  if (best > 0) { const int idx = 255 - (int)(key & 255u); br = idx >> 4; bc = idx & 15; }
  const int off = T / 2 + (T % 2 - 1);
  x_io = (x / T - 8 + bc) * T + off;                               // :1564-1566
  y_io = (y / T - 8 + br) * T + off;
  sim_out = __fdiv_rn(__fmul_rn((float)best, 100.f), (float)(4 * nf));
  return !(sim_out < threshold);
}

// a coarsest-level candidate of handle-local template t: refine it up the pyramid and, if it survives every level, append the
// final match to the global list (one returning atomic per surviving candidate)
__device__ __noinline__ void fl_refine_and_emit_warp(const fl_tdb& db, const fl_refine_args& ra, float threshold, int t, int cls, int x, int y,
                                                     float sim, fl_match_t* __restrict__ cand, int cap, int* __restrict__ d_count) {
  bool alive = true;
  for (int l = ra.n_levels - 2; l >= 0 && alive; --l) alive = fl_refine_candidate_warp(db, ra.g[l], l, ra.lm[l], threshold, t, x, y, sim);
  if (!alive) return;
  if ((threadIdx.x & 31) == 0) {
    const int slot = atomicAdd(d_count, 1);
    if (slot < cap) {
      fl_match_t mt;
      mt.x = x; mt.y = y; mt.similarity = sim; mt.class_idx = cls; mt.template_id = db.tid_of[t];
      cand[slot] = mt;
    }
  }
}
