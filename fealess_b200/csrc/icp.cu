// ICP pose refinement on sm_100a: depth back-projection, paired-crop extraction, exact nearest neighbours on a uniform
// grid, and the point-to-point (Kabsch/SVD) iteration of the reference, one CTA per pose hypothesis, all iterations
// inside one kernel launch.
//
// Reference semantics (paths relative to /root/reference):
//   cup_d2pc::depthTo3d ICP/depth_to_3d.cpp:99-137, 244-260      scale_mat_vec3f / is_vec3f_valid / matToVec ICP/common.cpp:261-266, 382-425
//   detection ICP/detection.cpp:11-254 (test_id == 2)             getMean / transformPoints / copyPoints / getL2distClouds ICP/ICP.cpp:8-111
//   PointsCorresponding (KD-tree 1-NN) ICP/ICP.cpp:193-279        icpCloudToCloud_Ex ICP/ICP.cpp:617-809       NMS ICP/NMS.cpp:6-39
//
// Numerics: the reference accumulates centroids, the (uncentred) covariance and the mean distance SEQUENTIALLY in fp32.
// Its rotation moves by up to ~2e-4 rad per iteration if those sums are reordered or widened (measured), which
// is more than the 1e-4 rad parity tolerance, so the sums are reproduced as ordered fp32 chains: one warp lane per
// accumulator walks the (order-preserving) correspondence list.  Everything that is order-free (back-projection,
// nearest-neighbour search, transforms, distances, compaction scans) runs on all threads of the CTA.
// Compiled with -fmad=false so fp32 products and sums round separately, as in the reference build.
#include "fl_internal.cuh"
#include <float.h>
#include <math.h>

#define ICP_THREADS 256
#define ICP_WARPS (ICP_THREADS / 32)

__device__ __forceinline__ bool pt_valid(float z) { return z <= 900.0f; }   // is_vec3f_valid (NaN fails), common.cpp:261-266

// depth (u16 mm) -> point in mm exactly as depthTo3d (metres, fp32) followed by scale_mat_vec3f(., 1000)
__device__ __forceinline__ float3 backproject_mm(uint16_t dv, int u, int v, float inv_fx, float inv_fy, float cx, float cy) {
  float z = dv == 0 ? __int_as_float(0x7fc00000) : __fmul_rn((float)dv, (float)(1 / 1000.0));
  float xc = __fmul_rn(__fsub_rn((float)u, cx), inv_fx);
  float yc = __fmul_rn(__fsub_rn((float)v, cy), inv_fy);
  return make_float3(__fmul_rn(__fmul_rn(xc, z), 1000.0f), __fmul_rn(__fmul_rn(yc, z), 1000.0f), __fmul_rn(z, 1000.0f));
}

__global__ void __launch_bounds__(256) k_depth_to_3d(const uint16_t* __restrict__ depth, int W, int H, fl_intrinsics_t K,
                                                     float* __restrict__ out3) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W * H) return;
  int u = i % W, v = i / W;
  const float inv_fx = __fdiv_rn(1.0f, K.fx), inv_fy = __fdiv_rn(1.0f, K.fy);
  uint16_t dv = depth[i];
  float z = dv == 0 ? __int_as_float(0x7fc00000) : __fmul_rn((float)dv, (float)(1 / 1000.0));
  out3[3 * (size_t)i + 0] = __fmul_rn(__fmul_rn(__fsub_rn((float)u, K.cx), inv_fx), z);
  out3[3 * (size_t)i + 1] = __fmul_rn(__fmul_rn(__fsub_rn((float)v, K.cy), inv_fy), z);
  out3[3 * (size_t)i + 2] = z;
}
void fl_launch_depth_to_3d(const uint16_t* depth, int W, int H, fl_intrinsics_t K, float* out3, cudaStream_t s) {
  k_depth_to_3d<<<(W * H + 255) / 256, 256, 0, s>>>(depth, W, H, K, out3);
}

// block-wide exclusive scan of one int per thread (ICP_THREADS threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < ICP_WARPS ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    if (lane < ICP_WARPS) s_warp[lane] = w;
  }
  __syncthreads();
  int base = warp > 0 ? s_warp[warp - 1] : 0;
  *total = s_warp[ICP_WARPS - 1];
  __syncthreads();
  return base + inc - v;
}

// Ordered fp32 chains, staged through shared memory.  The reference sums centroids (getMean, ICP.cpp:8-25) and the
// uncentred covariance (covariance += m * r^T, ICP.cpp:731-735) sequentially in fp32, and the result is order sensitive
// (file header), so each of the 15 sums is one dependent chain walked by one lane of warp 0: lanes 0..2 / 3..5 sum one
// component of the model / reference list, lanes 6..14 one product m[a] * r[b].  What made the chains slow was not the
// adds but the global-memory latency in front of each batch of them, so the other warps stream the two point lists into
// a double-buffered shared-memory stage (coalesced) one chunk ahead of the chain lanes.
// Must be called by all ICP_THREADS threads; ends with a barrier; s_sum[0..14] holds the sums afterwards.
#define ICP_CH 384                                     // points per stage buffer
__device__ __forceinline__ void block_chain15(const float* __restrict__ pm, int n_m, const float* __restrict__ pr, int n_r, int n_cov,
                                              float (*s_st)[2][ICP_CH * 3], float* s_sum) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nm_all = max(n_m, n_cov), nr_all = max(n_r, n_cov);      // the covariance reads both lists
  const int n_chunks = (max(nm_all, nr_all) + ICP_CH - 1) / ICP_CH;
  auto load_chunk = [&](int c, int t, int nt) {
    const int base = c * ICP_CH, buf = c & 1;
    const int cm = max(0, min(ICP_CH, nm_all - base)) * 3, cr = max(0, min(ICP_CH, nr_all - base)) * 3;
    for (int i = t; i < cm; i += nt) s_st[buf][0][i] = pm[(size_t)base * 3 + i];
    for (int i = t; i < cr; i += nt) s_st[buf][1][i] = pr[(size_t)base * 3 + i];
  };
  const int n_lane = lane < 3 ? n_m : (lane < 6 ? n_r : n_cov);
  const int ia = lane < 3 ? lane : (lane < 6 ? lane - 3 : (lane - 6) / 3), ib = (lane - 6) % 3;
  float acc = 0.f;
  if (n_chunks > 0) load_chunk(0, tid, ICP_THREADS);
  __syncthreads();
  for (int c = 0; c < n_chunks; ++c) {
    if (warp == 0) {
      if (lane < 15) {
        const int cnt = max(0, min(ICP_CH, n_lane - c * ICP_CH));
        const float* a = s_st[c & 1][lane >= 3 && lane < 6 ? 1 : 0] + ia;
        if (lane < 6) {
#pragma unroll 8
          for (int i = 0; i < cnt; ++i) acc = __fadd_rn(acc, a[3 * i]);
        } else {
          const float* b = s_st[c & 1][1] + ib;
#pragma unroll 8
          for (int i = 0; i < cnt; ++i) acc = __fadd_rn(acc, __fmul_rn(a[3 * i], b[3 * i]));
        }
      }
    } else if (c + 1 < n_chunks) {
      load_chunk(c + 1, tid - 32, ICP_THREADS - 32);
    }
    __syncthreads();
  }
  if (warp == 0 && lane < 15) s_sum[lane] = acc;
  __syncthreads();
}

__device__ __forceinline__ void matvec3(const float* R, const float* v, float* o) {   // Matx33f * Vec3f: s = 0; s += a*b
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float s = __fadd_rn(0.f, __fmul_rn(R[3 * i], v[0]));
    s = __fadd_rn(s, __fmul_rn(R[3 * i + 1], v[1]));
    s = __fadd_rn(s, __fmul_rn(R[3 * i + 2], v[2]));
    o[i] = s;
  }
}
__device__ __forceinline__ void matmul3(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float s = __fadd_rn(0.f, __fmul_rn(A[3 * i], B[j]));
      s = __fadd_rn(s, __fmul_rn(A[3 * i + 1], B[3 + j]));
      s = __fadd_rn(s, __fmul_rn(A[3 * i + 2], B[6 + j]));
      C[3 * i + j] = s;
    }
}

// R = V * U^T of the SVD of a 3x3 fp32 matrix: one-sided (Hestenes) Jacobi on the transposed matrix, fp32 rotations with
// fp64 dot products - the scheme of OpenCV's JacobiSVD that cv::SVD::compute (ICP.cpp:741-742) runs for small matrices.
__device__ void svd3_rot(const float* cov, float* R) {
  float At[3][3], Vt[3][3];
  double Wd[3];
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) { At[i][k] = cov[3 * k + i]; Vt[i][k] = (i == k) ? 1.f : 0.f; }
  for (int i = 0; i < 3; ++i) { double sd = 0; for (int k = 0; k < 3; ++k) sd += (double)At[i][k] * At[i][k]; Wd[i] = sd; }
  const float eps = FLT_EPSILON * 2;
  for (int iter = 0; iter < 30; ++iter) {
    bool changed = false;
    for (int i = 0; i < 2; ++i)
      for (int j = i + 1; j < 3; ++j) {
        double a = Wd[i], p = 0, b = Wd[j];
        for (int k = 0; k < 3; ++k) p += (double)At[i][k] * At[j][k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        double beta = a - b, gamma = sqrt(p * p + beta * beta);
        float c, s;
        if (beta < 0) { double delta = (gamma - beta) * 0.5; s = (float)sqrt(delta / gamma); c = (float)(p / (gamma * s * 2)); }
        else { c = (float)sqrt((gamma + beta) / (gamma * 2)); s = (float)(p / (gamma * c * 2)); }
        a = b = 0;
        for (int k = 0; k < 3; ++k) {
          float t0 = __fadd_rn(__fmul_rn(c, At[i][k]), __fmul_rn(s, At[j][k]));
          float t1 = __fadd_rn(__fmul_rn(-s, At[i][k]), __fmul_rn(c, At[j][k]));
          At[i][k] = t0; At[j][k] = t1;
          a += (double)t0 * t0; b += (double)t1 * t1;
        }
        Wd[i] = a; Wd[j] = b;
        changed = true;
        for (int k = 0; k < 3; ++k) {
          float t0 = __fadd_rn(__fmul_rn(c, Vt[i][k]), __fmul_rn(s, Vt[j][k]));
          float t1 = __fadd_rn(__fmul_rn(-s, Vt[i][k]), __fmul_rn(c, Vt[j][k]));
          Vt[i][k] = t0; Vt[j][k] = t1;
        }
      }
    if (!changed) break;
  }
  float U[3][3];
  for (int i = 0; i < 3; ++i) {
    double sd = 0; for (int k = 0; k < 3; ++k) sd += (double)At[i][k] * At[i][k];
    sd = sqrt(sd);
    float s = (float)(sd > DBL_MIN ? 1 / sd : 0.);
    for (int k = 0; k < 3; ++k) U[i][k] = __fmul_rn(At[i][k], s);
  }
  // Mat(vt.t() * u.t()) (ICP.cpp:744) is a cv::gemm; for CV_32F its kernel accumulates the products in double and rounds once
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    double s = 0.;
    for (int i = 0; i < 3; ++i) s += (double)Vt[i][r] * (double)U[i][c];   // fp32 x fp32 is exact in fp64, so contraction cannot change it
    R[3 * r + c] = (float)s;
  }
}

// ------------------------------------------------------------------------------------------------
// K9 prepare: back-project the two crops, keep pixel pairs valid in both (order preserved), centroid shift (detection.cpp:28-44,
// 112-114, 162-206).  One CTA per hypothesis.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ICP_THREADS) k_icp_prepare(const uint16_t* __restrict__ ref_depth, int W, int H, fl_intrinsics_t Kr,
                                                             const fl_icp_hyp* __restrict__ hyps, fl_icp_ws ws, float* __restrict__ t_init_out) {
  __shared__ int s_warp[ICP_WARPS];
  __shared__ float s_c[6];
  __shared__ float s_st[2][2][ICP_CH * 3];
  __shared__ float s_sum[16];
  const int h = blockIdx.x;
  const fl_icp_hyp hy = hyps[h];
  float* pr = ws.pts_ref + (size_t)h * ws.max_pts * 3;
  float* pm = ws.pts_mod + (size_t)h * ws.max_pts * 3;
  if (hy.status != FL_OK) { if (threadIdx.x == 0) { ws.n_ref[h] = 0; ws.n_mod[h] = 0; } return; }
  const fl_rect_t rr = hy.rect_ref, rm = hy.rect_model;
  const int total = min(min(rr.width * rr.height, rm.width * rm.height), ws.max_pts);
  const float inv_fxr = __fdiv_rn(1.0f, Kr.fx), inv_fyr = __fdiv_rn(1.0f, Kr.fy);
  const float inv_fm = __fdiv_rn(1.0f, 608.0f);                    // initInternalMat: fx = fy = 608, c = (320, 240) (common.cpp:358)
  int n = 0;
  for (int c0 = 0; c0 < total; c0 += ICP_THREADS) {
    int i = c0 + threadIdx.x;
    float3 a = make_float3(0, 0, 0), b = make_float3(0, 0, 0);
    int keep = 0;
    if (i < total) {
      int ur = rr.x + i % rr.width, vr = rr.y + i / rr.width;
      int um = rm.x + i % rm.width, vm = rm.y + i / rm.width;
      a = backproject_mm(ref_depth[(size_t)vr * W + ur], ur, vr, inv_fxr, inv_fyr, Kr.cx, Kr.cy);
      b = backproject_mm(hy.model_depth[i], um, vm, inv_fm, inv_fm, 320.0f, 240.0f);
      keep = (pt_valid(a.z) && pt_valid(b.z)) ? 1 : 0;             // paired matToVec, common.cpp:382-405
    }
    int tot;
    int pos = n + block_excl_scan(keep, s_warp, &tot);
    if (keep) {
      pr[3 * pos] = a.x; pr[3 * pos + 1] = a.y; pr[3 * pos + 2] = a.z;
      pm[3 * pos] = b.x; pm[3 * pos + 1] = b.y; pm[3 * pos + 2] = b.z;
    }
    n += tot;
  }
  __syncthreads();
  block_chain15(pm, n, pr, n, 0, s_st, s_sum);                    // getMean x2 (detection.cpp:165-166): ordered fp32 chains
  if (threadIdx.x < 6) s_c[threadIdx.x] = n > 0 ? __fdiv_rn(s_sum[threadIdx.x], (float)n) : 0.f;
  __syncthreads();
  float tt[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) tt[k] = __fsub_rn(s_c[3 + k], s_c[k]);   // t_match_tmp = r_centroid - m_centroid (:177)
  for (int i = threadIdx.x; i < n; i += ICP_THREADS) {             // transformPoints(pts_mod, I, t_tmp) (:206)
    if (!pt_valid(pm[3 * i + 2])) continue;
#pragma unroll
    for (int k = 0; k < 3; ++k) pm[3 * i + k] = __fadd_rn(pm[3 * i + k], tt[k]);
  }
  if (threadIdx.x == 0) {
    ws.n_ref[h] = n; ws.n_mod[h] = n;
#pragma unroll
    for (int k = 0; k < 3; ++k) t_init_out[3 * h + k] = __fadd_rn(tt[k], hy.t_match[k]);   // t_init (:199)
  }
}

void fl_launch_icp_prepare(const uint16_t* ref_depth, int W, int H, fl_intrinsics_t K_ref, const fl_icp_hyp* hyps, fl_icp_ws ws,
                           float* t_init_out, cudaStream_t s) {
  if (ws.n_hyp > 0) k_icp_prepare<<<ws.n_hyp, ICP_THREADS, 0, s>>>(ref_depth, W, H, K_ref, hyps, ws, t_init_out);
}

// ------------------------------------------------------------------------------------------------
// K10/K11 the ICP loop.  Exact 1-NN (cvflann KDTreeSingleIndex, eps = 0) is replaced by an exact search over a 64x64
// uniform x/y grid of the reference cloud: rings of cells around the query are scanned until the best squared distance
// is no larger than the distance to the unvisited region, or that region lies beyond the accept radius 3*dist_mean
// (pairs farther than that are discarded anyway, ICP.cpp:266-273), so the accepted pairs equal the KD-tree's.
// Squared distance is ((dx*dx + dy*dy) + dz*dz) in fp32 like L2_Simple; ties go to the lower point index.
// ------------------------------------------------------------------------------------------------
struct grid_info { float minx, miny, inv_cell, cell; };

__device__ __forceinline__ int cell_of(float v, float mn, float inv_cell) {
  int c = (int)floorf(__fmul_rn(__fsub_rn(v, mn), inv_cell));
  return min(max(c, 0), FL_ICP_GRID - 1);
}

__device__ __forceinline__ void nn_search(const float4* __restrict__ gp, const int* __restrict__ cs, grid_info gi, float qx, float qy,
                                          float qz, float thr, float* best_out, int* idx_out) {
  float best = FLT_MAX; int bi = -1;
  const int cx = cell_of(qx, gi.minx, gi.inv_cell), cy = cell_of(qy, gi.miny, gi.inv_cell);
  const int kmax = max(max(cx, FL_ICP_GRID - 1 - cx), max(cy, FL_ICP_GRID - 1 - cy));
  for (int k = 0; k <= kmax; ++k) {
    const int y0 = cy - k, y1 = cy + k, x0 = cx - k, x1 = cx + k;
    for (int yy = max(y0, 0); yy <= min(y1, FL_ICP_GRID - 1); ++yy) {
      const bool edge_row = (yy == y0 || yy == y1);
      const int step = edge_row ? 1 : max(x1 - x0, 1);             // interior rows of the ring: only the two end cells
      for (int xx = x0; xx <= x1; xx += step) {
        if (xx < 0 || xx >= FL_ICP_GRID) continue;
        const int c = yy * FL_ICP_GRID + xx;
        const int e = cs[c + 1];
        for (int p = cs[c]; p < e; ++p) {
          const float4 r = gp[p];
          const float d0 = __fsub_rn(qx, r.x), d1 = __fsub_rn(qy, r.y), d2 = __fsub_rn(qz, r.z);
          const float dd = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
          const int id = __float_as_int(r.w);
          if (dd < best || (dd == best && id < bi)) { best = dd; bi = id; }
        }
      }
    }
    const float bound = (float)k * gi.cell * 0.9999f;             // every unvisited point is farther than this in x/y
    const float b2 = bound * bound;
    if (best <= b2 || b2 > thr) break;
  }
  *best_out = best; *idx_out = bi;
}

__global__ void __launch_bounds__(ICP_THREADS) k_icp_run(fl_icp_ws ws, fl_icp_params_t prm, const fl_icp_hyp* __restrict__ hyps,
                                                         const float* __restrict__ t_init, fl_icp_result_t* __restrict__ results) {
  __shared__ int s_warp[ICP_WARPS];
  __shared__ int s_cnt[FL_ICP_CELLS];          // cell counters, then scatter cursors
  __shared__ float s_red[ICP_WARPS * 4];
  __shared__ float s_sum[16];                  // chain results
  __shared__ float s_st[2][2][ICP_CH * 3];     // staging of the ordered chains (block_chain15); [.][0] doubles as the distance stage
  __shared__ int s_cnt2[2];                    // paired-distance counters: {valid pairs, inliers}
  __shared__ float s_Ropt[9], s_Topt[3], s_R[9], s_T[3];
  __shared__ float s_dist_mean, s_dist_diff, s_ratio;
  __shared__ int s_iter, s_go, s_finite;
  __shared__ grid_info s_gi;
  const int h = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_ref = ws.n_ref[h], n_mod = ws.n_mod[h];
  float* pref = ws.pts_ref + (size_t)h * ws.max_pts * 3;
  float* tmp = ws.pts_mod + (size_t)h * ws.max_pts * 3;           // pts_model_tmp
  float* cor_m = ws.cor_m + (size_t)h * ws.max_pts * 3;
  float* cor_r = ws.cor_r + (size_t)h * ws.max_pts * 3;
  float4* gp = ws.grid_pts + (size_t)h * ws.max_pts;
  int* cs = ws.cell_start + (size_t)h * (FL_ICP_CELLS + 1);
  fl_icp_result_t* res = results + h;
  const int status = hyps ? hyps[h].status : FL_OK;

  if (status != FL_OK || n_mod < 3 || n_ref < 3) {                 // ICP.cpp:633-638: returns -1, R and T stay zero-initialised
    if (tid == 0) {
      fl_icp_result_t r;
      for (int k = 0; k < 9; ++k) r.R[k] = 0.f;
      for (int k = 0; k < 3; ++k) r.T[k] = 0.f;
      r.dist_mean = -1.f; r.inlier_ratio = 0.f; r.iterations = 0; r.n_points = n_mod; r.status = status;
      *res = r;
    }
    return;
  }

  // ---- grid over the reference cloud (replaces the KD-tree build, ICP.cpp:650-659) ----
  float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
  for (int i = tid; i < n_ref; i += ICP_THREADS) {
    float x = pref[3 * i], y = pref[3 * i + 1], z = pref[3 * i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) { mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  if (lane == 0) { s_red[warp * 4] = mnx; s_red[warp * 4 + 1] = mny; s_red[warp * 4 + 2] = mxx; s_red[warp * 4 + 3] = mxy; }
  for (int i = tid; i < FL_ICP_CELLS; i += ICP_THREADS) s_cnt[i] = 0;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < ICP_WARPS; ++w) {
      mnx = fminf(mnx, s_red[w * 4]); mny = fminf(mny, s_red[w * 4 + 1]); mxx = fmaxf(mxx, s_red[w * 4 + 2]); mxy = fmaxf(mxy, s_red[w * 4 + 3]);
    }
    float ext = fmaxf(fmaxf(mxx - mnx, mxy - mny), 1e-3f);
    grid_info gi; gi.minx = mnx; gi.miny = mny; gi.cell = ext / (float)FL_ICP_GRID * 1.0001f; gi.inv_cell = 1.0f / gi.cell;
    s_gi = gi;
    for (int k = 0; k < 9; ++k) s_R[k] = (k % 4 == 0) ? 1.f : 0.f;   // R = I, T = 0 (ICP.cpp:644-645)
    for (int k = 0; k < 3; ++k) s_T[k] = 0.f;
  }
  __syncthreads();
  const grid_info gi = s_gi;
  for (int i = tid; i < n_ref; i += ICP_THREADS) {
    float x = pref[3 * i], y = pref[3 * i + 1], z = pref[3 * i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) atomicAdd(&s_cnt[cell_of(y, gi.miny, gi.inv_cell) * FL_ICP_GRID + cell_of(x, gi.minx, gi.inv_cell)], 1);
  }
  __syncthreads();
  {   // exclusive scan of the 4096 counters: 16 per thread
    constexpr int PER = FL_ICP_CELLS / ICP_THREADS;
    int loc[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { loc[k] = s_cnt[tid * PER + k]; sum += loc[k]; }
    int tot;
    int base = block_excl_scan(sum, s_warp, &tot);
#pragma unroll
    for (int k = 0; k < PER; ++k) { cs[tid * PER + k] = base; s_cnt[tid * PER + k] = base; base += loc[k]; }
    if (tid == 0) cs[FL_ICP_CELLS] = tot;
  }
  __syncthreads();
  for (int i = tid; i < n_ref; i += ICP_THREADS) {
    float x = pref[3 * i], y = pref[3 * i + 1], z = pref[3 * i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      int slot = atomicAdd(&s_cnt[cell_of(y, gi.miny, gi.inv_cell) * FL_ICP_GRID + cell_of(x, gi.minx, gi.inv_cell)], 1);
      gp[slot] = make_float4(x, y, z, __int_as_float(i));
    }
  }
  // ---- copyPoints(pts_model, pts_model_tmp) (ICP.cpp:667): invalid points become (0,0,0) ----
  for (int i = tid; i < n_mod; i += ICP_THREADS)
    if (!pt_valid(tmp[3 * i + 2])) { tmp[3 * i] = 0.f; tmp[3 * i + 1] = 0.f; tmp[3 * i + 2] = 0.f; }
  __syncthreads();
  __threadfence_block();

  // getL2distClouds (ICP.cpp:68-111): paired distance of every index pair valid in both clouds, inliers = dist <= thr,
  // dist_mean = (sequential fp32 sum of the inlier distances) / inliers, ratio = inliers / pairs.  The distances are
  // computed by warps 1..7 one chunk ahead into shared memory while lane 0 of warp 0 walks the previous chunk in order;
  // the two counters are order free (integer) and are reduced in parallel.  All threads call it; ends with a barrier.
  auto distance_chain = [&](float thr) {
    float* s_d0 = &s_st[0][0][0];
    float* s_d1 = &s_st[1][0][0];
    if (tid < 2) s_cnt2[tid] = 0;
    const int n_chunks = (n_mod + ICP_CH - 1) / ICP_CH;
    auto fill_chunk = [&](int c, int t, int nt) {
      float* dst = (c & 1) ? s_d1 : s_d0;
      int cnt = 0, nin = 0;
      for (int k = t; k < ICP_CH; k += nt) {
        const int i = c * ICP_CH + k;
        float d = -1.f;                                              // -1 marks a skipped pair (a real distance is >= 0 or NaN)
        if (i < n_mod) {
          const float rz = pref[3 * i + 2], mz = tmp[3 * i + 2];
          if (pt_valid(rz) && pt_valid(mz)) {
            float dx = __fsub_rn(tmp[3 * i], pref[3 * i]), dy = __fsub_rn(tmp[3 * i + 1], pref[3 * i + 1]), dz = __fsub_rn(mz, rz);
            d = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);   // cv::norm(Vec3f) accumulates in double
            ++cnt; if (d <= thr) ++nin;
          }
        }
        dst[k] = d;
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt); nin = __reduce_add_sync(0xffffffffu, nin);
      if ((threadIdx.x & 31) == 0 && (cnt | nin)) { atomicAdd(&s_cnt2[0], cnt); atomicAdd(&s_cnt2[1], nin); }
    };
    __syncthreads();
    if (n_chunks > 0) fill_chunk(0, tid, ICP_THREADS);
    __syncthreads();
    float sum = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      if (warp == 0) {
        if (lane == 0) {
          const float* src = (c & 1) ? s_d1 : s_d0;
          const int cnt = min(ICP_CH, n_mod - c * ICP_CH);
#pragma unroll 8
          for (int k = 0; k < cnt; ++k) { const float d = src[k]; if (d >= 0.f && d <= thr) sum = __fadd_rn(sum, d); }
        }
      } else if (c + 1 < n_chunks) {
        fill_chunk(c + 1, tid - 32, ICP_THREADS - 32);
      }
      __syncthreads();
    }
    if (tid == 0) {
      const int counter = s_cnt2[0], nin = s_cnt2[1];
      if (counter > 0) { s_dist_mean = __fdiv_rn(sum, (float)nin); s_ratio = __fdiv_rn((float)nin, (float)counter); }
      else { s_dist_mean = FLT_MAX; s_ratio = 0.f; }
    }
    __syncthreads();
  };

  distance_chain(FLT_MAX);                                          // ICP.cpp:670
  if (tid == 0) { s_dist_diff = FLT_MAX; s_iter = 0; }
  __syncthreads();

  while (true) {
    if (tid == 0) s_go = (s_dist_mean > prm.dist_mean_thr && s_dist_diff > prm.dist_diff_thr && s_iter < prm.icp_it_thr) ? 1 : 0;   // :684
    __syncthreads();
    if (!s_go) break;
    if (tid == 0) ++s_iter;
    __syncthreads();
    const int iter = s_iter;
    int n_cm, n_cr;
    const float* cm; const float* cr;
    if (iter == 1) {                                                // :700-704 copies with invalid -> 0
      for (int i = tid; i < n_ref; i += ICP_THREADS) {
        bool v = pt_valid(pref[3 * i + 2]);
        cor_r[3 * i] = v ? pref[3 * i] : 0.f; cor_r[3 * i + 1] = v ? pref[3 * i + 1] : 0.f; cor_r[3 * i + 2] = v ? pref[3 * i + 2] : 0.f;
      }
      for (int i = tid; i < n_mod; i += ICP_THREADS) {
        bool v = pt_valid(tmp[3 * i + 2]);
        cor_m[3 * i] = v ? tmp[3 * i] : 0.f; cor_m[3 * i + 1] = v ? tmp[3 * i + 1] : 0.f; cor_m[3 * i + 2] = v ? tmp[3 * i + 2] : 0.f;
      }
      n_cm = n_mod; n_cr = n_ref; cm = cor_m; cr = cor_r;
    } else {                                                        // :708 -> PointsCorresponding :193-279
      const float thr = __fmul_rn(3.f, s_dist_mean);
      int n = 0;
      for (int c0 = 0; c0 < n_mod; c0 += ICP_THREADS) {
        int i = c0 + tid;
        int keep = 0, bi = -1;
        float qx = 0, qy = 0, qz = 0;
        if (i < n_mod) {
          qx = tmp[3 * i]; qy = tmp[3 * i + 1]; qz = tmp[3 * i + 2];
          float best;
          nn_search(gp, cs, gi, qx, qy, qz, thr, &best, &bi);
          keep = (bi >= 0 && best <= thr) ? 1 : 0;                  // squared distance against un-squared 3*dist_mean (:268)
        }
        int tot;
        int pos = n + block_excl_scan(keep, s_warp, &tot);
        if (keep) {
          cor_m[3 * pos] = qx; cor_m[3 * pos + 1] = qy; cor_m[3 * pos + 2] = qz;
          cor_r[3 * pos] = pref[3 * bi]; cor_r[3 * pos + 1] = pref[3 * bi + 1]; cor_r[3 * pos + 2] = pref[3 * bi + 2];
        }
        n += tot;
      }
      n_cm = n_cr = n; cm = cor_m; cr = cor_r;
    }
    __syncthreads();
    if (n_cr < 3 || n_cm < 3) {                                     // :711-715
      if (tid == 0) s_iter = prm.icp_it_thr;
      __syncthreads();
      continue;
    }
    block_chain15(cm, n_cm, cr, n_cr, n_cm, s_st, s_sum);           // centroids + covariance chains (:722-735)
    if (tid == 0) {
      float mc[3], rc[3], cov[9], rm[3];
      for (int k = 0; k < 3; ++k) { mc[k] = __fdiv_rn(s_sum[k], (float)n_cm); rc[k] = __fdiv_rn(s_sum[3 + k], (float)n_cr); }
      for (int k = 0; k < 9; ++k) cov[k] = s_sum[6 + k];
      svd3_rot(cov, s_Ropt);                                        // :741-744
      matvec3(s_Ropt, mc, rm);
      int fin = 1;
      for (int k = 0; k < 3; ++k) { s_Topt[k] = __fsub_rn(rc[k], rm[k]); if (!isfinite(s_Topt[k])) fin = 0; }   // :747
      for (int k = 0; k < 9; ++k) if (!isfinite(s_Ropt[k])) fin = 0;
      s_finite = fin;                                               // :748-749
    }
    __syncthreads();
    if (!s_finite) continue;
    {
      float Ro[9], To[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) Ro[k] = s_Ropt[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) To[k] = s_Topt[k];
      for (int i = tid; i < n_mod; i += ICP_THREADS) {              // transformPoints in place (:756)
        float p[3] = {tmp[3 * i], tmp[3 * i + 1], tmp[3 * i + 2]};
        if (!pt_valid(p[2])) continue;
        float o[3]; matvec3(Ro, p, o);
#pragma unroll
        for (int k = 0; k < 3; ++k) tmp[3 * i + k] = __fadd_rn(o[k], To[k]);
      }
    }
    __syncthreads();
    const float old_mean = s_dist_mean;
    distance_chain(__fmul_rn(3.f, old_mean));                       // :778-780 (starts with a barrier: every thread has read old_mean)
    if (tid == 0) {
      s_dist_diff = __fsub_rn(old_mean, s_dist_mean);
      float nT[3], nR[9];
      matvec3(s_Ropt, s_T, nT);                                     // :793-797
      for (int k = 0; k < 3; ++k) s_T[k] = __fadd_rn(nT[k], s_Topt[k]);
      matmul3(s_Ropt, s_R, nR);
      for (int k = 0; k < 9; ++k) s_R[k] = nR[k];
    }
    __syncthreads();
  }

  if (tid == 0) {
    fl_icp_result_t r;
    if (hyps) {                                                     // detection.cpp:232-234
      float ti[3] = {t_init[3 * h], t_init[3 * h + 1], t_init[3 * h + 2]}, rt[3];
      matvec3(s_R, ti, rt);
      for (int k = 0; k < 3; ++k) r.T[k] = __fadd_rn(rt[k], s_T[k]);
      matmul3(s_R, hyps[h].r_match, r.R);
    } else {
      for (int k = 0; k < 9; ++k) r.R[k] = s_R[k];
      for (int k = 0; k < 3; ++k) r.T[k] = s_T[k];
    }
    r.dist_mean = s_dist_mean; r.inlier_ratio = s_ratio; r.iterations = s_iter; r.n_points = n_mod; r.status = FL_OK;
    *res = r;
  }
}

void fl_launch_icp_run(fl_icp_ws ws, fl_icp_params_t p, const fl_icp_hyp* hyps_or_null, const float* t_init_or_null,
                       fl_icp_result_t* results, cudaStream_t s) {
  if (ws.n_hyp > 0) k_icp_run<<<ws.n_hyp, ICP_THREADS, 0, s>>>(ws, p, hyps_or_null, t_init_or_null, results);
}

// ------------------------------------------------------------------------------------------------
// K12 greedy NMS (order dependent, n is small): one thread walks the list exactly like NMS.cpp:6-39
// ------------------------------------------------------------------------------------------------
__global__ void k_nms(const float* __restrict__ t3, const int32_t* __restrict__ n_model, const float* __restrict__ icp_dist, int n, float th,
                      int32_t* __restrict__ out_idx, int32_t* __restrict__ out_count, uint8_t* __restrict__ done) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int cnt = 0;
  for (int i = 0; i < n; ++i) done[i] = 0;
  for (int i = 0; i < n; ++i) {
    if (done[i]) continue;
    int win = i;
    const int size_th = (int)((double)(float)n_model[i] * 0.85);   // static_cast<int>((float)size * 0.85) (:17)
    for (int j = i + 1; j < n; ++j) {
      if (done[j]) continue;
      double s = 0;
      for (int k = 0; k < 3; ++k) { double d = (double)__fsub_rn(t3[3 * win + k], t3[3 * j + k]); s += d * d; }   // cv::norm(Mat, Mat): fp32 difference, fp64 squares
      if (sqrt(s) < (double)th) {
        done[j] = 1;
        if (n_model[j] > size_th && icp_dist[j] < icp_dist[win]) win = j;
      }
    }
    out_idx[cnt++] = win;
  }
  *out_count = cnt;
}

void fl_launch_nms(const float* t3, const int32_t* n_model, const float* icp_dist, int n, float th, int32_t* out_idx, int32_t* out_count,
                   cudaStream_t s) {
  // scratch "done" flags live after out_idx (caller allocates n ints + n bytes)
  k_nms<<<1, 32, 0, s>>>(t3, n_model, icp_dist, n, th, out_idx, out_count, reinterpret_cast<uint8_t*>(out_idx + n));
}
