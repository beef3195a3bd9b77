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

// ------------------------------------------------------------------------------------------------
// K9-K11 fused: ONE persistent launch does, per hypothesis, everything detection() does after the two depth images are known:
// back-projection of the two crops + pairing (detection.cpp:28-44, 112-114), the centroid shift (:162-206) and the whole
// icpCloudToCloud_Ex loop (ICP.cpp:617-809).  One CTA of 1,024 threads works on one hypothesis at a time and takes the next one
// off a ticket counter, so a batch of any size keeps every SM busy and a small batch (Recognition: the top-5 matches) gets one
// SM per hypothesis.
//
// What bounds a hypothesis is NOT bandwidth but three serial fp32 sums per iteration that must round exactly like the
// reference's loops (file header): n dependent FADDs at 4 cycles each (~17 us for 8.4k points).  Everything else is arranged
// around them:
//   * the ordered sums run in warp 0, one lane per column of a record stream (3 + 3 centroid columns, 9 covariance columns, or
//     the single distance column); STAGER warps build the records (coalesced loads, the nine products m[a] * r[b]) into a
//     shared-memory ring one chunk ahead, handing buffers over with named barriers, so a step of the chain is one LDS + one
//     FADD and the chain runs at the FADD latency;
//   * the exact 1-NN search of iteration k + 1 (cvflann KD-tree in the reference) does not need the new dist_mean, only the
//     ACCEPT test does (ICP.cpp:266-273), so all other warps search while warp 0 sums the distances of iteration k; the
//     acceptance + order-preserving compaction happens once the sum is known;
//   * the reference cloud's 64x64 x/y grid (cell starts + points sorted by cell, 16 B each) lives in SHARED memory; a query
//     scans whole grid ROWS of its neighbourhood as contiguous runs (two cell-start reads per row instead of two per cell),
//     first the 3x3 block, then - only if the best distance or the accept radius demands it - the rows of the larger square
//     from the centre outwards, stopping as soon as the next row is farther than the best candidate.
// Squared distance is ((dx*dx + dy*dy) + dz*dz) in fp32 like L2_Simple; ties go to the lower point index; pairs beyond the
// accept radius are dropped by the reference anyway, so the accepted pairs equal the KD-tree's.
// ------------------------------------------------------------------------------------------------
#define ICP_NT 1024
#define ICP_NWARPS (ICP_NT / 32)
#define ICP_REC 15                               // floats per record: m.xyz, r.xyz, m[a] * r[b]
#define ICP_CH 256                               // records per ring buffer (record mode, 2 buffers)
#define ICP_STAGE_FLOATS (2 * ICP_CH * ICP_REC)  // 7,680 floats = 30,720 B
#define ICP_DCH 1024                             // floats per ring buffer in distance mode
#define ICP_DNB 7                                // ... and buffers (7 x 1,024 <= 7,680)
#define ICP_NSTAGE 8                             // stager warps in record mode (one record per thread and chunk)

__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#define ICP_BAR_FULL(b) (1 + (b))
#define ICP_BAR_EMPTY(b) (8 + (b))

// block-wide exclusive scan of one int per thread (ICP_NT threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    s_warp[lane] = w;
  }
  __syncthreads();
  int base = warp > 0 ? s_warp[warp - 1] : 0;
  *total = s_warp[ICP_NWARPS - 1];
  __syncthreads();
  return base + inc - v;
}

// The consumer half of an ordered chain: warp 0 sums NCOL columns of `n_steps` records, record i of chunk c at
// stage[(c % NB) * CHUNK * NCOL + i * NCOL + column]; lane j owns column j.  Called by the 32 lanes of warp 0.
template <int NCOL, int CHUNK, int NB>
__device__ __forceinline__ void chain_consume(int n_steps, const float* s_stage, float* s_sum, int bar_threads) {
  const int lane = threadIdx.x & 31;
  const int n_chunks = (n_steps + CHUNK - 1) / CHUNK;
  float acc = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    const int b = c % NB;
    nb_sync(ICP_BAR_FULL(b), bar_threads);
    const int m = min(CHUNK, n_steps - c * CHUNK);
    if (lane < NCOL) {
      const float* s = s_stage + (size_t)b * CHUNK * NCOL + lane;
#pragma unroll 16
      for (int i = 0; i < m; ++i) acc = __fadd_rn(acc, s[i * NCOL]);
    }
    __syncwarp();
    if (c + NB < n_chunks) nb_arrive(ICP_BAR_EMPTY(b), bar_threads);
  }
  if (lane < NCOL) s_sum[lane] = acc;
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


struct grid_info { float minx, miny, inv_cell, cell; };

__device__ __forceinline__ int cell_of(float v, float mn, float inv_cell) {
  int c = (int)floorf(__fmul_rn(__fsub_rn(v, mn), inv_cell));
  return min(max(c, 0), FL_ICP_GRID - 1);
}

// exact nearest neighbour of q among the grid points, as far as it can matter: a neighbour whose squared distance exceeds
// bound2 may be missed (the caller drops such pairs).  Returns the SLOT in gp (or -1) and the squared distance.
__device__ __forceinline__ void nn_query(const float4* __restrict__ gp, const int* __restrict__ cs, const grid_info gi, const float qx,
                                         const float qy, const float qz, const float bound2, float* best_out, int* slot_out) {
  float best = FLT_MAX; int bslot = -1, bidx = 0x7fffffff;
  const int cx = cell_of(qx, gi.minx, gi.inv_cell), cy = cell_of(qy, gi.miny, gi.inv_cell);
  auto scan_run = [&](int p0, int p1) {
    for (int p = p0; p < p1; ++p) {
      const float4 r = gp[p];
      const float d0 = __fsub_rn(qx, r.x), d1 = __fsub_rn(qy, r.y), d2 = __fsub_rn(qz, r.z);
      const float dd = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
      if (dd <= best) {                                            // rare after the first few points
        const int id = __float_as_int(r.w);
        if (dd < best || id < bidx) { bidx = id; bslot = p; }
        best = dd;
      }
    }
  };
  {   // the 3 x 3 block around the query's cell: three contiguous runs
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, FL_ICP_GRID - 1);
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, FL_ICP_GRID - 1); ++yy) scan_run(cs[yy * FL_ICP_GRID + x0], cs[yy * FL_ICP_GRID + x1 + 1]);
  }
  // every point outside the (2k+1)^2 block is farther than k * cell in x or y; 0.9999 absorbs the rounding of cell_of
  const float c1 = gi.cell * 0.9999f;
  float r2 = fminf(best, bound2);
  if (!(c1 * c1 >= r2)) {
    const float inv_c1 = 1.0f / c1;
    int k = (int)ceilf(sqrtf(r2) * inv_c1);
    k = max(k, 2);
    while (k < FL_ICP_GRID && (float)k * c1 * (float)k * c1 < r2) ++k;
    k = min(k, FL_ICP_GRID);
    for (int j = 0; j <= k; ++j) {
      r2 = fminf(best, bound2);
      if (j >= 2) { const float dyb = (float)(j - 1) * c1; if (dyb * dyb >= r2) break; }   // rows at offset j lie beyond (j-1) * cell in y
      int kx = (int)ceilf(sqrtf(r2) * inv_c1);
      while (kx < k && (float)kx * c1 * (float)kx * c1 < r2) ++kx;
      kx = min(max(kx, 1), k);
      const int x0 = max(cx - kx, 0), x1 = min(cx + kx, FL_ICP_GRID - 1);
      for (int sgn = 0; sgn < (j ? 2 : 1); ++sgn) {
        const int yy = sgn ? cy - j : cy + j;
        if (yy < 0 || yy >= FL_ICP_GRID) continue;
        if (j <= 1 && kx == 1) continue;                           // already scanned in the 3 x 3 pass
        scan_run(cs[yy * FL_ICP_GRID + x0], cs[yy * FL_ICP_GRID + x1 + 1]);
      }
    }
  }
  *best_out = best; *slot_out = bslot;
}

struct fl_icp_args {
  fl_icp_ws ws;
  fl_icp_params_t prm;
  const fl_icp_hyp* hyps;            // NULL: cloud mode (points and counts already in the workspace)
  const uint16_t* ref_depth; int W, H; fl_intrinsics_t Kr;
  fl_icp_result_t* results;
  int* ticket;                       // zero before the launch
  int smem_pts;                      // grid points that fit the shared-memory copy (0: the grid stays in global memory)
};

__global__ void __launch_bounds__(ICP_NT, 1) k_icp_fused(const __grid_constant__ fl_icp_args A) {
  extern __shared__ __align__(16) uint8_t s_dyn[];              // [smem_pts float4][4097 int cell starts][ICP_STAGE_FLOATS float]
  __shared__ int s_warp[ICP_NWARPS];
  __shared__ float s_red[ICP_NWARPS * 4];
  __shared__ float s_sum[16];
  __shared__ int s_cnt2[2];
  __shared__ float s_Ropt[9], s_Topt[3], s_R[9], s_T[3], s_tinit[3], s_tt[3];
  __shared__ float s_dist_mean, s_dist_diff, s_ratio;
  __shared__ int s_iter, s_go, s_finite, s_h;
  __shared__ grid_info s_gi;
  const fl_icp_ws& ws = A.ws;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* s_gp = reinterpret_cast<float4*>(s_dyn);
  int* s_cs = reinterpret_cast<int*>(s_dyn + (size_t)A.smem_pts * 16);
  float* s_stage = reinterpret_cast<float*>(s_cs + FL_ICP_CELLS + 4);
  int* s_cur = reinterpret_cast<int*>(s_stage);                 // scatter cursors of the grid build (the ring is idle then)

  for (;;) {
    if (tid == 0) s_h = atomicAdd(A.ticket, 1);
    __syncthreads();
    const int h = s_h;
    if (h >= ws.n_hyp) break;
    float* pref = ws.pts_ref + (size_t)h * ws.max_pts * 3;
    float* tmp = ws.pts_mod + (size_t)h * ws.max_pts * 3;         // pts_mod, then pts_model_tmp
    float* cor_m = ws.cor_m + (size_t)h * ws.max_pts * 3;
    float* cor_r = ws.cor_r + (size_t)h * ws.max_pts * 3;
    float* dist = ws.dist + (size_t)h * ws.max_pts;
    float* nn_d2 = ws.nn_d2 + (size_t)h * ws.max_pts;
    int* nn_slot = ws.nn_slot + (size_t)h * ws.max_pts;
    fl_icp_result_t* res = A.results + h;
    const int status = A.hyps ? A.hyps[h].status : FL_OK;
    int n_ref = 0, n_mod = 0;

    // ---- record-mode chain (centroids, covariance): warp 0 consumes, warps 1..ICP_NSTAGE stage, the rest wait at the barrier ----
    auto chain_records = [&](const float* pm, int n_m, const float* pr, int n_r, int n_cov) {
      const int n_steps = max(max(n_m, n_r), n_cov);
      const int bar_threads = 32 * (1 + ICP_NSTAGE);
      if (warp == 0) {
        if (n_cov > 0) chain_consume<ICP_REC, ICP_CH, 2>(n_steps, s_stage, s_sum, bar_threads);
        else chain_consume<6, ICP_CH, 2>(n_steps, s_stage, s_sum, bar_threads);
      } else if (warp <= ICP_NSTAGE) {
        const int k = tid - 32;                                     // record of the chunk this thread builds
        const int n_chunks = (n_steps + ICP_CH - 1) / ICP_CH;
        const int ncol = n_cov > 0 ? ICP_REC : 6;
        for (int c = 0; c < n_chunks; ++c) {
          const int b = c & 1, i = c * ICP_CH + k;
          float m0 = 0.f, m1 = 0.f, m2 = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
          if (i < n_m || i < n_cov) { m0 = pm[3 * (size_t)i]; m1 = pm[3 * (size_t)i + 1]; m2 = pm[3 * (size_t)i + 2]; }
          if (i < n_r || i < n_cov) { r0 = pr[3 * (size_t)i]; r1 = pr[3 * (size_t)i + 1]; r2 = pr[3 * (size_t)i + 2]; }
          if (c >= 2) nb_sync(ICP_BAR_EMPTY(b), bar_threads);       // (the loads above are already in flight)
          float* o = s_stage + (size_t)b * ICP_CH * ncol + k * ncol;
          const bool im = i < n_m, ir = i < n_r, ic = i < n_cov;
          o[0] = im ? m0 : 0.f; o[1] = im ? m1 : 0.f; o[2] = im ? m2 : 0.f;
          o[3] = ir ? r0 : 0.f; o[4] = ir ? r1 : 0.f; o[5] = ir ? r2 : 0.f;
          if (n_cov > 0) {                                          // covariance += m * r^T (ICP.cpp:731-735): Matx product, s = 0 + a * b
            const float mm[3] = {m0, m1, m2}, rr[3] = {r0, r1, r2};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int bb = 0; bb < 3; ++bb) o[6 + 3 * a + bb] = ic ? __fmul_rn(mm[a], rr[bb]) : 0.f;
          }
          nb_arrive(ICP_BAR_FULL(b), bar_threads);
        }
      }
      __syncthreads();
    };

    // ---- distance pass (getL2distClouds, ICP.cpp:68-111), first half: every index pair valid in both clouds, its distance
    // (cv::norm(Vec3f): squares accumulate in double), inlier = dist <= thr.  The inlier distances go to dist[] (0 for everything
    // else: adding +0 leaves the running sum unchanged), the two integer counters are order free.  All threads. ----
    auto distance_fill = [&](float thr) {
      if (tid < 2) s_cnt2[tid] = 0;
      __syncthreads();
      int cnt = 0, nin = 0;
      for (int i0 = 0; i0 < n_mod; i0 += ICP_NT) {
        const int i = i0 + tid;
        if (i < n_mod) {
          float d_eff = 0.f;
          const float rz = pref[3 * (size_t)i + 2], mz = tmp[3 * (size_t)i + 2];
          if (pt_valid(rz) && pt_valid(mz)) {
            const float dx = __fsub_rn(tmp[3 * (size_t)i], pref[3 * (size_t)i]), dy = __fsub_rn(tmp[3 * (size_t)i + 1], pref[3 * (size_t)i + 1]), dz = __fsub_rn(mz, rz);
            const float d = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
            ++cnt;
            if (d <= thr) { ++nin; d_eff = d; }
          }
          dist[i] = d_eff;
        }
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt); nin = __reduce_add_sync(0xffffffffu, nin);
      if (lane == 0 && (cnt | nin)) { atomicAdd(&s_cnt2[0], cnt); atomicAdd(&s_cnt2[1], nin); }
      __syncthreads();
    };
    // second half: warp 0 sums dist[] in order, warp 1 stages it through the ring; with `search`, the warps 2.. look up the
    // nearest reference point of every model point meanwhile (PointsCorresponding of the NEXT iteration, ICP.cpp:193-279)
    auto distance_chain = [&](bool search, float bound2, const float4* gp, const grid_info gi) {
      const int bar_threads = 64;
      if (warp == 0) {
        chain_consume<1, ICP_DCH, ICP_DNB>(n_mod, s_stage, s_sum, bar_threads);
      } else if (warp == 1) {
        const int n_chunks = (n_mod + ICP_DCH - 1) / ICP_DCH;
        for (int c = 0; c < n_chunks; ++c) {
          const int b = c % ICP_DNB;
          float4 v[ICP_DCH / 128];
#pragma unroll
          for (int j = 0; j < ICP_DCH / 128; ++j) {
            const int i = c * ICP_DCH + 4 * (lane + 32 * j);
            v[j] = i < n_mod ? *reinterpret_cast<const float4*>(dist + i) : make_float4(0.f, 0.f, 0.f, 0.f);   // max_pts is a multiple of 4
          }
          if (c >= ICP_DNB) nb_sync(ICP_BAR_EMPTY(b), bar_threads);
#pragma unroll
          for (int j = 0; j < ICP_DCH / 128; ++j) *reinterpret_cast<float4*>(s_stage + (size_t)b * ICP_DCH + 4 * (lane + 32 * j)) = v[j];
          nb_arrive(ICP_BAR_FULL(b), bar_threads);
        }
      } else if (search) {
        for (int i = tid - 64; i < n_mod; i += ICP_NT - 64) {
          float best; int slot;
          nn_query(gp, s_cs, gi, tmp[3 * (size_t)i], tmp[3 * (size_t)i + 1], tmp[3 * (size_t)i + 2], bound2, &best, &slot);
          nn_d2[i] = best; nn_slot[i] = slot;
        }
      }
      __syncthreads();
      if (tid == 0) {
        const int counter = s_cnt2[0], nin = s_cnt2[1];
        if (counter > 0) { s_dist_mean = __fdiv_rn(s_sum[0], (float)nin); s_ratio = __fdiv_rn((float)nin, (float)counter); }
        else { s_dist_mean = FLT_MAX; s_ratio = 0.f; }
      }
      __syncthreads();
    };

    // ---- prepare: back-project the two crops, keep pixel pairs valid in both (order preserved), centroid shift ----
    if (A.hyps) {
      if (status == FL_OK) {
        const fl_icp_hyp& hy = A.hyps[h];
        const fl_rect_t rr = hy.rect_ref, rm = hy.rect_model;
        const int total = min(min(rr.width * rr.height, rm.width * rm.height), ws.max_pts);
        const float inv_fxr = __fdiv_rn(1.0f, A.Kr.fx), inv_fyr = __fdiv_rn(1.0f, A.Kr.fy);
        const float inv_fm = __fdiv_rn(1.0f, 608.0f);                // initInternalMat: fx = fy = 608, c = (320, 240) (common.cpp:358)
        int n = 0;
        for (int c0 = 0; c0 < total; c0 += ICP_NT) {
          const int i = c0 + tid;
          float3 a = make_float3(0, 0, 0), b = make_float3(0, 0, 0);
          int keep = 0;
          if (i < total) {
            const int ur = rr.x + i % rr.width, vr = rr.y + i / rr.width;
            const int um = rm.x + i % rm.width, vm = rm.y + i / rm.width;
            a = backproject_mm(A.ref_depth[(size_t)vr * A.W + ur], ur, vr, inv_fxr, inv_fyr, A.Kr.cx, A.Kr.cy);
            b = backproject_mm(hy.model_depth[i], um, vm, inv_fm, inv_fm, 320.0f, 240.0f);
            keep = (pt_valid(a.z) && pt_valid(b.z)) ? 1 : 0;         // paired matToVec, common.cpp:382-405
          }
          int tot;
          const int pos = n + block_excl_scan(keep, s_warp, &tot);
          if (keep) {
            pref[3 * (size_t)pos] = a.x; pref[3 * (size_t)pos + 1] = a.y; pref[3 * (size_t)pos + 2] = a.z;
            tmp[3 * (size_t)pos] = b.x; tmp[3 * (size_t)pos + 1] = b.y; tmp[3 * (size_t)pos + 2] = b.z;
          }
          n += tot;
        }
        n_ref = n_mod = n;
        __syncthreads();
        chain_records(tmp, n, pref, n, 0);                          // getMean x2 (detection.cpp:165-166)
        if (tid < 3) {
          const float mc = n > 0 ? __fdiv_rn(s_sum[tid], (float)n) : 0.f, rc = n > 0 ? __fdiv_rn(s_sum[3 + tid], (float)n) : 0.f;
          const float tt = __fsub_rn(rc, mc);                        // t_match_tmp = r_centroid - m_centroid (:177)
          s_tt[tid] = tt;
          s_tinit[tid] = __fadd_rn(tt, hy.t_match[tid]);             // t_init (:199)
        }
        __syncthreads();
        const float t0 = s_tt[0], t1 = s_tt[1], t2 = s_tt[2];
        for (int i = tid; i < n; i += ICP_NT) {                       // transformPoints(pts_mod, I, t_tmp) (:206)
          if (!pt_valid(tmp[3 * (size_t)i + 2])) continue;
          tmp[3 * (size_t)i] = __fadd_rn(tmp[3 * (size_t)i], t0); tmp[3 * (size_t)i + 1] = __fadd_rn(tmp[3 * (size_t)i + 1], t1);
          tmp[3 * (size_t)i + 2] = __fadd_rn(tmp[3 * (size_t)i + 2], t2);
        }
      }
    } else {
      n_ref = ws.n_ref[h]; n_mod = ws.n_mod[h];
    }
    __syncthreads();

    if (status != FL_OK || n_mod < 3 || n_ref < 3) {               // ICP.cpp:633-638: returns -1, R and T stay zero-initialised
      if (tid == 0) {
        fl_icp_result_t r;
        for (int k = 0; k < 9; ++k) r.R[k] = 0.f;
        for (int k = 0; k < 3; ++k) r.T[k] = 0.f;
        r.dist_mean = -1.f; r.inlier_ratio = 0.f; r.iterations = 0; r.n_points = n_mod; r.status = status;
        *res = r;
      }
      continue;
    }

    // ---- grid over the reference cloud (replaces the KD-tree build, ICP.cpp:650-659) ----
    const bool grid_in_smem = n_ref <= A.smem_pts;
    float4* gp = grid_in_smem ? s_gp : ws.grid_pts + (size_t)h * ws.max_pts;
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    for (int i = tid; i < n_ref; i += ICP_NT) {
      const float x = pref[3 * (size_t)i], y = pref[3 * (size_t)i + 1], z = pref[3 * (size_t)i + 2];
      if (isfinite(x) && isfinite(y) && isfinite(z)) { mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
      mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane == 0) { s_red[warp * 4] = mnx; s_red[warp * 4 + 1] = mny; s_red[warp * 4 + 2] = mxx; s_red[warp * 4 + 3] = mxy; }
    for (int i = tid; i < FL_ICP_CELLS; i += ICP_NT) s_cur[i] = 0;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < ICP_NWARPS; ++w) {
        mnx = fminf(mnx, s_red[w * 4]); mny = fminf(mny, s_red[w * 4 + 1]); mxx = fmaxf(mxx, s_red[w * 4 + 2]); mxy = fmaxf(mxy, s_red[w * 4 + 3]);
      }
      const float ext = fmaxf(fmaxf(mxx - mnx, mxy - mny), 1e-3f);
      grid_info gi; gi.minx = mnx; gi.miny = mny; gi.cell = ext / (float)FL_ICP_GRID * 1.0001f; gi.inv_cell = 1.0f / gi.cell;
      s_gi = gi;
      for (int k = 0; k < 9; ++k) s_R[k] = (k % 4 == 0) ? 1.f : 0.f;   // R = I, T = 0 (ICP.cpp:644-645)
      for (int k = 0; k < 3; ++k) s_T[k] = 0.f;
    }
    __syncthreads();
    const grid_info gi = s_gi;
    for (int i = tid; i < n_ref; i += ICP_NT) {
      const float x = pref[3 * (size_t)i], y = pref[3 * (size_t)i + 1], z = pref[3 * (size_t)i + 2];
      if (isfinite(x) && isfinite(y) && isfinite(z)) atomicAdd(&s_cur[cell_of(y, gi.miny, gi.inv_cell) * FL_ICP_GRID + cell_of(x, gi.minx, gi.inv_cell)], 1);
    }
    __syncthreads();
    {   // exclusive scan of the 4,096 counters: 4 per thread
      constexpr int PER = FL_ICP_CELLS / ICP_NT;
      int loc[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; ++k) { loc[k] = s_cur[tid * PER + k]; sum += loc[k]; }
      int tot;
      int base = block_excl_scan(sum, s_warp, &tot);
#pragma unroll
      for (int k = 0; k < PER; ++k) { s_cs[tid * PER + k] = base; s_cur[tid * PER + k] = base; base += loc[k]; }
      if (tid == 0) s_cs[FL_ICP_CELLS] = tot;
    }
    __syncthreads();
    for (int i = tid; i < n_ref; i += ICP_NT) {
      const float x = pref[3 * (size_t)i], y = pref[3 * (size_t)i + 1], z = pref[3 * (size_t)i + 2];
      if (isfinite(x) && isfinite(y) && isfinite(z)) {
        const int slot = atomicAdd(&s_cur[cell_of(y, gi.miny, gi.inv_cell) * FL_ICP_GRID + cell_of(x, gi.minx, gi.inv_cell)], 1);
        gp[slot] = make_float4(x, y, z, __int_as_float(i));
      }
    }
    // ---- copyPoints(pts_model, pts_model_tmp) (ICP.cpp:667): invalid points become (0,0,0) ----
    for (int i = tid; i < n_mod; i += ICP_NT)
      if (!pt_valid(tmp[3 * (size_t)i + 2])) { tmp[3 * (size_t)i] = 0.f; tmp[3 * (size_t)i + 1] = 0.f; tmp[3 * (size_t)i + 2] = 0.f; }
    __syncthreads();

    distance_fill(FLT_MAX);                                         // ICP.cpp:670
    distance_chain(false, 0.f, gp, gi);
    if (tid == 0) { s_dist_diff = FLT_MAX; s_iter = 0; }
    __syncthreads();
    bool have_nn = false;                                           // nn_slot / nn_d2 hold the neighbours of the current tmp[]

    while (true) {
      if (tid == 0) s_go = (s_dist_mean > A.prm.dist_mean_thr && s_dist_diff > A.prm.dist_diff_thr && s_iter < A.prm.icp_it_thr) ? 1 : 0;   // :684
      __syncthreads();
      if (!s_go) break;
      if (tid == 0) ++s_iter;
      __syncthreads();
      const int iter = s_iter;
      int n_cm, n_cr;
      if (iter == 1) {                                              // :700-704 copies with invalid -> 0
        for (int i = tid; i < n_ref; i += ICP_NT) {
          const bool v = pt_valid(pref[3 * (size_t)i + 2]);
          cor_r[3 * (size_t)i] = v ? pref[3 * (size_t)i] : 0.f; cor_r[3 * (size_t)i + 1] = v ? pref[3 * (size_t)i + 1] : 0.f; cor_r[3 * (size_t)i + 2] = v ? pref[3 * (size_t)i + 2] : 0.f;
        }
        for (int i = tid; i < n_mod; i += ICP_NT) {
          const bool v = pt_valid(tmp[3 * (size_t)i + 2]);
          cor_m[3 * (size_t)i] = v ? tmp[3 * (size_t)i] : 0.f; cor_m[3 * (size_t)i + 1] = v ? tmp[3 * (size_t)i + 1] : 0.f; cor_m[3 * (size_t)i + 2] = v ? tmp[3 * (size_t)i + 2] : 0.f;
        }
        n_cm = n_mod; n_cr = n_ref;
      } else {                                                      // :708 -> PointsCorresponding :193-279
        const float thr = __fmul_rn(3.f, s_dist_mean);
        if (!have_nn) {                                             // (only when the overlapped search could not be used)
          for (int i = tid; i < n_mod; i += ICP_NT) {
            float best; int slot;
            nn_query(gp, s_cs, gi, tmp[3 * (size_t)i], tmp[3 * (size_t)i + 1], tmp[3 * (size_t)i + 2], thr, &best, &slot);
            nn_d2[i] = best; nn_slot[i] = slot;
          }
          __syncthreads();
        }
        int n = 0;
        for (int c0 = 0; c0 < n_mod; c0 += ICP_NT) {
          const int i = c0 + tid;
          int keep = 0, slot = -1;
          if (i < n_mod) { slot = nn_slot[i]; keep = (slot >= 0 && nn_d2[i] <= thr) ? 1 : 0; }   // squared distance against un-squared 3*dist_mean (:268)
          int tot;
          const int pos = n + block_excl_scan(keep, s_warp, &tot);
          if (keep) {
            const float4 r = gp[slot];
            cor_m[3 * (size_t)pos] = tmp[3 * (size_t)i]; cor_m[3 * (size_t)pos + 1] = tmp[3 * (size_t)i + 1]; cor_m[3 * (size_t)pos + 2] = tmp[3 * (size_t)i + 2];
            cor_r[3 * (size_t)pos] = r.x; cor_r[3 * (size_t)pos + 1] = r.y; cor_r[3 * (size_t)pos + 2] = r.z;
          }
          n += tot;
        }
        n_cm = n_cr = n;
      }
      __syncthreads();
      if (n_cr < 3 || n_cm < 3) {                                   // :711-715
        if (tid == 0) s_iter = A.prm.icp_it_thr;
        __syncthreads();
        continue;
      }
      chain_records(cor_m, n_cm, cor_r, n_cr, n_cm);                // centroids + covariance chains (:722-735)
      if (tid == 0) {
        float mc[3], rc[3], cov[9], rm[3];
        for (int k = 0; k < 3; ++k) { mc[k] = __fdiv_rn(s_sum[k], (float)n_cm); rc[k] = __fdiv_rn(s_sum[3 + k], (float)n_cr); }
        for (int k = 0; k < 9; ++k) cov[k] = s_sum[6 + k];
        svd3_rot(cov, s_Ropt);                                      // :741-744
        matvec3(s_Ropt, mc, rm);
        int fin = 1;
        for (int k = 0; k < 3; ++k) { s_Topt[k] = __fsub_rn(rc[k], rm[k]); if (!isfinite(s_Topt[k])) fin = 0; }   // :747
        for (int k = 0; k < 9; ++k) if (!isfinite(s_Ropt[k])) fin = 0;
        s_finite = fin;                                             // :748-749
      }
      __syncthreads();
      if (!s_finite) continue;                                      // (tmp[] unchanged: the neighbours found for it stay valid)
      {
        float Ro[9], To[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ro[k] = s_Ropt[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) To[k] = s_Topt[k];
        for (int i = tid; i < n_mod; i += ICP_NT) {                 // transformPoints in place (:756)
          float p[3] = {tmp[3 * (size_t)i], tmp[3 * (size_t)i + 1], tmp[3 * (size_t)i + 2]};
          if (!pt_valid(p[2])) continue;
          float o[3]; matvec3(Ro, p, o);
#pragma unroll
          for (int k = 0; k < 3; ++k) tmp[3 * (size_t)i + k] = __fadd_rn(o[k], To[k]);
        }
      }
      const float old_mean = s_dist_mean;
      distance_fill(__fmul_rn(3.f, old_mean));                      // :778-780 (its first barrier also orders the transform before the search)
      // The loop goes on only if dist_mean dropped by more than dist_diff_thr; with a non-negative threshold the next accept
      // radius 3 * dist_mean is therefore below 3 * old_mean, which bounds the search that runs under the distance sum.
      const bool more = iter < A.prm.icp_it_thr;
      const float bound2 = A.prm.dist_diff_thr >= 0.f ? __fmul_rn(3.f, old_mean) : FLT_MAX;
      distance_chain(more, bound2, gp, gi);
      have_nn = more;
      if (tid == 0) {
        s_dist_diff = __fsub_rn(old_mean, s_dist_mean);
        float nT[3], nR[9];
        matvec3(s_Ropt, s_T, nT);                                   // :793-797
        for (int k = 0; k < 3; ++k) s_T[k] = __fadd_rn(nT[k], s_Topt[k]);
        matmul3(s_Ropt, s_R, nR);
        for (int k = 0; k < 9; ++k) s_R[k] = nR[k];
      }
      __syncthreads();
    }

    if (tid == 0) {
      fl_icp_result_t r;
      if (A.hyps) {                                                 // detection.cpp:232-234
        float ti[3] = {s_tinit[0], s_tinit[1], s_tinit[2]}, rt[3];
        matvec3(s_R, ti, rt);
        for (int k = 0; k < 3; ++k) r.T[k] = __fadd_rn(rt[k], s_T[k]);
        matmul3(s_R, A.hyps[h].r_match, r.R);
      } else {
        for (int k = 0; k < 9; ++k) r.R[k] = s_R[k];
        for (int k = 0; k < 3; ++k) r.T[k] = s_T[k];
      }
      r.dist_mean = s_dist_mean; r.inlier_ratio = s_ratio; r.iterations = s_iter; r.n_points = n_mod; r.status = FL_OK;
      *res = r;
    }
    __syncthreads();
  }
}

// dynamic shared memory of k_icp_fused for a batch whose largest cloud has max_pts points (0 grid points if they do not fit)
static int icp_smem_plan(int max_pts, int* smem_pts) {
  const int fixed = (FL_ICP_CELLS + 4) * 4 + ICP_STAGE_FLOATS * 4;
  const int avail = 227 * 1024 - 2048 - fixed;                      // 2 KB: the kernel's static shared memory
  *smem_pts = (max_pts * 16 <= avail) ? ((max_pts + 3) & ~3) : 0;
  return fixed + *smem_pts * 16;
}

int fl_launch_icp(fl_icp_ws ws, fl_icp_params_t p, const fl_icp_hyp* hyps_or_null, const uint16_t* ref_depth, int W, int H, fl_intrinsics_t K_ref,
                  fl_icp_result_t* results, int* ticket, int n_sm, cudaStream_t s) {
  if (ws.n_hyp <= 0) return 0;
  fl_icp_args A;
  A.ws = ws; A.prm = p; A.hyps = hyps_or_null; A.ref_depth = ref_depth; A.W = W; A.H = H; A.Kr = K_ref; A.results = results; A.ticket = ticket;
  const int smem = icp_smem_plan(ws.max_pts, &A.smem_pts);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  static int configured[64];                                        // per device: the attribute belongs to the (device, function) pair
  if (dev < 0 || dev >= 64) return -1;
  if (smem > configured[dev]) {
    if (cudaFuncSetAttribute(k_icp_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    configured[dev] = smem;
  }
  if (cudaMemsetAsync(ticket, 0, sizeof(int), s) != cudaSuccess) return -1;
  k_icp_fused<<<min(ws.n_hyp, n_sm), ICP_NT, smem, s>>>(A);
  return cudaGetLastError() == cudaSuccess ? 1 : -1;
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
