// K6, staged variant: global similarity of all templates at the coarsest pyramid level with the linear memories staged
// through shared memory by TMA bulk copies (cp.async.bulk + mbarrier), accumulators in registers.
//
// Why: one template reads 62 byte streams of ~1,200 B from 62 of the 1,024 linear-memory rows (2 modalities x 8 labels x
// T*T rows); over 8,000 templates that is ~0.6 GB of byte traffic per frame against a 1.2 MB working set, so the bound is
// on-chip bandwidth, not HBM.  The working set does not fit one SM's shared memory, so the kernel walks it in PHASES
// (one phase = the T*T rows of one (modality, label), or a block of them), double buffered:
//
//   * a persistent CTA owns a batch of templates (one warp per template slot, TPW slots per warp) and keeps their
//     packed-u8 similarity accumulators in registers for the whole kernel;
//   * one elected thread streams phase p+2 into the free buffer with a single cp.async.bulk while all warps consume
//     phase p; completion is signalled on an mbarrier (expect_tx);
//   * features were sorted by phase once per frame geometry (k_pack_staged; the phase id rides in the top byte of the
//     feature word), and a warp keeps its templates' feature words in registers (lane k holds feature k), broadcasting
//     the next one with a shuffle, so the inner loop touches global memory only through the bulk copies;
//   * lane g owns the NW consecutive 32-bit words [g NW, (g+1) NW) of the similarity map (NW odd, so the 32 lanes of one
//     LDS fall into 32 different banks at any feature offset); a byte-misaligned window costs NW + 1 loads, not 2 NW: the
//     high word of one funnel shift is the low word of the next.
//
// Semantics are those of similarity() / addSimilarities / the scan in matchClass (reference linemod/linemod.cpp:1130-1214,
// 1322-1338, 1487-1506) including flat addressing past a row end.  Eligibility (checked on the host): every
// template has <= 63 coarsest-level features over all modalities (so one u8 lane holds the total, 63*4 = 252) and the same
// width/height for all modalities at that level (what cropTemplates :52-96 produces); otherwise the baseline kernel runs.
#include "fl_internal.cuh"

#define SS_MAXF 64                      // feature words per template (<= 63 used)
#define SS_MAX_PHASES 254               // the phase id travels in the top byte of a feature word
#define SS_NOFEAT 0xFFFFFFFFu           // padding word: phase 255 never comes up

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA: bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// per-geometry packing for the staged kernel: for every template, the coarsest-level features of all modalities sorted
// by phase, as byte offsets inside the phase buffer, plus the per-phase prefix offsets.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pack_staged(fl_tdb db, fl_level_geom g, fl_staged_plan plan) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= db.n_templates) return;
  const int level = db.L - 1;
  const fl_template_hdr_t* hdr = db.hdr + ((size_t)t * db.L + level) * db.M;
  uint32_t* out = plan.gfeat + (size_t)t * SS_MAXF;
  // counting sort by phase (n <= 63): pass 1 counts, pass 2 places
  uint8_t cur[SS_MAX_PHASES + 1];
  for (int p = 0; p <= plan.n_phases; ++p) cur[p] = 0;
  for (int m = 0; m < db.M; ++m)
    for (int k = 0; k < hdr[m].feature_count; ++k) {
      const fl_feature_t f = db.feat[hdr[m].feature_begin + k];
      if (f.x < 0 || f.y < 0 || f.x >= g.W || f.y >= g.H) continue;                     // linemod.cpp:1179
      const int row = (f.y % g.T) * g.T + (f.x % g.T);
      const int ph = (m * 8 + f.label) * plan.n_rowblocks + row / plan.phase_rows;
      ++cur[ph + 1];
    }
  for (int p = 0; p < plan.n_phases; ++p) cur[p + 1] = (uint8_t)(cur[p + 1] + cur[p]);   // cur[p] = first slot of phase p
  const int total = cur[plan.n_phases];
  for (int m = 0; m < db.M; ++m)
    for (int k = 0; k < hdr[m].feature_count; ++k) {
      const fl_feature_t f = db.feat[hdr[m].feature_begin + k];
      if (f.x < 0 || f.y < 0 || f.x >= g.W || f.y >= g.H) continue;
      const int row = (f.y % g.T) * g.T + (f.x % g.T);
      const int rb = row / plan.phase_rows;
      const int ph = (m * 8 + f.label) * plan.n_rowblocks + rb;
      // feature word: phase in the top byte, byte offset of the feature's first response inside the phase buffer below
      out[cur[ph]++] = ((uint32_t)ph << 24) | (uint32_t)((row - rb * plan.phase_rows) * g.cells + (f.y / g.T) * g.Wd + f.x / g.T);
    }
  for (int k = total; k < SS_MAXF; ++k) out[k] = SS_NOFEAT;
}

void fl_launch_pack_staged(fl_tdb db, fl_level_geom g, fl_staged_plan plan, cudaStream_t s) {
  if (db.n_templates > 0) k_pack_staged<<<(db.n_templates + 127) / 128, 128, 0, s>>>(db, g, plan);
}

// ------------------------------------------------------------------------------------------------
#define SS_NBUF_MAX 4

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One feature of one template: lane g owns the NW consecutive 32-bit words [g*NW, (g+1)*NW) of the similarity map, so
// its window is the NW + 1 consecutive words starting at its base + the feature's word offset: NW + 1 loads (instead of
// two per word), one funnel shift and one add per word.  NW is odd, so the 32 lanes of one load hit 32 different banks.
template <int NW>
__device__ __forceinline__ void accumulate_feature(const uint8_t* __restrict__ lane_base, uint32_t a, uint32_t (&acc)[NW]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(lane_base + (a & 0x00FFFFFCu));
  const uint32_t sh = a << 3;                                 // funnel shift uses the low 5 bits: (a & 3) * 8
  constexpr int CH = 12;                                      // loads in flight per chunk (register budget)
  uint32_t carry = w[0];
#pragma unroll
  for (int c = 0; c < NW; c += CH) {
    uint32_t v[CH + 1];
    v[0] = carry;
#pragma unroll
    for (int i = 0; i < CH; ++i) if (c + i < NW) v[i + 1] = w[c + i + 1];
#pragma unroll
    for (int i = 0; i < CH; ++i) if (c + i < NW) acc[c + i] += __funnelshift_r(v[i], v[i + 1], sh);   // four u8 lanes, no carry (sums <= 252)
    carry = v[(NW - c) < CH ? (NW - c) : CH];
  }
}

template <int NW, int TPW>
__global__ void __launch_bounds__(1024, 1) k_similarity_staged(fl_tdb db, fl_level_geom g, const uint8_t* __restrict__ lm_level, float threshold,
                                                               fl_match_t* __restrict__ cand, int cap, int* __restrict__ d_count,
                                                               fl_staged_plan plan) {
  extern __shared__ __align__(128) uint8_t s_buf[];          // plan.n_buf x plan.buf_bytes
  __shared__ __align__(8) uint64_t s_full[SS_NBUF_MAX], s_empty[SS_NBUF_MAX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_cwarps = (blockDim.x >> 5) - 1;                 // consumer warps; the last warp is the TMA producer
  const int level = db.L - 1;
  const int t_begin = blockIdx.x * plan.tpc;
  const int t_end = min(t_begin + plan.tpc, db.n_templates);
  const int n_phases = plan.n_phases, n_buf = plan.n_buf;
  // CTAs walk the phases in rotated order so that, at any moment, different SMs pull different linear-memory rows out of
  // L2 instead of all 148 hammering the same lines (the sums are order independent)
  const int rot = (int)((blockIdx.x * 7u) % (unsigned)n_phases);

  if (tid == 0) {
    for (int b = 0; b < n_buf; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], n_cwarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == n_cwarps) {
    // ===== producer: one elected lane streams phase q into buffer q % n_buf as soon as every consumer warp released it =====
    if (lane == 0) {
      for (int q = 0; q < n_phases; ++q) {
        const int b = q % n_buf;
        if (q >= n_buf) mbar_wait(&s_empty[b], ((q / n_buf) - 1) & 1);
        int p = q + rot; if (p >= n_phases) p -= n_phases;
        const int ml = p / plan.n_rowblocks, rb = p - ml * plan.n_rowblocks;
        const int m = ml >> 3, lab = ml & 7;
        const uint8_t* src = lm_level + (size_t)m * g.mod_stride + (size_t)lab * g.label_stride + (size_t)rb * plan.phase_rows * g.cells;
        const int rows = min(plan.phase_rows, g.T * g.T - rb * plan.phase_rows);
        const uint32_t bytes = (uint32_t)(((size_t)rows * g.cells + plan.halo_bytes + 15) & ~(size_t)15);
        mbar_expect_tx(&s_full[b], bytes);
        bulk_g2s(s_buf + (size_t)b * plan.buf_bytes, src, bytes, &s_full[b]);
      }
    }
    return;
  }

  // ===== consumers: this warp's templates.  Feature words (phase-sorted, phase id in the top byte) live in registers,
  // one per lane (fa: features 0-31, fb: 32-63); kcur walks them in this CTA's rotated phase order =====
  int tt[TPW], kcur[TPW], nfeat[TPW], nleft[TPW];
  uint32_t fa[TPW], fb[TPW], fnext[TPW];
  uint32_t acc[TPW][NW];
#pragma unroll
  for (int s = 0; s < TPW; ++s) {
    const int t = t_begin + warp + s * n_cwarps;
    tt[s] = (t < t_end && db.class_enabled[db.class_of[t]]) ? t : -1;
    fa[s] = fb[s] = fnext[s] = SS_NOFEAT;
    kcur[s] = nfeat[s] = nleft[s] = 0;
    if (tt[s] >= 0) {
      fa[s] = plan.gfeat[(size_t)t * SS_MAXF + lane];
      fb[s] = plan.gfeat[(size_t)t * SS_MAXF + 32 + lane];
      nfeat[s] = __popc(__ballot_sync(0xffffffffu, fa[s] != SS_NOFEAT)) + __popc(__ballot_sync(0xffffffffu, fb[s] != SS_NOFEAT));
      // first feature at or after phase `rot` (padding words carry phase 255 and never count)
      kcur[s] = __popc(__ballot_sync(0xffffffffu, (int)(fa[s] >> 24) < rot)) + __popc(__ballot_sync(0xffffffffu, (int)(fb[s] >> 24) < rot));
      if (kcur[s] >= nfeat[s]) kcur[s] = 0;
      nleft[s] = nfeat[s];
      if (nleft[s] > 0) fnext[s] = __shfl_sync(0xffffffffu, kcur[s] < 32 ? fa[s] : fb[s], kcur[s] & 31);
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[s][i] = 0;
  }

  int b = 0, p = rot;
  uint32_t par = 0;
  const uint8_t* lane_base0 = s_buf + lane * (NW * 4);
  for (int q = 0; q < n_phases; ++q) {
    mbar_wait(&s_full[b], par);
    const uint8_t* lane_base = lane_base0 + b * plan.buf_bytes;
#pragma unroll
    for (int s = 0; s < TPW; ++s) {
      while ((int)(fnext[s] >> 24) == p) {                    // warp-uniform
        accumulate_feature<NW>(lane_base, fnext[s], acc[s]);
        if (++kcur[s] == nfeat[s]) kcur[s] = 0;
        fnext[s] = --nleft[s] > 0 ? __shfl_sync(0xffffffffu, kcur[s] < 32 ? fa[s] : fb[s], kcur[s] & 31) : SS_NOFEAT;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[b]);                  // this warp is done with buffer b
    if (++b == n_buf) { b = 0; par ^= 1; }
    if (++p == n_phases) p = 0;
  }

  // threshold + candidate emission (matchClass :1487-1506)
  const int off = g.T / 2 + (g.T % 2 - 1);
#pragma unroll
  for (int s = 0; s < TPW; ++s) {
    const int t = tt[s];
    if (t < 0) continue;
    const fl_template_hdr_t* hdr = db.hdr + ((size_t)t * db.L + level) * db.M;
    int nf = 0;
    for (int m = 0; m < db.M; ++m) nf += hdr[m].feature_count;
    const int raw_thr = (int)__fadd_rn(__fadd_rn((float)(2 * nf), __fmul_rn(__fdiv_rn(threshold, 100.f), (float)(2 * nf))), 0.5f);
    if (raw_thr >= 255) continue;                             // a u8 total cannot exceed it
    const int wf = (hdr[0].width - 1) / g.T + 1, hf = (hdr[0].height - 1) / g.T + 1;
    const int tp = min((g.Hd - hf) * g.Wd + (g.Wd - wf) + 1, g.cells);               // template_positions :1155
    const int cls = db.class_of[t];
    const uint32_t thr4 = raw_thr < 0 ? 0u : (uint32_t)raw_thr * 0x01010101u;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int w = lane * NW + i;
      if (4 * w >= tp) continue;
      const uint32_t v = acc[s][i];
      if (raw_thr >= 0 && !__vcmpgtu4(v, thr4)) continue;
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int j = 4 * w + bb;
        const int raw = (v >> (8 * bb)) & 0xFF;
        if (j < tp && raw > raw_thr) {
          const int r = j / g.Wd, c = j - r * g.Wd;
          const int slot = atomicAdd(d_count, 1);
          if (slot < cap) {
            fl_match_t mt;
            mt.x = c * g.T + off; mt.y = r * g.T + off;
            mt.similarity = __fadd_rn(__fdiv_rn(__fmul_rn((float)raw, 100.f), (float)(4 * nf)), 0.5f);
            mt.class_idx = cls; mt.template_id = db.L == 1 ? db.tid_of[t] : t;
            cand[slot] = mt;
          }
        }
      }
    }
  }
}

template <int NW, int TPW>
static int launch_staged(fl_tdb db, fl_level_geom g, const uint8_t* lm_level, float threshold, fl_match_t* cand, int cap, int* d_count,
                         fl_staged_plan plan, cudaStream_t s) {
  auto kern = k_similarity_staged<NW, TPW>;
  size_t smem = (size_t)plan.n_buf * plan.buf_bytes;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    configured = smem;
  }
  kern<<<plan.n_cta, plan.block_threads, smem, s>>>(db, g, lm_level, threshold, cand, cap, d_count, plan);
  return 0;
}

// host planning: returns false if the geometry is not supported by the staged kernel.  max_positions = the largest
// template_positions (linemod.cpp:1155) over the uploaded templates: only that many cells of a similarity map are ever
// looked at, so the per-lane accumulator count NW is sized for it rather than for the whole grid.
bool fl_plan_staged(const fl_level_geom& g, int M, int n_templates, int max_positions, int n_sm, fl_staged_plan* plan) {
  static const int kNW[] = {5, 7, 9, 11, 13, 15, 19, 23, 29, 37};
  fl_staged_plan p;
  memset(&p, 0, sizeof p);
  const int T2 = g.T * g.T;
  if (max_positions < 1 || n_templates < 1) return false;
  if (max_positions > g.cells) max_positions = g.cells;
  p.n_words = (max_positions + 3) / 4;
  const int nw_min = (p.n_words + 31) / 32;
  p.nw_template = 0;
  for (int nw : kNW) if (nw >= nw_min) { p.nw_template = nw; break; }
  if (!p.nw_template) return false;                            // accumulators would not fit the register budget
  p.tpw = p.nw_template > 11 ? 1 : 2;                          // register budget: 64 per thread at 1,024 threads
  // bytes staged past the last row of a phase: a window starts at most at the last cell of the last row and the loads run
  // unconditionally over 32 * NW + 1 words
  p.halo_bytes = 32 * p.nw_template * 4 + 16;
  if (p.halo_bytes + 16 > FL_LM_PAD) return false;
  // rows per phase / number of buffers: 4 buffers of <= 48 KB when a (modality, label) splits evenly, else 2 of <= 100 KB;
  // a row block must start 16-byte aligned in global memory
  int pr = 0, nbuf = 0;
  for (int attempt = 0; attempt < 2 && !pr; ++attempt) {
    const size_t budget = attempt == 0 ? 48 * 1024 : 100 * 1024;
    for (int cand_pr = T2; cand_pr >= 1; --cand_pr) {
      if ((size_t)cand_pr * g.cells + p.halo_bytes > budget) continue;
      if (cand_pr != T2 && ((size_t)cand_pr * g.cells) % 16 != 0) continue;
      pr = cand_pr; nbuf = attempt == 0 ? 4 : 2;
      break;
    }
  }
  if (!pr) return false;
  p.phase_rows = pr;
  p.n_rowblocks = (T2 + pr - 1) / pr;
  p.n_phases = M * 8 * p.n_rowblocks;
  if (p.n_phases > SS_MAX_PHASES) return false;                // the phase id travels in one byte of the feature word
  if ((size_t)pr * g.cells + p.halo_bytes >= (1u << 24)) return false;
  p.n_buf = p.n_phases < nbuf ? p.n_phases : nbuf;
  p.buf_bytes = (int)((((size_t)pr * g.cells + p.halo_bytes + 15) & ~(size_t)15) + 127) & ~127;
  // CTAs: whole waves of the SM count, up to 31 consumer warps x TPW templates each (+ 1 producer warp)
  int per_cta_max = 31 * p.tpw;
  int n_cta = (n_templates + per_cta_max - 1) / per_cta_max;
  n_cta = ((n_cta + n_sm - 1) / n_sm) * n_sm;
  p.tpc = (n_templates + n_cta - 1) / n_cta;
  p.n_cta = (n_templates + p.tpc - 1) / p.tpc;
  int warps = (p.tpc + p.tpw - 1) / p.tpw;
  p.block_threads = 32 * ((warps < 1 ? 1 : warps) + 1);
  *plan = p;
  return true;
}

int fl_launch_similarity_staged(fl_tdb db, fl_level_geom g, const uint8_t* lm_level, float threshold, fl_match_t* cand, int cap,
                                int* d_count, fl_staged_plan plan, cudaStream_t s) {
  switch (plan.nw_template) {
    case 5: return launch_staged<5, 2>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 7: return launch_staged<7, 2>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 9: return launch_staged<9, 2>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 11: return launch_staged<11, 2>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 13: return launch_staged<13, 1>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 15: return launch_staged<15, 1>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 19: return launch_staged<19, 1>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 23: return launch_staged<23, 1>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 29: return launch_staged<29, 1>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
    case 37: return launch_staged<37, 1>(db, g, lm_level, threshold, cand, cap, d_count, plan, s);
  }
  return -1;
}
