// Cluster-split persistent COCG: one thread-block CLUSTER (C <= 8 CTAs = C SMs) per matrix, the whole Krylov solve
// of that matrix in one kernel, everything on chip.
//
// Why (round-1 profile of k_cocg_small, one CTA per matrix): the matrix (1.2 MB for WR-90) was streamed from L2/HBM
// every iteration -- 85 us per matrix-iteration, 3x above even its HBM floor, bound by the load path.  Here CTA c of
// the cluster owns a contiguous band of rows of the RCM-ordered free unknowns and keeps their values in SHARED
// MEMORY for the whole solve (loaded once per job); r, x, p of the own rows live in registers (thread per row); the
// SpMV input is a window of p in shared memory whose halo is pulled from the neighbours' shared memory through
// DSMEM; the auxiliary-space preconditioner keeps G^T r per CTA by recurrence (G^T r -= alpha G^T q) so that its
// nodal exchange rides on the same barrier as the p^T q reduction.  Two cluster barriers per iteration:
//   A  q = A p (own rows), partial p^T q, nodal partial of q          -> push partials          -> barrier 1
//   B  alpha; x += alpha p; r -= alpha q; g -= alpha G^T q (pull nodal partials); z = D^-1 r + G L^-1 g;
//      partial r^T z, |r|^2; z parked in shared memory                -> push partials          -> barrier 2
//   D  beta; p = z + beta p for own rows and (pulling z from the owners) for the halo of the window.
// Every CTA sums the partials of all CTAs in rank order, so all CTAs see bit-identical scalars and take identical
// branches.  One solve of WR-90 (4 308 free unknowns, ~300 iterations) drops from 22 ms on one SM to ~1.5 ms on 8.
// Host side of the split (RCM, partition, windows, halo and nodal lists): cluster_plan.hpp.
// Replaces Eigen's BiCGSTAB/ILUT behind solve_linear (src/solver.cpp:35-193) for complex symmetric systems.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <memory>
#include <mutex>

#include "cluster_plan.hpp"
#include "solve_internal.cuh"

namespace cg = cooperative_groups;

namespace efb {

#ifndef EFB_CL_LPN
#define EFB_CL_LPN 2
#endif
constexpr int CL_THREADS_MAX = 640;   // thread per row up to here at 96 registers
constexpr int CL_THREADS_WIDE = 768;  // wide CTAs (a 6-CTA split of WR-90 has 745 rows per CTA): 80 registers, a few spills, still
                                      // faster than two rows per thread (measured 2.9 against 3.8 ms per job)

struct ClusterDev {
  int C, mc, aux;
  int max_own, max_w, max_my, max_slots, max_halo, max_n2e, max_nsrc;
  int max_push;    // longest push list (own rows sent into other CTAs' halos)
  int max_vunits;  // value image of a CTA in 8-byte units (real blocks: 1 per slot, complex blocks: 2)
  int ndeg, sdeg;  // widths of the per-node ELL lists below (own edges at a node / CTAs touching a node)
  const uint16_t *n2e_ell;  // [C][ndeg][max_my] own local row << 1 | head, 0xffff = none
  const uint32_t *nsrc_ell; // [C][sdeg][max_my] CTA << 16 | that CTA's node slot, 0xffffffff = none
  long long *prof;  // [C][24] cycle counters per phase (EDGEFEM_B200_CLUSTER_PROF=1), else null
  const int32_t *cta_info, *row_edge;
  const uint16_t *row_ws, *row_n0, *row_n1;
  const int32_t *blk_off, *blk_voff, *slot_src;
  const uint16_t *slot_col, *halo_ws, *push_row;
  const uint32_t *halo_src, *push_dst;
  const int32_t *node_id, *n2e_ptr;
  const uint32_t *n2e_item;
  const int32_t *nsrc_ptr;
  const uint32_t *nsrc_item;
};

struct ClusterPlanDev {  // shared by the systems with the same pattern / flags / gradient and the plan cache
  ~ClusterPlanDev() {
    for (void *b : blocks) dfree(b);
  }
  ClusterPlanHost h;
  uint64_t flags_hash = 0;  // of the real/complex row flags the plan was built for
  std::vector<uint16_t> n2e_ell;
  std::vector<uint32_t> nsrc_ell;
  int ndeg = 0, sdeg = 0;
  ClusterDev d{};
  std::vector<void *> blocks;
};

static void plan_degrees(const ClusterPlanHost &H, int *ndeg_out, int *sdeg_out) {
  int ndeg = 0, sdeg = 0;
  for (int cta = 0; cta < H.C; ++cta) {
    const int32_t *I = &H.cta_info[(size_t)cta * CL_INFO_STRIDE];
    const int32_t *np_ = &H.n2e_ptr[I[CI_OFF_NODE] + cta], *sp_ = &H.nsrc_ptr[I[CI_OFF_NODE] + cta];
    for (int j = 0; j < I[CI_N_MY]; ++j) {
      ndeg = std::max(ndeg, np_[j + 1] - np_[j]);
      sdeg = std::max(sdeg, sp_[j + 1] - sp_[j]);
    }
  }
  *ndeg_out = ndeg;
  *sdeg_out = sdeg;
}

static size_t cluster_smem_bytes(int nr, const ClusterPlanHost &P) {
  int ndeg, sdeg;
  plan_degrees(P, &ndeg, &sdeg);
  size_t b = 0;
  b += ((size_t)P.max_vunits * 8 + 15) / 16 * 16;  // mat_d (real blocks: doubles, complex blocks: complex128)
  b += (size_t)P.max_w * nr * 16;           // p_w
  b += (size_t)P.max_own * nr * 16 * 2;     // q_own, z_own
  b += (size_t)std::max(P.max_my, 1) * nr * 16 * 3;  // wp, g, w
  b += (size_t)std::max(P.max_my, 1) * 16;  // linv
  b += 2 * CL_MAX_C * 8 * 8;                // partial banks
  b += 33 * 8 * 8;                          // block reduction scratch
  b += 16;                                  // job slot
  b += ((size_t)P.max_slots * 2 + 15) / 16 * 16;  // mat_c
  b += ((size_t)std::max(P.max_halo, 1) * 4 + 15) / 16 * 16;   // halo_src
  b += ((size_t)std::max(P.max_halo, 1) * 2 + 15) / 16 * 16;   // halo_ws
  b += (size_t)std::max(P.max_halo, 1) * nr * 16;               // zh: halo staging the owners push z into
  b += ((size_t)std::max(P.max_push, 1) * 4 + 15) / 16 * 16;   // push_dst
  b += ((size_t)std::max(P.max_push, 1) * 2 + 15) / 16 * 16;   // push_row
  b += ((size_t)ndeg * std::max(P.max_my, 1) * 2 + 15) / 16 * 16;  // n2e_ell
  b += ((size_t)sdeg * std::max(P.max_my, 1) * 4 + 15) / 16 * 16;  // nsrc_ell
  b += 24 * 8;                                                 // phase counters
  return b + 16;
}

__device__ __forceinline__ c128 ldg_stream16(const c128 *p) {
  c128 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// complex division through one reciprocal (the scalars of the recurrence are far from over/underflow; a zero or
// non-finite denominator is caught by the callers as a breakdown)
__device__ __forceinline__ c128 cdiv_fast(c128 a, c128 b) {
  const double inv = 1.0 / fma(b.x, b.x, b.y * b.y);
  return make_double2(fma(a.x, b.x, a.y * b.y) * inv, fma(a.y, b.x, -a.x * b.y) * inv);
}

// Block-wide sums of N doubles pushed to every CTA of the cluster: warp shuffles, one block barrier, then thread
// (c*N + k) adds the warp partials of sum k in warp order and stores the result into bank[crank*8 + k] of CTA c.
template <int N>
__device__ __forceinline__ void reduce_push(cg::cluster_group &cluster, double (&v)[N], double *red, double *bank, int C, int crank, long long *pp = nullptr) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  long long tq0 = 0, tq1 = 0, tq2 = 0;
  if (pp) tq0 = clock64();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double a = v[k];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[wid * N + k] = a;
  }
  if (pp) tq1 = clock64();
  __syncthreads();
  if (pp) tq2 = clock64();
  if ((int)threadIdx.x < N * C) {
    const int k = threadIdx.x % N, c = threadIdx.x / N;
    double a0 = 0.0, a1 = 0.0;
    int w = 0;
    for (; w + 2 <= nw; w += 2) {
      a0 += red[w * N + k];
      a1 += red[(w + 1) * N + k];
    }
    if (w < nw) a0 += red[w * N + k];
    double *dst = c == crank ? bank : cluster.map_shared_rank(bank, c);
    dst[crank * 8 + k] = a0 + a1;
  }
  if (pp) { pp[0] += tq1 - tq0; pp[1] += tq2 - tq1; pp[2] += clock64() - tq2; }
}

// Totals over the ranks of a bank [CL_MAX_C][8], the same fixed tree in every warp of every CTA (bit-identical
// scalars everywhere): lane l holds entries (l/8, l%8) and (4 + l/8, l%8), then xor-shuffles over 8 and 16.
template <int N>
__device__ __forceinline__ void bank_totals(const double *bank, int C, double (&tot)[N]) {
  const int lane = threadIdx.x & 31;
  double v = bank[lane] + bank[lane + 32];
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
#pragma unroll
  for (int k = 0; k < N; ++k) tot[k] = __shfl_sync(0xffffffffu, v, k);
}

template <int NR, int RPT, int NTMAX = CL_THREADS_MAX>
__global__ void __launch_bounds__(NTMAX, 1)
k_cocg_cluster(const SolveDev D, const ClusterDev K, int first_matrix, int n_jobs, int groups, int *job_counter, const c128 *__restrict__ bvec,
               c128 *xvec, int zero_x, int max_restarts) {
  static_assert(3 * NR <= 8, "partial banks hold 8 doubles per CTA");
  cg::cluster_group cluster = cg::this_cluster();
  const int C = K.C;
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
  extern __shared__ __align__(16) unsigned char sm[];
  double *mat_d = (double *)sm;  // value image: per ELL block either doubles (all rows real) or complex128
  c128 *p_w = (c128 *)(sm + ((size_t)K.max_vunits * 8 + 15) / 16 * 16);
  c128 *q_own = p_w + (size_t)K.max_w * NR;
  c128 *z_own = q_own + (size_t)K.max_own * NR;
  c128 *wp = z_own + (size_t)K.max_own * NR;
  c128 *g_s = wp + (size_t)K.max_my * NR;
  c128 *w_s = g_s + (size_t)K.max_my * NR;
  c128 *linv_s = w_s + (size_t)K.max_my * NR;
  double *part = (double *)(linv_s + K.max_my);  // [2][CL_MAX_C][8]
  double *red = part + 2 * CL_MAX_C * 8;         // [33*8]
  int *s_job = (int *)(red + 33 * 8);
  uint16_t *mat_c = (uint16_t *)(s_job + 4);
  auto up16 = [](size_t b) { return (b + 15) / 16 * 16; };
  // the index lists live in shared memory: every cluster barrier invalidates L1 (CCTL.IVALL behind the cluster-scope
  // acquire), so lists read from global memory would come from L2 again in every phase of every iteration
  uint32_t *halo_src_s = (uint32_t *)((unsigned char *)mat_c + up16((size_t)K.max_slots * 2));
  uint16_t *halo_ws_s = (uint16_t *)((unsigned char *)halo_src_s + up16((size_t)max(K.max_halo, 1) * 4));
  c128 *zh = (c128 *)((unsigned char *)halo_ws_s + up16((size_t)max(K.max_halo, 1) * 2));  // [max_halo][NR]
  uint32_t *push_dst_s = (uint32_t *)(zh + (size_t)max(K.max_halo, 1) * NR);
  uint16_t *push_row_s = (uint16_t *)((unsigned char *)push_dst_s + up16((size_t)max(K.max_push, 1) * 4));
  uint16_t *n2e_ell_s = (uint16_t *)((unsigned char *)push_row_s + up16((size_t)max(K.max_push, 1) * 2));
  uint32_t *nsrc_ell_s = (uint32_t *)((unsigned char *)n2e_ell_s + up16((size_t)K.ndeg * K.max_my * 2));
  long long *prof_s = (long long *)((unsigned char *)nsrc_ell_s + up16((size_t)K.sdeg * K.max_my * 4));
  double *bank0 = part, *bank1 = part + CL_MAX_C * 8;

  const int32_t *I = K.cta_info + crank * CL_INFO_STRIDE;
  const int n_own = I[CI_N_OWN], n_my = I[CI_N_MY], n_halo = I[CI_N_HALO], n_blk = I[CI_N_BLK], n_slots = I[CI_N_SLOTS];
  const int off_row = I[CI_OFF_ROW], off_slot = I[CI_OFF_SLOT], off_blk = I[CI_OFF_BLK], off_halo = I[CI_OFF_HALO];
  const int off_node = I[CI_OFF_NODE];
  const int n_push = I[CI_N_PUSH], off_push = I[CI_OFF_PUSH];
  const int m = D.m;

  // per-thread rows: local row t = u * nth + tid (a warp = one 32-row ELL block)
  bool valid[RPT];
  int edge[RPT], ws[RPT], n0[RPT], n1[RPT], base[RPT], width[RPT], vbase[RPT];
  bool cplx[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    const int t = u * nth + tid;
    valid[u] = t < n_own;
    edge[u] = valid[u] ? K.row_edge[off_row + t] : 0;
    ws[u] = valid[u] ? K.row_ws[off_row + t] : 0;
    n0[u] = valid[u] ? K.row_n0[off_row + t] : 0;
    n1[u] = valid[u] ? K.row_n1[off_row + t] : 0;
    const int b = t >> 5;
    base[u] = 0;
    width[u] = 0;
    vbase[u] = 0;
    cplx[u] = true;
    if (b < n_blk) {
      base[u] = K.blk_off[off_blk + b];
      width[u] = (K.blk_off[off_blk + b + 1] - base[u]) >> 5;
      vbase[u] = K.blk_voff[off_blk + b];
      cplx[u] = (K.blk_voff[off_blk + b + 1] - vbase[u]) == 2 * (width[u] << 5);  // warp-uniform: a warp is one block
    }
  }
  for (int i = tid; i < n_slots; i += nth) mat_c[i] = K.slot_col[off_slot + i];
  for (int i = tid; i < n_halo; i += nth) {
    halo_ws_s[i] = K.halo_ws[off_halo + i];
    halo_src_s[i] = K.halo_src[off_halo + i];
  }
  for (int i = tid; i < n_push; i += nth) {
    push_dst_s[i] = K.push_dst[off_push + i];
    push_row_s[i] = K.push_row[off_push + i];
  }
  for (int i = tid; i < K.ndeg * K.max_my; i += nth) n2e_ell_s[i] = K.n2e_ell[(size_t)crank * K.ndeg * K.max_my + i];
  for (int i = tid; i < K.sdeg * K.max_my; i += nth) nsrc_ell_s[i] = K.nsrc_ell[(size_t)crank * K.sdeg * K.max_my + i];
  for (int i = tid; i < 2 * CL_MAX_C * 8; i += nth) part[i] = 0.0;  // banks of absent ranks (C < 8) stay zero
  if (tid < 24) prof_s[tid] = 0;
  // every CTA of the cluster has started (and set up its lists) before anyone touches a neighbour's shared memory: the first
  // remote access is the job id that rank 0 writes below (compute-sanitizer racecheck flags it without this barrier)
  cluster.sync();
  long long t_last = 0;
  const bool prof_on = K.prof != nullptr && tid == 0;
  if (prof_on) t_last = clock64();
  auto PROF = [&](int k) {
    if (prof_on) {
      const long long t = clock64();
      prof_s[k] += t - t_last;
      t_last = t;
    }
  };

  // q = A p over the own rows (p from the window in shared memory)
  auto spmv = [&](c128 (&out)[RPT][NR]) {
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      c128 acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = cmake(0.0, 0.0);
      const uint16_t *mc_ = mat_c + base[u] + lane;
      if (cplx[u]) {
        const c128 *mv = (const c128 *)(mat_d + vbase[u]) + lane;
        int k = 0;
        for (; k + 4 <= width[u]; k += 4) {
          c128 a4[4];
          int c4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a4[j] = mv[(k + j) * 32];
            c4[j] = mc_[(k + j) * 32];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] = cfma(a4[j], p_w[c4[j] * NR + r], acc[r]);
        }
        for (; k < width[u]; ++k) {
          const c128 a = mv[k * 32];
          const int c = mc_[k * 32];
#pragma unroll
          for (int r = 0; r < NR; ++r) acc[r] = cfma(a, p_w[c * NR + r], acc[r]);
        }
      } else {
        // real block: 8-byte values (2 shared-memory wavefronts per 32 entries instead of 4), 2 FMAs per entry instead of 4
        const double *mv = mat_d + vbase[u] + lane;
        int k = 0;
        for (; k + 4 <= width[u]; k += 4) {
          double a4[4];
          int c4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a4[j] = mv[(k + j) * 32];
            c4[j] = mc_[(k + j) * 32];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              const c128 pv = p_w[c4[j] * NR + r];
              acc[r].x = fma(a4[j], pv.x, acc[r].x);
              acc[r].y = fma(a4[j], pv.y, acc[r].y);
            }
        }
        for (; k < width[u]; ++k) {
          const double a = mv[k * 32];
          const int c = mc_[k * 32];
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const c128 pv = p_w[c * NR + r];
            acc[r].x = fma(a, pv.x, acc[r].x);
            acc[r].y = fma(a, pv.y, acc[r].y);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) out[u][r] = acc[r];
    }
  };
  // window halo: p_w[h] = z(owner) (+ beta p_w[h])
  auto pull_halo = [&](const c128 (&beta)[NR], bool use_beta) {
    for (int h = tid; h < n_halo; h += nth) {
      const int hw = halo_ws_s[h];
      const uint32_t src = halo_src_s[h];
      const c128 *rz = cluster.map_shared_rank(z_own, src >> 16) + (size_t)(src & 0xffffu) * NR;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const c128 z = rz[r];
        p_w[hw * NR + r] = use_beta ? cfma(beta[r], p_w[hw * NR + r], z) : z;
      }
    }
  };
  // The iteration does not pull: before barrier 2 every owner PUSHES z of the rows its neighbours read into their halo
  // staging zh (remote stores are fire-and-forget, a remote load is a 215-cycle round trip), after the barrier the halo of
  // the window is updated from local shared memory: p_w[h] = zh[h] + beta p_w[h].
  auto push_halo = [&]() {
    for (int i = tid; i < n_push; i += nth) {
      const uint32_t dst = push_dst_s[i];
      const c128 *zsrc = z_own + (size_t)push_row_s[i] * NR;
      c128 *zd = cluster.map_shared_rank(zh, dst >> 16) + (size_t)(dst & 0xffffu) * NR;
#pragma unroll
      for (int r = 0; r < NR; ++r) zd[r] = zsrc[r];
    }
  };
  auto halo_from_staging = [&](const c128 (&beta)[NR]) {
    for (int h = tid; h < n_halo; h += nth) {
      const int hw = halo_ws_s[h];
#pragma unroll
      for (int r = 0; r < NR; ++r) p_w[hw * NR + r] = cfma(beta[r], p_w[hw * NR + r], zh[(size_t)h * NR + r]);
    }
  };
  // wp[n] = sum over the own edges at my node n of +-q_own: LPN lanes per node, lane l takes items l, l+LPN, ... of the
  // node's ELL list.  All item loads are issued first, then all q loads (no data-dependent loop exit in the common
  // case of <= 8 LPN own edges at a node), log2(LPN) shuffle steps, fixed order.
  constexpr int LPN = EFB_CL_LPN, NPW = 32 / LPN;  // lanes per node, nodes per warp step
  auto nodal_partial = [&]() {
    if (!K.aux) return;
    const int ll = lane & (LPN - 1);
    for (int jb = (tid >> 5) * NPW; jb < n_my; jb += (nth >> 5) * NPW) {
      const int j = jb + lane / LPN;
      c128 a[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) a[r] = cmake(0.0, 0.0);
      if (j < n_my) {
        uint32_t it[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = ll + LPN * u;
          it[u] = k < K.ndeg ? (uint32_t)n2e_ell_s[k * K.max_my + j] : 0xffffu;
        }
        c128 v[8][NR];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int r = 0; r < NR; ++r) v[u][r] = it[u] != 0xffffu ? q_own[(size_t)(it[u] >> 1) * NR + r] : cmake(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int r = 0; r < NR; ++r) a[r] = (it[u] & 1u) ? cadd(a[r], v[u][r]) : csub(a[r], v[u][r]);
        for (int k = ll + 8 * LPN; k < K.ndeg; k += LPN) {  // nodes with more than 8 LPN own edges (rare)
          const uint32_t itk = n2e_ell_s[k * K.max_my + j];
          if (itk == 0xffffu) break;
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const c128 vv = q_own[(size_t)(itk >> 1) * NR + r];
            a[r] = (itk & 1u) ? cadd(a[r], vv) : csub(a[r], vv);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int o = 1; o < LPN; o <<= 1) {
          a[r].x += __shfl_xor_sync(0xffffffffu, a[r].x, o);
          a[r].y += __shfl_xor_sync(0xffffffffu, a[r].y, o);
        }
      if (j < n_my && ll == 0)
#pragma unroll
        for (int r = 0; r < NR; ++r) wp[j * NR + r] = a[r];
    }
  };
  // g[n] (-)= alpha * sum over every CTA touching n of its partial; w[n] = linv[n] g[n]
  auto nodal_combine = [&](bool set, const c128 (&alpha)[NR]) {
    if (!K.aux) return;
    for (int j = tid; j < n_my; j += nth) {
      c128 a[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) a[r] = cmake(0.0, 0.0);
      {  // up to four sources with independent (remote) loads, ascending rank; more: the serial tail
        uint32_t it[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) it[u] = u < K.sdeg ? nsrc_ell_s[u * K.max_my + j] : 0xffffffffu;
        c128 v[4][NR];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const c128 *src = ((int)(it[u] >> 16) == crank || it[u] == 0xffffffffu ? wp : cluster.map_shared_rank(wp, it[u] >> 16)) +
                            (size_t)(it[u] == 0xffffffffu ? 0u : (it[u] & 0xffffu)) * NR;
#pragma unroll
          for (int r = 0; r < NR; ++r) v[u][r] = it[u] != 0xffffffffu ? src[r] : cmake(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int r = 0; r < NR; ++r) a[r] = cadd(a[r], v[u][r]);
      }
      for (int k = 4; k < K.sdeg; ++k) {
        const uint32_t it = nsrc_ell_s[k * K.max_my + j];
        if (it == 0xffffffffu) break;
        const c128 *v = ((int)(it >> 16) == crank ? wp : cluster.map_shared_rank(wp, it >> 16)) + (size_t)(it & 0xffffu) * NR;
#pragma unroll
        for (int r = 0; r < NR; ++r) a[r] = cadd(a[r], v[r]);
      }
      const c128 li = linv_s[j];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const c128 gn = set ? a[r] : cfma(cneg(alpha[r]), a[r], g_s[j * NR + r]);
        g_s[j * NR + r] = gn;
        w_s[j * NR + r] = cmul(li, gn);
      }
    }
  };

  for (;;) {  // jobs
    if (crank == 0 && tid == 0) {
      const int q = atomicAdd(job_counter, 1);
      const int jb = q < n_jobs ? job_counter[1 + q] : -1;
      for (int c = 0; c < C; ++c) *cluster.map_shared_rank(s_job, c) = jb;
    }
    cluster.sync();
    const int job = *s_job;
    if (job < 0) break;
    const int f = first_matrix + job / groups;
    const int s0 = f * D.n_rhs + (job % groups) * NR;
    const c128 *__restrict__ av = D.vals + (size_t)f * D.nnz;
    const c128 *__restrict__ dinv = D.dinv + (size_t)f * m;
    for (int b = 0; b < n_blk; ++b) {  // value image of the job, block by block (real blocks keep the real part only)
      const int o0 = K.blk_off[off_blk + b], o1 = K.blk_off[off_blk + b + 1], v0 = K.blk_voff[off_blk + b];
      const bool bc = (K.blk_voff[off_blk + b + 1] - v0) == 2 * (o1 - o0);
      for (int i = o0 + tid; i < o1; i += nth) {
        const int src = K.slot_src[off_slot + i];
        const c128 v = src >= 0 ? ldg_stream16(&av[src]) : cmake(0.0, 0.0);
        if (bc) ((c128 *)(mat_d + v0))[i - o0] = v;
        else mat_d[v0 + i - o0] = v.x;
      }
    }
    if (K.aux)
      for (int j = tid; j < n_my; j += nth) linv_s[j] = D.linv[(size_t)f * D.n_node + K.node_id[off_node + j]];
    const size_t off0 = (size_t)s0 * m;
    const c128 *bg = bvec + off0;
    c128 *xg = xvec + off0;
    // Dirichlet rows are decoupled identity rows: x_e = b_e / A_ee
    if (K.mc < m)
      for (int e = crank * nth + tid; e < m; e += C * nth)
        if (D.dir[e]) {
          const c128 de = dinv[e];
#pragma unroll
          for (int r = 0; r < NR; ++r) xg[(size_t)r * m + e] = cmul(de, bg[(size_t)r * m + e]);
        }
    c128 di[RPT], xr[RPT][NR], rv[RPT][NR], pr[RPT][NR];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      di[u] = valid[u] ? dinv[edge[u]] : cmake(0.0, 0.0);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        xr[u][r] = (valid[u] && !zero_x) ? xg[(size_t)r * m + edge[u]] : cmake(0.0, 0.0);
        rv[u][r] = cmake(0.0, 0.0);
        pr[u][r] = cmake(0.0, 0.0);
      }
    }
    int iters[NR];
    bool act[NR], conv[NR];
    double bb[NR], rrn[NR];
    c128 rho[NR], zero_nr[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      iters[r] = 0; act[r] = false; conv[r] = false; bb[r] = 0.0; rrn[r] = 0.0;
      rho[r] = cmake(0.0, 0.0);
      zero_nr[r] = cmake(0.0, 0.0);
    }
    __syncthreads();

    for (int cycle = 0;; ++cycle) {
      // (1) true residual r = b - A x of the current iterate
      c128 qv[RPT][NR];
      if (cycle == 0 && zero_x) {
#pragma unroll
        for (int u = 0; u < RPT; ++u)
#pragma unroll
          for (int r = 0; r < NR; ++r) qv[u][r] = cmake(0.0, 0.0);
      } else {
#pragma unroll
        for (int u = 0; u < RPT; ++u)
          if (valid[u])
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              p_w[ws[u] * NR + r] = xr[u][r];
              z_own[(size_t)(u * nth + tid) * NR + r] = xr[u][r];
            }
        cluster.sync();
        pull_halo(zero_nr, false);
        __syncthreads();
        spmv(qv);
      }
      {
        double d[2 * NR];
#pragma unroll
        for (int k = 0; k < 2 * NR; ++k) d[k] = 0.0;
#pragma unroll
        for (int u = 0; u < RPT; ++u)
          if (valid[u])
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              const c128 bi = bg[(size_t)r * m + edge[u]];
              const c128 ri = csub(bi, qv[u][r]);
              rv[u][r] = ri;
              q_own[(size_t)(u * nth + tid) * NR + r] = ri;
              d[2 * r] += cabs2(ri);
              d[2 * r + 1] += cabs2(bi);
            }
        __syncthreads();
        nodal_partial();
        reduce_push<2 * NR>(cluster, d, red, bank0, C, crank);
        cluster.sync();
        double tot[2 * NR];
        bank_totals<2 * NR>(bank0, C, tot);
        nodal_combine(true, zero_nr);
        bool any = false;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          rrn[r] = tot[2 * r];
          bb[r] = tot[2 * r + 1];
          conv[r] = rrn[r] <= D.tol2 * bb[r];
          act[r] = !conv[r] && iters[r] < D.max_it && cycle <= max_restarts && isfinite(rrn[r]);
          any |= act[r];
        }
        if (!any) break;
      }
      __syncthreads();
      // (2) z = M^-1 r, rho = r^T z, p = z
      c128 zv[RPT][NR];
      {
        double d[3 * NR];
#pragma unroll
        for (int k = 0; k < 3 * NR; ++k) d[k] = 0.0;
#pragma unroll
        for (int u = 0; u < RPT; ++u)
          if (valid[u])
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              c128 z = cmul(di[u], rv[u][r]);
              if (K.aux) z = cadd(z, csub(w_s[n1[u] * NR + r], w_s[n0[u] * NR + r]));
              zv[u][r] = z;
              z_own[(size_t)(u * nth + tid) * NR + r] = z;
              const c128 t = cmul(rv[u][r], z);
              d[3 * r] += t.x; d[3 * r + 1] += t.y;
              d[3 * r + 2] += cabs2(rv[u][r]);
            }
        reduce_push<3 * NR>(cluster, d, red, bank1, C, crank);
        cluster.sync();
        double tot[3 * NR];
        bank_totals<3 * NR>(bank1, C, tot);
#pragma unroll
        for (int r = 0; r < NR; ++r) rho[r] = cmake(tot[3 * r], tot[3 * r + 1]);
      }
#pragma unroll
      for (int u = 0; u < RPT; ++u)
        if (valid[u])
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            pr[u][r] = zv[u][r];
            p_w[ws[u] * NR + r] = zv[u][r];
          }
      pull_halo(zero_nr, false);
      __syncthreads();
      // (3) iterations until the recursive residual converges
      for (;;) {
        // A
        PROF(0);
        spmv(qv);
        PROF(1);
        {
          double d[2 * NR];
#pragma unroll
          for (int k = 0; k < 2 * NR; ++k) d[k] = 0.0;
#pragma unroll
          for (int u = 0; u < RPT; ++u)
            if (valid[u])
#pragma unroll
              for (int r = 0; r < NR; ++r) {
                q_own[(size_t)(u * nth + tid) * NR + r] = qv[u][r];
                const c128 t = cmul(pr[u][r], qv[u][r]);
                d[2 * r] += t.x; d[2 * r + 1] += t.y;
              }
          __syncthreads();
          PROF(2);
          nodal_partial();
          PROF(3);
          reduce_push<2 * NR>(cluster, d, red, bank0, C, crank, prof_on ? prof_s + 16 : nullptr);
          PROF(4);
        }
        cluster.sync();  // barrier 1
        PROF(5);
        // B
        c128 alpha[NR];
        {
          double tot[2 * NR];
          bank_totals<2 * NR>(bank0, C, tot);
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const c128 pq = cmake(tot[2 * r], tot[2 * r + 1]);
            const bool brk = (pq.x == 0.0 && pq.y == 0.0) || !(isfinite(pq.x) && isfinite(pq.y));
            if (brk) act[r] = false;
            alpha[r] = act[r] ? cdiv_fast(rho[r], pq) : cmake(0.0, 0.0);
            if (act[r]) iters[r] += 1;
          }
        }
        PROF(6);
        nodal_combine(false, alpha);
        PROF(7);
#pragma unroll
        for (int u = 0; u < RPT; ++u)
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            xr[u][r] = cfma(alpha[r], pr[u][r], xr[u][r]);
            rv[u][r] = cfma(cneg(alpha[r]), qv[u][r], rv[u][r]);
          }
        __syncthreads();
        PROF(8);
        {
          double d[3 * NR];
#pragma unroll
          for (int k = 0; k < 3 * NR; ++k) d[k] = 0.0;
#pragma unroll
          for (int u = 0; u < RPT; ++u)
            if (valid[u])
#pragma unroll
              for (int r = 0; r < NR; ++r) {
                c128 z = cmul(di[u], rv[u][r]);
                if (K.aux) z = cadd(z, csub(w_s[n1[u] * NR + r], w_s[n0[u] * NR + r]));
                zv[u][r] = z;
                z_own[(size_t)(u * nth + tid) * NR + r] = z;
                const c128 t = cmul(rv[u][r], z);
                d[3 * r] += t.x; d[3 * r + 1] += t.y;
                d[3 * r + 2] += cabs2(rv[u][r]);
              }
          PROF(9);
          reduce_push<3 * NR>(cluster, d, red, bank1, C, crank, prof_on ? prof_s + 19 : nullptr);
          push_halo();  // z_own is complete: reduce_push holds a block barrier after the z stores
          PROF(10);
        }
        cluster.sync();  // barrier 2
        PROF(11);
        // D
        c128 beta[NR];
        bool still = false;
        {
          double tot[3 * NR];
          bank_totals<3 * NR>(bank1, C, tot);
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const c128 rho_new = cmake(tot[3 * r], tot[3 * r + 1]);
            if (act[r]) {
              rrn[r] = tot[3 * r + 2];
              if (rrn[r] <= D.tol2 * bb[r] || iters[r] >= D.max_it || !isfinite(rrn[r])) act[r] = false;
            }
            beta[r] = (act[r] && (rho[r].x != 0.0 || rho[r].y != 0.0)) ? cdiv_fast(rho_new, rho[r]) : cmake(0.0, 0.0);
            rho[r] = rho_new;
            still |= act[r];
          }
        }
        if (!still) break;
        PROF(12);
#pragma unroll
        for (int u = 0; u < RPT; ++u)
          if (valid[u])
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              pr[u][r] = cfma(beta[r], pr[u][r], zv[u][r]);
              p_w[ws[u] * NR + r] = pr[u][r];
            }
        PROF(13);
        halo_from_staging(beta);
        PROF(14);
        __syncthreads();
      }
      __syncthreads();
    }
    // results
#pragma unroll
    for (int u = 0; u < RPT; ++u)
      if (valid[u])
#pragma unroll
        for (int r = 0; r < NR; ++r) xg[(size_t)r * m + edge[u]] = xr[u][r];
    if (crank == 0 && tid == 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        int32_t *st = D.state + (s0 + r) * 4;
        st[ST_ACTIVE] = 0; st[ST_ITERS] = iters[r]; st[ST_CONV] = conv[r] ? 1 : 0; st[ST_REC] = 0;
        c128 *sc = D.scal + (size_t)(s0 + r) * NSCAL;
        sc[S_RR] = cmake(rrn[r], 0.0);
        sc[S_BB] = cmake(bb[r], 0.0);
      }
    }
    // the next job's first barrier orders the reuse of the shared-memory buffers across the cluster
  }
  if (prof_on)
    for (int k = 0; k < 24; ++k) atomicAdd((unsigned long long *)&K.prof[crank * 24 + k], (unsigned long long)prof_s[k]);
}

// ---------------------------------------------------------------- host side
void cluster_plan_free(System *S) {
  if (!S->cl_plan) return;
  delete (std::shared_ptr<ClusterPlanDev> *)S->cl_plan;  // the device arrays go when the last holder (system or cache) lets go
  S->cl_plan = nullptr;
}

// flag[r] = 1 when row r holds a value with a non-zero imaginary part in any of the matrices [first, first+n): a warp per row
__global__ void __launch_bounds__(256) k_row_complex(const c128 *__restrict__ vals, long long nnz, const int32_t *__restrict__ rowptr, int m,
                                                    int first, int n, uint8_t *__restrict__ flag) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= m) return;
  const int k0 = rowptr[r], len = rowptr[r + 1] - k0;
  bool any = false;
  for (int i = lane; i < len * n && !any; i += 32) {
    const int f = i / len, k = i - f * len;
    any = vals[(size_t)(first + f) * (size_t)nnz + k0 + k].y != 0.0;
  }
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) flag[r] = any ? 1 : 0;
}

namespace {
struct PlanCacheEntry {
  int device, m, C;
  long long nnz;
  uint64_t hash;
  bool aux;
  std::shared_ptr<ClusterPlanDev> plan;
};
std::mutex g_plan_mu;
std::vector<PlanCacheEntry> g_plan_cache;  // most recent first, at most 6
}  // namespace

static std::shared_ptr<ClusterPlanDev> plan_cache_find(int device, uint64_t hash, int m, long long nnz, int C, bool aux) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  for (auto &e : g_plan_cache)
    if (e.device == device && e.hash == hash && e.m == m && e.nnz == nnz && e.C == C && e.aux == aux) return e.plan;
  return nullptr;
}
static void plan_cache_put(int device, uint64_t hash, int m, long long nnz, int C, bool aux, std::shared_ptr<ClusterPlanDev> p) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plan_cache.insert(g_plan_cache.begin(), PlanCacheEntry{device, m, C, nnz, hash, aux, std::move(p)});
  if (g_plan_cache.size() > 6) g_plan_cache.pop_back();
}
void cluster_plan_cache_clear() {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plan_cache.clear();
}

template <typename T>
static int up(Ctx *c, ClusterPlanDev *P, const T **dst, const std::vector<T> &v) {
  T *d = nullptr;
  int rc = dev_upload(c, &d, v.data(), std::max<size_t>(v.size(), 1));
  if (rc) return rc;
  P->blocks.push_back(d);
  *dst = d;
  return EFB_OK;
}

static int cluster_plan_upload(System *S, ClusterPlanDev *P) {
  Ctx *c = S->ctx;
  ClusterPlanHost &H = P->h;
  // uploads read the host vectors asynchronously: give every vector at least one element and sync at the end
  auto pad = [](auto &v) { if (v.empty()) v.resize(1); };
  pad(H.row_edge); pad(H.row_ws); pad(H.row_n0); pad(H.row_n1); pad(H.blk_off); pad(H.blk_voff); pad(H.slot_src); pad(H.slot_col);
  pad(H.halo_ws); pad(H.halo_src); pad(H.push_row); pad(H.push_dst); pad(H.node_id); pad(H.n2e_ptr); pad(H.n2e_item); pad(H.nsrc_ptr); pad(H.nsrc_item);
  ClusterDev &d = P->d;
  d.C = H.C; d.mc = H.mc; d.aux = H.aux ? 1 : 0;
  d.max_own = H.max_own; d.max_w = H.max_w; d.max_my = std::max(H.max_my, 1); d.max_slots = H.max_slots;
  d.max_vunits = H.max_vunits;
  d.max_push = H.max_push;
  d.max_halo = H.max_halo; d.max_n2e = H.max_n2e; d.max_nsrc = H.max_nsrc;
  {
    // per-node lists as ELL (thread per node in the kernel, independent loads): own edges at a node, CTAs touching a node
    int ndeg, sdeg;
    plan_degrees(H, &ndeg, &sdeg);
    P->ndeg = ndeg;
    P->sdeg = sdeg;
    P->n2e_ell.assign((size_t)H.C * ndeg * d.max_my + 1, (uint16_t)0xffffu);
    P->nsrc_ell.assign((size_t)H.C * sdeg * d.max_my + 1, 0xffffffffu);
    for (int cta = 0; cta < H.C; ++cta) {
      const int32_t *I = &H.cta_info[(size_t)cta * CL_INFO_STRIDE];
      const int32_t *np_ = &H.n2e_ptr[I[CI_OFF_NODE] + cta], *sp_ = &H.nsrc_ptr[I[CI_OFF_NODE] + cta];
      for (int j = 0; j < I[CI_N_MY]; ++j) {
        for (int k = np_[j]; k < np_[j + 1]; ++k)
          P->n2e_ell[((size_t)cta * ndeg + (k - np_[j])) * d.max_my + j] = (uint16_t)H.n2e_item[I[CI_OFF_N2E] + k];
        for (int k = sp_[j]; k < sp_[j + 1]; ++k)
          P->nsrc_ell[((size_t)cta * sdeg + (k - sp_[j])) * d.max_my + j] = H.nsrc_item[I[CI_OFF_NSRC] + k];
      }
    }
    d.ndeg = ndeg;
    d.sdeg = sdeg;
    int rce;
    if ((rce = up(c, P, &d.n2e_ell, P->n2e_ell))) return rce;
    if ((rce = up(c, P, &d.nsrc_ell, P->nsrc_ell))) return rce;
  }
  d.prof = nullptr;
  if (const char *e = getenv("EDGEFEM_B200_CLUSTER_PROF"))
    if (atoi(e) > 0) {
      long long *pb = nullptr;
      int rcp = dev_alloc(c, &pb, (size_t)CL_MAX_C * 24);
      if (rcp) return rcp;
      P->blocks.push_back(pb);
      d.prof = pb;
    }
  int rc;
  if ((rc = up(c, P, &d.cta_info, H.cta_info))) return rc;
  if ((rc = up(c, P, &d.row_edge, H.row_edge))) return rc;
  if ((rc = up(c, P, &d.row_ws, H.row_ws))) return rc;
  if ((rc = up(c, P, &d.row_n0, H.row_n0))) return rc;
  if ((rc = up(c, P, &d.row_n1, H.row_n1))) return rc;
  if ((rc = up(c, P, &d.blk_off, H.blk_off))) return rc;
  if ((rc = up(c, P, &d.blk_voff, H.blk_voff))) return rc;
  if ((rc = up(c, P, &d.slot_src, H.slot_src))) return rc;
  if ((rc = up(c, P, &d.slot_col, H.slot_col))) return rc;
  if ((rc = up(c, P, &d.halo_ws, H.halo_ws))) return rc;
  if ((rc = up(c, P, &d.halo_src, H.halo_src))) return rc;
  if ((rc = up(c, P, &d.push_row, H.push_row))) return rc;
  if ((rc = up(c, P, &d.push_dst, H.push_dst))) return rc;
  if ((rc = up(c, P, &d.node_id, H.node_id))) return rc;
  if ((rc = up(c, P, &d.n2e_ptr, H.n2e_ptr))) return rc;
  if ((rc = up(c, P, &d.n2e_item, H.n2e_item))) return rc;
  if ((rc = up(c, P, &d.nsrc_ptr, H.nsrc_ptr))) return rc;
  if ((rc = up(c, P, &d.nsrc_item, H.nsrc_item))) return rc;
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

static int g_cl_resident[CL_MAX_C + 1] = {0};  // measured resident clusters per cluster size (cudaOccupancyMaxActiveClusters)

template <int NR, int RPT, int NTMAX = CL_THREADS_MAX>
static int launch_cluster(Ctx *c, const SolveDev &D, const ClusterDev &K, int nth, size_t smem, int first_matrix, int n_jobs, int groups,
                          int *job_counter, const c128 *b, c128 *x, int zero_x, int mr, int *clusters_out) {
  auto kern = k_cocg_cluster<NR, RPT, NTMAX>;
  EFB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3((unsigned)nth, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)K.C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)K.C, 1, 1);
  int max_clusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (e != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    return fail(c, EFB_ERR_LIMIT, "cluster solver: no cluster of %d CTAs x %zu B of shared memory can be resident (%s)", K.C, smem,
                e != cudaSuccess ? cudaGetErrorString(e) : "0 clusters");
  }
  g_cl_resident[K.C] = max_clusters;
  const int n_clusters = std::max(1, std::min(n_jobs, max_clusters));
  cfg.gridDim = dim3((unsigned)(n_clusters * K.C), 1, 1);
  *clusters_out = n_clusters;
  EFB_CUDA(c, cudaLaunchKernelEx(&cfg, kern, D, K, first_matrix, n_jobs, groups, job_counter, b, x, zero_x, mr));
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

// Used whenever the system is symmetric (COCG), whole (not row-partitioned) and a split into <= 8 CTAs fits in shared
// memory.  EDGEFEM_B200_CLUSTER=0 disables it, =C forces the cluster size.
int run_krylov_cluster(SolvePlan &P, const efb_solve_opts *o, bool zero_x, bool *ran) {
  System *S = P.S;
  Ctx *c = S->ctx;
  *ran = false;
  int forced = -1;
  if (const char *e = getenv("EDGEFEM_B200_CLUSTER")) forced = atoi(e);
  if (forced == 0) return EFB_OK;
  if (S->m != S->m_global || S->m > 65535 * CL_MAX_C) return EFB_OK;
  if ((int)S->h_rowptr.size() != S->m + 1) return EFB_OK;
  const bool want_aux = P.aux && (int)S->h_edge_nodes.size() == 2 * S->m;
  if (P.aux && !want_aux) return EFB_OK;
  int dev_smem = 0;
  EFB_CUDA(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  SubTrace st;
  const int n_rhs = S->n_rhs;
  // One right-hand side per job.  Two per job (sharing the matrix reads and the barriers) is templated but NOT enabled:
  // next to a complex matrix slice it only fits without the auxiliary space, it was measured slower there (SpMV 2.3x for
  // two right-hand sides: 32-byte gathers halve the distinct bank groups), and it does not converge: with 440 bytes of
  // register spills a search-direction entry comes back as denormal garbage right after `p = z + beta p` on zeros (the
  // values printed before the update are correct) -- a spill-slot defect of that instantiation that was not resolved.
  const int nr_want = 1;
  ClusterPlanDev *PL = S->cl_plan ? ((std::shared_ptr<ClusterPlanDev> *)S->cl_plan)->get() : nullptr;
  const uint8_t *dirp = (int)S->h_dir.size() == S->m ? S->h_dir.data() : nullptr;
  auto fits = [&](const ClusterPlanHost &H, int nr, int *rpt_out, int *nth_out) {
    if (cluster_smem_bytes(nr, H) > (size_t)dev_smem) return false;
    for (int rpt : {1, 2, 4}) {
      const int nth = std::max(64, ((H.max_own + rpt - 1) / rpt + 31) / 32 * 32);
      if (nth <= (rpt == 1 ? CL_THREADS_WIDE : CL_THREADS_MAX)) {
        *rpt_out = rpt;
        *nth_out = nth;
        return true;
      }
    }
    return false;
  };
  int nr = 0, rpt = 1, nth = 0;
  const int n_matrix = P.n_matrix;
  // free unknowns / free entries: cheap size estimate of the smallest cluster that can hold the system
  int mc = 0;
  long long nnz_free = 0;
  for (int r = 0; r < S->m; ++r) {
    if (dirp && dirp[r]) continue;
    ++mc;
    for (int k = S->h_rowptr[r]; k < S->h_rowptr[r + 1]; ++k) nnz_free += !(dirp && dirp[S->h_colidx[k]]);
  }
  if (mc == 0) return EFB_OK;
  // per CTA: matrix slice (+10 % padding; 10 B per slot in real blocks, 18 B in complex ones) + window (own + ~0.8 own of
  // halo) + q, z + nodal arrays + lists
  auto est_bytes = [&](int C, double cplx_frac) {
    const double own = (double)mc / C, halo = C > 1 ? 0.8 * own + 64 : 0.0;
    return nnz_free * 1.1 * (10.0 + 8.0 * cplx_frac) / C + (own + halo) * 16.0 + own * 32.0 + (want_aux ? own * 0.4 * 128.0 : 0.0) + halo * 6.0 + 6000.0;
  };
  // resident clusters of C CTAs (one CTA per SM; a cluster lives inside one GPC of ~18 SMs): measured once a launch of that
  // size has happened, estimated before
  auto resident_est = [&](int C) {
    if (g_cl_resident[C] > 0) return g_cl_resident[C];
    const int per_gpc = std::max(1, 18 / C), gpcs = std::max(1, c->sm_count / 18);
    return std::max(1, std::min(c->sm_count / C, per_gpc * gpcs) - (C >= 5 ? 1 : 0));
  };
  // relative cost of a job on C CTAs: half of the 8-CTA iteration is barriers and reductions (fixed), half scales with the rows per CTA
  auto job_cost = [](int C) { return C >= 8 ? 1.0 : 0.5 + 0.75 * (8.0 / C - 1.0) + 0.5; };  // measured: 6 CTAs 1.22-1.25, 7 CTAs ~1.2 (two rounds of the same 15)
  const int n_jobs_all = n_matrix * n_rhs;
  auto batch_cost = [&](int C) { return std::ceil((double)n_jobs_all / resident_est(C)) * job_cost(C); };
  const int nn1 = want_aux ? S->n_node : 0;
  const bool one_cta_ok = S->m <= 16384 && ((size_t)mc * 16 + (size_t)nn1 * 16 + 4096) <= (size_t)dev_smem;
  // Latency or throughput?  A cluster job is ~10x faster than a one-CTA job of k_cocg_small (measured on WR-90:
  // 7.5 us per rhs-iteration on 8 SMs against 85 us per two-rhs iteration on one), but only ~15 clusters of 8 (~24 of 6) are
  // resident against 148 CTAs, and a one-CTA job carries both right-hand sides: big batches (the 256-point sweep)
  // stay on the one-CTA kernel when it applies, small ones (one frequency, the shards of a strong-scaled sweep) run
  // here.  Systems the one-CTA kernel cannot take (p does not fit in its shared memory) always run here.  Decided
  // BEFORE the value scan and any plan build (first with the most favourable storage, all rows real).
  // The one-CTA kernel in the same unit (one 8-CTA cluster job, 2.3 ms on WR-90): a two-rhs job is ~10; with up to half as
  // many matrices as SMs every matrix is split into one-rhs jobs and the launch lasts ~5.4 (64 matrices: 12.5 ms), between
  // sm_count / 2 and sm_count only the longest are split (128 matrices: 22.3 ms), beyond that it runs in rounds.
  double rounds_1;
  {
    const double half = 0.5 * c->sm_count;
    if (n_rhs < 2) rounds_1 = std::ceil((double)n_matrix / c->sm_count) * 5.4;
    else if (n_matrix <= half) rounds_1 = 5.4;  // resident one-rhs jobs: 12.5 ms
    else if (n_matrix <= c->sm_count) rounds_1 = 7.8 + 2.2 * (n_matrix - half) / half;  // 80 matrices 18.1 ms ... 128: 22.3 ms
    else rounds_1 = std::ceil((double)n_matrix / c->sm_count) * 10.0;
  }
  auto candidates = [&](double cplx_frac) {
    std::vector<int> cs;
    if (forced > 0) {
      cs.push_back(std::min(forced, CL_MAX_C));
      return cs;
    }
    for (int C = 1; C <= CL_MAX_C; ++C)
      if (est_bytes(C, cplx_frac) <= 0.97 * dev_smem && (mc + C - 1) / C <= 4 * CL_THREADS_MAX) cs.push_back(C);
    std::stable_sort(cs.begin(), cs.end(), [&](int x, int y) {
      const double cx = batch_cost(x), cy = batch_cost(y);
      return cx != cy ? cx < cy : x > y;
    });
    return cs;
  };
  {
    const std::vector<int> c0 = candidates(0.0);
    if (c0.empty()) return EFB_OK;  // does not fit in 8 SMs: the other solver paths take it
    if (forced < 0 && one_cta_ok && batch_cost(c0[0]) > rounds_1) return EFB_OK;
  }
  // Rows with a value that is not real in any matrix of the range (a lossless system: only the rows of the port faces):
  // their ELL blocks store complex128, the others doubles.
  std::vector<uint8_t> row_cplx((size_t)S->m, 1);
  const bool no_real = getenv("EDGEFEM_B200_CLUSTER_NO_REAL") != nullptr;
  long long nnz_cplx = 0;
  if (!no_real) {
    uint8_t *d_flag = nullptr;
    int rcf = dev_alloc(c, &d_flag, (size_t)S->m);
    if (rcf) return rcf;
    k_row_complex<<<(unsigned)((S->m + 7) / 8), 256, 0, c->stream>>>(S->d_vals, (long long)S->nnz, S->d_rowptr, S->m, P.first_matrix, n_matrix, d_flag);
    EFB_CHECK_LAUNCH(c);
    EFB_CUDA(c, cudaMemcpyAsync(row_cplx.data(), d_flag, (size_t)S->m, cudaMemcpyDeviceToHost, c->stream));
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    dfree(d_flag);
  }
  for (int r = 0; r < S->m; ++r)
    if (row_cplx[r] && !(dirp && dirp[r])) nnz_cplx += S->h_rowptr[r + 1] - S->h_rowptr[r];
  const double cplx_frac = std::min(1.0, (double)nnz_cplx / std::max<long long>(1, nnz_free));
  uint64_t flags_hash = 1469598103934665603ull;
  for (int r = 0; r < S->m; ++r) flags_hash = (flags_hash ^ row_cplx[r]) * 1099511628211ull;
  const std::vector<int> cand = candidates(cplx_frac);
  if (cand.empty()) return EFB_OK;
  if (forced < 0 && one_cta_ok && batch_cost(cand[0]) > rounds_1) return EFB_OK;
  st.mark("  cluster solver: value scan + shape choice");
  const bool reuse = !S->cl_dirty && PL && PL->h.aux == want_aux && PL->flags_hash == flags_hash && PL->h.C == cand[0];
  if (!reuse) {
    cluster_plan_free(S);
    PL = nullptr;
    // plans are shared between systems with the same pattern, Dirichlet flags, gradient and real/complex rows (a driver that
    // creates one system per call -- calculate_sparams_eigenmode in a frequency loop -- builds the plan once)
    uint64_t hsh = flags_hash;
    auto mix = [&hsh](const void *ptr, size_t bytes) {
      const uint32_t *w = (const uint32_t *)ptr;
      for (size_t i = 0; i < bytes / 4; ++i) {
        hsh ^= w[i];
        hsh *= 1099511628211ull;
      }
      const unsigned char *t = (const unsigned char *)ptr + (bytes / 4) * 4;
      for (size_t i = 0; i < bytes % 4; ++i) {
        hsh ^= t[i];
        hsh *= 1099511628211ull;
      }
    };
    mix(S->h_rowptr.data(), S->h_rowptr.size() * 4);
    mix(S->h_colidx.data(), S->h_colidx.size() * 4);
    if (dirp) mix(dirp, (size_t)S->m);
    if (want_aux) mix(S->h_edge_nodes.data(), S->h_edge_nodes.size() * 4);
    std::shared_ptr<ClusterPlanDev> sp;
    for (size_t ci = 0; ci < cand.size() && !sp; ++ci) {
      const int C = cand[ci];
      sp = plan_cache_find(c->device, hsh, S->m, (long long)S->nnz, C, want_aux);
      if (sp) break;
      ClusterPlanHost H;
      if (!build_cluster_plan(S->m, S->h_rowptr.data(), S->h_colidx.data(), dirp, want_aux ? S->n_node : 0,
                              want_aux ? S->h_edge_nodes.data() : nullptr, C, H, row_cplx.data()))
        continue;
      int r1, t1;
      if (!fits(H, 1, &r1, &t1)) continue;
      sp = std::make_shared<ClusterPlanDev>();
      sp->h = std::move(H);
      sp->flags_hash = flags_hash;
      int rc = cluster_plan_upload(S, sp.get());
      if (rc) return rc;
      plan_cache_put(c->device, hsh, S->m, (long long)S->nnz, C, want_aux, sp);
      st.mark("  cluster plan (RCM, partition, windows, lists)");
      if (st.on)
        for (auto &tm : sp->h.timing) fprintf(stderr, "[edgefem-b200 trace]     .     plan: %-32s %.3f ms\n", tm.first, tm.second);
    }
    if (!sp) return EFB_OK;  // does not fit: the other solver paths take it
    S->cl_plan = new std::shared_ptr<ClusterPlanDev>(sp);
    PL = sp.get();
    S->cl_dirty = false;
  }
  nr = nr_want;
  if (!fits(PL->h, nr, &rpt, &nth)) {
    nr = 1;
    if (!fits(PL->h, nr, &rpt, &nth)) return EFB_OK;
  }
  const size_t smem = cluster_smem_bytes(nr, PL->h);
  const int groups = n_rhs / nr;
  const int n_jobs = P.n_matrix * groups;
  {
    // queue order = longest expected job first (iteration counts of the previous solve, else frequency)
    std::vector<double> w((size_t)n_jobs, 0.0);
    const bool have_it = (int)S->last_iters.size() == S->n_sys;
    const bool have_om = (int)S->last_omega.size() == S->n_matrix;
    for (int j = 0; j < n_jobs; ++j) {
      const int f = P.first_matrix + j / groups, s0 = f * n_rhs + (j % groups) * nr;
      if (have_it)
        for (int k = 0; k < nr; ++k) w[j] = std::max(w[j], (double)S->last_iters[s0 + k]);
      else if (have_om)
        w[j] = S->last_omega[f];
    }
    std::vector<int32_t> order((size_t)n_jobs);
    for (int j = 0; j < n_jobs; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] > w[b]; });
    if ((size_t)n_jobs + 1 > (size_t)S->n_sys + 1) return fail(c, EFB_ERR_STATE, "cluster solver: job queue overflow");
    S->h_job_order.resize((size_t)n_jobs + 1);
    S->h_job_order[0] = 0;
    for (int j = 0; j < n_jobs; ++j) S->h_job_order[1 + j] = order[j];
    EFB_CUDA(c, cudaMemcpyAsync(S->d_job, S->h_job_order.data(), S->h_job_order.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  }
  if (!S->ev_s0) {
    EFB_CUDA(c, cudaEventCreate(&S->ev_s0));
    EFB_CUDA(c, cudaEventCreate(&S->ev_s1));
  }
  if (PL->d.prof) EFB_CUDA(c, cudaMemsetAsync(PL->d.prof, 0, (size_t)CL_MAX_C * 24 * sizeof(long long), c->stream));
  EFB_CUDA(c, cudaEventRecord(S->ev_s0, c->stream));
  const int mr = o->max_restarts > 0 ? o->max_restarts : 3;
  int rc = EFB_OK, ncl = 0;
#define EFB_CL(NRV, RPTV)                                                                                                             \
  rc = launch_cluster<NRV, RPTV>(c, P.D, PL->d, nth, smem, P.first_matrix, n_jobs, groups, S->d_job, S->d_b, S->d_x, zero_x ? 1 : 0, mr, &ncl)
  if (rpt == 1 && nth > CL_THREADS_MAX)
    rc = launch_cluster<1, 1, CL_THREADS_WIDE>(c, P.D, PL->d, nth, smem, P.first_matrix, n_jobs, groups, S->d_job, S->d_b, S->d_x, zero_x ? 1 : 0, mr, &ncl);
  else if (rpt == 1) EFB_CL(1, 1); else if (rpt == 2) EFB_CL(1, 2); else EFB_CL(1, 4);
#undef EFB_CL
  if (rc) return rc;
  EFB_CUDA(c, cudaEventRecord(S->ev_s1, c->stream));
  if (PL->d.prof) {  // EDGEFEM_B200_CLUSTER_PROF=1: cycles of thread 0 of every CTA rank between the phase marks
    std::vector<long long> hp((size_t)CL_MAX_C * 24);
    EFB_CUDA(c, cudaMemcpyAsync(hp.data(), PL->d.prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    static const char *names[24] = {"D->A", "spmv", "q store+bar", "nodal partial", "reduce+push 1", "cluster barrier 1", "totals+alpha",
                                    "nodal combine", "x,r update+bar", "z+dots", "reduce+push 2", "cluster barrier 2", "totals+beta",
                                    "p update", "halo pull", "", " r1: shuffles", " r1: block barrier", " r1: sum + push", " r2: shuffles", " r2: block barrier", " r2: sum + push", "", ""};
    for (int k = 0; k < 22; ++k) {
      if (!names[k][0]) continue;
      fprintf(stderr, "[cluster prof] %-18s", names[k]);
      for (int r = 0; r < PL->h.C; ++r) fprintf(stderr, " %10lld", hp[(size_t)r * 24 + k]);
      fprintf(stderr, "\n");
    }
  }
  S->small_timed = true;
  S->last_cluster_c = PL->h.C;
  S->last_cluster_nr = nr;
  S->last_cluster_n = ncl;
  *ran = true;
  return EFB_OK;
}

}  // namespace efb

using namespace efb;

// ---------------------------------------------------------------- debug / test hooks (host only, no GPU needed)
extern "C" {

struct efb_cluster_plan_dbg {
  ClusterPlanHost h;
};

int efb_debug_cluster_plan_build(int32_t m, const int32_t *rowptr, const int32_t *colidx, const uint8_t *dir, int32_t n_node,
                                 const int32_t *edge_nodes, const uint8_t *row_complex, int32_t C, void **out) {
  if (!rowptr || !colidx || !out || m <= 0) return EFB_ERR_INVALID;
  auto *d = new efb_cluster_plan_dbg();
  if (!build_cluster_plan(m, rowptr, colidx, dir, n_node, edge_nodes, C, d->h, row_complex)) {
    set_global_error("efb_debug_cluster_plan_build: " + d->h.error);
    delete d;
    return EFB_ERR_LIMIT;
  }
  *out = d;
  return EFB_OK;
}

void efb_debug_cluster_plan_free(void *p) { delete (efb_cluster_plan_dbg *)p; }

// copies the named array (converted to int64) into buf (capacity cap); returns its length, or -1 for an unknown name
void efb_clear_caches(void) { cluster_plan_cache_clear(); }

int64_t efb_debug_cluster_plan_get(void *p, const char *name, int64_t *buf, int64_t cap) {
  if (!p || !name) return -1;
  const ClusterPlanHost &H = ((efb_cluster_plan_dbg *)p)->h;
  std::vector<int64_t> v;
  auto put = [&](const auto &src) { v.assign(src.begin(), src.end()); };
  const std::string n = name;
  if (n == "dims") v = {H.C, H.mc, H.m, H.aux ? 1 : 0, H.max_own, H.max_w, H.max_my, H.max_slots, H.max_halo,
                        (int64_t)cluster_smem_bytes(1, H), (int64_t)cluster_smem_bytes(2, H), H.max_vunits};
  else if (n == "c_orig") put(H.c_orig);
  else if (n == "cta_info") put(H.cta_info);
  else if (n == "row_edge") put(H.row_edge);
  else if (n == "row_ws") put(H.row_ws);
  else if (n == "row_n0") put(H.row_n0);
  else if (n == "row_n1") put(H.row_n1);
  else if (n == "blk_off") put(H.blk_off);
  else if (n == "blk_voff") put(H.blk_voff);
  else if (n == "slot_src") put(H.slot_src);
  else if (n == "slot_col") put(H.slot_col);
  else if (n == "halo_ws") put(H.halo_ws);
  else if (n == "push_row") put(H.push_row);
  else if (n == "push_dst") put(H.push_dst);
  else if (n == "halo_src") put(H.halo_src);
  else if (n == "node_id") put(H.node_id);
  else if (n == "n2e_ptr") put(H.n2e_ptr);
  else if (n == "n2e_item") put(H.n2e_item);
  else if (n == "nsrc_ptr") put(H.nsrc_ptr);
  else if (n == "nsrc_item") put(H.nsrc_item);
  else return -1;
  if (buf)
    for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) buf[i] = v[i];
  return (int64_t)v.size();
}

}  // extern "C"
