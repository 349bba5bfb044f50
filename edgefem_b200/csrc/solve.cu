// K3 CSR complex128 SpMV (batched: many matrices on one pattern, several right-hand sides per
// matrix) and K4 fused Krylov phases: COCG (complex symmetric A) and BiCGSTAB (general A) with
// Jacobi or Jacobi + nodal gradient-space ("auxiliary space", Hiptmair) preconditioning.
// Replaces Eigen's BiCGSTAB/IncompleteLUT/SparseLU behind solve_linear (src/solver.cpp:35-193);
// the contract kept is SolveResult's (iterations, TRUE relative residual, converged).
// Reductions are deterministic: per-block partials + last-block finalisation in fixed order;
// scalars (alpha, beta, omega, rho) never leave the device inside the iteration loop.
#include <algorithm>
#include <cstdlib>

#include <queue>

#include "common.cuh"
#include "solve_internal.cuh"
#include "spmv_tma.cuh"

namespace efb {

constexpr int VEC_THREADS = 256;

// ---------------------------------------------------------------- reductions
template <int NV>
__device__ bool reduce_and_ticket(double (&v)[NV], double *partial_sys, unsigned *counter_sys, double (&tot)[NV]) {
  __shared__ double sh[4][32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = v[k];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) sh[k][wid] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = 0.0;
      for (int w = 0; w < nw; ++w) a += sh[k][w];
      partial_sys[blockIdx.x * 4 + k] = a;
    }
    __threadfence();
    const unsigned t = atomicAdd(counter_sys, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double loc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) loc[k] = 0.0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; ++k) loc[k] += __ldcg(&partial_sys[b * 4 + k]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = loc[k];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) sh[k][wid] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = 0.0;
      for (int w = 0; w < nw; ++w) a += sh[k][w];
      tot[k] = a;
    }
    *counter_sys = 0u;
    return true;
  }
  return false;
}

__device__ __forceinline__ c128 *scal_of(const SolveDev &D, int s) { return D.scal + (size_t)s * NSCAL; }
__device__ __forceinline__ double *partial_of(const SolveDev &D, int s) { return D.partial + (size_t)s * RED_MAX_BLOCKS * 4; }

// ---------------------------------------------------------------- K3: SpMV
// y[s] = A[f] x[s] for the NR right-hand sides s = f*n_rhs + z*NR + {0..NR-1}.
// LPR lanes cooperate on a row (coalesced 16-byte value loads, warp-shuffle row reduction).
// DOT: 0 none | 1: scal[slot0] = sum w.y (unconjugated) | 2: scal[slot0] = sum conj(w) y
//      3: scal[slot0] = sum conj(y) w , scal[slot1] = sum |y|^2
constexpr int SPMV_UNR = 2;  // rows in flight per lane group (memory-level parallelism)

__device__ __forceinline__ c128 ldg_stream(const c128 *p) {  // streaming 16-byte load: read-only path, do not pollute L1
  c128 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_stream(const int32_t *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int LPR, int NR, int DOT>
__global__ void __launch_bounds__(256)
k_spmv(SolveDev D, int first_matrix, const c128 *__restrict__ x, c128 *__restrict__ y, const c128 *__restrict__ wv,
       int slot0, int slot1, int use_active) {
  const int f = first_matrix + blockIdx.y;
  const int s0 = f * D.n_rhs + blockIdx.z * NR;
  bool any = false;
#pragma unroll
  for (int r = 0; r < NR; ++r) any |= (!use_active) || D.state[(s0 + r) * 4 + ST_ACTIVE];
  if (!any) return;
  const c128 *__restrict__ av = D.vals + (size_t)f * D.nnz;
  constexpr int RPB = 256 / LPR;
  const int lane = threadIdx.x % LPR, rl = threadIdx.x / LPR;
  double dots[NR][4];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) dots[r][k] = 0.0;
  const int ntile = (D.m + RPB - 1) / RPB;
  for (int tile = blockIdx.x * SPMV_UNR; tile < ntile; tile += gridDim.x * SPMV_UNR) {
    int row[SPMV_UNR], kb[SPMV_UNR], ke[SPMV_UNR];
    int len = 0;
#pragma unroll
    for (int u = 0; u < SPMV_UNR; ++u) {
      row[u] = (tile + u) * RPB + rl;
      kb[u] = 0; ke[u] = 0;
      if (tile + u < ntile && row[u] < D.m) {
        kb[u] = __ldg(&D.rowptr[row[u]]);
        ke[u] = __ldg(&D.rowptr[row[u] + 1]);
      }
      len = max(len, ke[u] - kb[u]);
    }
    c128 acc[SPMV_UNR][NR];
#pragma unroll
    for (int u = 0; u < SPMV_UNR; ++u)
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[u][r] = cmake(0.0, 0.0);
    for (int off = lane; off < len; off += LPR) {
      c128 a[SPMV_UNR];
      int c[SPMV_UNR];
#pragma unroll
      for (int u = 0; u < SPMV_UNR; ++u) {  // issue every independent load first
        const int k = kb[u] + off;
        const bool in = k < ke[u];
        a[u] = in ? ldg_stream(&av[k]) : cmake(0.0, 0.0);
        c[u] = in ? ldg_stream(&D.colidx[k]) : 0;
      }
#pragma unroll
      for (int u = 0; u < SPMV_UNR; ++u)
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[u][r] = cfma(a[u], __ldg(&x[(size_t)(s0 + r) * D.m + c[u]]), acc[u][r]);
    }
#pragma unroll
    for (int u = 0; u < SPMV_UNR; ++u)
#pragma unroll
      for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
          acc[u][r].x += __shfl_xor_sync(0xffffffffu, acc[u][r].x, o);
          acc[u][r].y += __shfl_xor_sync(0xffffffffu, acc[u][r].y, o);
        }
      }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < SPMV_UNR; ++u) {
        if (!(tile + u < ntile && row[u] < D.m)) continue;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const size_t idx = (size_t)(s0 + r) * D.m + row[u];
          y[idx] = acc[u][r];
          if (DOT == 1) {
            const c128 q = cmul(wv[idx], acc[u][r]);
            dots[r][0] += q.x; dots[r][1] += q.y;
          } else if (DOT == 2) {
            const c128 q = cmulconj(wv[idx], acc[u][r]);
            dots[r][0] += q.x; dots[r][1] += q.y;
          } else if (DOT == 3) {
            const c128 q = cmulconj(acc[u][r], wv[idx]);
            dots[r][0] += q.x; dots[r][1] += q.y;
            dots[r][2] += cabs2(acc[u][r]);
          }
        }
      }
    }
  }
  if (DOT != 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int s = s0 + r;
      double tot[4];
      if (reduce_and_ticket<4>(dots[r], partial_of(D, s), D.counter + s, tot)) {
        c128 *sc = scal_of(D, s);
        sc[slot0] = cmake(tot[0], tot[1]);
        if (DOT == 3) sc[slot1] = cmake(tot[2], 0.0);
      }
    }
  }
}

// CSR-stream SpMV (large matrices): a warp owns a row-aligned chunk of <= SPMV_STREAM_W entries and at
// most 32 rows.  Phase 1 streams the chunk's values/columns with fully coalesced, independent
// 16-byte loads (8 in flight per lane), gathers x and parks the products in shared memory;
// phase 2 gives one lane per row and sums that row's products in order (deterministic).
template <int NR, int DOT>
__global__ void __launch_bounds__(256)
k_spmv_stream(SolveDev D, const int32_t *__restrict__ sp_chunk, int n_chunks, int first_matrix, const c128 *__restrict__ x,
              c128 *__restrict__ y, const c128 *__restrict__ wv, int slot0, int slot1, int use_active) {
  __shared__ c128 prod[8][NR][SPMV_SLOTS];
  const int f = first_matrix + blockIdx.y;
  const int s0 = f * D.n_rhs + blockIdx.z * NR;
  bool any = false;
#pragma unroll
  for (int r = 0; r < NR; ++r) any |= (!use_active) || D.state[(s0 + r) * 4 + ST_ACTIVE];
  if (!any) return;
  const c128 *__restrict__ av = D.vals + (size_t)f * D.nnz;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double dots[NR][4];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) dots[r][k] = 0.0;
  for (int ch = blockIdx.x * 8 + wid; ch < n_chunks; ch += gridDim.x * 8) {
    const int r0 = __ldg(&sp_chunk[ch]), r1 = __ldg(&sp_chunk[ch + 1]);
    const int nrow = r1 - r0;
    int rs = 0, re = 0;
    if (lane <= nrow) rs = __ldg(&D.rowptr[r0 + lane]);
    const int k0 = __shfl_sync(0xffffffffu, rs, 0);
    const int k1 = __shfl_sync(0xffffffffu, rs, nrow);
    re = __shfl_down_sync(0xffffffffu, rs, 1);
    c128 a[SPMV_STREAM_W / 32];
    int c[SPMV_STREAM_W / 32];
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      const int k = k0 + lane + 32 * j;
      const bool in = k < k1;
      a[j] = in ? ldg_stream(&av[k]) : cmake(0.0, 0.0);
      c[j] = in ? ldg_stream(&D.colidx[k]) : -1;
    }
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      if (c[j] >= 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) prod[wid][r][spmv_slot(lane + 32 * j)] = cmul(a[j], __ldg(&x[(size_t)(s0 + r) * D.m + c[j]]));
      }
    }
    __syncwarp();
    if (lane < nrow) {
      const int row = r0 + lane;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        c128 acc = cmake(0.0, 0.0);
        for (int k = rs - k0; k < re - k0; ++k) acc = cadd(acc, prod[wid][r][spmv_slot(k)]);
        const size_t idx = (size_t)(s0 + r) * D.m + row;
        y[idx] = acc;
        if (DOT == 1) {
          const c128 q = cmul(wv[idx], acc);
          dots[r][0] += q.x; dots[r][1] += q.y;
        } else if (DOT == 2) {
          const c128 q = cmulconj(wv[idx], acc);
          dots[r][0] += q.x; dots[r][1] += q.y;
        } else if (DOT == 3) {
          const c128 q = cmulconj(acc, wv[idx]);
          dots[r][0] += q.x; dots[r][1] += q.y;
          dots[r][2] += cabs2(acc);
        }
      }
    }
    __syncwarp();
  }
  if (DOT != 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int s = s0 + r;
      double tot[4];
      if (reduce_and_ticket<4>(dots[r], partial_of(D, s), D.counter + s, tot)) {
        c128 *sc = scal_of(D, s);
        sc[slot0] = cmake(tot[0], tot[1]);
        if (DOT == 3) sc[slot1] = cmake(tot[2], 0.0);
      }
    }
  }
}

// CSR-stream SpMV with the matrix stream moved by the TMA engine (the default for large matrices).  The SpMV is bound
// by the L1 data pipe (LSU wavefronts: 82 % of peak in ncu), not by HBM: the x gather costs one wavefront per distinct
// 128-byte line of every load, the products and row sums go through shared memory, and in k_spmv_stream the value and
// column streams use the same pipe (a cp.async variant of the stream was measured no faster: LDGSTS still goes through
// L1).  Here lane 0 of a warp
// issues two bulk copies per chunk (cp.async.bulk global -> shared: <= 4 KB of values, <= 1 KB of columns as the
// 16-byte-aligned superset) that complete on an mbarrier; the copies bypass L1 entirely, run one chunk ahead of the
// warp (2-stage ring) and leave the LSU to the x gather, the in-place products and the in-order row sums.
// Same chunks, products and summation order as k_spmv_stream: bit-identical y.
template <int DOT>
__global__ void __launch_bounds__(256, 2)
k_spmv_tma(SolveDev D, const int32_t *__restrict__ sp_chunk, int n_chunks, int first_matrix, const c128 *__restrict__ x,
           c128 *__restrict__ y, const c128 *__restrict__ wv, int slot0, int slot1, int use_active) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  const int f = first_matrix + blockIdx.y;
  const int s0 = f * D.n_rhs;
  if (use_active && !D.state[s0 * 4 + ST_ACTIVE]) return;
  const c128 *__restrict__ av = D.vals + (size_t)f * D.nnz;
  const c128 *__restrict__ xs = x + (size_t)s0 * D.m;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char *wbase = tma_smem + (size_t)wid * 2 * SPMV_TMA_STAGE;
  const unsigned wbase_s = (unsigned)__cvta_generic_to_shared(wbase);
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(tma_smem + (size_t)8 * 2 * SPMV_TMA_STAGE) + (unsigned)wid * 16u;
  if (lane == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  double dots[4] = {0.0, 0.0, 0.0, 0.0};
  const int stride = gridDim.x * 8;
  int ch = blockIdx.x * 8 + wid;

  // lane 0: one arrive.expect_tx and two bulk copies per chunk (an exhausted descriptor copies nothing)
  auto issue = [&](int stage, const ChunkDesc &d) {
    if (lane == 0 && d.nrow > 0) {
      const unsigned sb = wbase_s + (unsigned)stage * SPMV_TMA_STAGE;
      const unsigned bar = bar_s + 8u * (unsigned)stage;
      const int ka = d.k0 & ~3;
      const unsigned vbytes = (unsigned)(d.k1 - d.k0) * 16u;
      const unsigned cbytes = (unsigned)((d.k1 - ka + 3) >> 2) * 16u;
      fence_proxy_async();  // the stage was last touched by ordinary shared-memory accesses of this warp
      mbar_expect_tx(bar, vbytes + cbytes);
      if (vbytes) bulk_g2s(sb, av + d.k0, vbytes, bar);
      if (cbytes) bulk_g2s(sb + SPMV_VAL_BYTES, D.colidx + ka, cbytes, bar);
    }
  };

  ChunkDesc cur = load_chunk_desc(sp_chunk, D.rowptr, ch, n_chunks);
  ChunkDesc nxt = load_chunk_desc(sp_chunk, D.rowptr, ch + stride, n_chunks);
  int rs_cur = (lane <= cur.nrow && cur.nrow > 0) ? __ldg(&D.rowptr[cur.r0 + lane]) : 0;
  issue(0, cur);
  int st = 0;
  unsigned phase0 = 0, phase1 = 0;  // parity each stage's barrier completes next
  for (; ch < n_chunks; ch += stride) {
    issue(st ^ 1, nxt);
    const ChunkDesc nn = load_chunk_desc(sp_chunk, D.rowptr, ch + 2 * stride, n_chunks);
    const int rs_nxt = (lane <= nxt.nrow && nxt.nrow > 0) ? __ldg(&D.rowptr[nxt.r0 + lane]) : 0;
    {
      const unsigned bar = bar_s + 8u * (unsigned)st;
      const unsigned par = st ? phase1 : phase0;
      while (!mbar_try_wait(bar, par)) {
      }
      if (st) phase1 ^= 1u; else phase0 ^= 1u;
    }
    unsigned char *sb = wbase + (size_t)st * SPMV_TMA_STAGE;
    c128 *sv = (c128 *)sb;
    const int32_t *sc = (const int32_t *)(sb + SPMV_VAL_BYTES) + (cur.k0 & 3);
    const int n = cur.k1 - cur.k0;
    int c[SPMV_STREAM_W / 32];
    c128 xv[SPMV_STREAM_W / 32];
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      const int i = lane + 32 * j;
      c[j] = (i < n) ? sc[i] : -1;
    }
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) xv[j] = (c[j] >= 0) ? __ldg(&xs[c[j]]) : cmake(0.0, 0.0);
    // products in place, at the skewed slots: every lane reads its values first (slot(i) >= i would run into values
    // another lane has not read yet)
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      const int i = lane + 32 * j;
      if (c[j] >= 0) xv[j] = cmul(sv[i], xv[j]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      if (c[j] >= 0) sv[spmv_slot(lane + 32 * j)] = xv[j];
    }
    __syncwarp();
    const int re_cur = __shfl_down_sync(0xffffffffu, rs_cur, 1);
    if (lane < cur.nrow) {
      const int row = cur.r0 + lane;
      const int rs = rs_cur - cur.k0, re = re_cur - cur.k0;
      c128 acc = cmake(0.0, 0.0);
      for (int k = rs; k < re; ++k) acc = cadd(acc, sv[spmv_slot(k)]);
      const size_t idx = (size_t)s0 * D.m + row;
      y[idx] = acc;
      if (DOT == 1) {
        const c128 q = cmul(wv[idx], acc);
        dots[0] += q.x; dots[1] += q.y;
      } else if (DOT == 2) {
        const c128 q = cmulconj(wv[idx], acc);
        dots[0] += q.x; dots[1] += q.y;
      } else if (DOT == 3) {
        const c128 q = cmulconj(acc, wv[idx]);
        dots[0] += q.x; dots[1] += q.y;
        dots[2] += cabs2(acc);
      }
    }
    __syncwarp();
    cur = nxt;
    nxt = nn;
    rs_cur = rs_nxt;
    st ^= 1;
  }
  if (DOT != 0) {
    double tot[4];
    if (reduce_and_ticket<4>(dots, partial_of(D, s0), D.counter + s0, tot)) {
      c128 *sc = scal_of(D, s0);
      sc[slot0] = cmake(tot[0], tot[1]);
      if (DOT == 3) sc[slot1] = cmake(tot[2], 0.0);
    }
  }
}

// ---------------------------------------------------------------- preconditioner pieces
// dinv = 1/diag(A) (1 where the diagonal is missing or zero)
__global__ void k_dinv(SolveDev D, int first_matrix, const int32_t *__restrict__ diag_pos, int jacobi) {
  const int f = first_matrix + blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
    c128 v = cmake(1.0, 0.0);
    if (jacobi) {
      const int p = diag_pos[i];
      if (p >= 0) {
        const c128 a = D.vals[(size_t)f * D.nnz + p];
        if (a.x != 0.0 || a.y != 0.0) v = cdiv(cmake(1.0, 0.0), a);
      }
    }
    D.dinv[(size_t)f * D.m + i] = v;
  }
}

// nodal diagonal of G^T A G (G = discrete gradient, -1 at the tail node, +1 at the head node): L[n] = sum over the free edges
// e, e2 at node n of G[e][n] A[e][e2] G[e2][n].  One thread per node walks its incident edges (list order) and their rows
// (column order): a fixed summation order, so the preconditioner -- and with it every iterate -- is bit-reproducible from
// run to run (the first version scattered per-edge partial sums with fp64 atomics).
__global__ void k_nodal_diag(SolveDev D, int first_matrix) {
  const int f = first_matrix + blockIdx.y;
  c128 *L = D.linv + (size_t)f * D.n_node;
  const c128 *__restrict__ av = D.vals + (size_t)f * D.nnz;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < D.n_node; n += gridDim.x * blockDim.x) {
    c128 acc = cmake(0.0, 0.0);
    for (int q = D.n2e_ptr[n]; q < D.n2e_ptr[n + 1]; ++q) {
      const int it = D.n2e_item[q];
      if (it & 2) continue;  // Dirichlet edge
      const int e = it >> 2;
      c128 row = cmake(0.0, 0.0);
      for (int k = D.rowptr[e]; k < D.rowptr[e + 1]; ++k) {
        const int e2 = D.colidx[k];
        if (D.dir[e2]) continue;
        const int2 cd = D.edge_nodes[e2];
        if (cd.y == n) row = cadd(row, av[k]);
        else if (cd.x == n) row = csub(row, av[k]);
      }
      acc = (it & 1) ? cadd(acc, row) : csub(acc, row);
    }
    L[n] = acc;
  }
}

__global__ void k_linv(SolveDev D, int first_matrix) {
  const int f = first_matrix + blockIdx.y;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < D.n_node; n += gridDim.x * blockDim.x) {
    c128 *p = D.linv + (size_t)f * D.n_node + n;
    const c128 a = *p;
    c128 v = cmake(0.0, 0.0);
    if (!D.node_dir[n] && (a.x != 0.0 || a.y != 0.0)) v = cdiv(cmake(1.0, 0.0), a);
    *p = v;
  }
}

// w[s][n] = linv[f][n] * sum_e G[e][n] in[s][e]
__global__ void k_prec_node(SolveDev D, int first_sys, const c128 *__restrict__ in) {
  const int s = first_sys + blockIdx.y;
  if (!D.state[s * 4 + ST_ACTIVE]) return;
  const int f = s / D.n_rhs;
  const c128 *__restrict__ v = in + (size_t)s * D.m;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < D.n_node; n += gridDim.x * blockDim.x) {
    c128 acc = cmake(0.0, 0.0);
    const c128 li = D.linv[(size_t)f * D.n_node + n];
    if (li.x != 0.0 || li.y != 0.0) {
      for (int k = D.n2e_ptr[n]; k < D.n2e_ptr[n + 1]; ++k) {
        const int it = D.n2e_item[k];
        if (it & 2) continue;  // Dirichlet edge
        const int e = it >> 2;
        const c128 a = v[e];
        acc = (it & 1) ? cadd(acc, a) : csub(acc, a);
      }
      acc = cmul(li, acc);
    }
    D.w[(size_t)s * D.n_node + n] = acc;
  }
}

// out = dinv .* in (+ G w when aux) ; DOT 1: scal[slot] = sum in.out (unconjugated) ; COPY: p = out too
template <int DOT, int COPY>
__global__ void __launch_bounds__(VEC_THREADS)
k_prec_edge(SolveDev D, int first_sys, const c128 *__restrict__ in, c128 *__restrict__ out, c128 *__restrict__ out2, int aux, int slot) {
  const int s = first_sys + blockIdx.y;
  if (!D.state[s * 4 + ST_ACTIVE]) return;
  const int f = s / D.n_rhs;
  const size_t off = (size_t)s * D.m;
  double d[2] = {0.0, 0.0};
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < D.m; e += gridDim.x * blockDim.x) {
    const c128 a = in[off + e];
    c128 z = cmul(D.dinv[(size_t)f * D.m + e], a);
    if (aux && !D.dir[e]) {
      const int2 ab = D.edge_nodes[e];
      const c128 *w = D.w + (size_t)s * D.n_node;
      z = cadd(z, csub(w[ab.y], w[ab.x]));
    }
    out[off + e] = z;
    if (COPY) out2[off + e] = z;
    if (DOT) {
      const c128 q = cmul(a, z);
      d[0] += q.x; d[1] += q.y;
    }
  }
  if (DOT) {
    double tot[2];
    if (reduce_and_ticket<2>(d, partial_of(D, s), D.counter + s, tot)) scal_of(D, s)[slot] = cmake(tot[0], tot[1]);
  }
}

// ---------------------------------------------------------------- init / residual
// r = b - y (y = A x), rr = |r|^2, bb = |b|^2.  Finalise: true-residual convergence test.
// BICG: also r0 = r and rho = r0^H r = rr.
template <int BICG>
__global__ void __launch_bounds__(VEC_THREADS)
k_init_residual(SolveDev D, int first_sys, const c128 *__restrict__ b, const c128 *__restrict__ y, c128 *__restrict__ r,
                c128 *__restrict__ r0, int zero_x) {
  const int s = first_sys + blockIdx.y;
  if (!D.state[s * 4 + ST_ACTIVE]) return;
  const size_t off = (size_t)s * D.m;
  double d[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
    const c128 bi = b[off + i];
    const c128 ri = zero_x ? bi : csub(bi, y[off + i]);
    r[off + i] = ri;
    if (BICG) r0[off + i] = ri;
    d[0] += cabs2(ri);
    d[1] += cabs2(bi);
  }
  double tot[2];
  if (reduce_and_ticket<2>(d, partial_of(D, s), D.counter + s, tot)) {
    c128 *sc = scal_of(D, s);
    sc[S_RR] = cmake(tot[0], 0.0);
    sc[S_BB] = cmake(tot[1], 0.0);
    sc[S_RHO0] = cmake(tot[0], 0.0);  // BiCGSTAB rho = r0^H r (COCG overwrites it with r^T z)
    sc[S_ALPHA] = cmake(1.0, 0.0);
    sc[S_OMEGA] = cmake(1.0, 0.0);
    int32_t *st = D.state + s * 4;
    st[ST_REC] = 0;
    if (tot[0] <= D.tol2 * tot[1]) {  // also covers b == 0
      st[ST_ACTIVE] = 0;
      st[ST_CONV] = 1;
    }
  }
}

__global__ void k_fill_zero(c128 *x, int m, int first_sys) {
  const size_t off = (size_t)(first_sys + blockIdx.y) * m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) x[off + i] = cmake(0.0, 0.0);
}

__global__ void k_state_begin(int32_t *state, int first_sys, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t *st = state + (first_sys + i) * 4;
  st[ST_ACTIVE] = 1; st[ST_ITERS] = 0; st[ST_CONV] = 0; st[ST_REC] = 0;
}

// after an iteration phase: systems whose recursive residual converged are re-activated so the
// next k_init_residual verifies the TRUE residual; systems that ran out of iterations stop.
__global__ void k_state_reactivate(int32_t *state, int first_sys, int n, int max_it) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t *st = state + (first_sys + i) * 4;
  if (st[ST_CONV]) { st[ST_ACTIVE] = 0; return; }
  if (st[ST_REC]) { st[ST_ACTIVE] = 1; return; }
  if (st[ST_ITERS] >= max_it) st[ST_ACTIVE] = 0;
}

// ---------------------------------------------------------------- COCG phases
// x += alpha p ; r -= alpha q ; rr = |r|^2      alpha = rho / (p^T q)
__global__ void __launch_bounds__(VEC_THREADS)
k_cocg_update(SolveDev D, int first_sys, c128 *__restrict__ x, c128 *__restrict__ r, const c128 *__restrict__ p,
              const c128 *__restrict__ q, int par) {
  const int s = first_sys + blockIdx.y;
  int32_t *st = D.state + s * 4;
  if (!st[ST_ACTIVE]) return;
  c128 *sc = scal_of(D, s);
  const c128 rho = sc[S_RHO0 + par], pq = sc[S_PQ];
  const bool brk = (pq.x == 0.0 && pq.y == 0.0) || !(isfinite(pq.x) && isfinite(pq.y));
  const c128 alpha = brk ? cmake(0.0, 0.0) : cdiv(rho, pq);
  const size_t off = (size_t)s * D.m;
  double d[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
    x[off + i] = cfma(alpha, p[off + i], x[off + i]);
    const c128 ri = cfma(cneg(alpha), q[off + i], r[off + i]);
    r[off + i] = ri;
    d[0] += cabs2(ri);
  }
  double tot[2];
  if (reduce_and_ticket<2>(d, partial_of(D, s), D.counter + s, tot)) {
    sc[S_RR] = cmake(tot[0], 0.0);
    st[ST_ITERS] += 1;
    if (tot[0] <= D.tol2 * sc[S_BB].x) { st[ST_ACTIVE] = 0; st[ST_REC] = 1; }
    else if (brk || st[ST_ITERS] >= D.max_it) st[ST_ACTIVE] = 0;
  }
}

// p = z + beta p     beta = rho_new / rho
__global__ void __launch_bounds__(VEC_THREADS)
k_cocg_p(SolveDev D, int first_sys, c128 *__restrict__ p, const c128 *__restrict__ z, int par) {
  const int s = first_sys + blockIdx.y;
  if (!D.state[s * 4 + ST_ACTIVE]) return;
  const c128 *sc = scal_of(D, s);
  const c128 rho = sc[S_RHO0 + par], rho_new = sc[S_RHO0 + (par ^ 1)];
  const c128 beta = (rho.x == 0.0 && rho.y == 0.0) ? cmake(0.0, 0.0) : cdiv(rho_new, rho);
  const size_t off = (size_t)s * D.m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x)
    p[off + i] = cfma(beta, p[off + i], z[off + i]);
}

// ---------------------------------------------------------------- BiCGSTAB phases
// p = r + beta (p - omega v)   beta = (rho_new/rho_old)(alpha/omega)   (first: p = r)
__global__ void __launch_bounds__(VEC_THREADS)
k_bicg_p(SolveDev D, int first_sys, c128 *__restrict__ p, const c128 *__restrict__ r, const c128 *__restrict__ v, int par, int first_it) {
  const int s = first_sys + blockIdx.y;
  if (!D.state[s * 4 + ST_ACTIVE]) return;
  const c128 *sc = scal_of(D, s);
  const size_t off = (size_t)s * D.m;
  if (first_it) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) p[off + i] = r[off + i];
    return;
  }
  const c128 rho_new = sc[S_RHO0 + par], rho_old = sc[S_RHO0 + (par ^ 1)];
  const c128 omega = sc[S_OMEGA], alpha = sc[S_ALPHA];
  c128 beta = cmake(0.0, 0.0);
  if ((rho_old.x != 0.0 || rho_old.y != 0.0) && (omega.x != 0.0 || omega.y != 0.0)) beta = cmul(cdiv(rho_new, rho_old), cdiv(alpha, omega));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
    const c128 t = cfma(cneg(omega), v[off + i], p[off + i]);
    p[off + i] = cfma(beta, t, r[off + i]);
  }
}

// s = r - alpha v (in place in r)     alpha = rho / (r0^H v)
__global__ void __launch_bounds__(VEC_THREADS)
k_bicg_s(SolveDev D, int first_sys, c128 *__restrict__ r, const c128 *__restrict__ v, int par) {
  const int s = first_sys + blockIdx.y;
  if (!D.state[s * 4 + ST_ACTIVE]) return;
  const c128 *sc = scal_of(D, s);
  const c128 rho = sc[S_RHO0 + par], r0v = sc[S_R0V];
  const c128 alpha = (r0v.x == 0.0 && r0v.y == 0.0) ? cmake(0.0, 0.0) : cdiv(rho, r0v);
  const size_t off = (size_t)s * D.m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x)
    r[off + i] = cfma(cneg(alpha), v[off + i], r[off + i]);
}

// omega = (t^H s)/(t^H t) ; x += alpha y + omega z ; r = s - omega t ; rr, rho_new = r0^H r
__global__ void __launch_bounds__(VEC_THREADS)
k_bicg_x(SolveDev D, int first_sys, c128 *__restrict__ x, c128 *__restrict__ r, const c128 *__restrict__ y,
         const c128 *__restrict__ z, const c128 *__restrict__ t, const c128 *__restrict__ r0, int par) {
  const int s = first_sys + blockIdx.y;
  int32_t *st = D.state + s * 4;
  if (!st[ST_ACTIVE]) return;
  c128 *sc = scal_of(D, s);
  const c128 rho = sc[S_RHO0 + par], r0v = sc[S_R0V], ts = sc[S_TS];
  const double tt = sc[S_TT].x;
  const bool brk = (r0v.x == 0.0 && r0v.y == 0.0) || !(isfinite(r0v.x) && isfinite(r0v.y));
  const c128 alpha = brk ? cmake(0.0, 0.0) : cdiv(rho, r0v);
  const c128 omega = (tt > 0.0) ? cmake(ts.x / tt, ts.y / tt) : cmake(0.0, 0.0);
  const size_t off = (size_t)s * D.m;
  double d[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
    c128 xi = cfma(alpha, y[off + i], x[off + i]);
    x[off + i] = cfma(omega, z[off + i], xi);
    const c128 ri = cfma(cneg(omega), t[off + i], r[off + i]);
    r[off + i] = ri;
    d[0] += cabs2(ri);
    const c128 q = cmulconj(r0[off + i], ri);
    d[1] += q.x; d[2] += q.y;
  }
  double tot[4];
  if (reduce_and_ticket<4>(d, partial_of(D, s), D.counter + s, tot)) {
    sc[S_RR] = cmake(tot[0], 0.0);
    sc[S_RHO0 + (par ^ 1)] = cmake(tot[1], tot[2]);
    sc[S_ALPHA] = alpha;
    sc[S_OMEGA] = omega;
    st[ST_ITERS] += 1;
    if (tot[0] <= D.tol2 * sc[S_BB].x) { st[ST_ACTIVE] = 0; st[ST_REC] = 1; }
    else if (brk || (omega.x == 0.0 && omega.y == 0.0) || st[ST_ITERS] >= D.max_it) st[ST_ACTIVE] = 0;
  }
}

// ---------------------------------------------------------------- persistent small-system COCG
// Sweep regime (SURVEY hard part 2): hundreds of small systems (m ~ 6e3) sharing one pattern.  One
// CTA owns one matrix and NR of its right-hand sides for the WHOLE solve: the search direction p
// lives in shared memory (SpMV gathers hit smem instead of L2), r/q/x live in this CTA's private
// slice of global memory (L2 resident), matrix values stream from HBM once per iteration, every
// reduction is CTA-local (no tickets, no grid-wide barriers, no kernel launches inside the loop)
// and CTAs pull the next matrix from an atomic queue when theirs has converged.
constexpr int SMALL_LPR = 16;
#ifndef EFB_SMALL_GL
#define EFB_SMALL_GL 2  // lanes per node of the nodal gather
#endif
#ifndef EFB_SMALL_GB
#define EFB_SMALL_GB 4  // items per lane and batch (independent loads)
#endif

// 32-byte load (LDG.256, sm_100): both right-hand sides of one interleaved r entry in ONE request -- a random gather costs a
// wavefront per distinct line whatever its width, so this halves the load-pipe cost of the nodal gather
__device__ __forceinline__ void ldg256(const c128 *p, c128 &a, c128 &b) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p) : "memory");
}

// Index lists of the compact numbering (free unknown -> edge id, node -> incident free edges).  They are the same for every
// job of a launch; when they fit beside p and w they are staged ONCE per CTA in shared memory as 16-bit values: the nodal
// gather's chain pointer -> item -> r and the chains orig -> x / 1/diag then cost one L2 round trip instead of two or three
// (the gather alone was 29 % of the iteration from global memory).
struct SmallLists {
  long long *prof;  // EDGEFEM_B200_SMALL_PROF=1: cycles per phase of CTA 0 (thread 0), else NULL
  const int32_t *g_orig, *g_ptr, *g_item;
  const uint16_t *s_orig, *s_ptr, *s_item;
  int staged;
  __device__ __forceinline__ int orig(int i) const { return staged ? (int)s_orig[i] : __ldg(&g_orig[i]); }
  __device__ __forceinline__ int ptr(int n) const { return staged ? (int)s_ptr[n] : __ldg(&g_ptr[n]); }
  __device__ __forceinline__ int item(int k) const { return staged ? (int)s_item[k] : __ldg(&g_item[k]); }
};
#define EFB_SPROF(k)                                   \
  do {                                                 \
    if (L.prof && blockIdx.x == 0 && tid == 0) {       \
      const long long t_ = clock64();                  \
      L.prof[k] += t_ - xt;                            \
      xt = t_;                                         \
    }                                                  \
  } while (0)
static size_t small_lists_bytes(int mc, int nn) { return ((size_t)mc + (nn ? (size_t)nn + 1 + 2 * (size_t)mc : 0) + 8) * sizeof(uint16_t); }

// One job of the persistent solver: matrix f, the NR right-hand sides starting at system s0, solved to convergence by
// the whole CTA (see k_cocg_small below).
template <int NR, int SPD, bool DB, bool RES = false>
__device__ __forceinline__ void cocg_small_job(const SolveDev &D, const int32_t *__restrict__ sell_ptr, const int32_t *__restrict__ sell_col,
                                               const int32_t *__restrict__ sell_perm, const c128 *__restrict__ sell_vals, long long sell_total,
                                               int n_slices, const c128 *__restrict__ bvec, c128 *xvec, c128 *rvec, c128 *qvec, int aux, int zero_x,
                                               int max_restarts, int mc, const SmallLists &L,
                                               const int2 *__restrict__ c_edge_nodes, unsigned char *sm_raw, int f, int s0) {
  const int m = D.m, nn = aux ? D.n_node : 0;
  // RES (one right-hand side): r and q live in shared memory next to p and w -- the vector passes and the nodal gather of
  // the preconditioner then never leave the SM (x, 1/diag and the matrix stream are what is left of the L2 traffic)
  static_assert(!RES || NR == 1, "resident r, q: one right-hand side per job");
  c128 *p_s = (c128 *)sm_raw;                 // [NR][mc]
  c128 *w_s = p_s + (size_t)NR * mc;          // [NR][nn]
  c128 *r_s = w_s + (size_t)NR * nn;          // RES: [mc]
  c128 *q_s = r_s + (RES ? mc : 0);           // RES: [mc]
  double *red = (double *)(q_s + (RES ? mc : 0));  // [33*8]
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nwarp = nth >> 5;
  const c128 *__restrict__ av = D.vals + (size_t)f * D.nnz;
  const c128 *__restrict__ dinv = D.dinv + (size_t)f * m;
  const c128 *__restrict__ linv = D.linv + (size_t)f * (aux ? D.n_node : 0);
  // per-rhs vectors are addressed as base + r*m (one base pointer per vector instead of NR pointers)
  struct VRef {
    c128 *base;
    int m;
    __device__ __forceinline__ c128 *operator[](int r) const { return base + (size_t)r * m; }
  };
  const size_t off0 = (size_t)s0 * m;
  const VRef xg{xvec + off0, m}, qg{RES ? q_s : qvec + off0, m}, bg{const_cast<c128 *>(bvec) + off0, m};
  // r is interleaved by right-hand side, r[i][NR]: the nodal gather of the preconditioner reads it at random, and the NR values of
  // an edge then share one 32-byte sector instead of wasting half of NR sectors
  struct RRef {
    c128 *base;
    __device__ __forceinline__ c128 &operator()(int r, int i) const { return base[(size_t)i * NR + r]; }
  };
  const RRef rg{RES ? r_s : rvec + off0};
  // block-uniform scalars live in shared memory (written by thread 0 between barriers): bb, rr and a
  // double-buffered rho per right-hand side
  double *sc_bb = red + 33 * 8, *sc_rr = sc_bb + NR, *sc_rho = sc_rr + NR;  // sc_rho[parity][r][2]
  int par = 0;
  int iters[NR];
  bool act[NR], conv[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) { iters[r] = 0; act[r] = false; conv[r] = false; }
  if (tid < 8 * NR) sc_bb[tid] = 0.0;
  // Dirichlet rows are decoupled identity rows: x_e = b_e / A_ee, once
  if (mc < m)
    for (int e = tid; e < m; e += nth)
      if (D.dir[e]) {
        const c128 de = __ldg(&dinv[e]);
#pragma unroll
        for (int r = 0; r < NR; ++r) xg[r][e] = cmul(de, bg[r][e]);
      }
  __syncthreads();

  // q = A * (vector in p_s), optional dot p.q.  SELL-32: a lane owns a row, 32 rows form a slice stored
  // column-major.  A warp owns a CONTIGUOUS range of slices (balanced by entry count), so its values and
  // columns are one flat stream: every step is a coalesced 512-byte value load + 128-byte column load,
  // SPD steps are in flight per lane and the next batch is requested before the current one is consumed
  // (double buffering in registers); slice ends only flush the row accumulators.  No shuffles.
  const c128 *__restrict__ sv = sell_vals + (size_t)f * (size_t)sell_total;
  int s_lo, s_hi;
  {
    const long long t0 = sell_total * wid / nwarp, t1 = sell_total * (wid + 1) / nwarp;
    auto lower = [&](long long t) {
      int lo = 0, hi = n_slices;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)__ldg(&sell_ptr[mid]) < t) lo = mid + 1; else hi = mid;
      }
      return lo;
    };
    s_lo = lower(t0);
    s_hi = (wid == nwarp - 1) ? n_slices : lower(t1);
  }
  // per-slice variant (SPD == 0): slices dealt round-robin to the warps, batches of 4 independent loads
  auto spmv_slices = [&](bool want_dot, double (&dots)[2 * NR]) {
#pragma unroll
    for (int k = 0; k < 2 * NR; ++k) dots[k] = 0.0;
    for (int sl = wid; sl < n_slices; sl += nwarp) {
      const int base = __ldg(&sell_ptr[sl]);
      const int width = (__ldg(&sell_ptr[sl + 1]) - base) >> 5;
      const int row = __ldg(&sell_perm[sl * 32 + lane]);
      const c128 *vp = sv + base + lane;
      const int32_t *cp = sell_col + base + lane;
      c128 acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = cmake(0.0, 0.0);
      int j = 0;
      for (; j + 4 <= width; j += 4) {
        c128 a4[4];
        int c4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a4[u] = ldg_stream(vp + 32 * (j + u));
          c4[u] = ldg_stream(cp + 32 * (j + u));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int r = 0; r < NR; ++r) acc[r] = cfma(a4[u], p_s[(size_t)r * mc + c4[u]], acc[r]);
      }
      for (; j < width; ++j) {
        const c128 a = ldg_stream(vp + 32 * j);
        const int c = ldg_stream(cp + 32 * j);
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] = cfma(a, p_s[(size_t)r * mc + c], acc[r]);
      }
      if (row >= 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          qg[r][row] = acc[r];
          if (want_dot) {
            const c128 t = cmul(p_s[(size_t)r * mc + row], acc[r]);
            dots[2 * r] += t.x; dots[2 * r + 1] += t.y;
          }
        }
      }
    }
  };
  auto spmv_flat = [&](bool want_dot, double (&dots)[2 * NR]) {
#pragma unroll
    for (int k = 0; k < 2 * NR; ++k) dots[k] = 0.0;
    if (s_lo >= s_hi) return;
    const c128 *vp = sv + lane;
    const int32_t *cp = sell_col + lane;
    int step = __ldg(&sell_ptr[s_lo]) >> 5;
    const int step_end = __ldg(&sell_ptr[s_hi]) >> 5;
    int sl = s_lo;
    int next_b = __ldg(&sell_ptr[sl + 1]) >> 5;
    int row = __ldg(&sell_perm[sl * 32 + lane]);
    c128 acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = cmake(0.0, 0.0);
    auto flush = [&]() {
      if (row >= 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          qg[r][row] = acc[r];
          if (want_dot) {
            const c128 t = cmul(p_s[(size_t)r * mc + row], acc[r]);
            dots[2 * r] += t.x; dots[2 * r + 1] += t.y;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = cmake(0.0, 0.0);
      ++sl;
      if (sl < s_hi) {
        next_b = __ldg(&sell_ptr[sl + 1]) >> 5;
        row = __ldg(&sell_perm[sl * 32 + lane]);
      }
    };
    while (sl < s_hi && next_b == step) flush();  // leading empty slices (rows without entries)
    constexpr int SPDX = SPD > 0 ? SPD : 1;
    c128 a0[SPDX];
    int c0[SPDX];
#pragma unroll
    for (int u = 0; u < SPD; ++u) {
      const int st = step + u;
      a0[u] = cmake(0.0, 0.0);
      c0[u] = 0;
      if (st < step_end) {
        a0[u] = ldg_stream(vp + (size_t)32 * st);
        c0[u] = ldg_stream(cp + (size_t)32 * st);
      }
    }
    while (step < step_end) {
      c128 a1[DB ? SPDX : 1];
      int c1[DB ? SPDX : 1];
      if (DB) {
#pragma unroll
        for (int u = 0; u < SPD; ++u) {
          const int st = step + SPD + u;
          a1[DB ? u : 0] = cmake(0.0, 0.0);
          c1[DB ? u : 0] = 0;
          if (st < step_end) {
            a1[DB ? u : 0] = ldg_stream(vp + (size_t)32 * st);
            c1[DB ? u : 0] = ldg_stream(cp + (size_t)32 * st);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < SPD; ++u) {
        const int st = step + u;
        if (st < step_end) {
#pragma unroll
          for (int r = 0; r < NR; ++r) acc[r] = cfma(a0[u], p_s[(size_t)r * mc + c0[u]], acc[r]);
          while (sl < s_hi && next_b == st + 1) flush();
        }
      }
      step += SPD;
#pragma unroll
      for (int u = 0; u < SPD; ++u) {
        if (DB) {
          a0[u] = a1[DB ? u : 0];
          c0[u] = c1[DB ? u : 0];
        } else {
          const int st = step + u;
          a0[u] = cmake(0.0, 0.0);
          c0[u] = 0;
          if (st < step_end) {
            a0[u] = ldg_stream(vp + (size_t)32 * st);
            c0[u] = ldg_stream(cp + (size_t)32 * st);
          }
        }
      }
    }
    while (sl < s_hi) flush();  // trailing empty slices
  };

  auto spmv = [&](bool want_dot, double (&dots)[2 * NR]) {
    if (SPD == 0) spmv_slices(want_dot, dots);
    else spmv_flat(want_dot, dots);
  };

  for (int cycle = 0;; ++cycle) {
    // (1) true residual r = b - A x from the current iterate
    double d4[2 * NR];
    if (cycle == 0 && zero_x) {
      for (int i = tid; i < mc; i += nth) {
        const int e = L.orig(i);
#pragma unroll
        for (int r = 0; r < NR; ++r) { xg[r][e] = cmake(0.0, 0.0); qg[r][i] = cmake(0.0, 0.0); }
      }
    } else {
      for (int i = tid; i < mc; i += nth) {
        const int e = L.orig(i);
#pragma unroll
        for (int r = 0; r < NR; ++r) p_s[(size_t)r * mc + i] = xg[r][e];
      }
      __syncthreads();
      spmv(false, d4);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2 * NR; ++k) d4[k] = 0.0;
    for (int i = tid; i < mc; i += nth) {
      const int e = L.orig(i);
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const c128 bi = bg[r][e];
        const c128 ri = csub(bi, qg[r][i]);
        rg(r, i) = ri;
        d4[2 * r] += cabs2(ri);
        d4[2 * r + 1] += cabs2(bi);
      }
    }
    block_allreduce<2 * NR>(d4, red);
    bool any = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const double rr_r = d4[2 * r], bb_r = d4[2 * r + 1];
      if (tid == 0) { sc_rr[r] = rr_r; sc_bb[r] = bb_r; }
      conv[r] = (rr_r <= D.tol2 * bb_r);
      act[r] = !conv[r] && iters[r] < D.max_it && cycle <= max_restarts && isfinite(rr_r);
      any |= act[r];
    }
    __syncthreads();
    if (!any) break;

    // (2)+(3) preconditioned COCG until the recursive residual converges.  Two block reductions per
    // iteration: (rho_new = r^T z, |r|^2) after the preconditioner and p^T A p after the SpMV.
    bool fresh = true;  // p = z on entry, p = z + beta p afterwards
    long long xt = L.prof ? clock64() : 0;
    for (;;) {
      EFB_SPROF(7);
      // z = M^-1 r  (z parked in q), rho_new = r^T z, rr = |r|^2
      if (aux) {
        // nodal gather w = diag(G^T A G)^-1 G^T r.  GL lanes per node, 32/GL nodes per warp step: the whole node set is
        // covered in nn / (warps * 32/GL) steps (WR-90: 972 nodes, 32 warps, GL = 4 -> 4 steps; the 16-lanes-per-node form
        // needed 16 and each step is a chain of L2 round trips: 22-29 % of the iteration).  A lane takes items l, l+GL, ...
        // of its node, GB at a time with independent loads (both right-hand sides of an edge in one LDG.256), then
        // log2(GL) shuffle steps in a fixed order.
        constexpr int GL = EFB_SMALL_GL, GB = EFB_SMALL_GB, NPWS = 32 / GL;
        const int gl = tid & (GL - 1), wsub = (tid & 31) / GL;
        for (int n0 = (tid >> 5) * NPWS; n0 < nn; n0 += (nth >> 5) * NPWS) {
          const int n = n0 + wsub;
          const bool on = n < nn;
          const int kb = on ? L.ptr(n) : 0, ke = on ? L.ptr(n + 1) : 0;
          c128 li = cmake(0.0, 0.0);
          if (on && gl == 0) li = __ldg(&linv[n]);
          c128 a2[NR];
#pragma unroll
          for (int r = 0; r < NR; ++r) a2[r] = cmake(0.0, 0.0);
          for (int k0 = kb + gl; k0 < ke; k0 += GL * GB) {
            int it[GB];
#pragma unroll
            for (int u = 0; u < GB; ++u) it[u] = k0 + GL * u < ke ? L.item(k0 + GL * u) : -1;  // compact edge << 1 | head
            c128 v[GB][NR];
#pragma unroll
            for (int u = 0; u < GB; ++u) {
#pragma unroll
              for (int r = 0; r < NR; ++r) v[u][r] = cmake(0.0, 0.0);
              if (it[u] >= 0) {
                if constexpr (NR == 2 && !RES) {
                  ldg256(&rg(0, it[u] >> 1), v[u][0], v[u][NR - 1]);
                } else {
#pragma unroll
                  for (int r = 0; r < NR; ++r) v[u][r] = rg(r, it[u] >> 1);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < GB; ++u)
#pragma unroll
              for (int r = 0; r < NR; ++r) a2[r] = (it[u] & 1) ? cadd(a2[r], v[u][r]) : csub(a2[r], v[u][r]);
          }
#pragma unroll
          for (int o = GL / 2; o > 0; o >>= 1)
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              a2[r].x += __shfl_xor_sync(0xffffffffu, a2[r].x, o);
              a2[r].y += __shfl_xor_sync(0xffffffffu, a2[r].y, o);
            }
          if (on && gl == 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r) w_s[(size_t)r * nn + n] = cmul(li, a2[r]);
          }
        }
        __syncthreads();
      }
      EFB_SPROF(0);
      double dz[3 * NR];
#pragma unroll
      for (int k = 0; k < 3 * NR; ++k) dz[k] = 0.0;
      for (int e = tid; e < mc; e += nth) {
        const c128 di = __ldg(&dinv[L.orig(e)]);
        int2 ab = make_int2(0, 0);
        const bool g = aux != 0;
        if (g) ab = __ldg(&c_edge_nodes[e]);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const c128 ri = rg(r, e);
          c128 z = cmul(di, ri);
          if (g) z = cadd(z, csub(w_s[(size_t)r * nn + ab.y], w_s[(size_t)r * nn + ab.x]));
          qg[r][e] = z;
          const c128 t = cmul(ri, z);
          dz[3 * r] += t.x; dz[3 * r + 1] += t.y;
          dz[3 * r + 2] += cabs2(ri);
        }
      }
      EFB_SPROF(1);
      block_allreduce<3 * NR>(dz, red);
      EFB_SPROF(2);
      bool still = false;
      c128 beta[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const c128 rho_new = cmake(dz[3 * r], dz[3 * r + 1]);
        const c128 rho_old = cmake(sc_rho[(par * NR + r) * 2], sc_rho[(par * NR + r) * 2 + 1]);
        beta[r] = (fresh || (rho_old.x == 0.0 && rho_old.y == 0.0)) ? cmake(0.0, 0.0) : cdiv(rho_new, rho_old);
        if (tid == 0) {
          sc_rho[((par ^ 1) * NR + r) * 2] = rho_new.x;
          sc_rho[((par ^ 1) * NR + r) * 2 + 1] = rho_new.y;
        }
        if (act[r]) {
          const double rr_r = dz[3 * r + 2];
          if (tid == 0) sc_rr[r] = rr_r;
          if (rr_r <= D.tol2 * sc_bb[r] || iters[r] >= D.max_it || !isfinite(rr_r)) act[r] = false;
        }
        still |= act[r];
      }
      par ^= 1;
      if (!still) break;
      for (int e = tid; e < mc; e += nth)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const c128 z = qg[r][e];
          p_s[(size_t)r * mc + e] = fresh ? z : cfma(beta[r], p_s[(size_t)r * mc + e], z);
        }
      fresh = false;
      __syncthreads();
      EFB_SPROF(3);
      // q = A p ; alpha = rho / p^T q
      double dq[2 * NR];
      spmv(true, dq);
      EFB_SPROF(4);
      block_allreduce<2 * NR>(dq, red);
      EFB_SPROF(5);
      c128 alpha[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const c128 pq = cmake(dq[2 * r], dq[2 * r + 1]);
        const bool brk = (pq.x == 0.0 && pq.y == 0.0) || !(isfinite(pq.x) && isfinite(pq.y));
        if (brk) act[r] = false;
        const c128 rho_r = cmake(sc_rho[(par * NR + r) * 2], sc_rho[(par * NR + r) * 2 + 1]);
        alpha[r] = act[r] ? cdiv(rho_r, pq) : cmake(0.0, 0.0);
        if (act[r]) iters[r] += 1;
      }
      // x += alpha p ; r -= alpha q   (|r|^2 is accumulated by the next preconditioner pass)
      for (int i = tid; i < mc; i += nth) {
        const int e = L.orig(i);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          if (!act[r]) continue;
          xg[r][e] = cfma(alpha[r], p_s[(size_t)r * mc + i], xg[r][e]);
          rg(r, i) = cfma(cneg(alpha[r]), qg[r][i], rg(r, i));
        }
      }
      bool any2 = false;
#pragma unroll
      for (int r = 0; r < NR; ++r) any2 |= act[r];
      __syncthreads();
      EFB_SPROF(6);
      if (!any2) break;
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      int32_t *st = D.state + (s0 + r) * 4;
      st[ST_ACTIVE] = 0; st[ST_ITERS] = iters[r]; st[ST_CONV] = conv[r] ? 1 : 0; st[ST_REC] = 0;
      c128 *sc = scal_of(D, s0 + r);
      sc[S_RR] = cmake(sc_rr[r], 0.0);
      sc[S_BB] = cmake(sc_bb[r], 0.0);
    }
  }
  __syncthreads();
}

template <int NR, int NT, int SPD, bool DB, int MINB = 1, bool RES = false>
__global__ void __launch_bounds__(NT, MINB)
k_cocg_small(SolveDev D, const int32_t *__restrict__ sell_ptr, const int32_t *__restrict__ sell_col, const int32_t *__restrict__ sell_perm,
             const c128 *__restrict__ sell_vals, long long sell_total, int n_slices, int first_matrix, int n_jobs, int groups_per_matrix, int mixed,
             int *job_counter, const c128 *__restrict__ bvec, c128 *xvec, c128 *rvec, c128 *qvec, int aux, int zero_x, int max_restarts,
             int mc, const int32_t *__restrict__ c_orig, const int2 *__restrict__ c_edge_nodes, const int32_t *__restrict__ c_n2e_ptr,
             const int32_t *__restrict__ c_n2e_item, int stage_off, long long *prof) {
  // The solver runs on the mc FREE unknowns ("compact" ids, c_orig maps them to edge ids): r, q, p and the SELL
  // structure are compact, b, x, dinv keep the original layout (stride m).
  // Job codes: plain index (matrix = code / groups, NR right-hand sides of group code % groups), or -- mixed mode, two
  // right-hand sides per matrix -- matrix * 4 + kind: kind 0 = both right-hand sides together, 1 / 2 = one of them alone.
  // The host splits the matrices expected to finish last into single jobs so that the SMs that would idle through the
  // second round of jobs share its tail (run_cocg_small).
  extern __shared__ __align__(16) unsigned char sm_raw[];
  __shared__ int s_job;
  const int tid = threadIdx.x;
  SmallLists L{prof, c_orig, c_n2e_ptr, c_n2e_item, nullptr, nullptr, nullptr, 0};
  if (stage_off > 0) {  // 16-bit copies of the lists behind the solver's own shared memory (host checked the ranges)
    uint16_t *so = (uint16_t *)(sm_raw + stage_off), *sp = so + mc, *si = sp + (aux ? D.n_node + 1 : 0);
    for (int i = tid; i < mc; i += NT) so[i] = (uint16_t)__ldg(&c_orig[i]);
    if (aux) {
      for (int i = tid; i <= D.n_node; i += NT) sp[i] = (uint16_t)__ldg(&c_n2e_ptr[i]);
      const int n_item = __ldg(&c_n2e_ptr[D.n_node]);
      for (int i = tid; i < n_item; i += NT) si[i] = (uint16_t)__ldg(&c_n2e_item[i]);
    }
    L.s_orig = so;
    L.s_ptr = sp;
    L.s_item = si;
    L.staged = 1;
  }
  for (;;) {
    if (tid == 0) {
      const int q = atomicAdd(job_counter, 1);
      s_job = q < n_jobs ? job_counter[1 + q] : -1;  // queue position -> job (longest expected first)
    }
    __syncthreads();
    const int job = s_job;
    __syncthreads();
    if (job < 0) break;
    if (mixed && NR == 2) {
      const int f = first_matrix + (job >> 2), kind = job & 3;
      if (kind == 0)
        cocg_small_job<2, SPD, DB>(D, sell_ptr, sell_col, sell_perm, sell_vals, sell_total, n_slices, bvec, xvec, rvec, qvec, aux, zero_x,
                                   max_restarts, mc, L, c_edge_nodes, sm_raw, f, f * D.n_rhs);
      else
        cocg_small_job<1, SPD, DB>(D, sell_ptr, sell_col, sell_perm, sell_vals, sell_total, n_slices, bvec, xvec, rvec, qvec, aux, zero_x,
                                   max_restarts, mc, L, c_edge_nodes, sm_raw, f, f * D.n_rhs + kind - 1);
    } else {
      const int f = first_matrix + job / groups_per_matrix;
      cocg_small_job<NR, SPD, DB, RES>(D, sell_ptr, sell_col, sell_perm, sell_vals, sell_total, n_slices, bvec, xvec, rvec, qvec, aux, zero_x,
                                  max_restarts, mc, L, c_edge_nodes, sm_raw, f,
                                  f * D.n_rhs + (job % groups_per_matrix) * NR);
    }
  }
}

// vals (CSR order) -> SELL-32 order of the compact pattern: a pure gather through the slot -> CSR position map
// (-1 = padding); grid (slices, matrices), 256 threads
__global__ void k_csr_to_sell(const c128 *__restrict__ vals, long long nnz, const int32_t *__restrict__ sell_ptr,
                              const int32_t *__restrict__ sell_src, c128 *__restrict__ sell_vals, long long sell_total, int first_matrix) {
  const int sl = blockIdx.x, f = first_matrix + blockIdx.y;
  const int base = sell_ptr[sl], cnt = sell_ptr[sl + 1] - base;
  const c128 *__restrict__ src = vals + (size_t)f * nnz;
  c128 *dst = sell_vals + (size_t)f * sell_total + base;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int k = sell_src[base + i];
    dst[i] = k >= 0 ? src[k] : cmake(0.0, 0.0);
  }
}

static size_t small_smem_bytes_res(int m, int nn) { return ((size_t)3 * m + (size_t)nn) * sizeof(c128) + (33 * 8 + 8) * sizeof(double) + 64; }
static size_t small_smem_bytes(int nr, int m, int nn) { return ((size_t)nr * m + (size_t)nr * nn) * sizeof(c128) + (33 * 8 + 8 * nr) * sizeof(double) + 64; }

// FP64 FMA throughput probe (roofline denominator for the assembly kernel; MEASURED_PEAKS.json has
// no FP64 figure): 8 independent chains x FP64_PROBE_ITERS FMAs per thread.
constexpr int FP64_PROBE_ITERS = 4096;
__global__ void __launch_bounds__(256) k_fp64_probe(double *out, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
  for (int i = 0; i < FP64_PROBE_ITERS; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (r == 123.456) out[0] = r;  // never true: keeps the chains alive
}

// ---------------------------------------------------------------- host side
// blocks along x for an elementwise / reduction kernel over `m` items launched for `n_par`
// independent systems (grid.y): enough CTAs to fill the chip a few times over, but no more --
// every extra CTA costs a ticketed partial reduction, which dominates for small batched systems.
static int vec_grid(const Ctx *c, int m, int n_par = 1) {
  const int nb_max = (m + VEC_THREADS - 1) / VEC_THREADS;
  const int target = c->sm_count * 16;
  const int want = (target + std::max(1, n_par) - 1) / std::max(1, n_par);
  return std::max(1, std::min(std::min(nb_max, want), std::min(RED_MAX_BLOCKS, c->sm_count * 8)));
}

static int pick_lpr(const System *S) {
  const double avg = (double)S->nnz / std::max(1, S->m);
  if (avg <= 6.0) return 4;
  if (avg <= 12.0) return 8;
  if (avg <= 48.0) return 16;
  return 32;
}

int solver_free(System *S) {
  dfree(S->d_work); dfree(S->d_dinv); dfree(S->d_linv); dfree(S->d_w); dfree(S->d_scal);
  dfree(S->d_partial); dfree(S->d_counter); dfree(S->d_state); dfree(S->d_flag); dfree(S->d_job);
  if (S->ev_s0) cudaEventDestroy(S->ev_s0);
  if (S->ev_s1) cudaEventDestroy(S->ev_s1);
  S->ev_s0 = nullptr; S->ev_s1 = nullptr;
  S->d_work = nullptr; S->d_dinv = nullptr; S->d_linv = nullptr; S->d_w = nullptr; S->d_scal = nullptr;
  S->d_partial = nullptr; S->d_counter = nullptr; S->d_state = nullptr; S->d_flag = nullptr; S->d_job = nullptr;
  return EFB_OK;
}

static int solver_alloc(System *S) {
  Ctx *c = S->ctx;
  int rc;
  if (!S->d_work) {
    if ((rc = dev_alloc(c, &S->d_work, (size_t)V_NUM * S->n_sys * S->m))) return rc;
    S->n_work_vec = V_NUM;
    if ((rc = dev_alloc(c, &S->d_dinv, (size_t)S->n_matrix * S->m))) return rc;
    if ((rc = dev_alloc(c, &S->d_scal, (size_t)S->n_sys * NSCAL))) return rc;
    if ((rc = dev_alloc(c, &S->d_partial, (size_t)S->n_sys * RED_MAX_BLOCKS * 4))) return rc;
    if ((rc = dev_alloc(c, &S->d_counter, (size_t)S->n_sys))) return rc;
    if ((rc = dev_alloc(c, &S->d_state, (size_t)S->n_sys * 4))) return rc;
    if ((rc = dev_alloc(c, &S->d_job, (size_t)1 + S->n_sys))) return rc;
    EFB_CUDA(c, cudaMemsetAsync(S->d_counter, 0, (size_t)S->n_sys * sizeof(unsigned), c->stream));
    EFB_CUDA(c, cudaMemsetAsync(S->d_scal, 0, (size_t)S->n_sys * NSCAL * sizeof(c128), c->stream));
    EFB_CUDA(c, cudaMemsetAsync(S->d_state, 0, (size_t)S->n_sys * 4 * sizeof(int32_t), c->stream));
  }
  if (S->n_node > 0 && !S->d_linv) {
    if ((rc = dev_alloc(c, &S->d_linv, (size_t)S->n_matrix * S->n_node))) return rc;
    if ((rc = dev_alloc(c, &S->d_w, (size_t)S->n_sys * S->n_node))) return rc;
  }
  return EFB_OK;
}

static SolveDev make_dev(System *S, double tol, int max_it) {
  SolveDev D;
  D.m = S->m; D.n_rhs = S->n_rhs; D.n_node = S->n_node; D.nnz = (long long)S->nnz;
  D.rowptr = S->d_rowptr; D.colidx = S->d_colidx; D.vals = S->d_vals;
  D.dir = S->d_dir; D.node_dir = S->d_node_dir; D.edge_nodes = S->d_edge_nodes;
  D.n2e_ptr = S->d_n2e_ptr; D.n2e_item = S->d_n2e_item;
  D.dinv = S->d_dinv; D.linv = S->d_linv; D.w = S->d_w;
  D.scal = S->d_scal; D.partial = S->d_partial; D.counter = S->d_counter; D.state = S->d_state;
  D.tol2 = tol * tol; D.max_it = max_it;
  return D;
}

template <int NR, int DOT>
static void launch_spmv_lpr(Ctx *c, const SolveDev &D, int lpr, dim3 grid, int first_matrix, const c128 *x, c128 *y, const c128 *w,
                            int slot0, int slot1, int use_active) {
  switch (lpr) {
    case 4: k_spmv<4, NR, DOT><<<grid, 256, 0, c->stream>>>(D, first_matrix, x, y, w, slot0, slot1, use_active); break;
    case 8: k_spmv<8, NR, DOT><<<grid, 256, 0, c->stream>>>(D, first_matrix, x, y, w, slot0, slot1, use_active); break;
    case 16: k_spmv<16, NR, DOT><<<grid, 256, 0, c->stream>>>(D, first_matrix, x, y, w, slot0, slot1, use_active); break;
    default: k_spmv<32, NR, DOT><<<grid, 256, 0, c->stream>>>(D, first_matrix, x, y, w, slot0, slot1, use_active); break;
  }
}

// y = A x over matrices [first, first+count) and all their right-hand sides
static int launch_spmv(System *S, const SolveDev &D, int first, int count, const c128 *x, c128 *y, const c128 *w, int dot, int slot0,
                       int slot1, int use_active) {
  Ctx *c = S->ctx;
  if (S->d_sp_chunk && S->n_rhs == 1 && !getenv("EDGEFEM_B200_NO_STREAM_SPMV")) {
    static const bool use_regs = [] {  // EDGEFEM_B200_SPMV_KERNEL=regs: the register-streamed kernel
      const char *e = getenv("EDGEFEM_B200_SPMV_KERNEL");
      return e && strcmp(e, "regs") == 0;
    }();
    if (!use_regs) {
      // persistent: 2 CTAs per SM (shared-memory ring of 2 stages per warp), every warp walks chunks with the grid stride
      const size_t smem = (size_t)8 * 2 * SPMV_TMA_STAGE + 8 * 16;
      const int nb = std::max(1, std::min((S->n_sp_chunks + 7) / 8, std::min(RED_MAX_BLOCKS, c->sm_count * 2)));
      dim3 grid((unsigned)nb, (unsigned)count, 1u);
#define EFB_SPMV_TMA(DOT)                                                                                         \
  EFB_CUDA(c, cudaFuncSetAttribute(k_spmv_tma<DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
  k_spmv_tma<DOT><<<grid, 256, smem, c->stream>>>(D, S->d_sp_chunk, S->n_sp_chunks, first, x, y, w, slot0, slot1, use_active)
      switch (dot) {
        case 0: EFB_SPMV_TMA(0); break;
        case 1: EFB_SPMV_TMA(1); break;
        case 2: EFB_SPMV_TMA(2); break;
        default: EFB_SPMV_TMA(3); break;
      }
#undef EFB_SPMV_TMA
      EFB_CHECK_LAUNCH(c);
      return EFB_OK;
    }
    const int nb = std::max(1, std::min((S->n_sp_chunks + 7) / 8, std::min(RED_MAX_BLOCKS, c->sm_count * 8)));
    dim3 grid((unsigned)nb, (unsigned)count, 1u);
    switch (dot) {
      case 0: k_spmv_stream<1, 0><<<grid, 256, 0, c->stream>>>(D, S->d_sp_chunk, S->n_sp_chunks, first, x, y, w, slot0, slot1, use_active); break;
      case 1: k_spmv_stream<1, 1><<<grid, 256, 0, c->stream>>>(D, S->d_sp_chunk, S->n_sp_chunks, first, x, y, w, slot0, slot1, use_active); break;
      case 2: k_spmv_stream<1, 2><<<grid, 256, 0, c->stream>>>(D, S->d_sp_chunk, S->n_sp_chunks, first, x, y, w, slot0, slot1, use_active); break;
      default: k_spmv_stream<1, 3><<<grid, 256, 0, c->stream>>>(D, S->d_sp_chunk, S->n_sp_chunks, first, x, y, w, slot0, slot1, use_active); break;
    }
    EFB_CHECK_LAUNCH(c);
    return EFB_OK;
  }
  const int lpr = pick_lpr(S);
  const int rpb = 256 / lpr;
  const bool two = (S->n_rhs % 2 == 0);
  const int n_par = count * (two ? S->n_rhs / 2 : S->n_rhs);
  const int want = (c->sm_count * 16 + n_par - 1) / n_par;
  int nbx = std::max(1, std::min(std::min((S->m + rpb - 1) / rpb, want), std::min(RED_MAX_BLOCKS, c->sm_count * 8)));
  dim3 grid((unsigned)nbx, (unsigned)count, (unsigned)(two ? S->n_rhs / 2 : S->n_rhs));
#define EFB_SPMV(NR)                                                                              \
  switch (dot) {                                                                                  \
    case 0: launch_spmv_lpr<NR, 0>(c, D, lpr, grid, first, x, y, w, slot0, slot1, use_active); break; \
    case 1: launch_spmv_lpr<NR, 1>(c, D, lpr, grid, first, x, y, w, slot0, slot1, use_active); break; \
    case 2: launch_spmv_lpr<NR, 2>(c, D, lpr, grid, first, x, y, w, slot0, slot1, use_active); break; \
    default: launch_spmv_lpr<NR, 3>(c, D, lpr, grid, first, x, y, w, slot0, slot1, use_active); break; \
  }
  if (two) { EFB_SPMV(2) } else { EFB_SPMV(1) }
#undef EFB_SPMV
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

static int apply_precond(SolvePlan &P, const c128 *in, c128 *out, c128 *out2, int dot_slot) {
  Ctx *c = P.S->ctx;
  if (P.aux) {
    k_prec_node<<<P.ngrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, in);
    EFB_CHECK_LAUNCH(c);
  }
  if (dot_slot >= 0) {
    if (out2) k_prec_edge<1, 1><<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, in, out, out2, P.aux, dot_slot);
    else k_prec_edge<1, 0><<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, in, out, out2, P.aux, dot_slot);
  } else {
    k_prec_edge<0, 0><<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, in, out, out2, P.aux, 0);
  }
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

static int setup_precond(SolvePlan &P) {
  System *S = P.S;
  Ctx *c = S->ctx;
  dim3 g((unsigned)vec_grid(c, S->m, P.n_matrix), (unsigned)P.n_matrix);
  k_dinv<<<g, VEC_THREADS, 0, c->stream>>>(P.D, P.first_matrix, S->d_diag_pos, P.precond != EFB_PRECOND_NONE);
  EFB_CHECK_LAUNCH(c);
  if (P.aux) {
    dim3 gn((unsigned)vec_grid(c, S->n_node, P.n_matrix), (unsigned)P.n_matrix);
    k_nodal_diag<<<gn, VEC_THREADS, 0, c->stream>>>(P.D, P.first_matrix);
    EFB_CHECK_LAUNCH(c);
    k_linv<<<gn, VEC_THREADS, 0, c->stream>>>(P.D, P.first_matrix);
    EFB_CHECK_LAUNCH(c);
  }
  return EFB_OK;
}

// r = b - A x (true residual), convergence test, method-specific start vectors
static int init_cycle(SolvePlan &P, bool zero_x) {
  System *S = P.S;
  Ctx *c = S->ctx;
  int rc;
  if (!zero_x)
    if ((rc = launch_spmv(S, P.D, P.first_matrix, P.n_matrix, S->d_x, P.vec[V_Q], nullptr, 0, 0, 0, 1))) return rc;
  if (P.method == EFB_METHOD_BICGSTAB)
    k_init_residual<1><<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, S->d_b, P.vec[V_Q], P.vec[V_R], P.vec[V_R0], zero_x);
  else
    k_init_residual<0><<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, S->d_b, P.vec[V_Q], P.vec[V_R], P.vec[V_R0], zero_x);
  EFB_CHECK_LAUNCH(c);
  if (P.method == EFB_METHOD_COCG) {
    // z = M^-1 r ; p = z ; rho(par 0) = r^T z
    if ((rc = apply_precond(P, P.vec[V_R], P.vec[V_Z], P.vec[V_P], S_RHO0))) return rc;
  }
  return EFB_OK;
}

static int cocg_iteration(SolvePlan &P, int par) {
  System *S = P.S;
  Ctx *c = S->ctx;
  int rc;
  if ((rc = launch_spmv(S, P.D, P.first_matrix, P.n_matrix, P.vec[V_P], P.vec[V_Q], P.vec[V_P], 1, S_PQ, 0, 1))) return rc;
  k_cocg_update<<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, S->d_x, P.vec[V_R], P.vec[V_P], P.vec[V_Q], par);
  EFB_CHECK_LAUNCH(c);
  if ((rc = apply_precond(P, P.vec[V_R], P.vec[V_Z], nullptr, S_RHO0 + (par ^ 1)))) return rc;
  k_cocg_p<<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, P.vec[V_P], P.vec[V_Z], par);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

static int bicg_iteration(SolvePlan &P, int par, bool first_it) {
  System *S = P.S;
  Ctx *c = S->ctx;
  int rc;
  // V_Q holds v, V_P holds p, V_Y = M^-1 p, V_Z = M^-1 s, V_T = t, s lives in V_R
  k_bicg_p<<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, P.vec[V_P], P.vec[V_R], P.vec[V_Q], par, first_it ? 1 : 0);
  EFB_CHECK_LAUNCH(c);
  if ((rc = apply_precond(P, P.vec[V_P], P.vec[V_Y], nullptr, -1))) return rc;
  if ((rc = launch_spmv(S, P.D, P.first_matrix, P.n_matrix, P.vec[V_Y], P.vec[V_Q], P.vec[V_R0], 2, S_R0V, 0, 1))) return rc;
  k_bicg_s<<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, P.vec[V_R], P.vec[V_Q], par);
  EFB_CHECK_LAUNCH(c);
  if ((rc = apply_precond(P, P.vec[V_R], P.vec[V_Z], nullptr, -1))) return rc;
  if ((rc = launch_spmv(S, P.D, P.first_matrix, P.n_matrix, P.vec[V_Z], P.vec[V_T], P.vec[V_R], 3, S_TS, S_TT, 1))) return rc;
  k_bicg_x<<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, S->d_x, P.vec[V_R], P.vec[V_Y], P.vec[V_Z], P.vec[V_T], P.vec[V_R0], par);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

static int make_plan(System *S, int first_matrix, int n_matrix, const efb_solve_opts *o, SolvePlan &P) {
  Ctx *c = S->ctx;
  int rc = solver_alloc(S);
  if (rc) return rc;
  P.S = S;
  P.first_matrix = first_matrix;
  P.n_matrix = n_matrix;
  P.first_sys = first_matrix * S->n_rhs;
  P.n_sys = n_matrix * S->n_rhs;
  P.method = o->method == EFB_METHOD_AUTO ? (o->symmetric_hint ? EFB_METHOD_COCG : EFB_METHOD_BICGSTAB) : o->method;
  P.precond = o->precond;
  P.aux = (o->precond == EFB_PRECOND_AUX) && S->n_node > 0 && S->d_edge_nodes;
  if (o->precond == EFB_PRECOND_AUX && !P.aux) P.precond = EFB_PRECOND_JACOBI;
  P.D = make_dev(S, o->tolerance, o->max_iterations);
  for (int v = 0; v < V_NUM; ++v) P.vec[v] = S->d_work + (size_t)v * S->n_sys * S->m;
  P.vgrid = dim3((unsigned)vec_grid(c, S->m, P.n_sys), (unsigned)P.n_sys);
  P.ngrid = dim3((unsigned)vec_grid(c, std::max(1, S->n_node), P.n_sys), (unsigned)P.n_sys);
  return EFB_OK;
}

// Persistent path: used whenever p (+ nodal scratch) of NR right-hand sides fits in shared memory.
static int run_cocg_small(SolvePlan &P, const efb_solve_opts *o, bool zero_x, bool *ran) {
  System *S = P.S;
  Ctx *c = S->ctx;
  *ran = false;
  const char *off = getenv("EDGEFEM_B200_NO_PERSISTENT");
  if (off && off[0] == '1') return EFB_OK;
  int dev_smem = 0;
  EFB_CUDA(c, cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  const int nn = P.aux ? S->n_node : 0;
  int variant = 1;
  if (const char *v = getenv("EDGEFEM_B200_SMALL_VARIANT")) variant = atoi(v);
  if (S->m > 16384 || S->m != S->m_global) return EFB_OK;  // large or row-partitioned: generic multi-kernel path
  if (P.aux && (int)S->h_edge_nodes.size() != 2 * S->m) return EFB_OK;
  SubTrace st;
  if (S->small_dirty) {
    int rcs = build_small_structs(S);
    if (rcs) return rcs;
    st.mark("  small structs (compact numbering, SELL, lists)");
  }
  const int mc = S->m_c;
  if (mc == 0) return EFB_OK;
  int nr = (S->n_rhs % 2 == 0) ? 2 : 1;
  if (small_smem_bytes(nr, mc, nn) > (size_t)dev_smem) nr = 1;
  if (variant >= 4) nr = 1;  // one right-hand side per CTA, two CTAs per SM
  // Resident mode: one right-hand side per job with r and q in shared memory too (WR-90: 32 us per iteration against 35 us
  // with r, q in L2 -- the matrix stream, not shared with a second right-hand side, is half of it).  Used when every
  // right-hand side gets its own SM anyway (the all-split case of the head split below: 64 matrices 13.9 -> 12.5 ms); with
  // more work than SMs the two-rhs jobs are cheaper per right-hand side (256 matrices: 39.5 against 46.1 ms).
  const bool res_mode = variant < 4 && !getenv("EDGEFEM_B200_NO_RESIDENT") && small_smem_bytes_res(mc, nn) <= (size_t)dev_smem &&
                        (long long)P.n_matrix * S->n_rhs <= c->sm_count;
  if (res_mode) nr = 1;
  size_t smem = res_mode ? small_smem_bytes_res(mc, nn) : small_smem_bytes(nr, mc, nn);
  if (smem > (size_t)dev_smem || !S->d_sell_ptr) return EFB_OK;  // system too large for one CTA: generic multi-kernel path
  long long *d_prof = nullptr;
  static const bool prof_on = getenv("EDGEFEM_B200_SMALL_PROF") != nullptr;
  if (prof_on) {
    int rcp = dev_alloc(c, &d_prof, 8);
    if (rcp) return rcp;
    EFB_CUDA(c, cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), c->stream));
  }
  int stage_off = 0;  // 16-bit index lists behind p, w and the reduction scratch, when they fit (SmallLists)
  if (S->m <= 65535 && 2 * mc <= 65535 && variant < 4 && !getenv("EDGEFEM_B200_NO_STAGE") &&
      ((smem + 15) / 16 * 16) + small_lists_bytes(mc, nn) <= (size_t)dev_smem) {
    stage_off = (int)((smem + 15) / 16 * 16);
    smem = (size_t)stage_off + small_lists_bytes(mc, nn);
  }
  if (!S->d_sell_vals) {
    int rc0 = dev_alloc(c, &S->d_sell_vals, (size_t)S->n_matrix * (size_t)S->sell_total);
    if (rc0) return rc0;
  }
  {
    dim3 g((unsigned)S->n_slices, (unsigned)P.n_matrix);
    k_csr_to_sell<<<g, 256, 0, c->stream>>>(S->d_vals, (long long)S->nnz, S->d_sell_ptr, S->d_sell_src, S->d_sell_vals, S->sell_total, P.first_matrix);
    EFB_CHECK_LAUNCH(c);
  }
  const int groups = S->n_rhs / nr;
  int n_jobs = P.n_matrix * groups;
  int mixed = 0;
  {
    // Queue order = longest expected job first (LPT list scheduling): with n_jobs between 1x and 2x the SM count the
    // makespan is the sum of two jobs on one SM, so long jobs must start first and short ones fill in behind them.
    // Predictor: iteration counts of the previous solve of this system if there was one, else the assembly frequency
    // (the indefinite shift k0^2 M grows with frequency and so does the iteration count), else submission order.
    std::vector<double> w((size_t)n_jobs, 0.0);
    const bool have_it = (int)S->last_iters.size() == S->n_sys;
    const bool have_om = (int)S->last_omega.size() == S->n_matrix;
    for (int j = 0; j < n_jobs; ++j) {
      const int f = P.first_matrix + j / groups, s0 = f * S->n_rhs + (j % groups) * nr;
      if (have_it) {
        for (int k = 0; k < nr; ++k) w[j] = std::max(w[j], (double)S->last_iters[s0 + k]);
      } else if (have_om) {
        w[j] = S->last_omega[f];
      }
    }
    std::vector<int32_t> order((size_t)n_jobs);
    for (int j = 0; j < n_jobs; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return w[a] > w[b]; });
    std::vector<int32_t> codes(order);
    // Tail split (two right-hand sides per matrix, more matrices than SMs): 256 two-rhs jobs on 148 SMs leave 40 SMs
    // idle through the second round.  The X matrices expected to finish last are queued as two one-rhs jobs each (a
    // one-rhs job streams the matrix for itself: cost rho ~ 0.7-0.8 of a two-rhs job), X chosen by simulating the queue.
    static const bool no_split = [] {  // EDGEFEM_B200_TAIL_SPLIT=0|1 (default below)
      const char *e = getenv("EDGEFEM_B200_TAIL_SPLIT");
      return e ? atoi(e) == 0 : false;
    }();
    // Only with iteration counts of a previous solve of this system: with frequencies as the only predictor the split
    // was measured slower than no split (54.5 against 53.7 ms per end-to-end sweep).
    if (nr == 2 && S->n_rhs == 2 && n_jobs > c->sm_count && variant < 4 && !no_split && have_it) {
      double rho = 0.8;  // WR-90, same box: no split 48.0 / 49.7 ms, rho 0.5 57.1, 0.6 46.7, 0.7 46.4 / 46.7, 0.8 46.7 ms
      if (const char *e = getenv("EDGEFEM_B200_TAIL_RHO")) rho = atof(e);
      std::vector<double> cost((size_t)n_jobs);
      double wmin = 0.0;
      for (int j = 0; j < n_jobs; ++j) wmin = (j == 0) ? w[j] : std::min(wmin, w[j]);
      // weights are only a ranking when they are frequencies: map them to a plausible cost spread (+-15 %)
      double wmax = 0.0;
      for (int j = 0; j < n_jobs; ++j) wmax = std::max(wmax, w[j]);
      for (int j = 0; j < n_jobs; ++j)
        cost[j] = have_it ? std::max(1.0, w[j]) : (wmax > wmin ? 0.85 + 0.3 * (w[j] - wmin) / (wmax - wmin) : 1.0);
      auto makespan = [&](int X) {
        // queue: the n_jobs - X heaviest as two-rhs jobs, then 2X one-rhs jobs, all in descending cost; every SM takes
        // the next job when it becomes free (min-heap of SM finish times)
        std::vector<double> q;
        q.reserve((size_t)n_jobs + X);
        for (int i = 0; i < n_jobs - X; ++i) q.push_back(cost[order[i]]);
        for (int i = n_jobs - X; i < n_jobs; ++i) {
          q.push_back(rho * cost[order[i]]);
          q.push_back(rho * cost[order[i]]);
        }
        std::sort(q.begin(), q.end(), [](double a, double b) { return a > b; });
        std::priority_queue<double, std::vector<double>, std::greater<double>> busy;
        for (int i = 0; i < c->sm_count; ++i) busy.push(0.0);
        double end = 0.0;
        for (double t : q) {
          const double b = busy.top() + t;
          busy.pop();
          busy.push(b);
          end = std::max(end, b);
        }
        return end;
      };
      int bestX = 0;
      double best = makespan(0);
      for (int X = 8; X <= std::min(n_jobs, c->sm_count); X += 8) {
        const double t = makespan(X);
        if (t < best * 0.999) {
          best = t;
          bestX = X;
        }
      }
      if (bestX > 0) {
        mixed = 1;
        struct Q {
          double cost;
          int32_t code;
        };
        std::vector<Q> q;
        for (int i = 0; i < n_jobs - bestX; ++i) q.push_back({cost[order[i]], (int32_t)(order[i] * 4)});
        for (int i = n_jobs - bestX; i < n_jobs; ++i) {
          q.push_back({rho * cost[order[i]], (int32_t)(order[i] * 4 + 1)});
          q.push_back({rho * cost[order[i]], (int32_t)(order[i] * 4 + 2)});
        }
        std::stable_sort(q.begin(), q.end(), [](const Q &a, const Q &b) { return a.cost > b.cost; });
        codes.resize(q.size());
        for (size_t i = 0; i < q.size(); ++i) codes[i] = q[i].code;
        n_jobs = (int)codes.size();
      }
    }
    // Head split (one round: fewer matrices than SMs, e.g. the 128- or 64-point shard of a sweep over 2 or 4 GPUs): the launch
    // lasts as long as its longest job while sm_count - n_jobs SMs idle.  The X = idle-SM-count matrices expected to take
    // longest are queued as two one-rhs jobs each; with no more than sm_count / 2 matrices every one is split.  A one-rhs job
    // running next to idle SMs costs ~0.5 of the two-rhs job (measured: 64 matrices 13.9 ms all split, 128 matrices 22.3 ms
    // with 20 split, 26 ms unsplit).
    if (!mixed && nr == 2 && S->n_rhs == 2 && n_jobs <= c->sm_count && variant < 4 && !no_split && (have_it || have_om)) {
      const int X = std::min(c->sm_count - n_jobs, n_jobs);
      if (X > 0) {
        mixed = 1;
        codes.clear();
        for (int i = 0; i < X; ++i) {
          codes.push_back((int32_t)(order[i] * 4 + 1));
          codes.push_back((int32_t)(order[i] * 4 + 2));
        }
        for (int i = X; i < n_jobs; ++i) codes.push_back((int32_t)(order[i] * 4));
        n_jobs = (int)codes.size();
      }
    }
    if ((size_t)n_jobs + 1 > (size_t)S->n_sys + 1) return fail(c, EFB_ERR_STATE, "persistent solver: job queue overflow");
    S->h_job_order.resize((size_t)n_jobs + 1);
    S->h_job_order[0] = 0;
    for (int j = 0; j < n_jobs; ++j) S->h_job_order[1 + j] = codes[j];
    EFB_CUDA(c, cudaMemcpyAsync(S->d_job, S->h_job_order.data(), S->h_job_order.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  }
  if (!S->ev_s0) {
    EFB_CUDA(c, cudaEventCreate(&S->ev_s0));
    EFB_CUDA(c, cudaEventCreate(&S->ev_s1));
  }
  EFB_CUDA(c, cudaEventRecord(S->ev_s0, c->stream));
  const bool two_per_sm = variant >= 4 && 2 * (smem + 1024) <= 228 * 1024;
  const int grid = std::max(1, std::min(n_jobs, (two_per_sm ? 2 : 1) * c->sm_count));
  // kernel shape: threads per CTA, loads in flight per lane, register double buffering
  const int mr = o->max_restarts > 0 ? o->max_restarts : 3;
#define EFB_SMALL_LAUNCH(NRV, NT, SPDV, DBV)                                                                                              \
  do {                                                                                                                                    \
    EFB_CUDA(c, cudaFuncSetAttribute(k_cocg_small<NRV, NT, SPDV, DBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
    k_cocg_small<NRV, NT, SPDV, DBV><<<grid, NT, smem, c->stream>>>(P.D, S->d_sell_ptr, S->d_sell_col, S->d_sell_perm, S->d_sell_vals,    \
                                                                      S->sell_total, S->n_slices, P.first_matrix, n_jobs, groups, mixed, S->d_job, \
                                                                      S->d_b, S->d_x, P.vec[V_R], P.vec[V_Q], P.aux ? 1 : 0, zero_x ? 1 : 0, mr, \
                                                                      mc, S->d_c_orig, S->d_c_edge_nodes, S->d_c_n2e_ptr, S->d_c_n2e_item, stage_off, d_prof); \
  } while (0)
#define EFB_SMALL_LAUNCH2(NRV, NT, SPDV, DBV)                                                                                             \
  do {                                                                                                                                    \
    EFB_CUDA(c, cudaFuncSetAttribute(k_cocg_small<NRV, NT, SPDV, DBV, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    k_cocg_small<NRV, NT, SPDV, DBV, 2><<<grid, NT, smem, c->stream>>>(P.D, S->d_sell_ptr, S->d_sell_col, S->d_sell_perm, S->d_sell_vals, \
                                                                      S->sell_total, S->n_slices, P.first_matrix, n_jobs, groups, mixed, S->d_job, \
                                                                      S->d_b, S->d_x, P.vec[V_R], P.vec[V_Q], P.aux ? 1 : 0, zero_x ? 1 : 0, mr, \
                                                                      mc, S->d_c_orig, S->d_c_edge_nodes, S->d_c_n2e_ptr, S->d_c_n2e_item, stage_off, d_prof); \
  } while (0)
  if (res_mode) {
    auto kern = k_cocg_small<1, 1024, 4, false, 1, true>;
    EFB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 1024, smem, c->stream>>>(P.D, S->d_sell_ptr, S->d_sell_col, S->d_sell_perm, S->d_sell_vals, S->sell_total, S->n_slices, P.first_matrix,
                                          n_jobs, groups, mixed, S->d_job, S->d_b, S->d_x, P.vec[V_R], P.vec[V_Q], P.aux ? 1 : 0, zero_x ? 1 : 0, mr, mc,
                                          S->d_c_orig, S->d_c_edge_nodes, S->d_c_n2e_ptr, S->d_c_n2e_item, stage_off, d_prof);
  } else if (nr == 2) {
    switch (variant) {
      case 0: EFB_SMALL_LAUNCH(2, 512, 8, true); break;
      case 2: EFB_SMALL_LAUNCH(2, 1024, 0, false); break;
      case 3: EFB_SMALL_LAUNCH(2, 1024, 2, true); break;
      default: EFB_SMALL_LAUNCH(2, 1024, 4, false); break;
    }
  } else {
    switch (variant) {
      case 0: EFB_SMALL_LAUNCH(1, 512, 8, true); break;
      case 2: EFB_SMALL_LAUNCH(1, 1024, 0, false); break;
      case 3: EFB_SMALL_LAUNCH(1, 1024, 2, true); break;
      case 4: EFB_SMALL_LAUNCH2(1, 512, 4, false); break;
      case 5: EFB_SMALL_LAUNCH2(1, 512, 2, false); break;
      default: EFB_SMALL_LAUNCH(1, 1024, 4, false); break;
    }
  }
#undef EFB_SMALL_LAUNCH
#undef EFB_SMALL_LAUNCH2
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cudaEventRecord(S->ev_s1, c->stream));
  if (d_prof) {
    long long h[8];
    EFB_CUDA(c, cudaMemcpyAsync(h, d_prof, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    dfree(d_prof);
    double tot = 0;
    for (int k = 0; k < 8; ++k) tot += (double)h[k];
    tot = std::max(tot, 1.0);
    fprintf(stderr, "[edgefem-b200 small prof] share of CTA 0's iteration cycles: nodal gather %.3f, z pass %.3f, reduce %.3f, p update %.3f, SpMV %.3f, reduce (+ warp imbalance) %.3f, x/r update %.3f, loop %.3f (total %.0f cycles, lists %s)\n",
            h[0] / tot, h[1] / tot, h[2] / tot, h[3] / tot, h[4] / tot, h[5] / tot, h[6] / tot, h[7] / tot, tot, stage_off ? "in shared memory" : "in global memory");
  }
  S->small_timed = true;
  *ran = true;
  return EFB_OK;
}

int solver_alloc_public(System *S) { return solver_alloc(S); }

// ||b - A x|| / ||b|| of every system of the matrices [first, first+n): fills residual and converged (iters, method and
// precond are the caller's)
int true_residuals(System *S, int first_matrix, int n_matrix, double tol, efb_solve_result *results) {
  Ctx *c = S->ctx;
  int rc = solver_alloc(S);
  if (rc) return rc;
  efb_solve_opts o;
  memset(&o, 0, sizeof o);
  o.tolerance = tol;
  o.max_iterations = 0;
  o.precond = EFB_PRECOND_JACOBI;
  o.method = EFB_METHOD_COCG;
  SolvePlan P;
  if ((rc = make_plan(S, first_matrix, n_matrix, &o, P))) return rc;
  const int nsys = P.n_sys;
  std::vector<int32_t> act((size_t)nsys * 4);
  for (int i = 0; i < nsys; ++i) { act[i * 4 + ST_ACTIVE] = 1; act[i * 4 + ST_ITERS] = 0; act[i * 4 + ST_CONV] = 0; act[i * 4 + ST_REC] = 0; }
  EFB_CUDA(c, cudaMemcpyAsync(S->d_state + (size_t)P.first_sys * 4, act.data(), act.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  if ((rc = launch_spmv(S, P.D, P.first_matrix, P.n_matrix, S->d_x, P.vec[V_Q], nullptr, 0, 0, 0, 0))) return rc;
  k_init_residual<0><<<P.vgrid, VEC_THREADS, 0, c->stream>>>(P.D, P.first_sys, S->d_b, P.vec[V_Q], P.vec[V_R], P.vec[V_R0], 0);
  EFB_CHECK_LAUNCH(c);
  std::vector<c128> hscal((size_t)nsys * NSCAL);
  EFB_CUDA(c, cudaMemcpyAsync(hscal.data(), S->d_scal + (size_t)P.first_sys * NSCAL, hscal.size() * sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < nsys; ++i) {
    const double rr = hscal[(size_t)i * NSCAL + S_RR].x, bb = hscal[(size_t)i * NSCAL + S_BB].x;
    results[i].residual = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
    results[i].converged = (rr <= P.D.tol2 * bb * (1.0 + 1e-6)) && std::isfinite(rr) ? 1 : 0;
  }
  return EFB_OK;
}

}  // namespace efb

using namespace efb;

extern "C" {

int efb_solve(efb_system *sys_, int32_t first_matrix, int32_t n_matrix, const efb_solve_opts *opts, efb_solve_result *results) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_solve");
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_solve: NULL system");
  Ctx *c = S->ctx;
  if (!opts || !results || first_matrix < 0 || n_matrix <= 0 || first_matrix + n_matrix > S->n_matrix)
    return fail(c, EFB_ERR_INVALID, "efb_solve: bad arguments");
  if (!S->assembled) return fail(c, EFB_ERR_STATE, "efb_solve: matrix values were never assembled or set");
  if (!(opts->tolerance > 0.0) || opts->max_iterations < 0) return fail(c, EFB_ERR_INVALID, "efb_solve: bad tolerance / max_iterations");
  if (opts->method < 0 || opts->method > EFB_METHOD_COCG || opts->precond < 0 || opts->precond > EFB_PRECOND_NONE)
    return fail(c, EFB_ERR_INVALID, "efb_solve: unknown method / preconditioner");
  EFB_CUDA(c, cudaSetDevice(c->device));
  SubTrace st;
  SolvePlan P;
  int rc = make_plan(S, first_matrix, n_matrix, opts, P);
  if (rc) return rc;
  st.mark("solve: plan + workspace");
  const int check_every = opts->check_every > 0 ? opts->check_every : 32;
  const int max_restarts = opts->max_restarts > 0 ? opts->max_restarts : 3;
  const int nsys = P.n_sys;
  std::vector<int32_t> hstate((size_t)nsys * 4);
  Timed tm(c);
  k_state_begin<<<(nsys + 127) / 128, 128, 0, c->stream>>>(S->d_state, P.first_sys, nsys);
  EFB_CHECK_LAUNCH(c);
  if ((rc = setup_precond(P))) return rc;
  bool zero_x = opts->zero_initial_guess != 0;
  if (zero_x) {
    dim3 g((unsigned)vec_grid(c, S->m, nsys), (unsigned)nsys);
    k_fill_zero<<<g, VEC_THREADS, 0, c->stream>>>(S->d_x, S->m, P.first_sys);
    EFB_CHECK_LAUNCH(c);
  }
  auto read_state = [&]() -> int {
    EFB_CUDA(c, cudaMemcpyAsync(hstate.data(), S->d_state + (size_t)P.first_sys * 4, hstate.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    return EFB_OK;
  };
  auto any_active = [&]() {
    for (int i = 0; i < nsys; ++i)
      if (hstate[i * 4 + ST_ACTIVE]) return true;
    return false;
  };
  bool ran_small = false;
  S->small_timed = false;
  st.mark("solve: preconditioner launches");
  if (P.method == EFB_METHOD_COCG && opts->max_iterations > 0) {
    S->last_cluster_c = 0;
    if ((rc = run_krylov_cluster(P, opts, zero_x, &ran_small))) return rc;
    if (!ran_small && (rc = run_cocg_small(P, opts, zero_x, &ran_small))) return rc;
    st.mark("solve: persistent-solver set-up + launch (host)");
    if (ran_small && (rc = read_state())) return rc;
    st.mark("solve: wait for the device");
  }
  // Multi-kernel path: 5-7 launches per iteration.  For mid-size systems (10^4..10^6 unknowns) every kernel runs for a few
  // microseconds and the loop is bound by launch latency, so a batch of `check_every` iterations is captured ONCE into a
  // CUDA graph and replayed (scalars, activity flags and the convergence test live on the device; an inactive system makes
  // its kernels return at once).  Parity of the rho double buffer repeats with the (even) batch length; the first two
  // iterations of a cycle run outside the graph (BiCGSTAB's first iteration is special).  EDGEFEM_B200_NO_GRAPH=1: plain launches.
  static const bool no_graph = getenv("EDGEFEM_B200_NO_GRAPH") != nullptr;
  const int batch = (check_every + 1) & ~1;
  cudaGraphExec_t gexec = nullptr;
  long long graph_launches = 0;
  auto iteration = [&](int it) -> int {
    return P.method == EFB_METHOD_COCG ? cocg_iteration(P, it & 1) : bicg_iteration(P, it & 1, it == 0);
  };
  auto run_batch = [&]() -> int {  // `batch` iterations starting at an even iteration number
    if (no_graph) {
      for (int k = 0; k < batch; ++k) {
        const int rcb = iteration(2 + k);
        if (rcb) return rcb;
      }
      return EFB_OK;
    }
    if (!gexec) {
      const long long l0 = c->launches;
      EFB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      int rcb = EFB_OK;
      for (int k = 0; k < batch && !rcb; ++k) rcb = iteration(2 + k);
      cudaGraph_t g = nullptr;
      const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
      graph_launches = c->launches - l0;
      c->launches = l0;  // captured, not executed
      if (rcb) {
        if (g) cudaGraphDestroy(g);
        return rcb;
      }
      if (e != cudaSuccess) return fail(c, EFB_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
      const cudaError_t e2 = cudaGraphInstantiate(&gexec, g, 0);
      cudaGraphDestroy(g);
      if (e2 != cudaSuccess) return fail(c, EFB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
    }
    EFB_CUDA(c, cudaGraphLaunch(gexec, c->stream));
    c->launches += graph_launches;
    return EFB_OK;
  };
  struct GraphGuard {
    cudaGraphExec_t &g;
    ~GraphGuard() {
      if (g) cudaGraphExecDestroy(g);
    }
  } graph_guard{gexec};
  for (int cycle = 0; !ran_small && cycle <= max_restarts; ++cycle) {
    if ((rc = init_cycle(P, zero_x))) return rc;
    zero_x = false;
    if ((rc = read_state())) return rc;
    if (!any_active() || opts->max_iterations == 0) break;
    if ((rc = iteration(0)) || (rc = iteration(1))) return rc;
    while (true) {
      if ((rc = run_batch())) return rc;
      if ((rc = read_state())) return rc;
      if (!any_active()) break;
    }
    k_state_reactivate<<<(nsys + 127) / 128, 128, 0, c->stream>>>(S->d_state, P.first_sys, nsys, opts->max_iterations);
    EFB_CHECK_LAUNCH(c);
    if ((rc = read_state())) return rc;
    if (!any_active()) break;
    if (cycle == max_restarts) {
      // final verification pass for the systems re-activated just now
      if ((rc = init_cycle(P, false))) return rc;
      if ((rc = read_state())) return rc;
    }
  }
  // final true residuals: r = b - A x for every system (cheap, and independent of the path taken)
  for (int i = 0; i < nsys; ++i) {
    results[i].iters = hstate[i * 4 + ST_ITERS];
    results[i].method = P.method;
    results[i].precond = P.aux ? EFB_PRECOND_AUX : P.precond;
  }
  if ((rc = true_residuals(S, first_matrix, n_matrix, opts->tolerance, results))) return rc;
  st.mark("solve: final residuals + read-back");
  if ((int)S->last_iters.size() != S->n_sys) S->last_iters.assign((size_t)S->n_sys, 0);
  for (int i = 0; i < nsys; ++i) S->last_iters[(size_t)P.first_sys + i] = results[i].iters;
  return EFB_OK;
}

int efb_system_last_solve_kernel_ms(efb_system *sys_, double *ms) {
  System *S = (System *)sys_;
  if (!S || !ms) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_system_last_solve_kernel_ms: bad arguments");
  *ms = -1.0;
  if (!S->small_timed) return EFB_OK;  // the last solve did not use the persistent kernel
  EFB_CUDA(S->ctx, cudaEventSynchronize(S->ev_s1));
  float f = 0.f;
  EFB_CUDA(S->ctx, cudaEventElapsedTime(&f, S->ev_s0, S->ev_s1));
  *ms = (double)f;
  return EFB_OK;
}

int efb_system_last_solve_shape(efb_system *sys_, int32_t *cluster_ctas, int32_t *rhs_per_job, int32_t *n_clusters) {
  System *S = (System *)sys_;
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_system_last_solve_shape: NULL system");
  if (cluster_ctas) *cluster_ctas = S->last_cluster_c;
  if (rhs_per_job) *rhs_per_job = S->last_cluster_c ? S->last_cluster_nr : 0;
  if (n_clusters) *n_clusters = S->last_cluster_c ? S->last_cluster_n : 0;
  return EFB_OK;
}

int efb_spmv_host(efb_system *sys_, int32_t matrix, const double *x, double *y) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_spmv_host");
  if (!S || !x || !y || matrix < 0 || matrix >= S->n_matrix) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_spmv_host: bad arguments");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  int rc = solver_alloc(S);
  if (rc) return rc;
  SolveDev D = make_dev(S, 1e-10, 0);
  // use system slot matrix*n_rhs of work vectors P (in) and Q (out)
  c128 *vin = S->d_work + (size_t)V_P * S->n_sys * S->m, *vout = S->d_work + (size_t)V_Q * S->n_sys * S->m;
  const size_t off = (size_t)matrix * S->n_rhs * S->m;
  EFB_CUDA(c, cudaMemcpyAsync(vin + off, x, (size_t)S->m * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
  {
    Timed tm(c);
    if ((rc = launch_spmv(S, D, matrix, 1, vin, vout, nullptr, 0, 0, 0, 0))) return rc;
  }
  EFB_CUDA(c, cudaMemcpyAsync(y, vout + off, (size_t)S->m * sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_bench_kernel(efb_system *sys_, int32_t which, int32_t reps, double *avg_ms) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_bench_kernel");
  if (!S || !avg_ms || reps <= 0 || which < 0 || which > 4) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_bench_kernel: bad arguments");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  if (which == 4) {  // FP64 FMA probe: returns ms per launch of sm_count*8 CTAs x 256 threads x 8 chains x 4096 FMAs
    int rc0 = solver_alloc(S);
    if (rc0) return rc0;
    cudaEvent_t a0, a1;
    EFB_CUDA(c, cudaEventCreate(&a0));
    EFB_CUDA(c, cudaEventCreate(&a1));
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) EFB_CUDA(c, cudaEventRecord(a0, c->stream));
      for (int i = 0; i < (pass == 0 ? 2 : reps); ++i) {
        k_fp64_probe<<<c->sm_count * 8, 256, 0, c->stream>>>(S->d_partial, 1.0000001, 1e-9);
        EFB_CHECK_LAUNCH(c);
      }
      if (pass == 1) EFB_CUDA(c, cudaEventRecord(a1, c->stream));
    }
    EFB_CUDA(c, cudaEventSynchronize(a1));
    float msf = 0.f;
    EFB_CUDA(c, cudaEventElapsedTime(&msf, a0, a1));
    cudaEventDestroy(a0);
    cudaEventDestroy(a1);
    *avg_ms = (double)msf / reps;
    return EFB_OK;
  }
  if (!S->assembled) return fail(c, EFB_ERR_STATE, "efb_bench_kernel: assemble first");
  int rc;
  efb_solve_opts o;
  memset(&o, 0, sizeof o);
  o.tolerance = 1e-300;  // never converges: the timed iterations stay active
  o.max_iterations = 1 << 30;
  o.precond = EFB_PRECOND_JACOBI;
  o.method = which == 2 ? EFB_METHOD_COCG : EFB_METHOD_BICGSTAB;
  SolvePlan P;
  if (which != 3) {
    if ((rc = make_plan(S, 0, S->n_matrix, &o, P))) return rc;
    k_state_begin<<<(P.n_sys + 127) / 128, 128, 0, c->stream>>>(S->d_state, 0, P.n_sys);
    EFB_CHECK_LAUNCH(c);
    if ((rc = setup_precond(P))) return rc;
    if (which != 0 && (rc = init_cycle(P, false))) return rc;
  } else if (!S->mesh || S->last_mat_blob.empty()) {
    return fail(c, EFB_ERR_STATE, "efb_bench_kernel: no previous efb_assemble_volume call to replay");
  }
  cudaEvent_t e0, e1;
  EFB_CUDA(c, cudaEventCreate(&e0));
  EFB_CUDA(c, cudaEventCreate(&e1));
  const int count = (int)S->last_omega.size();
  for (int pass = 0; pass < 2; ++pass) {  // pass 0 = warm-up
    const int n = pass == 0 ? std::min(reps, 3) : reps;
    if (pass == 1) EFB_CUDA(c, cudaEventRecord(e0, c->stream));
    for (int i = 0; i < n; ++i) {
      if (which == 0) rc = launch_spmv(S, P.D, 0, S->n_matrix, S->d_x, P.vec[V_Q], nullptr, 0, 0, 0, 0);
      else if (which == 1) rc = bicg_iteration(P, i & 1, false);
      else if (which == 2) rc = cocg_iteration(P, i & 1);
      else rc = assemble_launch(S, 0, count, S->last_mode);
      if (rc) return rc;
    }
    if (pass == 1) EFB_CUDA(c, cudaEventRecord(e1, c->stream));
  }
  EFB_CUDA(c, cudaEventSynchronize(e1));
  float ms = 0.f;
  EFB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *avg_ms = (double)ms / reps;
  return EFB_OK;
}

}  // extern "C"
