// Row-partitioned solve of ONE large system over the GPUs of a node (second bullet: the single-large-mesh
// case C5).  One process per GPU; rank r owns the contiguous rows [r*chunk, min(m, (r+1)*chunk)) of the global edge
// space (efb_dist_row_range), assembles them with the ordinary volume kernel (efb_system_create_rows +
// efb_assemble_volume: the row-gather assembly needs no communication) and solves with COCG + Jacobi.
//
// The exchange step of the path is the SpMV input vector.  It is NOT exchanged by a collective: every rank maps the
// peers' copy of the vector through CUDA IPC and the SpMV kernel loads the off-rank entries it needs straight over
// NVLink (k_dist_spmv, halo_mode 0) -- compute and transfer are one kernel, only the halo entries move (a slab
// partition of a first-seen-by-tet numbering touches ~1-2 % remote columns).  halo_mode 1 is the library baseline
// beside it: ncclAllGather of the whole vector, then a local SpMV.
//
// Cross-rank ordering comes from the two scalar all-reduces a CG iteration needs anyway; the recurrence is arranged so
// that no third synchronisation is required: the SpMV is applied to z (complete on every rank before the all-reduce of
// r^T z returns) and  q = A z + beta q,  p = z + beta p  are formed in its epilogue.
//   K1  t = A z (remote z), q = t + beta q, p = z + beta p, partial p^T q        -> F1 -> all-reduce {p^T q}
//   K2  alpha = rho / p^T q, x += alpha p, r -= alpha q, z = D^-1 r, partial r^T z, |r|^2 -> F2 -> all-reduce {rho, rr}
// NCCL is resolved at run time (dlopen libnccl.so.2 -- the copy torch already loaded when the caller is Python), so the
// library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>

#include "common.cuh"
#include "spmv_tma.cuh"

namespace efb {

namespace {

struct NcclApi {
  void *lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi *nccl_api(std::string &err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // EDGEFEM_B200_NCCL_LIB overrides; otherwise the soname -- which resolves to the copy already loaded in the
    // process (torch's bundled NCCL when the caller is Python, see cabi._preload_nccl) or the system library
    const char *names[] = {getenv("EDGEFEM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n || !n[0]) continue;
      api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.AllGather || !api.CommDestroy) {
    err = "libnccl.so.2 not found or incomplete";
    return nullptr;
  }
  return &api;
}

struct Dist {  // per context
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
};

constexpr int DIST_MAX_WORLD = 8;
constexpr int DIST_RED_BLOCKS = 1184;  // 8 x 148 CTAs carry reduction partials
constexpr int DIST_VEC_THREADS = 256;

struct XView {  // how a kernel reads entry `col` (global edge id) of the distributed SpMV input
  const c128 *peer[DIST_MAX_WORLD];  // peer[o] = rank o's copy of its own rows (CUDA IPC mapping)
  const c128 *local;                 // this rank's own rows
  const c128 *full;                  // halo_mode 1: the all-gathered vector (index = global id)
  int chunk, row0, m_loc, mode;
};

__device__ __forceinline__ c128 xload(const XView &v, int col) {
  if (v.mode == 1) return v.full[col];
  const unsigned l = (unsigned)(col - v.row0);
  if (l < (unsigned)v.m_loc) return __ldg(&v.local[l]);  // on-rank (the vector is constant during the kernel)
  const int o = col / v.chunk;                            // block partition: owner by division
  // plain (coherent) load through the peer mapping: one NVLink read per halo entry
  const c128 *p = v.peer[o] + (col - o * v.chunk);
  c128 r;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ c128 ld_stream(const c128 *p) {
  c128 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int32_t *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// device scalars of a distributed solve (doubles)
enum {
  DS_PQ = 0,        // [2] p^T q (+1 pad: partial triples are written whole)   (all-reduced after K1)
  DS_RHO = 3,       // [2] r^T z, then [1] |r|^2   (all-reduced after K2)
  DS_RR = 5,
  DS_RHO_PREV = 6,  // [2] rho the current p was built with
  DS_BB = 8,        // [1] |b|^2 (+2 pad)  (all-reduced once)
  DS_SYNC = 12,     // [1] always 0: operand of the all-reduces that only order the ranks
  DS_RR2 = 13,      // [1] |r|^2 of the aux path (all-reduced on its own: orders the ranks before the nodal gather)
  DS_NUM = 16
};

// block-level sum of N doubles per thread -> partial[blockIdx][N]   (deterministic: fixed tree)
template <int N>
__device__ __forceinline__ void block_partial(double (&v)[N], double *__restrict__ partial) {
  __shared__ double red[8][N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) red[wid][k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double a = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w][k];
      partial[(size_t)blockIdx.x * N + k] = a;
    }
  }
}

// one CTA: sums the block partials in block order, writes them to sc[dst..dst+N); optionally rho -> rho_prev
template <int N>
__global__ void __launch_bounds__(256) k_dist_finish(const double *__restrict__ partial, int n_blocks, double *sc, int dst, int save_rho) {
  __shared__ double red[256][N];
  double a[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a[k] = 0.0;
  for (int b = threadIdx.x; b < n_blocks; b += 256)
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] += partial[(size_t)b * N + k];
#pragma unroll
  for (int k = 0; k < N; ++k) red[threadIdx.x][k] = a[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < N; ++k) red[threadIdx.x][k] += red[threadIdx.x + s][k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) sc[dst + k] = red[0][k];
    if (save_rho) {  // K1 of this iteration has consumed rho: it becomes rho_prev for alpha and the next beta
      sc[DS_RHO_PREV] = sc[DS_RHO];
      sc[DS_RHO_PREV + 1] = sc[DS_RHO + 1];
    }
  }
}

// K1 (EPI 1): t = A z over the view, q = t + beta q, p = z + beta p, partial p^T q.   beta = rho / rho_prev (0 if first)
// EPI 0:      r = b - A x over the view, z = dinv r, partial {r^T z, |r|^2}         (true residual / start of a cycle)
// CSR-stream mapping (see k_spmv_stream in solve.cu): a warp owns <= 256 entries of whole rows.
template <int EPI>
__global__ void __launch_bounds__(256)
k_dist_spmv(const int32_t *__restrict__ sp_chunk, int n_chunks, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
            const c128 *__restrict__ vals, const __grid_constant__ XView xv, const c128 *__restrict__ bvec, const c128 *__restrict__ dinv, c128 *__restrict__ zloc,
            c128 *__restrict__ r, c128 *__restrict__ p, c128 *__restrict__ q, const double *__restrict__ sc, int first, double *__restrict__ partial) {
  // products parked at skewed slots i + i/16 (the row sums read at a stride of one row: see spmv_slot in solve.cu)
  __shared__ c128 prod[8][SPMV_STREAM_W + SPMV_STREAM_W / 16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  c128 beta = cmake(0.0, 0.0);
  if (EPI == 1 && !first) {
    const c128 rho = cmake(sc[DS_RHO], sc[DS_RHO + 1]), rp = cmake(sc[DS_RHO_PREV], sc[DS_RHO_PREV + 1]);
    if (rp.x != 0.0 || rp.y != 0.0) beta = cdiv(rho, rp);
  }
  double d[3] = {0.0, 0.0, 0.0};
  for (int ch = blockIdx.x * 8 + wid; ch < n_chunks; ch += gridDim.x * 8) {
    const int r0 = __ldg(&sp_chunk[ch]), r1 = __ldg(&sp_chunk[ch + 1]);
    const int nrow = r1 - r0;
    int rs = 0;
    if (lane <= nrow) rs = __ldg(&rowptr[r0 + lane]);
    const int k0 = __shfl_sync(0xffffffffu, rs, 0);
    const int k1 = __shfl_sync(0xffffffffu, rs, nrow);
    const int re = __shfl_down_sync(0xffffffffu, rs, 1);
    c128 a[SPMV_STREAM_W / 32];
    int c[SPMV_STREAM_W / 32];
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      const int k = k0 + lane + 32 * j;
      const bool in = k < k1;
      a[j] = in ? ld_stream(&vals[k]) : cmake(0.0, 0.0);
      c[j] = in ? ld_stream(&colidx[k]) : -1;
    }
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j)
      if (c[j] >= 0) prod[wid][(lane + 32 * j) + ((lane + 32 * j) >> 4)] = cmul(a[j], xload(xv, c[j]));
    __syncwarp();
    if (lane < nrow) {
      const int row = r0 + lane;
      c128 acc = cmake(0.0, 0.0);
      for (int k = rs - k0; k < re - k0; ++k) acc = cadd(acc, prod[wid][k + (k >> 4)]);
      if (EPI == 1) {
        const c128 zi = zloc[row];
        c128 qi = acc, pi = zi;
        if (!first) {
          qi = cfma(beta, q[row], acc);
          pi = cfma(beta, p[row], zi);
        }
        q[row] = qi;
        p[row] = pi;
        const c128 t = cmul(pi, qi);
        d[0] += t.x; d[1] += t.y;
      } else {
        const c128 ri = csub(bvec[row], acc);
        const c128 zi = cmul(dinv[row], ri);
        r[row] = ri;
        q[row] = zi;  // parked: z may still be read by peers if it aliases the SpMV input; the caller copies q -> z
        const c128 t = cmul(ri, zi);
        d[0] += t.x; d[1] += t.y;
        d[2] += cabs2(ri);
      }
    }
    __syncwarp();
  }
  block_partial<3>(d, partial);
}

// The same SpMV with the matrix stream moved by the TMA engine (see k_spmv_tma in solve.cu): lane 0 of a warp issues two
// bulk copies per chunk (values, columns) that complete on an mbarrier one chunk ahead of the warp (2-stage ring in shared
// memory), so the LSU only sees the gather of the input vector -- on-rank entries from local memory, halo entries straight
// from the peers' memory -- the in-place products and the in-order row sums.  Persistent: 2 CTAs x 8 warps per SM.
// Same chunks, products and summation order as k_dist_spmv: bit-identical results.
template <int EPI>
__global__ void __launch_bounds__(256, 2)
k_dist_spmv_tma(const int32_t *__restrict__ sp_chunk, int n_chunks, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                const c128 *__restrict__ vals, const __grid_constant__ XView xv, const c128 *__restrict__ bvec, const c128 *__restrict__ dinv,
                c128 *__restrict__ zloc, c128 *__restrict__ r, c128 *__restrict__ p, c128 *__restrict__ q, const double *__restrict__ sc, int first,
                double *__restrict__ partial) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
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
  c128 beta = cmake(0.0, 0.0);
  if (EPI == 1 && !first) {
    const c128 rho = cmake(sc[DS_RHO], sc[DS_RHO + 1]), rp = cmake(sc[DS_RHO_PREV], sc[DS_RHO_PREV + 1]);
    if (rp.x != 0.0 || rp.y != 0.0) beta = cdiv(rho, rp);
  }
  double d[3] = {0.0, 0.0, 0.0};
  const int stride = gridDim.x * 8;
  int ch = blockIdx.x * 8 + wid;
  auto issue = [&](int stage, const ChunkDesc &cd) {
    if (lane == 0 && cd.nrow > 0) {
      const unsigned sb = wbase_s + (unsigned)stage * SPMV_TMA_STAGE;
      const unsigned bar = bar_s + 8u * (unsigned)stage;
      const int ka = cd.k0 & ~3;
      const unsigned vbytes = (unsigned)(cd.k1 - cd.k0) * 16u;
      const unsigned cbytes = (unsigned)((cd.k1 - ka + 3) >> 2) * 16u;
      fence_proxy_async();
      mbar_expect_tx(bar, vbytes + cbytes);
      if (vbytes) bulk_g2s(sb, vals + cd.k0, vbytes, bar);
      if (cbytes) bulk_g2s(sb + SPMV_VAL_BYTES, colidx + ka, cbytes, bar);
    }
  };
  ChunkDesc cur = load_chunk_desc(sp_chunk, rowptr, ch, n_chunks);
  ChunkDesc nxt = load_chunk_desc(sp_chunk, rowptr, ch + stride, n_chunks);
  int rs_cur = (lane <= cur.nrow && cur.nrow > 0) ? __ldg(&rowptr[cur.r0 + lane]) : 0;
  issue(0, cur);
  int st = 0;
  unsigned phase0 = 0, phase1 = 0;
  for (; ch < n_chunks; ch += stride) {
    issue(st ^ 1, nxt);
    const ChunkDesc nn = load_chunk_desc(sp_chunk, rowptr, ch + 2 * stride, n_chunks);
    const int rs_nxt = (lane <= nxt.nrow && nxt.nrow > 0) ? __ldg(&rowptr[nxt.r0 + lane]) : 0;
    {
      const unsigned bar = bar_s + 8u * (unsigned)st;
      const unsigned par = st ? phase1 : phase0;
      while (!mbar_try_wait(bar, par)) {
      }
      if (st) phase1 ^= 1u; else phase0 ^= 1u;
    }
    unsigned char *sb = wbase + (size_t)st * SPMV_TMA_STAGE;
    c128 *sv = (c128 *)sb;
    const int32_t *scol = (const int32_t *)(sb + SPMV_VAL_BYTES) + (cur.k0 & 3);
    const int n = cur.k1 - cur.k0;
    int c[SPMV_STREAM_W / 32];
    c128 xv8[SPMV_STREAM_W / 32];
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      const int i = lane + 32 * j;
      c[j] = (i < n) ? scol[i] : -1;
    }
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) xv8[j] = (c[j] >= 0) ? xload(xv, c[j]) : cmake(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j) {
      const int i = lane + 32 * j;
      if (c[j] >= 0) xv8[j] = cmul(sv[i], xv8[j]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SPMV_STREAM_W / 32; ++j)
      if (c[j] >= 0) sv[spmv_slot(lane + 32 * j)] = xv8[j];
    __syncwarp();
    const int re_cur = __shfl_down_sync(0xffffffffu, rs_cur, 1);
    if (lane < cur.nrow) {
      const int row = cur.r0 + lane;
      c128 acc = cmake(0.0, 0.0);
      for (int k = rs_cur - cur.k0; k < re_cur - cur.k0; ++k) acc = cadd(acc, sv[spmv_slot(k)]);
      if (EPI == 1) {
        const c128 zi = zloc[row];
        c128 qi = acc, pi = zi;
        if (!first) {
          qi = cfma(beta, q[row], acc);
          pi = cfma(beta, p[row], zi);
        }
        q[row] = qi;
        p[row] = pi;
        const c128 t = cmul(pi, qi);
        d[0] += t.x; d[1] += t.y;
      } else {
        const c128 ri = csub(bvec[row], acc);
        const c128 zi = cmul(dinv[row], ri);
        r[row] = ri;
        q[row] = zi;
        const c128 t = cmul(ri, zi);
        d[0] += t.x; d[1] += t.y;
        d[2] += cabs2(ri);
      }
    }
    __syncwarp();
    cur = nxt;
    nxt = nn;
    rs_cur = rs_nxt;
    st ^= 1;
  }
  block_partial<3>(d, partial);
}

// K2: alpha = rho_prev / p^T q ; x += alpha p ; r -= alpha q ; z = dinv r ; partial {r^T z, |r|^2}
__global__ void __launch_bounds__(DIST_VEC_THREADS)
k_dist_update(int m, const double *__restrict__ sc, const c128 *__restrict__ dinv, const c128 *__restrict__ p, const c128 *__restrict__ q,
              c128 *__restrict__ x, c128 *__restrict__ r, c128 *__restrict__ z, double *__restrict__ partial) {
  const c128 pq = cmake(sc[DS_PQ], sc[DS_PQ + 1]), rho = cmake(sc[DS_RHO_PREV], sc[DS_RHO_PREV + 1]);
  c128 alpha = cmake(0.0, 0.0);
  if ((pq.x != 0.0 || pq.y != 0.0) && isfinite(pq.x) && isfinite(pq.y)) alpha = cdiv(rho, pq);
  double d[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    x[i] = cfma(alpha, p[i], x[i]);
    const c128 ri = cfma(cneg(alpha), q[i], r[i]);
    const c128 zi = cmul(dinv[i], ri);
    r[i] = ri;
    z[i] = zi;
    const c128 t = cmul(ri, zi);
    d[0] += t.x; d[1] += t.y;
    d[2] += cabs2(ri);
  }
  block_partial<3>(d, partial);
}

__global__ void k_dist_dinv(int m, const int32_t *__restrict__ diag_pos, const c128 *__restrict__ vals, c128 *__restrict__ dinv, int jacobi) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    c128 v = cmake(1.0, 0.0);
    const int pp = diag_pos[i];
    if (jacobi && pp >= 0) {
      const c128 a = vals[pp];
      if (a.x != 0.0 || a.y != 0.0) v = cdiv(cmake(1.0, 0.0), a);
    }
    dinv[i] = v;
  }
}

__global__ void k_dist_copy(int m, const c128 *__restrict__ src, c128 *__restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void k_dist_norm2(int m, const c128 *__restrict__ v, double *__restrict__ partial) {
  double d[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) d[0] += cabs2(v[i]);
  block_partial<3>(d, partial);
}

// ---- auxiliary-space pieces
// K2a: alpha = rho_prev / p^T q ; x += alpha p ; r -= alpha q ; partial |r|^2   (r lives in the exported buffer)
__global__ void __launch_bounds__(DIST_VEC_THREADS)
k_dist_update_r(int m, const double *__restrict__ sc, const c128 *__restrict__ p, const c128 *__restrict__ q, c128 *__restrict__ x,
                c128 *__restrict__ r, double *__restrict__ partial) {
  const c128 pq = cmake(sc[DS_PQ], sc[DS_PQ + 1]), rho = cmake(sc[DS_RHO_PREV], sc[DS_RHO_PREV + 1]);
  c128 alpha = cmake(0.0, 0.0);
  if ((pq.x != 0.0 || pq.y != 0.0) && isfinite(pq.x) && isfinite(pq.y)) alpha = cdiv(rho, pq);
  double d[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    x[i] = cfma(alpha, p[i], x[i]);
    const c128 ri = cfma(cneg(alpha), q[i], r[i]);
    r[i] = ri;
    d[0] += cabs2(ri);
  }
  block_partial<3>(d, partial);
}

// N1: w[slot] = linv[slot] * sum over the free edges at my node of +-r (off-rank entries through the peer mappings)
__global__ void __launch_bounds__(DIST_VEC_THREADS)
k_dist_node(int n_my, const int32_t *__restrict__ mn_ptr, const int32_t *__restrict__ mn_item, const c128 *__restrict__ linv,
            const __grid_constant__ XView rv, c128 *__restrict__ w) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_my; j += gridDim.x * blockDim.x) {
    const c128 li = linv[j];
    c128 a = cmake(0.0, 0.0);
    if (li.x != 0.0 || li.y != 0.0) {
      for (int k = mn_ptr[j]; k < mn_ptr[j + 1]; ++k) {
        const int it = mn_item[k];
        const c128 v = xload(rv, it >> 1);
        a = (it & 1) ? cadd(a, v) : csub(a, v);
      }
      a = cmul(li, a);
    }
    w[j] = a;
  }
}

// E1: z = dinv r + (w[head] - w[tail]) on free edges ; partial {r^T z}
__global__ void __launch_bounds__(DIST_VEC_THREADS)
k_dist_edge(int m, const c128 *__restrict__ dinv, const c128 *__restrict__ r, const int2 *__restrict__ slot, const uint8_t *__restrict__ dir,
            const c128 *__restrict__ w, c128 *__restrict__ z, double *__restrict__ partial) {
  double d[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const c128 ri = r[i];
    c128 zi = cmul(dinv[i], ri);
    if (!dir[i]) {
      const int2 ab = slot[i];
      zi = cadd(zi, csub(w[ab.y], w[ab.x]));
    }
    z[i] = zi;
    const c128 t = cmul(ri, zi);
    d[0] += t.x; d[1] += t.y;
  }
  block_partial<3>(d, partial);
}

// r = b - t (t = A x parked in q by k_dist_spmv<2>) ; partial |r|^2
__global__ void __launch_bounds__(DIST_VEC_THREADS)
k_dist_resid(int m, const c128 *__restrict__ b, const c128 *__restrict__ t, c128 *__restrict__ r, double *__restrict__ partial) {
  double d[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const c128 ri = csub(b[i], t[i]);
    r[i] = ri;
    d[0] += cabs2(ri);
  }
  block_partial<3>(d, partial);
}

// nodal diagonal of G^T A G restricted to the local rows, accumulated into the full nodal array (summed over ranks by NCCL)
__global__ void k_dist_nodal_diag(int m_loc, int row0, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const c128 *__restrict__ vals,
                                  const uint8_t *__restrict__ dir_all, const int2 *__restrict__ en_all, c128 *L) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m_loc; i += gridDim.x * blockDim.x) {
    const int e = row0 + i;
    if (dir_all[e]) continue;
    const int2 ab = en_all[e];
    c128 la = cmake(0.0, 0.0), lb = cmake(0.0, 0.0);
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int e2 = colidx[k];
      if (dir_all[e2]) continue;
      const int2 cd = en_all[e2];
      const c128 a = vals[k];
      if (ab.x == cd.x) la = cadd(la, a);
      if (ab.x == cd.y) la = csub(la, a);
      if (ab.y == cd.y) lb = cadd(lb, a);
      if (ab.y == cd.x) lb = csub(lb, a);
    }
    atomicAdd(&L[ab.x].x, la.x); atomicAdd(&L[ab.x].y, la.y);
    atomicAdd(&L[ab.y].x, lb.x); atomicAdd(&L[ab.y].y, lb.y);
  }
}
__global__ void k_dist_linv(int n_my, const int32_t *__restrict__ my_node, const uint8_t *__restrict__ my_dir, const c128 *__restrict__ L, c128 *linv) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_my; j += gridDim.x * blockDim.x) {
    const c128 a = L[my_node[j]];
    c128 v = cmake(0.0, 0.0);
    if (!my_dir[j] && (a.x != 0.0 || a.y != 0.0)) v = cdiv(cmake(1.0, 0.0), a);
    linv[j] = v;
  }
}

struct DistState {  // per row-partitioned system
  int world = 1, rank = 0, chunk = 0;
  c128 *d_z = nullptr;      // [chunk] SpMV input of the iteration (IPC-shared)
  c128 *d_xs = nullptr;     // [chunk] solution copy the peers read for the true residual (IPC-shared)
  c128 *d_r = nullptr, *d_p = nullptr, *d_q = nullptr, *d_dinv = nullptr;  // [m_loc]
  c128 *d_full = nullptr;   // [world*chunk] halo_mode 1 gather buffer (lazy)
  double *d_sc = nullptr;   // [DS_NUM]
  double *d_partial = nullptr;  // [DIST_RED_BLOCKS*3]
  unsigned char *d_handles = nullptr;  // [world][2][64] IPC handles (all-gathered)
  c128 *peer_z[DIST_MAX_WORLD] = {}, *peer_x[DIST_MAX_WORLD] = {}, *peer_r[DIST_MAX_WORLD] = {};
  bool opened = false;
  // auxiliary-space preconditioner z = D^-1 r + G L^-1 G^T r on the row partition: every rank keeps the nodes of its
  // own edges ("my nodes": owned by it or shared with a neighbour), gathers G^T r for them over ALL their incident
  // edges -- off-rank residual entries are read from the peers' exported r -- and applies G w with purely local data.
  bool aux = false;
  c128 *d_rexp = nullptr;        // [chunk] residual (IPC-shared): replaces d_r when aux
  int n_my = 0;
  int32_t *d_my_node = nullptr;  // [n_my] global node id
  int32_t *d_mn_ptr = nullptr;   // [n_my+1]
  int32_t *d_mn_item = nullptr;  // global edge id << 1 | head, free edges only
  int2 *d_edge_slot = nullptr;   // [m_loc] my-node slots of (tail, head)
  uint8_t *d_my_dir = nullptr;   // [n_my] node touches a Dirichlet edge
  c128 *d_linv = nullptr;        // [n_my]
  c128 *d_wloc = nullptr;        // [n_my]
  c128 *d_Lfull = nullptr;       // [n_node] nodal diagonal of G^T A G, summed over the ranks (set-up only)
  int2 *d_en_all = nullptr;      // [m_global] end nodes of every edge (set-up only)
};

}  // namespace

static int nccl_fail(Ctx *c, NcclApi *api, ncclResult_t r, const char *what) {
  return fail(c, EFB_ERR_CUDA, "%s failed: %s", what, api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
}

#define EFB_NCCL(c, api, expr)                                   \
  do {                                                           \
    ncclResult_t _r = (expr);                                    \
    if (_r != ncclSuccess) return nccl_fail((c), (api), _r, #expr); \
  } while (0)

void dist_free(System *S) {
  DistState *T = (DistState *)S->dist_state;
  if (!T) return;
  cudaStreamSynchronize(S->ctx->stream);
  if (T->opened)
    for (int o = 0; o < T->world; ++o)
      if (o != T->rank) {
        if (T->peer_z[o]) cudaIpcCloseMemHandle(T->peer_z[o]);
        if (T->peer_x[o]) cudaIpcCloseMemHandle(T->peer_x[o]);
        if (T->peer_r[o]) cudaIpcCloseMemHandle(T->peer_r[o]);
      }
  // IPC-exported allocations go straight back to the driver (a pooled block must not be re-issued while a peer maps it)
  cudaFree(T->d_z);
  cudaFree(T->d_xs);
  cudaFree(T->d_rexp);
  dfree(T->d_my_node); dfree(T->d_mn_ptr); dfree(T->d_mn_item); dfree(T->d_edge_slot); dfree(T->d_my_dir); dfree(T->d_linv); dfree(T->d_wloc);
  dfree(T->d_Lfull); dfree(T->d_en_all);
  dfree(T->d_r); dfree(T->d_p); dfree(T->d_q); dfree(T->d_dinv); dfree(T->d_full); dfree(T->d_sc); dfree(T->d_partial);
  dfree(T->d_handles);
  delete T;
  S->dist_state = nullptr;
}

static int dist_prepare(System *S, DistState **out, bool want_aux = false) {
  Ctx *c = S->ctx;
  Dist *dd = (Dist *)c->dist;
  if (!dd) return fail(c, EFB_ERR_STATE, "efb_dist_*: call efb_dist_init on the context first");
  if (S->n_matrix != 1 || S->n_rhs != 1) return fail(c, EFB_ERR_INVALID, "efb_dist_*: row-partitioned systems carry one matrix and one right-hand side");
  if (!S->d_sp_chunk) return fail(c, EFB_ERR_LIMIT, "efb_dist_*: a row has more than %d entries", SPMV_STREAM_W);
  if (S->dist_state && want_aux && !((DistState *)S->dist_state)->aux) dist_free(S);  // rebuild with the nodal lists
  if (S->dist_state) {
    *out = (DistState *)S->dist_state;
    return EFB_OK;
  }
  if (want_aux && (!S->mesh || (int)S->mesh->h_edge_nodes.size() != 2 * S->m_global || (int)S->h_dir_all.size() != S->m_global))
    return fail(c, EFB_ERR_STATE, "efb_dist_solve: the auxiliary-space preconditioner needs a mesh-born system with Dirichlet flags set");
  std::string err;
  NcclApi *api = nccl_api(err);
  if (!api) return fail(c, EFB_ERR_STATE, "efb_dist_*: %s", err.c_str());
  const int world = dd->world, rank = dd->rank;
  const int chunk = (S->m_global + world - 1) / world;
  if (S->row0 != std::min(S->m_global, rank * chunk) || S->row0 + S->m != std::min(S->m_global, (rank + 1) * chunk))
    return fail(c, EFB_ERR_INVALID, "efb_dist_*: rank %d must own rows [%d, %d) (efb_dist_row_range), system has [%d, %d)", rank,
                std::min(S->m_global, rank * chunk), std::min(S->m_global, (rank + 1) * chunk), S->row0, S->row0 + S->m);
  DistState *T = new DistState();
  S->dist_state = T;
  T->world = world;
  T->rank = rank;
  T->chunk = chunk;
  int rc;
  // exported vectors: their own cudaMalloc blocks (never pooled), padded to `chunk` so an all-gather has equal counts
  EFB_CUDA(c, cudaMalloc((void **)&T->d_z, (size_t)chunk * sizeof(c128)));
  EFB_CUDA(c, cudaMalloc((void **)&T->d_xs, (size_t)chunk * sizeof(c128)));
  EFB_CUDA(c, cudaMalloc((void **)&T->d_rexp, (size_t)chunk * sizeof(c128)));
  EFB_CUDA(c, cudaMemsetAsync(T->d_rexp, 0, (size_t)chunk * sizeof(c128), c->stream));
  EFB_CUDA(c, cudaMemsetAsync(T->d_z, 0, (size_t)chunk * sizeof(c128), c->stream));
  EFB_CUDA(c, cudaMemsetAsync(T->d_xs, 0, (size_t)chunk * sizeof(c128), c->stream));
  if ((rc = dev_alloc(c, &T->d_r, (size_t)S->m))) return rc;
  if ((rc = dev_alloc(c, &T->d_p, (size_t)S->m))) return rc;
  if ((rc = dev_alloc(c, &T->d_q, (size_t)S->m))) return rc;
  if ((rc = dev_alloc(c, &T->d_dinv, (size_t)S->m))) return rc;
  if ((rc = dev_alloc(c, &T->d_sc, (size_t)DS_NUM))) return rc;
  if ((rc = dev_alloc(c, &T->d_partial, (size_t)DIST_RED_BLOCKS * 3))) return rc;
  if ((rc = dev_alloc(c, &T->d_handles, (size_t)world * 3 * sizeof(cudaIpcMemHandle_t)))) return rc;
  if (want_aux) {
    // my nodes = nodes of the own edges; their lists hold ALL incident free edges (global ids)
    const Mesh *M = S->mesh;
    const int nn = M->n_node, mg = S->m_global;
    const int32_t *en = M->h_edge_nodes.data();
    const uint8_t *dirg = S->h_dir_all.data();
    std::vector<int32_t> slot_of((size_t)nn, -1), my_node;
    for (int i = 0; i < S->m; ++i)
      for (int k = 0; k < 2; ++k) {
        const int nd = en[2 * (size_t)(S->row0 + i) + k];
        if (slot_of[nd] < 0) slot_of[nd] = 0;
      }
    for (int nd = 0; nd < nn; ++nd)
      if (slot_of[nd] == 0) {
        slot_of[nd] = (int32_t)my_node.size();
        my_node.push_back(nd);
      }
    const int n_my = (int)my_node.size();
    std::vector<int32_t> ptr((size_t)n_my + 1, 0);
    std::vector<uint8_t> my_dir((size_t)std::max(n_my, 1), 0);
    for (int e = 0; e < mg; ++e)
      for (int k = 0; k < 2; ++k) {
        const int sl = slot_of[en[2 * (size_t)e + k]];
        if (sl < 0) continue;
        if (dirg[e]) my_dir[sl] = 1;
        else ptr[sl + 1]++;
      }
    for (int j = 0; j < n_my; ++j) ptr[j + 1] += ptr[j];
    std::vector<int32_t> item((size_t)std::max(ptr[n_my], 1)), cur(ptr.begin(), ptr.end() - 1);
    for (int e = 0; e < mg; ++e) {
      if (dirg[e]) continue;
      for (int k = 0; k < 2; ++k) {
        const int sl = slot_of[en[2 * (size_t)e + k]];
        if (sl >= 0) item[cur[sl]++] = (e << 1) | k;  // k = 1: head (+1)
      }
    }
    std::vector<int2> eslot((size_t)std::max(S->m, 1));
    for (int i = 0; i < S->m; ++i) eslot[i] = make_int2(slot_of[en[2 * (size_t)(S->row0 + i)]], slot_of[en[2 * (size_t)(S->row0 + i) + 1]]);
    T->n_my = n_my;
    if ((rc = dev_upload(c, &T->d_my_node, my_node.data(), (size_t)std::max(n_my, 1)))) return rc;
    if ((rc = dev_upload(c, &T->d_mn_ptr, ptr.data(), ptr.size()))) return rc;
    if ((rc = dev_upload(c, &T->d_mn_item, item.data(), item.size()))) return rc;
    if ((rc = dev_upload(c, &T->d_edge_slot, eslot.data(), eslot.size()))) return rc;
    if ((rc = dev_upload(c, &T->d_my_dir, my_dir.data(), my_dir.size()))) return rc;
    if ((rc = dev_upload(c, &T->d_en_all, (const int2 *)en, (size_t)mg))) return rc;
    if ((rc = dev_alloc(c, &T->d_linv, (size_t)std::max(n_my, 1)))) return rc;
    if ((rc = dev_alloc(c, &T->d_wloc, (size_t)std::max(n_my, 1)))) return rc;
    if ((rc = dev_alloc(c, &T->d_Lfull, (size_t)nn))) return rc;
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vectors above are locals
    T->aux = true;
  }
  EFB_CUDA(c, cudaMemsetAsync(T->d_sc, 0, DS_NUM * sizeof(double), c->stream));
  T->peer_z[rank] = T->d_z;
  T->peer_x[rank] = T->d_xs;
  T->peer_r[rank] = T->d_rexp;
  if (world > 1) {
    cudaIpcMemHandle_t mine[3];
    EFB_CUDA(c, cudaIpcGetMemHandle(&mine[0], T->d_z));
    EFB_CUDA(c, cudaIpcGetMemHandle(&mine[1], T->d_xs));
    EFB_CUDA(c, cudaIpcGetMemHandle(&mine[2], T->d_rexp));
    EFB_CUDA(c, cudaMemcpyAsync(T->d_handles + (size_t)rank * sizeof(mine), mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    EFB_NCCL(c, api, api->AllGather(T->d_handles + (size_t)rank * sizeof(mine), T->d_handles, sizeof(mine), ncclUint8, dd->comm, c->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)world * 3);
    EFB_CUDA(c, cudaMemcpyAsync(all.data(), T->d_handles, all.size() * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, c->stream));
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int o = 0; o < world; ++o) {
      if (o == rank) continue;
      EFB_CUDA(c, cudaIpcOpenMemHandle((void **)&T->peer_z[o], all[(size_t)o * 3], cudaIpcMemLazyEnablePeerAccess));
      EFB_CUDA(c, cudaIpcOpenMemHandle((void **)&T->peer_x[o], all[(size_t)o * 3 + 1], cudaIpcMemLazyEnablePeerAccess));
      EFB_CUDA(c, cudaIpcOpenMemHandle((void **)&T->peer_r[o], all[(size_t)o * 3 + 2], cudaIpcMemLazyEnablePeerAccess));
    }
    T->opened = true;
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = T;
  return EFB_OK;
}

static XView make_view(const System *S, const DistState *T, c128 *const *peers, const c128 *local, int mode) {
  XView v;
  for (int o = 0; o < DIST_MAX_WORLD; ++o) v.peer[o] = o < T->world ? peers[o] : nullptr;
  v.local = local;
  v.full = T->d_full;
  v.chunk = T->chunk;
  v.row0 = S->row0;
  v.m_loc = S->m;
  v.mode = mode;
  return v;
}


// grid of the distributed SpMV and its launch (TMA ring by default; EDGEFEM_B200_SPMV_KERNEL=regs: register streaming)
static bool dist_spmv_regs() {
  static const bool v = [] {
    const char *e = getenv("EDGEFEM_B200_SPMV_KERNEL");
    return e && strcmp(e, "regs") == 0;
  }();
  return v;
}
static int spmv_grid(const Ctx *c, int n_chunks) {
  if (!dist_spmv_regs()) return std::max(1, std::min((n_chunks + 7) / 8, std::min(DIST_RED_BLOCKS, c->sm_count * 2)));
  return std::max(1, std::min((n_chunks + 7) / 8, DIST_RED_BLOCKS));
}
template <int EPI>
static int launch_dist_spmv(System *S, DistState *T, const XView &xv, c128 *r, int first) {
  Ctx *c = S->ctx;
  const int nb = spmv_grid(c, S->n_sp_chunks);
  if (dist_spmv_regs()) {
    k_dist_spmv<EPI><<<nb, 256, 0, c->stream>>>(S->d_sp_chunk, S->n_sp_chunks, S->d_rowptr, S->d_colidx, S->d_vals, xv, S->d_b, T->d_dinv, T->d_z,
                                                r, T->d_p, T->d_q, T->d_sc, first, T->d_partial);
  } else {
    const size_t smem = (size_t)8 * 2 * SPMV_TMA_STAGE + 8 * 16;
    EFB_CUDA(c, cudaFuncSetAttribute(k_dist_spmv_tma<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_dist_spmv_tma<EPI><<<nb, 256, smem, c->stream>>>(S->d_sp_chunk, S->n_sp_chunks, S->d_rowptr, S->d_colidx, S->d_vals, xv, S->d_b, T->d_dinv,
                                                       T->d_z, r, T->d_p, T->d_q, T->d_sc, first, T->d_partial);
  }
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}
static int vec_blocks(const Ctx *c, int m) { return std::max(1, std::min((m + DIST_VEC_THREADS - 1) / DIST_VEC_THREADS, DIST_RED_BLOCKS)); }

// halo_mode 1: gather the exported vector of every rank into d_full
static int gather_full(System *S, DistState *T, NcclApi *api, const c128 *mine) {
  Ctx *c = S->ctx;
  Dist *dd = (Dist *)c->dist;
  if (!T->d_full) {
    int rc = dev_alloc(c, &T->d_full, (size_t)T->world * T->chunk);
    if (rc) return rc;
  }
  EFB_NCCL(c, api, api->AllGather(mine, T->d_full, (size_t)T->chunk * 2, ncclFloat64, dd->comm, c->stream));
  return EFB_OK;
}

struct DistRun {
  System *S;
  DistState *T;
  NcclApi *api;
  Dist *dd;
  int mode;
  bool aux = false;
};

// z = D^-1 r + G L^-1 G^T r with r complete on every rank: nodal gather over the peers' residuals, then the local edge pass
// (z goes to the exported buffer), all-reduce {r^T z} -- which also orders the ranks before the next SpMV reads z
static int dist_aux_precond(const DistRun &R) {
  System *S = R.S;
  DistState *T = R.T;
  Ctx *c = S->ctx;
  const int vb = vec_blocks(c, S->m), nb = vec_blocks(c, std::max(T->n_my, 1));
  if (R.mode == 1) {
    int rc = gather_full(S, T, R.api, T->d_rexp);
    if (rc) return rc;
  }
  const XView rv = make_view(S, T, T->peer_r, T->d_rexp, R.mode);
  k_dist_node<<<nb, DIST_VEC_THREADS, 0, c->stream>>>(T->n_my, T->d_mn_ptr, T->d_mn_item, T->d_linv, rv, T->d_wloc);
  EFB_CHECK_LAUNCH(c);
  k_dist_edge<<<vb, DIST_VEC_THREADS, 0, c->stream>>>(S->m, T->d_dinv, T->d_rexp, T->d_edge_slot, S->d_dir, T->d_wloc, T->d_z, T->d_partial);
  EFB_CHECK_LAUNCH(c);
  k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, vb, T->d_sc, DS_RHO, 0);  // partial[2] is zero: DS_RR is restored below
  EFB_CHECK_LAUNCH(c);
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_RHO, T->d_sc + DS_RHO, 2, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  EFB_CUDA(c, cudaMemcpyAsync(T->d_sc + DS_RR, T->d_sc + DS_RR2, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return EFB_OK;
}

// r = b - A x (x read through the view of the exported solution copy), z = dinv r, all-reduce {rho, rr}
static int dist_residual(const DistRun &R) {
  System *S = R.S;
  DistState *T = R.T;
  Ctx *c = S->ctx;
  const int nb = spmv_grid(c, S->n_sp_chunks);
  k_dist_copy<<<vec_blocks(c, S->m), DIST_VEC_THREADS, 0, c->stream>>>(S->m, S->d_x, T->d_xs);
  EFB_CHECK_LAUNCH(c);
  // every rank's copy must be complete before anybody reads it: the {bb}/{rho,rr} all-reduce of the previous step
  // orders the kernels of one rank, an explicit tiny all-reduce orders the ranks
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_SYNC, T->d_sc + DS_SYNC, 1, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  if (R.mode == 1) {
    int rc = gather_full(S, T, R.api, T->d_xs);
    if (rc) return rc;
  }
  const XView xv = make_view(S, T, T->peer_x, T->d_xs, R.mode);
  if (R.aux) {
    // r = b - A x into the exported residual, |r|^2 all-reduced (every rank's r is complete after it), then z = M^-1 r
    { int rcl = launch_dist_spmv<0>(S, T, xv, T->d_rexp, 1); if (rcl) return rcl; }
    k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, nb, T->d_sc, DS_RHO, 0);
    EFB_CHECK_LAUNCH(c);
    EFB_CUDA(c, cudaMemcpyAsync(T->d_sc + DS_RR2, T->d_sc + DS_RR, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_RR2, T->d_sc + DS_RR2, 1, ncclFloat64, ncclSum, R.dd->comm, c->stream));
    return dist_aux_precond(R);
  }
  { int rcl = launch_dist_spmv<0>(S, T, xv, T->d_r, 1); if (rcl) return rcl; }
  k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, nb, T->d_sc, DS_RHO, 0);
  EFB_CHECK_LAUNCH(c);
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_RHO, T->d_sc + DS_RHO, 3, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  // z (parked in q by the kernel) -> exported z; peers finished reading the old z before the all-reduce returned
  k_dist_copy<<<vec_blocks(c, S->m), DIST_VEC_THREADS, 0, c->stream>>>(S->m, T->d_q, T->d_z);
  EFB_CHECK_LAUNCH(c);
  // ... and the new z must be complete everywhere before the first K1 reads it
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_SYNC, T->d_sc + DS_SYNC, 1, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  return EFB_OK;
}

// one COCG iteration (K1, F1, all-reduce, K2, F2, all-reduce)
static int dist_iteration(const DistRun &R, int first) {
  System *S = R.S;
  DistState *T = R.T;
  Ctx *c = S->ctx;
  const int nb = spmv_grid(c, S->n_sp_chunks), vb = vec_blocks(c, S->m);
  if (R.mode == 1) {
    int rc = gather_full(S, T, R.api, T->d_z);
    if (rc) return rc;
  }
  const XView zv = make_view(S, T, T->peer_z, T->d_z, R.mode);
  { int rcl = launch_dist_spmv<1>(S, T, zv, T->d_r, first); if (rcl) return rcl; }
  k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, nb, T->d_sc, DS_PQ, 1);  // partial[2] is unused by K1 (zero)
  EFB_CHECK_LAUNCH(c);
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_PQ, T->d_sc + DS_PQ, 2, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  if (R.aux) {
    k_dist_update_r<<<vb, DIST_VEC_THREADS, 0, c->stream>>>(S->m, T->d_sc, T->d_p, T->d_q, S->d_x, T->d_rexp, T->d_partial);
    EFB_CHECK_LAUNCH(c);
    k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, vb, T->d_sc, DS_RR2 - 0, 0);  // writes DS_RR2 .. DS_RR2+2 (pads)
    EFB_CHECK_LAUNCH(c);
    EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_RR2, T->d_sc + DS_RR2, 1, ncclFloat64, ncclSum, R.dd->comm, c->stream));
    return dist_aux_precond(R);
  }
  k_dist_update<<<vb, DIST_VEC_THREADS, 0, c->stream>>>(S->m, T->d_sc, T->d_dinv, T->d_p, T->d_q, S->d_x, T->d_r, T->d_z, T->d_partial);
  EFB_CHECK_LAUNCH(c);
  k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, vb, T->d_sc, DS_RHO, 0);
  EFB_CHECK_LAUNCH(c);
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_RHO, T->d_sc + DS_RHO, 3, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  return EFB_OK;
}

// nodal diagonal L = diag(G^T A G): local rows -> full nodal array -> summed over the ranks -> inverted on my nodes
static int dist_aux_setup(const DistRun &R) {
  System *S = R.S;
  DistState *T = R.T;
  Ctx *c = S->ctx;
  const int nn = S->mesh->n_node;
  EFB_CUDA(c, cudaMemsetAsync(T->d_Lfull, 0, (size_t)nn * sizeof(c128), c->stream));
  k_dist_nodal_diag<<<vec_blocks(c, S->m), DIST_VEC_THREADS, 0, c->stream>>>(S->m, S->row0, S->d_rowptr, S->d_colidx, S->d_vals, S->d_dir_all,
                                                                             T->d_en_all, T->d_Lfull);
  EFB_CHECK_LAUNCH(c);
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_Lfull, T->d_Lfull, (size_t)nn * 2, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  k_dist_linv<<<vec_blocks(c, std::max(T->n_my, 1)), DIST_VEC_THREADS, 0, c->stream>>>(T->n_my, T->d_my_node, T->d_my_dir, T->d_Lfull, T->d_linv);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

}  // namespace efb

using namespace efb;

extern "C" {

int efb_dist_unique_id(uint8_t *id128) {
  if (!id128) return EFB_ERR_INVALID;
  std::string err;
  NcclApi *api = nccl_api(err);
  if (!api) return fail(nullptr, EFB_ERR_STATE, "efb_dist_unique_id: %s", err.c_str());
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(nullptr, api, r, "ncclGetUniqueId");
  memcpy(id128, &id, 128);
  return EFB_OK;
}

int efb_dist_init(efb_ctx *ctx_, int32_t rank, int32_t world, const uint8_t *id128) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || !id128 || world < 1 || world > DIST_MAX_WORLD || rank < 0 || rank >= world)
    return fail(c, EFB_ERR_INVALID, "efb_dist_init: bad arguments (world <= %d)", DIST_MAX_WORLD);
  if (c->dist) return fail(c, EFB_ERR_STATE, "efb_dist_init: context already initialised");
  std::string err;
  NcclApi *api = nccl_api(err);
  if (!api) return fail(c, EFB_ERR_STATE, "efb_dist_init: %s", err.c_str());
  EFB_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  Dist *dd = new Dist();
  dd->rank = rank;
  dd->world = world;
  ncclResult_t r = api->CommInitRank(&dd->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete dd;
    return nccl_fail(c, api, r, "ncclCommInitRank");
  }
  c->dist = dd;
  return EFB_OK;
}

void efb_dist_finalize(efb_ctx *ctx_) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || !c->dist) return;
  Dist *dd = (Dist *)c->dist;
  std::string err;
  NcclApi *api = nccl_api(err);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (api && dd->comm) api->CommDestroy(dd->comm);
  delete dd;
  c->dist = nullptr;
}

int efb_dist_row_range(int32_t m, int32_t rank, int32_t world, int32_t *row_begin, int32_t *row_end) {
  if (m <= 0 || world < 1 || rank < 0 || rank >= world || !row_begin || !row_end) return EFB_ERR_INVALID;
  const int chunk = (m + world - 1) / world;
  *row_begin = std::min(m, rank * chunk);
  *row_end = std::min(m, (rank + 1) * chunk);
  return EFB_OK;
}

int efb_dist_solve(efb_system *sys_, const efb_solve_opts *opts, efb_solve_result *result, int32_t halo_mode) {
  System *S = (System *)sys_;
  if (!S || !opts || !result) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_dist_solve: NULL argument");
  Ctx *c = S->ctx;
  if (!S->assembled) return fail(c, EFB_ERR_STATE, "efb_dist_solve: matrix values were never assembled or set");
  if (!(opts->tolerance > 0.0) || opts->max_iterations < 0 || halo_mode < 0 || halo_mode > 1) return fail(c, EFB_ERR_INVALID, "efb_dist_solve: bad options");
  if (opts->method != EFB_METHOD_COCG && opts->method != EFB_METHOD_AUTO)
    return fail(c, EFB_ERR_INVALID, "efb_dist_solve: the row-partitioned solver is COCG (complex symmetric systems)");
  EFB_CUDA(c, cudaSetDevice(c->device));
  const bool aux = opts->precond == EFB_PRECOND_AUX;
  DistState *T = nullptr;
  int rc = dist_prepare(S, &T, aux);
  if (rc) return rc;
  std::string err;
  DistRun R{S, T, nccl_api(err), (Dist *)c->dist, halo_mode, aux};
  const int vb = vec_blocks(c, S->m);
  k_dist_dinv<<<vb, DIST_VEC_THREADS, 0, c->stream>>>(S->m, S->d_diag_pos, S->d_vals, T->d_dinv, opts->precond != EFB_PRECOND_NONE ? 1 : 0);
  EFB_CHECK_LAUNCH(c);
  if (aux && (rc = dist_aux_setup(R))) return rc;
  if (opts->zero_initial_guess) EFB_CUDA(c, cudaMemsetAsync(S->d_x, 0, (size_t)S->m * sizeof(c128), c->stream));
  // |b|^2
  k_dist_norm2<<<vb, DIST_VEC_THREADS, 0, c->stream>>>(S->m, S->d_b, T->d_partial);
  EFB_CHECK_LAUNCH(c);
  k_dist_finish<3><<<1, 256, 0, c->stream>>>(T->d_partial, vb, T->d_sc, DS_BB, 0);
  EFB_CHECK_LAUNCH(c);
  EFB_NCCL(c, R.api, R.api->AllReduce(T->d_sc + DS_BB, T->d_sc + DS_BB, 1, ncclFloat64, ncclSum, R.dd->comm, c->stream));
  const double tol2 = opts->tolerance * opts->tolerance;
  const int check_every = opts->check_every > 0 ? opts->check_every : 32;
  const int max_restarts = opts->max_restarts > 0 ? opts->max_restarts : 3;
  int iters = 0;
  double rr = 0.0, bb = 0.0;
  bool conv = false;
  double h[DS_NUM];
  auto read_scalars = [&]() -> int {
    EFB_CUDA(c, cudaMemcpyAsync(h, T->d_sc, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    rr = h[DS_RR];
    bb = h[DS_BB];
    return EFB_OK;
  };
  for (int cycle = 0; cycle <= max_restarts; ++cycle) {
    // true residual of the current iterate (every rank takes the same branch: the scalars are all-reduced)
    if ((rc = dist_residual(R))) return rc;
    if ((rc = read_scalars())) return rc;
    const double lim = bb > 0.0 ? tol2 * bb : tol2;
    conv = rr <= lim;
    if (conv || iters >= opts->max_iterations || !std::isfinite(rr)) break;
    bool first = true, stop = false;
    while (!stop) {
      const int n = std::min(check_every, opts->max_iterations - iters);
      for (int k = 0; k < n; ++k) {
        if ((rc = dist_iteration(R, first ? 1 : 0))) return rc;
        first = false;
      }
      iters += n;
      if ((rc = read_scalars())) return rc;
      stop = rr <= lim || iters >= opts->max_iterations || !std::isfinite(rr) || n == 0;
    }
  }
  result->iters = iters;
  result->method = EFB_METHOD_COCG;
  result->precond = aux ? EFB_PRECOND_AUX : (opts->precond == EFB_PRECOND_NONE ? EFB_PRECOND_NONE : EFB_PRECOND_JACOBI);
  result->residual = bb > 0.0 ? sqrt(rr / bb) : sqrt(rr);
  result->converged = conv ? 1 : 0;
  return EFB_OK;
}

// which 0: K1 alone (distributed SpMV + fused epilogue; halo_mode 1 includes the all-gather); 1: one full COCG iteration
int efb_dist_bench(efb_system *sys_, int32_t which, int32_t reps, int32_t halo_mode, double *avg_ms) {
  System *S = (System *)sys_;
  if (!S || !avg_ms || reps <= 0 || which < 0 || which > 2 || halo_mode < 0 || halo_mode > 1)
    return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_dist_bench: bad arguments");
  Ctx *c = S->ctx;
  if (!S->assembled) return fail(c, EFB_ERR_STATE, "efb_dist_bench: matrix values were never assembled or set");
  EFB_CUDA(c, cudaSetDevice(c->device));
  DistState *T = nullptr;
  int rc = dist_prepare(S, &T, which == 2);
  if (rc) return rc;
  std::string err;
  DistRun R{S, T, nccl_api(err), (Dist *)c->dist, halo_mode, which == 2};
  const int vb = vec_blocks(c, S->m), nb = spmv_grid(c, S->n_sp_chunks);
  k_dist_dinv<<<vb, DIST_VEC_THREADS, 0, c->stream>>>(S->m, S->d_diag_pos, S->d_vals, T->d_dinv, 1);
  EFB_CHECK_LAUNCH(c);
  if (which == 2 && (rc = dist_aux_setup(R))) return rc;
  if ((rc = dist_residual(R))) return rc;
  cudaEvent_t e0, e1;
  EFB_CUDA(c, cudaEventCreate(&e0));
  EFB_CUDA(c, cudaEventCreate(&e1));
  for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
    const int n = pass == 0 ? std::min(reps, 3) : reps;
    EFB_CUDA(c, cudaEventRecord(e0, c->stream));
    for (int k = 0; k < n; ++k) {
      if (which >= 1) {
        if ((rc = dist_iteration(R, 0))) return rc;
      } else {
        if (halo_mode == 1 && (rc = gather_full(S, T, R.api, T->d_z))) return rc;
        const XView zv = make_view(S, T, T->peer_z, T->d_z, halo_mode);
        if ((rc = launch_dist_spmv<1>(S, T, zv, T->d_r, 1))) return rc;
      }
    }
    EFB_CUDA(c, cudaEventRecord(e1, c->stream));
    EFB_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    EFB_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    *avg_ms = (double)ms / n;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return EFB_OK;
}

}  // extern "C"
