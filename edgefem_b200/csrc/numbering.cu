// global edge numbering on the device, bit-exact with the reference's sequential hash-map walk
// (src/mesh_gmsh.cpp:104-146): tets in order x local pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3), then tris x pairs
// (0,1)(1,2)(2,0); an edge gets the next id the first time its key (min<<32 | max of the node IDS) is seen;
// orient = +1 iff conn[a] < conn[b]; edges[id] = (min, max).
//
// "First seen" is a sort: position p = 6*tet + k (tris continue after the tets) is the visiting order, so
//   1. sort (key, p) by key with a STABLE radix sort       -> every run of equal keys starts with its first visit
//   2. run heads give (first_p, run); sort those by first_p -> the rank of a run in that order IS its edge id
//   3. scatter id back through p.
// Two cub::DeviceRadixSort passes and three small kernels; 121 M positions (the 20 M-tet cube) take milliseconds
// where the host's unordered_map walk takes seconds.  CUB is library code (ships with the CUDA toolkit).
#include <cub/cub.cuh>

#include "common.cuh"

namespace efb {
namespace {

__global__ void k_edge_keys(const long long *__restrict__ tet_conn, long long n_tet, const long long *__restrict__ tri_conn, long long n_tri,
                            unsigned long long *__restrict__ keys, unsigned *__restrict__ pos, int8_t *__restrict__ tet_orient,
                            int8_t *__restrict__ tri_orient) {
  const long long N = 6 * n_tet + 3 * n_tri;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= N) return;
  long long a, b;
  if (p < 6 * n_tet) {
    const long long t = p / 6;
    const int k = (int)(p - 6 * t);
    const int la = (k < 3) ? 0 : (k < 5 ? 1 : 2);
    const int lb = (k == 0) ? 1 : ((k == 1 || k == 3) ? 2 : 3);
    a = tet_conn[4 * t + la];
    b = tet_conn[4 * t + lb];
    tet_orient[p] = a < b ? 1 : -1;
  } else {
    const long long q = p - 6 * n_tet, t = q / 3;
    const int k = (int)(q - 3 * t);
    a = tri_conn[3 * t + k];
    b = tri_conn[3 * t + (k + 1) % 3];
    tri_orient[q] = a < b ? 1 : -1;
  }
  const unsigned long long lo = (unsigned long long)(a < b ? a : b), hi = (unsigned long long)(a < b ? b : a);
  keys[p] = (lo << 32) | hi;  // == (min << 32) ^ max for ids below 2^32 (make_edge_key, mesh.hpp:56-60)
  pos[p] = (unsigned)p;
}

// head[i] = 1 where a run of equal keys starts
__global__ void k_run_heads(const unsigned long long *__restrict__ keys, long long N, unsigned *__restrict__ head) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// run index (inclusive scan of head - 1) -> at run heads: first_p[run] = pos (the first visit), run_id[run] = run
__global__ void k_run_first(const unsigned *__restrict__ head, const unsigned *__restrict__ scan, const unsigned *__restrict__ pos, long long N,
                            unsigned *__restrict__ first_p, unsigned *__restrict__ run_id) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N || !head[i]) return;
  const unsigned run = scan[i] - 1;
  first_p[run] = pos[i];
  run_id[run] = run;
}

// sorted_run[id] = run with the id-th smallest first visit  ->  id_of_run[run] = id
__global__ void k_invert(const unsigned *__restrict__ sorted_run, long long m, unsigned *__restrict__ id_of_run) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < m) id_of_run[sorted_run[i]] = (unsigned)i;
}

__global__ void k_scatter_ids(const unsigned long long *__restrict__ keys, const unsigned *__restrict__ head, const unsigned *__restrict__ scan,
                              const unsigned *__restrict__ pos, const unsigned *__restrict__ id_of_run, long long N, long long n_tet6,
                              int32_t *__restrict__ tet_edges, int32_t *__restrict__ tri_edges, long long *__restrict__ edges) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned id = id_of_run[scan[i] - 1];
  const unsigned p = pos[i];
  if (p < n_tet6) tet_edges[p] = (int32_t)id;
  else tri_edges[p - n_tet6] = (int32_t)id;
  if (head[i]) {
    edges[2 * (long long)id] = (long long)(keys[i] >> 32);
    edges[2 * (long long)id + 1] = (long long)(keys[i] & 0xffffffffull);
  }
}

template <typename T>
struct Scratch {  // plain cudaMalloc scratch (GBs at the 20 M-tet size: not worth caching in the pool)
  T *p = nullptr;
  ~Scratch() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

}  // namespace
}  // namespace efb

using namespace efb;

extern "C" int efb_build_edges(efb_ctx *ctx_, int64_t n_tet, const int64_t *tet_conn, int64_t n_tri, const int64_t *tri_conn,
                               int32_t *tet_edges, int8_t *tet_orient, int32_t *tri_edges, int8_t *tri_orient, int64_t *n_edges,
                               int64_t *edges, int64_t edges_capacity) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || n_tet < 0 || n_tri < 0 || (n_tet > 0 && (!tet_conn || !tet_edges || !tet_orient)) ||
      (n_tri > 0 && (!tri_conn || !tri_edges || !tri_orient)) || !n_edges)
    return fail(c, EFB_ERR_INVALID, "efb_build_edges: bad arguments");
  const long long N = 6 * n_tet + 3 * n_tri;
  *n_edges = 0;
  if (N == 0) return EFB_OK;
  if (N >= (1ll << 31)) return fail(c, EFB_ERR_LIMIT, "efb_build_edges: more than 2^31 element edges");
  for (long long i = 0; i < 4 * n_tet; ++i)
    if (tet_conn[i] < 0 || tet_conn[i] >= (1ll << 32)) return fail(c, EFB_ERR_INVALID, "efb_build_edges: node id outside [0, 2^32) (the edge key packs two ids in 64 bits)");
  for (long long i = 0; i < 3 * n_tri; ++i)
    if (tri_conn[i] < 0 || tri_conn[i] >= (1ll << 32)) return fail(c, EFB_ERR_INVALID, "efb_build_edges: node id outside [0, 2^32)");
  EFB_CUDA(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  Scratch<long long> d_tet, d_tri, d_edges;
  Scratch<unsigned long long> k0, k1;
  Scratch<unsigned> v0, v1, head, scan, first_p, run_id, first_sorted, run_sorted, id_of_run;
  Scratch<int8_t> d_to, d_ro;
  Scratch<int32_t> d_te, d_re;
  Scratch<unsigned char> tmp;
  EFB_CUDA(c, d_tet.alloc(4 * n_tet));
  EFB_CUDA(c, d_tri.alloc(3 * n_tri));
  EFB_CUDA(c, k0.alloc(N)); EFB_CUDA(c, k1.alloc(N)); EFB_CUDA(c, v0.alloc(N)); EFB_CUDA(c, v1.alloc(N));
  EFB_CUDA(c, head.alloc(N)); EFB_CUDA(c, scan.alloc(N));
  EFB_CUDA(c, d_to.alloc(6 * n_tet)); EFB_CUDA(c, d_ro.alloc(3 * n_tri)); EFB_CUDA(c, d_te.alloc(6 * n_tet)); EFB_CUDA(c, d_re.alloc(3 * n_tri));
  if (n_tet) EFB_CUDA(c, cudaMemcpyAsync(d_tet.p, tet_conn, 4 * n_tet * sizeof(long long), cudaMemcpyHostToDevice, st));
  if (n_tri) EFB_CUDA(c, cudaMemcpyAsync(d_tri.p, tri_conn, 3 * n_tri * sizeof(long long), cudaMemcpyHostToDevice, st));
  Timed tm(c);
  const unsigned nb = (unsigned)((N + 255) / 256);
  k_edge_keys<<<nb, 256, 0, st>>>(d_tet.p, n_tet, d_tri.p, n_tri, k0.p, v0.p, d_to.p, d_ro.p);
  EFB_CHECK_LAUNCH(c);
  // 1. stable sort by key (ids < 2^32 on both halves: all 64 bits matter only up to the top set bit; sort them all)
  cub::DoubleBuffer<unsigned long long> kb(k0.p, k1.p);
  cub::DoubleBuffer<unsigned> vb(v0.p, v1.p);
  size_t tb = 0, tb2 = 0, tb3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (int)N, 0, 64, st);
  cub::DeviceScan::InclusiveSum(nullptr, tb2, head.p, scan.p, (int)N, st);
  {
    cub::DoubleBuffer<unsigned> a(nullptr, nullptr), b(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, tb3, a, b, (int)N, 0, 32, st);
  }
  EFB_CUDA(c, tmp.alloc(std::max(tb, std::max(tb2, tb3))));
  EFB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, kb, vb, (int)N, 0, 64, st));
  c->launches++;
  const unsigned long long *keys = kb.Current();
  const unsigned *pos = vb.Current();
  // 2. run heads, run indices, number of distinct edges
  k_run_heads<<<nb, 256, 0, st>>>(keys, N, head.p);
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cub::DeviceScan::InclusiveSum(tmp.p, tb2, head.p, scan.p, (int)N, st));
  c->launches++;
  unsigned m_u = 0;
  EFB_CUDA(c, cudaMemcpyAsync(&m_u, scan.p + (N - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  const long long m = m_u;
  *n_edges = m;
  if (m >= (1ll << 31)) return fail(c, EFB_ERR_LIMIT, "efb_build_edges: more than 2^31 edges");
  if (edges && edges_capacity < m) return fail(c, EFB_ERR_INVALID, "efb_build_edges: edges buffer holds %lld edges, need %lld", (long long)edges_capacity, m);
  EFB_CUDA(c, first_p.alloc(m)); EFB_CUDA(c, run_id.alloc(m)); EFB_CUDA(c, first_sorted.alloc(m)); EFB_CUDA(c, run_sorted.alloc(m));
  EFB_CUDA(c, id_of_run.alloc(m)); EFB_CUDA(c, d_edges.alloc(2 * m));
  k_run_first<<<nb, 256, 0, st>>>(head.p, scan.p, pos, N, first_p.p, run_id.p);
  EFB_CHECK_LAUNCH(c);
  // 3. order the runs by their first visit: the rank is the edge id
  cub::DoubleBuffer<unsigned> fb(first_p.p, first_sorted.p), rb(run_id.p, run_sorted.p);
  EFB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb3, fb, rb, (int)m, 0, 32, st));
  c->launches++;
  const unsigned mb = (unsigned)((m + 255) / 256);
  k_invert<<<mb, 256, 0, st>>>(rb.Current(), m, id_of_run.p);
  EFB_CHECK_LAUNCH(c);
  k_scatter_ids<<<nb, 256, 0, st>>>(keys, head.p, scan.p, pos, id_of_run.p, N, 6 * n_tet, d_te.p, d_re.p, d_edges.p);
  EFB_CHECK_LAUNCH(c);
  if (n_tet) {
    EFB_CUDA(c, cudaMemcpyAsync(tet_edges, d_te.p, 6 * n_tet * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    EFB_CUDA(c, cudaMemcpyAsync(tet_orient, d_to.p, 6 * n_tet, cudaMemcpyDeviceToHost, st));
  }
  if (n_tri) {
    EFB_CUDA(c, cudaMemcpyAsync(tri_edges, d_re.p, 3 * n_tri * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    EFB_CUDA(c, cudaMemcpyAsync(tri_orient, d_ro.p, 3 * n_tri, cudaMemcpyDeviceToHost, st));
  }
  if (edges) EFB_CUDA(c, cudaMemcpyAsync(edges, d_edges.p, 2 * m * sizeof(long long), cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  return EFB_OK;
}

// =====================================================================================================
// Large-mesh set-up on the device: edge -> (tet, local) incidence lists and the CSR pattern + position map
// of a row block.  Same results as the host code in abi.cu (tested bit for bit); used from efb_mesh_create /
// system_create_rows when the mesh is large (device_setup_enabled()).
// =====================================================================================================
namespace efb {
namespace {

__global__ void k_e2t_pairs(const int32_t *__restrict__ tet_edges, long long n6, unsigned *__restrict__ key, unsigned *__restrict__ item) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n6) return;
  key[p] = (unsigned)tet_edges[p];
  item[p] = (unsigned)(((p / 6) << 3) | (p % 6));
}

__global__ void k_count_keys(const unsigned *__restrict__ key, long long n6, int32_t *__restrict__ cnt) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p < n6) atomicAdd(&cnt[key[p]], 1);  // integer counts: order-independent
}

constexpr int PAT_MAX_ROW = 256;  // longest row the device pattern builder handles (longer: host path)

// sorted, duplicate-free column list of one row in thread-local storage; returns its length or -1 on overflow
__device__ __forceinline__ int row_columns(const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ e2t_item,
                                           const int32_t *__restrict__ tet_edges, int g, int32_t (&buf)[PAT_MAX_ROW]) {
  int L = 0;
  for (int k = e2t_ptr[g]; k < e2t_ptr[g + 1]; ++k) {
    const int32_t *e6 = tet_edges + 6ll * (e2t_item[k] >> 3);
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
      const int32_t cnew = e6[j];
      int lo = 0, hi = L;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (buf[mid] < cnew) lo = mid + 1; else hi = mid;
      }
      if (lo < L && buf[lo] == cnew) continue;
      if (L == PAT_MAX_ROW) return -1;
      for (int i = L; i > lo; --i) buf[i] = buf[i - 1];
      buf[lo] = cnew;
      ++L;
    }
  }
  return L;
}

__global__ void __launch_bounds__(128)
k_pattern_count(const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ e2t_item, const int32_t *__restrict__ tet_edges, int row0,
                int m_loc, int32_t *__restrict__ rowlen, int32_t *__restrict__ overflow) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_loc) return;
  int32_t buf[PAT_MAX_ROW];
  const int L = row_columns(e2t_ptr, e2t_item, tet_edges, row0 + r, buf);
  if (L < 0) {
    atomicExch(overflow, 1);
    rowlen[r] = 0;
  } else {
    rowlen[r] = L;
  }
}

__global__ void __launch_bounds__(128)
k_pattern_fill(const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ e2t_item, const int32_t *__restrict__ tet_edges, int row0,
               int m_loc, const int32_t *__restrict__ rowptr, int32_t *__restrict__ colidx, uint16_t *__restrict__ pos, long long kpos0) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m_loc) return;
  int32_t buf[PAT_MAX_ROW];
  const int g = row0 + r;
  const int L = row_columns(e2t_ptr, e2t_item, tet_edges, g, buf);
  int32_t *dst = colidx + rowptr[r];
  for (int i = 0; i < L; ++i) dst[i] = buf[i];
  for (int k = e2t_ptr[g]; k < e2t_ptr[g + 1]; ++k) {
    const int32_t *e6 = tet_edges + 6ll * (e2t_item[k] >> 3);
    for (int j = 0; j < 6; ++j) {
      int lo = 0, hi = L;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (buf[mid] < e6[j]) lo = mid + 1; else hi = mid;
      }
      pos[((long long)k - kpos0) * 6 + j] = (uint16_t)lo;
    }
  }
}

}  // namespace

bool device_setup_enabled(long long n_tet) {
  if (const char *e = getenv("EDGEFEM_B200_DEVICE_SETUP")) return e[0] == '1';
  return n_tet >= 200000;
}

// fills M->d_tet_edges, d_e2t_ptr, d_e2t_item and M->h_e2t_ptr (h_e2t_item stays empty: mesh_host_e2t_item fetches it)
int device_e2t(Mesh *M, const int32_t *h_tet_edges, long long nt) {
  Ctx *c = M->ctx;
  cudaStream_t st = c->stream;
  const long long n6 = 6 * nt;
  int rc;
  if ((rc = dev_upload(c, &M->d_tet_edges, h_tet_edges, (size_t)n6))) return rc;
  if ((rc = dev_alloc(c, &M->d_e2t_ptr, (size_t)M->m + 1))) return rc;
  if ((rc = dev_alloc(c, &M->d_e2t_item, (size_t)std::max<long long>(n6, 1)))) return rc;
  Scratch<unsigned> k0, k1, v0;
  Scratch<int32_t> cnt;
  Scratch<unsigned char> tmp;
  EFB_CUDA(c, k0.alloc(n6)); EFB_CUDA(c, k1.alloc(n6)); EFB_CUDA(c, v0.alloc(n6)); EFB_CUDA(c, cnt.alloc((size_t)M->m + 1));
  const unsigned nb = (unsigned)((n6 + 255) / 256);
  k_e2t_pairs<<<nb, 256, 0, st>>>(M->d_tet_edges, n6, k0.p, v0.p);
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cudaMemsetAsync(cnt.p, 0, ((size_t)M->m + 1) * sizeof(int32_t), st));
  k_count_keys<<<nb, 256, 0, st>>>(k0.p, n6, cnt.p);
  EFB_CHECK_LAUNCH(c);
  int bits = 1;
  while ((1ll << bits) < (long long)M->m) ++bits;
  size_t tb = 0, tb2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, v0.p, (unsigned *)M->d_e2t_item, (int)n6, 0, bits, st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb2, cnt.p, M->d_e2t_ptr, M->m + 1, st);
  EFB_CUDA(c, tmp.alloc(std::max(tb, tb2)));
  // stable sort by edge id: items of an edge stay in ascending (tet, local) order = the host's counting sort
  EFB_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tb, k0.p, k1.p, v0.p, (unsigned *)M->d_e2t_item, (int)n6, 0, bits, st));
  EFB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb2, cnt.p, M->d_e2t_ptr, M->m + 1, st));
  c->launches += 2;
  M->h_e2t_ptr.resize((size_t)M->m + 1);
  EFB_CUDA(c, cudaMemcpyAsync(M->h_e2t_ptr.data(), M->d_e2t_ptr, ((size_t)M->m + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  return EFB_OK;
}

int mesh_host_e2t_item(Mesh *M) {  // host copy of the incidence items for the host-side pattern builder
  if (!M->h_e2t_item.empty() || M->n_tet == 0) return EFB_OK;
  Ctx *c = M->ctx;
  M->h_e2t_item.resize((size_t)6 * M->n_tet);
  EFB_CUDA(c, cudaMemcpyAsync(M->h_e2t_item.data(), M->d_e2t_item, M->h_e2t_item.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

// CSR pattern + position map of the rows [S->row0, S->row0 + S->m) on the device; fills d_rowptr, d_colidx, d_e2t_pos,
// h_rowptr, h_colidx, nnz.  *done = false (nothing allocated) when a row is too long for the device builder.
int device_pattern(System *S, bool *done) {
  Ctx *c = S->ctx;
  Mesh *M = S->mesh;
  cudaStream_t st = c->stream;
  *done = false;
  const int m = S->m;
  Scratch<int32_t> rowlen, ovf;
  Scratch<unsigned char> tmp;
  EFB_CUDA(c, rowlen.alloc((size_t)m + 1)); EFB_CUDA(c, ovf.alloc(1));
  EFB_CUDA(c, cudaMemsetAsync(ovf.p, 0, sizeof(int32_t), st));
  EFB_CUDA(c, cudaMemsetAsync(rowlen.p + m, 0, sizeof(int32_t), st));
  const unsigned nb = (unsigned)((m + 127) / 128);
  k_pattern_count<<<nb, 128, 0, st>>>(M->d_e2t_ptr, M->d_e2t_item, M->d_tet_edges, S->row0, m, rowlen.p, ovf.p);
  EFB_CHECK_LAUNCH(c);
  int32_t h_ovf = 0;
  EFB_CUDA(c, cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  if (h_ovf) return EFB_OK;
  int rc;
  if ((rc = dev_alloc(c, &S->d_rowptr, (size_t)m + 1))) return rc;
  size_t tb = 0;
  // 64-bit safe: row lengths sum below 2^31 is checked right after
  cub::DeviceScan::ExclusiveSum(nullptr, tb, rowlen.p, S->d_rowptr, m + 1, st);
  EFB_CUDA(c, tmp.alloc(tb));
  EFB_CUDA(c, cub::DeviceScan::ExclusiveSum(tmp.p, tb, rowlen.p, S->d_rowptr, m + 1, st));
  c->launches++;
  S->h_rowptr.resize((size_t)m + 1);
  EFB_CUDA(c, cudaMemcpyAsync(S->h_rowptr.data(), S->d_rowptr, ((size_t)m + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  // overflow of the int32 scan shows up as a non-monotone / negative total
  const long long nnz = S->h_rowptr[m];
  if (nnz < 0) return fail(c, EFB_ERR_LIMIT, "efb_system_create: nnz >= 2^31 (int32 CSR like Eigen's default index)");
  S->nnz = nnz;
  const long long kpos0 = M->h_e2t_ptr[S->row0], n_inc = M->h_e2t_ptr[S->row0 + m] - kpos0;
  if ((rc = dev_alloc(c, &S->d_colidx, (size_t)std::max<long long>(nnz, 1)))) return rc;
  if ((rc = dev_alloc(c, &S->d_e2t_pos, (size_t)std::max<long long>(n_inc * 6, 1)))) return rc;
  k_pattern_fill<<<nb, 128, 0, st>>>(M->d_e2t_ptr, M->d_e2t_item, M->d_tet_edges, S->row0, m, S->d_rowptr, S->d_colidx, S->d_e2t_pos, kpos0);
  EFB_CHECK_LAUNCH(c);
  S->h_colidx.resize((size_t)nnz);
  EFB_CUDA(c, cudaMemcpyAsync(S->h_colidx.data(), S->d_colidx, (size_t)nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EFB_CUDA(c, cudaStreamSynchronize(st));
  *done = true;
  return EFB_OK;
}

}  // namespace efb
