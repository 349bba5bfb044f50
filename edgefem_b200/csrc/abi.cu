// Context, mesh and system handles: upload, CSR pattern construction, gather maps.
// Pattern rules follow the reference bit-exactly:
// union of the 6x6 edge cliques of all tets (Eigen setFromTriplets,
// src/assemble_maxwell.cpp:205) plus caller-listed extra entries (coeffRef insertions),
// row-major, columns sorted ascending.
#include <stdarg.h>

#include <algorithm>
#include <functional>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "common.cuh"

namespace efb {

static std::mutex g_err_mu;
static std::string g_err;

void set_global_error(const std::string &s) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = s;
}

int fail(Ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  set_global_error(buf);
  return code;
}

// ---------------------------------------------------------------- caching device allocator
namespace {
struct PoolBlock {
  void *p;
  size_t bytes;
};
std::mutex g_pool_mu;
std::unordered_map<void *, std::pair<int, size_t>> g_live;       // every dev_alloc'ed pointer -> (device, bytes)
std::unordered_map<int, std::vector<PoolBlock>> g_free;          // cached blocks per device
std::unordered_map<int, size_t> g_free_bytes;
constexpr size_t POOL_MIN_BLOCK = 1 << 20;                       // smaller requests are rounded up to a power of two
constexpr size_t POOL_MAX_BLOCKS = 4096;                         // cap of cached blocks per device
constexpr size_t POOL_MAX_BYTES = 16ull << 30;                   // cap of cached bytes per device
}  // namespace

size_t pool_round(size_t bytes) {
  if (bytes >= POOL_MIN_BLOCK) return bytes;
  size_t r = 512;
  while (r < bytes) r <<= 1;
  return r;
}

void *pool_get(int device, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  auto &fl = g_free[device];
  // small requests were rounded up to a power of two by pool_round(); large ones accept up to 25% + 1 MiB of slack
  const size_t hi = bytes < POOL_MIN_BLOCK ? bytes : bytes + bytes / 4 + (1 << 20);
  int best = -1;
  for (int i = 0; i < (int)fl.size(); ++i)
    if (fl[i].bytes >= bytes && fl[i].bytes <= hi && (best < 0 || fl[i].bytes < fl[best].bytes)) best = i;
  if (best < 0) return nullptr;
  void *p = fl[best].p;
  g_free_bytes[device] -= fl[best].bytes;
  g_live[p] = {device, fl[best].bytes};
  fl.erase(fl.begin() + best);
  return p;
}

void pool_register(void *p, int device, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_live[p] = {device, bytes};
}

void dfree(void *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_live.find(p);
    if (it != g_live.end()) {
      const int device = it->second.first;
      const size_t bytes = it->second.second;
      g_live.erase(it);
      if (g_free_bytes[device] + bytes <= POOL_MAX_BYTES && g_free[device].size() < POOL_MAX_BLOCKS) {
        g_free[device].push_back({p, bytes});
        g_free_bytes[device] += bytes;
        return;
      }
    }
  }
  cudaFree(p);
}

void pool_trim() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto &kv : g_free) {
    for (auto &b : kv.second) cudaFree(b.p);
    kv.second.clear();
    g_free_bytes[kv.first] = 0;
  }
}

// EDGEFEM_B200_TRACE=2: host-side stage times of system creation on stderr
template <typename F>
static void parallel_for(int64_t n, F f) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<int64_t>(std::max(1u, std::min(hw, 32u)), std::max<int64_t>(1, n / 512));
  if (nt <= 1) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  int64_t step = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    int64_t a = t * step, b = std::min<int64_t>(n, a + step);
    if (a >= b) break;
    th.emplace_back([=] { f(a, b); });
  }
  for (auto &x : th) x.join();
}

// ---------------------------------------------------------------- kernels used here
__global__ void k_slot_bbox(const int4 *__restrict__ tet_nodes, const uint8_t *__restrict__ tet_slot,
                            const double4 *__restrict__ xyz, int n_tet, unsigned long long *bbox_enc) {
  // order-preserving encoding of doubles into uint64 so atomicMin/Max work
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tet) return;
  int4 nd = tet_nodes[t];
  int s = tet_slot[t];
  int ids[4] = {nd.x, nd.y, nd.z, nd.w};
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double4 p = xyz[ids[i]];
    mn[0] = fmin(mn[0], p.x); mx[0] = fmax(mx[0], p.x);
    mn[1] = fmin(mn[1], p.y); mx[1] = fmax(mx[1], p.y);
    mn[2] = fmin(mn[2], p.z); mx[2] = fmax(mx[2], p.z);
  }
  auto enc = [](double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
  };
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    atomicMin(&bbox_enc[s * 6 + a], enc(mn[a]));
    atomicMax(&bbox_enc[s * 6 + 3 + a], enc(mx[a]));
  }
}

__global__ void k_bbox_decode(unsigned long long *enc, double *out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long u = enc[i];
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
  out[i] = __longlong_as_double((long long)u);
}

// bake the Dirichlet COLUMN mask into the assembly map: bit 15 of an incidence's column offset says
// "this column is a Dirichlet edge" (the entry then stays an explicit zero)
__global__ void k_pos_dirichlet(const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                const uint8_t *__restrict__ dir, int m, uint16_t *pos, long long pos_off) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  const int rb = rowptr[r];
  for (int k = e2t_ptr[r]; k < e2t_ptr[r + 1]; ++k)
    for (int j = 0; j < 6; ++j) {
      const uint16_t p = pos[(size_t)k * 6 + j - pos_off] & 0x7fffu;
      pos[(size_t)k * 6 + j - pos_off] = p | (dir[colidx[rb + p]] ? 0x8000u : 0u);
    }
}

__global__ void k_n2e_dirichlet(int32_t *n2e_item, long long n, const uint8_t *__restrict__ dir) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int it = n2e_item[i] & ~2;
  n2e_item[i] = it | (dir[it >> 2] ? 2 : 0);
}

__global__ void k_node_dir(const int2 *__restrict__ edge_nodes, const uint8_t *__restrict__ dir, int m,
                           uint8_t *node_dir) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  if (dir[e]) {
    int2 n = edge_nodes[e];
    node_dir[n.x] = 1;
    node_dir[n.y] = 1;
  }
}

}  // namespace efb

using namespace efb;

extern "C" {

int efb_abi_version(void) { return EFB_ABI_VERSION; }

int efb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *efb_last_error(const efb_ctx *ctx) {
  if (ctx) return ((const Ctx *)ctx)->err.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_err;
  return copy.c_str();
}

int efb_ctx_create(int device, efb_ctx **out) {
  if (!out) return fail(nullptr, EFB_ERR_INVALID, "efb_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, EFB_ERR_CUDA,
                "efb_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n) return fail(nullptr, EFB_ERR_INVALID, "efb_ctx_create: device %d out of range [0,%d)", device, n);
  Ctx *c = new Ctx();
  c->device = device;
  EFB_CUDA(c, cudaSetDevice(device));
  EFB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  EFB_CUDA(c, cudaEventCreate(&c->ev0));
  EFB_CUDA(c, cudaEventCreate(&c->ev1));
  EFB_CUDA(c, cudaEventCreate(&c->tm0));
  EFB_CUDA(c, cudaEventCreate(&c->tm1));
  cudaDeviceProp prop;
  EFB_CUDA(c, cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  *out = (efb_ctx *)c;
  return EFB_OK;
}

void efb_ctx_destroy(efb_ctx *ctx_) {
  Ctx *c = (Ctx *)ctx_;
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  efb_dist_finalize(ctx_);
  pool_trim();
  if (c->d_flush) cudaFree(c->d_flush);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->tm0) cudaEventDestroy(c->tm0);
  if (c->tm1) cudaEventDestroy(c->tm1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// Measurement helper: overwrite a 256 MB scratch buffer on the context's stream, i.e. evict the 126 MB L2 (used by bench.py
// between timed steps whose working set would otherwise stay L2-resident from one step to the next).
int efb_l2_flush(efb_ctx *ctx_) {
  Ctx *c = (Ctx *)ctx_;
  if (!c) return EFB_ERR_INVALID;
  constexpr size_t FLUSH_BYTES = (size_t)256 << 20;
  EFB_CUDA(c, cudaSetDevice(c->device));
  if (!c->d_flush) EFB_CUDA(c, cudaMalloc(&c->d_flush, FLUSH_BYTES));
  EFB_CUDA(c, cudaMemsetAsync(c->d_flush, 0x5a, FLUSH_BYTES, c->stream));
  return EFB_OK;
}

int efb_ctx_sync(efb_ctx *ctx_) {
  Ctx *c = (Ctx *)ctx_;
  if (!c) return EFB_ERR_INVALID;
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

double efb_last_kernel_ms(const efb_ctx *ctx_) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || !c->ev_valid) return -1.0;
  if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
  return (double)ms;
}

int64_t efb_launch_count(const efb_ctx *ctx_) { return ctx_ ? ((const Ctx *)ctx_)->launches : 0; }

int efb_timer_start(efb_ctx *ctx_) {
  Ctx *c = (Ctx *)ctx_;
  if (!c) return EFB_ERR_INVALID;
  EFB_CUDA(c, cudaSetDevice(c->device));
  EFB_CUDA(c, cudaEventRecord(c->tm0, c->stream));
  return EFB_OK;
}

int efb_timer_stop(efb_ctx *ctx_, double *ms) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || !ms) return EFB_ERR_INVALID;
  EFB_CUDA(c, cudaSetDevice(c->device));
  EFB_CUDA(c, cudaEventRecord(c->tm1, c->stream));
  EFB_CUDA(c, cudaEventSynchronize(c->tm1));
  float f = 0.f;
  EFB_CUDA(c, cudaEventElapsedTime(&f, c->tm0, c->tm1));
  *ms = (double)f;
  return EFB_OK;
}

// ------------------------------------------------------------------ mesh
int efb_mesh_create(efb_ctx *ctx_, const efb_mesh_desc *d, efb_mesh **out) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || !d || !out) return fail(c, EFB_ERR_INVALID, "efb_mesh_create: NULL argument");
  *out = nullptr;
  if (d->n_node <= 0 || d->n_tet < 0 || d->n_edge <= 0 || !d->xyz || !d->edge_nodes ||
      (d->n_tet > 0 && (!d->tet_nodes || !d->tet_edges || !d->tet_orient || !d->tet_phys)))
    return fail(c, EFB_ERR_INVALID, "efb_mesh_create: bad descriptor");
  if ((int64_t)d->n_tet >= (1ll << 28)) return fail(c, EFB_ERR_LIMIT, "efb_mesh_create: n_tet >= 2^28");
  EFB_CUDA(c, cudaSetDevice(c->device));
  Mesh *M = new Mesh();
  M->ctx = c;
  M->n_node = d->n_node;
  M->n_tet = d->n_tet;
  M->m = d->n_edge;
  const int64_t nt = d->n_tet;
  // validate indices
  for (int64_t i = 0; i < 4 * nt; ++i)
    if (d->tet_nodes[i] < 0 || d->tet_nodes[i] >= d->n_node) {
      delete M;
      return fail(c, EFB_ERR_INVALID, "efb_mesh_create: tet_nodes[%lld] out of range", (long long)i);
    }
  for (int64_t i = 0; i < 6 * nt; ++i)
    if (d->tet_edges[i] < 0 || d->tet_edges[i] >= d->n_edge) {
      delete M;
      return fail(c, EFB_ERR_INVALID, "efb_mesh_create: tet_edges[%lld] out of range", (long long)i);
    }
  // slots = distinct tags ascending
  {
    std::vector<int32_t> tags;  // a handful of distinct values: one pass with a last-tag cache, not a 20 M-element sort
    int32_t last = 0;
    bool have_last = false;
    for (int64_t t = 0; t < nt && tags.size() <= (size_t)MAX_SLOTS; ++t) {
      const int32_t g = d->tet_phys[t];
      if (have_last && g == last) continue;
      last = g;
      have_last = true;
      auto it = std::lower_bound(tags.begin(), tags.end(), g);
      if (it == tags.end() || *it != g) tags.insert(it, g);
    }
    if ((int)tags.size() > MAX_SLOTS) {
      delete M;
      return fail(c, EFB_ERR_LIMIT, "efb_mesh_create: %zu distinct physical tags > %d", tags.size(), MAX_SLOTS);
    }
    M->slot_tags = tags;
    M->n_slots = (int)tags.size();
  }
  std::vector<uint8_t> slot(nt), sign(nt);
  parallel_for(nt, [&](int64_t ta, int64_t tb) {
    for (int64_t t = ta; t < tb; ++t) {
      slot[t] = (uint8_t)(std::lower_bound(M->slot_tags.begin(), M->slot_tags.end(), d->tet_phys[t]) - M->slot_tags.begin());
      uint8_t s = 0;
      for (int k = 0; k < 6; ++k)
        if (d->tet_orient[6 * t + k] < 0) s |= (uint8_t)(1u << k);
      sign[t] = s;
    }
  });
  M->h_tet_edges.assign(d->tet_edges, d->tet_edges + 6 * nt);
  M->h_edge_nodes.assign(d->edge_nodes, d->edge_nodes + 2 * (int64_t)d->n_edge);
  // edge -> incident (tet, local) lists, ascending tet (deterministic summation order): counting sort on the host,
  // or a stable radix sort on the device for large meshes (same lists)
  const bool dev_setup = device_setup_enabled(nt) && nt > 0;
  if (!dev_setup) {
    M->h_e2t_ptr.assign((size_t)M->m + 1, 0);
    for (int64_t i = 0; i < 6 * nt; ++i) M->h_e2t_ptr[d->tet_edges[i] + 1]++;
    for (int e = 0; e < M->m; ++e) M->h_e2t_ptr[e + 1] += M->h_e2t_ptr[e];
    M->h_e2t_item.resize((size_t)6 * nt);
    std::vector<int32_t> cur(M->h_e2t_ptr.begin(), M->h_e2t_ptr.end() - 1);
    for (int64_t t = 0; t < nt; ++t)
      for (int k = 0; k < 6; ++k) M->h_e2t_item[cur[d->tet_edges[6 * t + k]]++] = (int32_t)((t << 3) | k);
  }
  std::vector<double4> xyz4(d->n_node);
  for (int i = 0; i < d->n_node; ++i) xyz4[i] = make_double4(d->xyz[3 * i], d->xyz[3 * i + 1], d->xyz[3 * i + 2], 0.0);
  int rc;
  if ((rc = dev_upload(c, &M->d_xyz, xyz4.data(), xyz4.size()))) return rc;
  if ((rc = dev_upload(c, &M->d_tet_nodes, (const int4 *)d->tet_nodes, (size_t)nt))) return rc;
  if ((rc = dev_upload(c, &M->d_tet_sign, sign.data(), sign.size()))) return rc;
  if ((rc = dev_upload(c, &M->d_tet_slot, slot.data(), slot.size()))) return rc;
  if (dev_setup) {
    if ((rc = device_e2t(M, d->tet_edges, nt))) return rc;
  } else {
    if ((rc = dev_upload(c, &M->d_e2t_ptr, M->h_e2t_ptr.data(), M->h_e2t_ptr.size()))) return rc;
    if ((rc = dev_upload(c, &M->d_e2t_item, M->h_e2t_item.data(), M->h_e2t_item.size()))) return rc;
  }
  if (asm_use_row_kernel()) {
    if ((rc = dev_alloc(c, &M->d_geom, (size_t)std::max<int64_t>(1, nt)))) return rc;
  }
  if ((rc = dev_alloc(c, &M->d_rec, (size_t)std::max<int64_t>(1, nt)))) return rc;
  if ((rc = dev_alloc(c, &M->d_e2t_ss, (size_t)std::max<int64_t>(1, 6 * nt)))) return rc;
  if ((rc = launch_tet_geometry(M))) return rc;
  // per-slot bounding boxes (PML profile, src/assemble_maxwell.cpp:66-89) on device
  if ((rc = dev_alloc(c, &M->d_slot_bbox, (size_t)std::max(1, M->n_slots) * 6))) return rc;
  if (nt > 0) {
    unsigned long long *enc = nullptr;
    if ((rc = dev_alloc(c, &enc, (size_t)M->n_slots * 6))) return rc;
    std::vector<unsigned long long> init((size_t)M->n_slots * 6);
    for (int s = 0; s < M->n_slots; ++s)
      for (int a = 0; a < 6; ++a) init[s * 6 + a] = a < 3 ? ~0ull : 0ull;
    EFB_CUDA(c, cudaMemcpyAsync(enc, init.data(), init.size() * 8, cudaMemcpyHostToDevice, c->stream));
    k_slot_bbox<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(M->d_tet_nodes, M->d_tet_slot, M->d_xyz, (int)nt, enc);
    EFB_CHECK_LAUNCH(c);
    k_bbox_decode<<<(M->n_slots * 6 + 63) / 64, 64, 0, c->stream>>>(enc, M->d_slot_bbox, M->n_slots * 6);
    EFB_CHECK_LAUNCH(c);
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    dfree(enc);
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = (efb_mesh *)M;
  return EFB_OK;
}

void efb_mesh_destroy(efb_mesh *mesh_) {
  Mesh *M = (Mesh *)mesh_;
  if (!M) return;
  cudaSetDevice(M->ctx->device);
  cudaStreamSynchronize(M->ctx->stream);  // pooled blocks may be handed out again immediately
  dfree(M->d_xyz); dfree(M->d_tet_nodes); dfree(M->d_tet_sign); dfree(M->d_tet_slot);
  dfree(M->d_e2t_ptr); dfree(M->d_e2t_item); dfree(M->d_slot_bbox); dfree(M->d_geom); dfree(M->d_rec); dfree(M->d_e2t_ss); dfree(M->d_tet_edges);
  delete M;
}

int efb_mesh_num_slots(const efb_mesh *mesh_) { return mesh_ ? ((const Mesh *)mesh_)->n_slots : 0; }

int efb_mesh_get_slot_tags(const efb_mesh *mesh_, int32_t *tags) {
  const Mesh *M = (const Mesh *)mesh_;
  if (!M || !tags) return EFB_ERR_INVALID;
  for (int i = 0; i < M->n_slots; ++i) tags[i] = M->slot_tags[i];
  return EFB_OK;
}

// ------------------------------------------------------------------ system
static int system_alloc_common(System *S) {
  Ctx *c = S->ctx;
  int rc;
  S->n_sys = S->n_matrix * S->n_rhs;
  if (!S->d_rowptr) {  // (the device pattern builder leaves both arrays in place)
    if ((rc = dev_upload(c, &S->d_rowptr, S->h_rowptr.data(), S->h_rowptr.size()))) return rc;
    if ((rc = dev_upload(c, &S->d_colidx, S->h_colidx.data(), S->h_colidx.size()))) return rc;
  }
  std::vector<int32_t> diag(S->m, -1);
  parallel_for(S->m, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; ++r) {
      const int32_t *beg = S->h_colidx.data() + S->h_rowptr[r], *end = S->h_colidx.data() + S->h_rowptr[r + 1];
      const int32_t g = (int32_t)(r + S->row0);  // columns are global ids
      const int32_t *it = std::lower_bound(beg, end, g);
      if (it != end && *it == g) diag[r] = (int32_t)(it - S->h_colidx.data());
    }
  });
  if ((rc = dev_upload(c, &S->d_diag_pos, diag.data(), diag.size()))) return rc;
  {  // CSR-stream chunks: consecutive whole rows with <= SPMV_STREAM_W entries per warp
    std::vector<int32_t> ch{0};
    bool ok = true;
    int acc = 0;
    for (int r = 0; r < S->m; ++r) {
      const int len = S->h_rowptr[r + 1] - S->h_rowptr[r];
      if (len > SPMV_STREAM_W) { ok = false; break; }
      if (acc + len > SPMV_STREAM_W || r - ch.back() >= 31) {  // <= 31 rows: lanes 0..nrow hold rowptr[r0..r1]
        ch.push_back(r);
        acc = 0;
      }
      acc += len;
    }
    if (ok) {
      ch.push_back(S->m);
      S->n_sp_chunks = (int)ch.size() - 1;
      if ((rc = dev_upload(c, &S->d_sp_chunk, ch.data(), ch.size()))) return rc;
    }
  }
  if ((rc = dev_alloc(c, &S->d_vals, (size_t)S->n_matrix * S->nnz))) return rc;
  if ((rc = dev_alloc(c, &S->d_b, (size_t)S->n_sys * S->m))) return rc;
  if ((rc = dev_alloc(c, &S->d_x, (size_t)S->n_sys * S->m))) return rc;
  if ((rc = dev_alloc(c, &S->d_dir_all, (size_t)S->m_global))) return rc;  // column flags are global, rows see the local slice
  S->d_dir = S->d_dir_all + S->row0;
  if ((rc = dev_alloc(c, &S->d_flag, (size_t)1))) return rc;
  EFB_CUDA(c, cudaMemsetAsync(S->d_flag, 0, sizeof(int32_t), c->stream));
  EFB_CUDA(c, cudaMemsetAsync(S->d_vals, 0, (size_t)S->n_matrix * S->nnz * sizeof(c128), c->stream));
  EFB_CUDA(c, cudaMemsetAsync(S->d_b, 0, (size_t)S->n_sys * S->m * sizeof(c128), c->stream));
  EFB_CUDA(c, cudaMemsetAsync(S->d_x, 0, (size_t)S->n_sys * S->m * sizeof(c128), c->stream));
  EFB_CUDA(c, cudaMemsetAsync(S->d_dir_all, 0, (size_t)S->m_global, c->stream));
  return EFB_OK;
}

}  // extern "C"

// Slots of the SELL-32 structure of the persistent small-system solver.  The order of the entries inside a SELL row is
// free.  The solver gathers p from shared memory with one 16-byte load per entry, and the 8 lanes of a quarter warp are
// served together only if their columns fall in 8 different bank groups (column mod 8): in CSR order they collide like
// random numbers (2.64 wavefronts per quarter on WR-90).  One thread per (slice, quarter): at every step the 8 rows pick,
// fewest choices first, an unused bank group among the entries they have left (their fullest one), else their fullest
// group (1.65 wavefronts; a maximum matching per step gives 1.56 at ten times the work).  Rows longer than
// SELL_FILL_MAX entries keep the CSR order.
constexpr int SELL_FILL_MAX = 160;
__global__ void __launch_bounds__(64)
k_sell_fill(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const int32_t *__restrict__ comp,
            const int32_t *__restrict__ orig, const int32_t *__restrict__ perm, const int32_t *__restrict__ sptr, int n_slices,
            int32_t *__restrict__ scol, int32_t *__restrict__ ssrc) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_slices * 4) return;
  const int sl = id >> 2, q = id & 3;
  const int base = sptr[sl], width = (sptr[sl + 1] - base) >> 5;
  int16_t ent[8][SELL_FILL_MAX];   // CSR offsets (relative to the row start) of the free entries, bucketed by bank group
  int16_t ecol[8][SELL_FILL_MAX];  // their compact columns (m_c <= 16384)
  int beg[8][8], cnt[8][8], left[8], rstart[8];
  bool fits = true;
  // The CSR walks are chains of dependent loads (column -> compact id): four entries are fetched per step so that the
  // loads of a step overlap.
  for (int t = 0; t < 8; ++t) {
    const int i = perm[sl * 32 + 8 * q + t];
    left[t] = 0;
    rstart[t] = 0;
    for (int g = 0; g < 8; ++g) cnt[t][g] = 0;
    if (i < 0) continue;
    const int r = orig[i];
    const int k0 = rowptr[r], k1 = rowptr[r + 1];
    rstart[t] = k0;
    for (int k = k0; k < k1; k += 4) {
      int c4[4], m4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c4[u] = (k + u < k1) ? colidx[k + u] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u) m4[u] = (c4[u] >= 0) ? comp[c4[u]] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (m4[u] >= 0) {
          cnt[t][m4[u] & 7]++;
          left[t]++;
        }
    }
    if (left[t] > SELL_FILL_MAX || k1 - k0 > 32767) fits = false;
  }
  if (!fits) {  // CSR order
    for (int t = 0; t < 8; ++t) {
      const int i = perm[sl * 32 + 8 * q + t];
      int j = 0;
      if (i >= 0) {
        const int r = orig[i];
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
          const int cc = comp[colidx[k]];
          if (cc < 0) continue;
          scol[base + j * 32 + 8 * q + t] = cc;
          ssrc[base + j * 32 + 8 * q + t] = k;
          ++j;
        }
      }
      for (; j < width; ++j) {
        scol[base + j * 32 + 8 * q + t] = 0;
        ssrc[base + j * 32 + 8 * q + t] = -1;
      }
    }
    return;
  }
  for (int t = 0; t < 8; ++t) {
    int run = 0;
    for (int g = 0; g < 8; ++g) {
      beg[t][g] = run;
      run += cnt[t][g];
      cnt[t][g] = 0;
    }
    if (left[t] == 0) continue;
    const int r = orig[perm[sl * 32 + 8 * q + t]];
    const int k0 = rowptr[r], k1 = rowptr[r + 1];
    for (int k = k0; k < k1; k += 4) {
      int c4[4], m4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) c4[u] = (k + u < k1) ? colidx[k + u] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u) m4[u] = (c4[u] >= 0) ? comp[c4[u]] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (m4[u] >= 0) {
          const int g = m4[u] & 7, at = beg[t][g] + cnt[t][g]++;
          ent[t][at] = (int16_t)(k + u - rstart[t]);
          ecol[t][at] = (int16_t)m4[u];
        }
    }
  }
  for (int j = 0; j < width; ++j) {
    unsigned used = 0, served = 0;
    int ng[8];
    for (int t = 0; t < 8; ++t) {
      ng[t] = 0;
      for (int g = 0; g < 8; ++g) ng[t] += cnt[t][g] > 0;
    }
    for (int u = 0; u < 8; ++u) {
      int t = -1;  // rows with fewer bank groups left choose first
      for (int v = 0; v < 8; ++v)
        if (!(served >> v & 1u) && left[v] > 0 && (t < 0 || ng[v] < ng[t])) t = v;
      if (t < 0) break;
      int best = -1, bc = 0, any = -1, ac = 0;
      for (int g = 0; g < 8; ++g) {
        const int cg = cnt[t][g];
        if (cg > ac) { ac = cg; any = g; }
        if (!(used >> g & 1u) && cg > bc) { bc = cg; best = g; }
      }
      const int g = best >= 0 ? best : any;
      used |= 1u << g;
      served |= 1u << t;
      const int at = beg[t][g] + --cnt[t][g];
      left[t]--;
      scol[base + j * 32 + 8 * q + t] = ecol[t][at];
      ssrc[base + j * 32 + 8 * q + t] = rstart[t] + ent[t][at];
    }
    for (int t = 0; t < 8; ++t)
      if (!(served >> t & 1u)) {
        scol[base + j * 32 + 8 * q + t] = 0;
        ssrc[base + j * 32 + 8 * q + t] = -1;
      }
  }
}

// Host build of the persistent small-system solver's structures (see System): compact numbering of the free
// unknowns, SELL-32 pattern over them, compact gradient lists.  Cheap (m <= 16384), runs once per system state.
int efb::build_small_structs(System *S) {
  Ctx *c = S->ctx;
  const int m = S->m;
  cudaStreamSynchronize(c->stream);
  dfree(S->d_sell_ptr); dfree(S->d_sell_col); dfree(S->d_sell_src); dfree(S->d_sell_perm); dfree(S->d_sell_vals);
  dfree(S->d_c_orig); dfree(S->d_c_edge_nodes); dfree(S->d_c_n2e_ptr); dfree(S->d_c_n2e_item);
  S->d_sell_ptr = S->d_sell_col = S->d_sell_src = S->d_sell_perm = nullptr;
  S->d_sell_vals = nullptr;
  S->d_c_orig = nullptr; S->d_c_edge_nodes = nullptr; S->d_c_n2e_ptr = nullptr; S->d_c_n2e_item = nullptr;
  const bool have_dir = (int)S->h_dir.size() == m;
  auto is_dir = [&](int e) { return have_dir && S->h_dir[e] != 0; };
  std::vector<int32_t> orig, comp((size_t)m, -1);
  orig.reserve(m);
  for (int r = 0; r < m; ++r)
    if (!is_dir(r)) {
      comp[r] = (int32_t)orig.size();
      orig.push_back(r);
    }
  const int mc = (int)orig.size();
  S->m_c = mc;
  std::vector<int32_t> len((size_t)mc, 0);
  for (int i = 0; i < mc; ++i) {
    const int r = orig[i];
    for (int k = S->h_rowptr[r]; k < S->h_rowptr[r + 1]; ++k) len[i] += !is_dir(S->h_colidx[k]);
  }
  std::vector<int32_t> order((size_t)mc);
  for (int i = 0; i < mc; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] > len[b]; });
  const int ns = (mc + 31) / 32;
  std::vector<int32_t> sptr((size_t)ns + 1, 0), perm((size_t)std::max(ns, 1) * 32, -1);
  for (int i = 0; i < mc; ++i) perm[i] = order[i];
  for (int sl = 0; sl < ns; ++sl) {
    int w = 0;
    for (int l = 0; l < 32; ++l) {
      const int i = perm[(size_t)sl * 32 + l];
      if (i >= 0) w = std::max(w, len[i]);
    }
    sptr[sl + 1] = sptr[sl] + 32 * w;
  }
  S->n_slices = ns;
  S->sell_total = sptr[ns];
  int rc;
  if ((rc = dev_upload(c, &S->d_c_orig, orig.data(), (size_t)std::max(mc, 1)))) return rc;
  if ((rc = dev_upload(c, &S->d_sell_ptr, sptr.data(), sptr.size()))) return rc;
  if ((rc = dev_upload(c, &S->d_sell_perm, perm.data(), perm.size()))) return rc;
  {
    // slots of the SELL structure on the device (bank-aware order of the entries inside a row: k_sell_fill)
    const size_t total = (size_t)std::max(sptr[ns], 1);
    if ((rc = dev_alloc(c, &S->d_sell_col, total))) return rc;
    if ((rc = dev_alloc(c, &S->d_sell_src, total))) return rc;
    int32_t *d_comp = nullptr;
    if ((rc = dev_upload(c, &d_comp, comp.data(), (size_t)std::max(m, 1)))) return rc;
    if (ns > 0) {
      k_sell_fill<<<(ns * 4 + 63) / 64, 64, 0, c->stream>>>(S->d_rowptr, S->d_colidx, d_comp, S->d_c_orig, S->d_sell_perm, S->d_sell_ptr, ns,
                                                          S->d_sell_col, S->d_sell_src);
      EFB_CHECK_LAUNCH(c);
    } else {
      EFB_CUDA(c, cudaMemsetAsync(S->d_sell_col, 0, sizeof(int32_t), c->stream));
      EFB_CUDA(c, cudaMemsetAsync(S->d_sell_src, 0xff, sizeof(int32_t), c->stream));
    }
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    dfree(d_comp);
  }
  if (S->n_node > 0 && (int)S->h_edge_nodes.size() == 2 * m) {
    const int nn = S->n_node;
    std::vector<int32_t> ptr((size_t)nn + 1, 0), item((size_t)std::max(2 * mc, 1));
    std::vector<int2> en((size_t)std::max(mc, 1));
    for (int i = 0; i < mc; ++i) {
      const int a = S->h_edge_nodes[2 * (size_t)orig[i]], b = S->h_edge_nodes[2 * (size_t)orig[i] + 1];
      en[i] = make_int2(a, b);
      ptr[a + 1]++;
      ptr[b + 1]++;
    }
    for (int n = 0; n < nn; ++n) ptr[n + 1] += ptr[n];
    std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < mc; ++i) {
      item[cur[en[i].x]++] = (i << 1);      // tail: G = -1
      item[cur[en[i].y]++] = (i << 1) | 1;  // head: G = +1
    }
    if ((rc = dev_upload(c, &S->d_c_edge_nodes, en.data(), en.size()))) return rc;
    if ((rc = dev_upload(c, &S->d_c_n2e_ptr, ptr.data(), ptr.size()))) return rc;
    if ((rc = dev_upload(c, &S->d_c_n2e_item, item.data(), item.size()))) return rc;
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vectors above are locals
  S->small_dirty = false;
  return EFB_OK;
}

extern "C" {

static int system_set_gradient(System *S, int n_node, const int32_t *edge_nodes) {
  Ctx *c = S->ctx;
  int rc;
  S->n_node = n_node;
  S->h_edge_nodes.assign(edge_nodes, edge_nodes + 2 * (size_t)S->m);
  S->small_dirty = true;
  S->cl_dirty = true;
  cudaStreamSynchronize(c->stream);  // blocks go back to a pool shared by every context on the device
  dfree(S->d_edge_nodes); dfree(S->d_n2e_ptr); dfree(S->d_n2e_item); dfree(S->d_node_dir);
  S->d_edge_nodes = nullptr; S->d_n2e_ptr = nullptr; S->d_n2e_item = nullptr; S->d_node_dir = nullptr;
  for (int64_t i = 0; i < 2 * (int64_t)S->m; ++i)
    if (edge_nodes[i] < 0 || edge_nodes[i] >= n_node) return fail(c, EFB_ERR_INVALID, "edge_nodes[%lld] out of range", (long long)i);
  if ((rc = dev_upload(c, &S->d_edge_nodes, (const int2 *)edge_nodes, (size_t)S->m))) return rc;
  std::vector<int32_t> ptr((size_t)n_node + 1, 0), item((size_t)2 * S->m);
  for (int64_t i = 0; i < 2 * (int64_t)S->m; ++i) ptr[edge_nodes[i] + 1]++;
  for (int n = 0; n < n_node; ++n) ptr[n + 1] += ptr[n];
  {
    std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
    for (int e = 0; e < S->m; ++e) {
      item[cur[edge_nodes[2 * e]]++] = (e << 2);          // tail: G = -1
      item[cur[edge_nodes[2 * e + 1]]++] = (e << 2) | 1;  // head: G = +1   (bit 1: Dirichlet edge, set by k_n2e_dirichlet)
    }
  }
  if ((rc = dev_upload(c, &S->d_n2e_ptr, ptr.data(), ptr.size()))) return rc;
  if ((rc = dev_upload(c, &S->d_n2e_item, item.data(), item.size()))) return rc;
  if ((rc = dev_alloc(c, &S->d_node_dir, (size_t)n_node))) return rc;
  EFB_CUDA(c, cudaMemsetAsync(S->d_node_dir, 0, (size_t)n_node, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

}  // extern "C"

// System over the rows [row0, row1) of the mesh's edge space (all rows for the ordinary single-GPU system).  Rows are
// indexed locally (0 .. row1-row0), columns keep their global edge ids.
static int system_create_rows(Mesh *M, int row0, int row1, int64_t n_extra, const int32_t *extra_rows, const int32_t *extra_cols,
                              int32_t n_matrix, int32_t n_rhs, efb_system **out) {
  Ctx *c = M->ctx;
  const int m = row1 - row0;
  SubTrace st;
  System *S = new System();
  S->ctx = c;
  S->mesh = M;
  S->m = m;
  S->m_global = M->m;
  S->row0 = row0;
  S->n_matrix = n_matrix;
  S->n_rhs = n_rhs;
  // large meshes: pattern + position map on the device (same arrays as the host code below)
  bool dev_done = false;
  if (n_extra == 0 && device_setup_enabled(M->n_tet) && M->d_tet_edges) {
    int rcd = device_pattern(S, &dev_done);
    if (rcd) return rcd;
  }
  if (dev_done) {
    st.mark("device pattern + position map");
    int32_t maxrow = 0;
    for (int r = 0; r < m; ++r) maxrow = std::max(maxrow, S->h_rowptr[r + 1] - S->h_rowptr[r]);
    int lim_nnz, lim_rows;
    asm_chunk_limits(&lim_nnz, &lim_rows);
    S->asm_row_kernel = asm_use_row_kernel();
    if (maxrow > lim_nnz || maxrow > 32767) {
      delete S;
      return fail(c, EFB_ERR_LIMIT, "efb_system_create: a row has %d entries (limit %d)", maxrow, std::min(lim_nnz, 32767));
    }
    std::vector<int32_t> chunk{0};
    int64_t acc = 0;
    int rows = 0;
    for (int r = 0; r < m; ++r) {
      const int len = S->h_rowptr[r + 1] - S->h_rowptr[r];
      if (acc + len > lim_nnz || rows >= lim_rows) {
        chunk.push_back(r);
        acc = 0;
        rows = 0;
      }
      acc += len;
      rows++;
    }
    chunk.push_back(m);
    S->n_chunks = (int)chunk.size() - 1;
    int rc;
    if ((rc = system_alloc_common(S))) return rc;
    st.mark("diag/stream structures + device buffers");
    if ((rc = dev_upload(c, &S->d_chunk_row, chunk.data(), chunk.size()))) return rc;
    if (m == M->m && (rc = system_set_gradient(S, M->n_node, M->h_edge_nodes.data()))) return rc;
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));
    st.mark("gradient lists + sync");
    *out = (efb_system *)S;
    return EFB_OK;
  }
  {
    int rch = mesh_host_e2t_item(M);
    if (rch) return rch;
  }
  // The pattern of a (mesh, extras) pair is cached on the mesh: drivers that create one system per call (a Python loop
  // over calculate_sparams_eigenmode, one call per frequency) rebuild nothing but the value arrays.
  uint64_t xhash = 1469598103934665603ull;
  for (int64_t i = 0; i < n_extra; ++i) {
    xhash = (xhash ^ (uint64_t)(uint32_t)extra_rows[i]) * 1099511628211ull;
    xhash = (xhash ^ (uint64_t)(uint32_t)extra_cols[i]) * 1099511628211ull;
  }
  Mesh::PatternCache &PC = M->pat_cache;
  const bool pat_hit = PC.valid && PC.row0 == row0 && PC.row1 == row1 && PC.n_extra == n_extra && PC.hash == xhash;
  std::vector<int32_t> rowlen(m);
  std::vector<uint16_t> pos;
  struct RangeCols {
    int64_t a = 0, b = 0;
    std::vector<int32_t> cols;
  };
  std::vector<RangeCols> ranges(64);
  std::atomic<int> n_ranges{0};
  if (pat_hit) {
    pos = PC.pos;
    for (int r = 0; r < m; ++r) rowlen[r] = PC.rowptr[r + 1] - PC.rowptr[r];
  } else {
  // extras bucketed by row
  std::vector<int64_t> xptr((size_t)m + 1, 0);
  for (int64_t i = 0; i < n_extra; ++i) xptr[extra_rows[i] + 1]++;
  for (int r = 0; r < m; ++r) xptr[r + 1] += xptr[r];
  std::vector<int32_t> xcol((size_t)n_extra);
  {
    std::vector<int64_t> cur(xptr.begin(), xptr.end() - 1);
    for (int64_t i = 0; i < n_extra; ++i) xcol[cur[extra_rows[i]]++] = extra_cols[i];
  }
  // pass 1: row lengths ; pass 2: fill
  const int32_t *te = M->h_tet_edges.data();
  const int64_t kpos0 = M->h_e2t_ptr[row0];  // first incidence of the local rows: origin of the position map
  auto row_cols = [&](int r, std::vector<int32_t> &buf) {
    buf.clear();
    for (int32_t k = M->h_e2t_ptr[row0 + r]; k < M->h_e2t_ptr[row0 + r + 1]; ++k) {
      const int32_t *e6 = te + 6 * (int64_t)(M->h_e2t_item[k] >> 3);
      buf.insert(buf.end(), e6, e6 + 6);
    }
    buf.insert(buf.end(), xcol.begin() + xptr[r], xcol.begin() + xptr[r + 1]);
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
  };
  // one pass: every worker builds the sorted column lists of its row range into a private buffer and fills the
  // tet-local -> row-local position map; the buffers are then concatenated at the row offsets
  pos.resize((size_t)(M->h_e2t_ptr[row1] - kpos0) * 6);
  parallel_for(m, [&](int64_t a, int64_t b) {
    RangeCols &rc_ = ranges[n_ranges.fetch_add(1)];
    rc_.a = a;
    rc_.b = b;
    rc_.cols.reserve((size_t)(b - a) * 20);
    std::vector<int32_t> buf;
    for (int64_t r = a; r < b; ++r) {
      row_cols((int)r, buf);
      rowlen[r] = (int32_t)buf.size();
      rc_.cols.insert(rc_.cols.end(), buf.begin(), buf.end());
      for (int32_t k = M->h_e2t_ptr[row0 + r]; k < M->h_e2t_ptr[row0 + r + 1]; ++k) {
        const int32_t *e6 = te + 6 * (int64_t)(M->h_e2t_item[k] >> 3);
        for (int j = 0; j < 6; ++j)
          pos[(size_t)(k - kpos0) * 6 + j] = (uint16_t)(std::lower_bound(buf.begin(), buf.end(), e6[j]) - buf.begin());
      }
    }
  });
  }
  S->h_rowptr.assign((size_t)m + 1, 0);
  int64_t nnz = 0;
  int32_t maxrow = 0;
  for (int r = 0; r < m; ++r) {
    nnz += rowlen[r];
    maxrow = std::max(maxrow, rowlen[r]);
  }
  if (nnz >= (1ll << 31)) {
    delete S;
    return fail(c, EFB_ERR_LIMIT, "efb_system_create: nnz %lld >= 2^31 (int32 CSR like Eigen's default index)", (long long)nnz);
  }
  int lim_nnz, lim_rows;
  asm_chunk_limits(&lim_nnz, &lim_rows);
  S->asm_row_kernel = asm_use_row_kernel();
  if (maxrow > lim_nnz || maxrow > 32767) {
    delete S;
    return fail(c, EFB_ERR_LIMIT, "efb_system_create: a row has %d entries (limit %d)", maxrow, std::min(lim_nnz, 32767));
  }
  st.mark("row column lists + pos map");
  S->nnz = nnz;
  for (int r = 0; r < m; ++r) S->h_rowptr[r + 1] = S->h_rowptr[r] + rowlen[r];
  S->h_colidx.resize((size_t)nnz);
  if (pat_hit) {
    S->h_colidx = PC.colidx;
  } else {
    const int nr = n_ranges.load();
    std::vector<std::thread> th;
    for (int i = 0; i < nr; ++i) {
      auto cp = [&, i] { std::copy(ranges[i].cols.begin(), ranges[i].cols.end(), S->h_colidx.data() + S->h_rowptr[ranges[i].a]); };
      if (nr > 1 && ranges[i].cols.size() > (1u << 20)) th.emplace_back(cp); else cp();
    }
    for (auto &x : th) x.join();
    if (m <= 400000) {  // small and mid-size meshes only: the cache holds host copies
      PC.valid = true;
      PC.row0 = row0; PC.row1 = row1; PC.n_extra = n_extra; PC.hash = xhash;
      PC.rowptr = S->h_rowptr;
      PC.colidx = S->h_colidx;
      PC.pos = pos;
    }
  }
  st.mark("concatenate");
  // assembly chunks: consecutive rows with <= lim_nnz entries and <= lim_rows rows
  std::vector<int32_t> chunk{0};
  {
    int64_t acc = 0;
    int rows = 0;
    for (int r = 0; r < m; ++r) {
      if (acc + rowlen[r] > lim_nnz || rows >= lim_rows) {
        chunk.push_back(r);
        acc = 0;
        rows = 0;
      }
      acc += rowlen[r];
      rows++;
    }
    chunk.push_back(m);
  }
  S->n_chunks = (int)chunk.size() - 1;
  int rc;
  if ((rc = system_alloc_common(S))) return rc;
  st.mark("diag/SELL/stream structures + device buffers");
  if ((rc = dev_upload(c, &S->d_e2t_pos, pos.data(), pos.size()))) return rc;
  if ((rc = dev_upload(c, &S->d_chunk_row, chunk.data(), chunk.size()))) return rc;

  // the discrete gradient (auxiliary-space preconditioner) is only kept for whole-mesh systems
  if (m == M->m && (rc = system_set_gradient(S, M->n_node, M->h_edge_nodes.data()))) return rc;
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  st.mark("gradient lists + sync");
  *out = (efb_system *)S;
  return EFB_OK;
}

extern "C" {

int efb_system_create(efb_mesh *mesh_, int64_t n_extra, const int32_t *extra_rows, const int32_t *extra_cols,
                      int32_t n_matrix, int32_t n_rhs, efb_system **out) {
  Mesh *M = (Mesh *)mesh_;
  if (!M || !out) return fail(M ? M->ctx : nullptr, EFB_ERR_INVALID, "efb_system_create: NULL argument");
  Ctx *c = M->ctx;
  *out = nullptr;
  if (n_matrix <= 0 || n_rhs <= 0 || n_extra < 0 || (n_extra > 0 && (!extra_rows || !extra_cols)))
    return fail(c, EFB_ERR_INVALID, "efb_system_create: bad arguments");
  EFB_CUDA(c, cudaSetDevice(c->device));
  const int m = M->m;
  for (int64_t i = 0; i < n_extra; ++i)
    if (extra_rows[i] < 0 || extra_rows[i] >= m || extra_cols[i] < 0 || extra_cols[i] >= m)
      return fail(c, EFB_ERR_INVALID, "efb_system_create: extra entry %lld out of range", (long long)i);
  return system_create_rows(M, 0, m, n_extra, extra_rows, extra_cols, n_matrix, n_rhs, out);
}

int efb_system_create_rows(efb_mesh *mesh_, int32_t row_begin, int32_t row_end, int32_t n_matrix, int32_t n_rhs, efb_system **out) {
  Mesh *M = (Mesh *)mesh_;
  if (!M || !out) return fail(M ? M->ctx : nullptr, EFB_ERR_INVALID, "efb_system_create_rows: NULL argument");
  Ctx *c = M->ctx;
  *out = nullptr;
  if (n_matrix <= 0 || n_rhs <= 0 || row_begin < 0 || row_end > M->m || row_begin >= row_end)
    return fail(c, EFB_ERR_INVALID, "efb_system_create_rows: bad arguments");
  EFB_CUDA(c, cudaSetDevice(c->device));
  return system_create_rows(M, row_begin, row_end, 0, nullptr, nullptr, n_matrix, n_rhs, out);
}

int efb_system_create_csr(efb_ctx *ctx_, int32_t m, int64_t nnz, const int32_t *rowptr, const int32_t *colidx,
                          const double *vals, int32_t n_matrix, int32_t n_rhs, efb_system **out) {
  Ctx *c = (Ctx *)ctx_;
  if (!c || !out || !rowptr || !colidx) return fail(c, EFB_ERR_INVALID, "efb_system_create_csr: NULL argument");
  *out = nullptr;
  if (m <= 0 || nnz < 0 || nnz >= (1ll << 31) || n_matrix <= 0 || n_rhs <= 0 || rowptr[0] != 0 || rowptr[m] != nnz)
    return fail(c, EFB_ERR_INVALID, "efb_system_create_csr: bad dimensions");
  for (int r = 0; r < m; ++r) {
    if (rowptr[r + 1] < rowptr[r]) return fail(c, EFB_ERR_INVALID, "efb_system_create_csr: rowptr not monotone");
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      if (colidx[k] < 0 || colidx[k] >= m) return fail(c, EFB_ERR_INVALID, "efb_system_create_csr: column out of range");
      if (k > rowptr[r] && colidx[k] <= colidx[k - 1])
        return fail(c, EFB_ERR_INVALID, "efb_system_create_csr: columns must be strictly ascending per row");
    }
  }
  EFB_CUDA(c, cudaSetDevice(c->device));
  System *S = new System();
  S->ctx = c;
  S->m = m;
  S->m_global = m;
  S->nnz = nnz;
  S->n_matrix = n_matrix;
  S->n_rhs = n_rhs;
  S->h_rowptr.assign(rowptr, rowptr + m + 1);
  S->h_colidx.assign(colidx, colidx + nnz);
  int rc;
  if ((rc = system_alloc_common(S))) return rc;
  if (vals)
    for (int f = 0; f < n_matrix; ++f)
      EFB_CUDA(c, cudaMemcpyAsync(S->d_vals + (size_t)f * nnz, vals, (size_t)nnz * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  S->assembled = vals != nullptr;
  *out = (efb_system *)S;
  return EFB_OK;
}

void efb_system_destroy(efb_system *sys_) {
  System *S = (System *)sys_;
  if (!S) return;
  cudaSetDevice(S->ctx->device);
  cudaStreamSynchronize(S->ctx->stream);
  dist_free(S);
  solver_free(S);
  cluster_plan_free(S);
  dfree(S->d_rowptr); dfree(S->d_colidx); dfree(S->d_diag_pos); dfree(S->d_vals);
  dfree(S->d_b); dfree(S->d_x); dfree(S->d_dir_all); dfree(S->d_e2t_pos); dfree(S->d_chunk_row); dfree(S->d_sp_chunk);
  dfree(S->d_sch_item); dfree(S->d_sch_ss); dfree(S->d_sch_row); dfree(S->d_sch_pos); dfree(S->d_sch_sec); dfree(S->d_sch_flag);
  dfree(S->d_sell_ptr); dfree(S->d_sell_col); dfree(S->d_sell_perm); dfree(S->d_sell_vals); dfree(S->d_sell_src);
  dfree(S->d_c_orig); dfree(S->d_c_edge_nodes); dfree(S->d_c_n2e_ptr); dfree(S->d_c_n2e_item);
  dfree(S->d_edge_nodes); dfree(S->d_n2e_ptr); dfree(S->d_n2e_item); dfree(S->d_node_dir);
  dfree(S->d_mat_blob);
  delete S;
}

int efb_system_dims(const efb_system *sys_, int32_t *m, int64_t *nnz, int32_t *n_matrix, int32_t *n_rhs) {
  const System *S = (const System *)sys_;
  if (!S) return EFB_ERR_INVALID;
  if (m) *m = S->m;
  if (nnz) *nnz = S->nnz;
  if (n_matrix) *n_matrix = S->n_matrix;
  if (n_rhs) *n_rhs = S->n_rhs;
  return EFB_OK;
}

int efb_system_get_pattern(const efb_system *sys_, int32_t *rowptr, int32_t *colidx) {
  const System *S = (const System *)sys_;
  if (!S) return EFB_ERR_INVALID;
  if (rowptr) memcpy(rowptr, S->h_rowptr.data(), S->h_rowptr.size() * sizeof(int32_t));
  if (colidx) memcpy(colidx, S->h_colidx.data(), S->h_colidx.size() * sizeof(int32_t));
  return EFB_OK;
}

int efb_system_get_values(efb_system *sys_, int32_t f, double *vals) {
  System *S = (System *)sys_;
  if (!S || !vals || f < 0 || f >= S->n_matrix) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_system_get_values: bad argument");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  EFB_CUDA(c, cudaMemcpyAsync(vals, S->d_vals + (size_t)f * S->nnz, (size_t)S->nnz * sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_system_set_values(efb_system *sys_, int32_t f, const double *vals) {
  System *S = (System *)sys_;
  if (!S || !vals || f < 0 || f >= S->n_matrix) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_system_set_values: bad argument");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  EFB_CUDA(c, cudaMemcpyAsync(S->d_vals + (size_t)f * S->nnz, vals, (size_t)S->nnz * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  S->assembled = true;
  return EFB_OK;
}

int efb_system_set_dirichlet(efb_system *sys_, const uint8_t *flags) {
  System *S = (System *)sys_;
  if (!S || !flags) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_system_set_dirichlet: bad argument");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  // flags cover ALL edges of the mesh (m_global entries): rows read the local slice, columns the whole array
  EFB_CUDA(c, cudaMemcpyAsync(S->d_dir_all, flags, (size_t)S->m_global, cudaMemcpyHostToDevice, c->stream));
  S->has_dir = true;
  S->h_dir.assign(flags + S->row0, flags + S->row0 + S->m);
  S->h_dir_all.assign(flags, flags + S->m_global);
  if (S->dist_state) dist_free(S);  // the distributed solver's nodal lists depend on the flags
  S->small_dirty = true;
  S->cl_dirty = true;
  S->sched_dirty = true;
  if (S->d_e2t_pos && S->mesh) {
    k_pos_dirichlet<<<(S->m + 127) / 128, 128, 0, c->stream>>>(S->mesh->d_e2t_ptr + S->row0, S->d_rowptr, S->d_colidx, S->d_dir_all, S->m,
                                                               S->d_e2t_pos, (long long)S->mesh->h_e2t_ptr[S->row0] * 6);
    EFB_CHECK_LAUNCH(c);
  }
  if (S->d_node_dir) {
    EFB_CUDA(c, cudaMemsetAsync(S->d_node_dir, 0, (size_t)S->n_node, c->stream));
    k_node_dir<<<(S->m + 255) / 256, 256, 0, c->stream>>>(S->d_edge_nodes, S->d_dir, S->m, S->d_node_dir);
    EFB_CHECK_LAUNCH(c);
    k_n2e_dirichlet<<<(2 * S->m + 255) / 256, 256, 0, c->stream>>>(S->d_n2e_item, 2ll * S->m, S->d_dir);
    EFB_CHECK_LAUNCH(c);
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_system_set_gradient(efb_system *sys_, int32_t n_node, const int32_t *edge_nodes) {
  System *S = (System *)sys_;
  if (!S || !edge_nodes || n_node <= 0) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_system_set_gradient: bad argument");
  EFB_CUDA(S->ctx, cudaSetDevice(S->ctx->device));
  int rc = system_set_gradient(S, n_node, edge_nodes);
  if (rc) return rc;
  if (S->has_dir) {
    k_node_dir<<<(S->m + 255) / 256, 256, 0, S->ctx->stream>>>(S->d_edge_nodes, S->d_dir, S->m, S->d_node_dir);
    EFB_CHECK_LAUNCH(S->ctx);
    k_n2e_dirichlet<<<(2 * S->m + 255) / 256, 256, 0, S->ctx->stream>>>(S->d_n2e_item, 2ll * S->m, S->d_dir);
    EFB_CHECK_LAUNCH(S->ctx);
    EFB_CUDA(S->ctx, cudaStreamSynchronize(S->ctx->stream));
  }
  return EFB_OK;
}

// ------------------------------------------------------------------ rhs / x
static int vec_io(System *S, c128 *base, int32_t idx, double *out, const double *in, const char *who) {
  if (!S || idx < 0 || idx >= S->n_sys || (!out && !in)) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "%s: bad argument", who);
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  c128 *p = base + (size_t)idx * S->m;
  if (out)
    EFB_CUDA(c, cudaMemcpyAsync(out, p, (size_t)S->m * sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
  else
    EFB_CUDA(c, cudaMemcpyAsync(p, in, (size_t)S->m * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_rhs_zero(efb_system *sys_, int32_t rhs) {
  System *S = (System *)sys_;
  if (!S || rhs < 0 || rhs >= S->n_sys) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_rhs_zero: bad argument");
  EFB_CUDA(S->ctx, cudaSetDevice(S->ctx->device));
  EFB_CUDA(S->ctx, cudaMemsetAsync(S->d_b + (size_t)rhs * S->m, 0, (size_t)S->m * sizeof(c128), S->ctx->stream));
  return EFB_OK;
}
int efb_rhs_set(efb_system *s, int32_t rhs, const double *b) { return vec_io((System *)s, s ? ((System *)s)->d_b : nullptr, rhs, nullptr, b, "efb_rhs_set"); }
int efb_rhs_get(efb_system *s, int32_t rhs, double *b) { return vec_io((System *)s, s ? ((System *)s)->d_b : nullptr, rhs, b, nullptr, "efb_rhs_get"); }
int efb_x_get(efb_system *s, int32_t rhs, double *x) { return vec_io((System *)s, s ? ((System *)s)->d_x : nullptr, rhs, x, nullptr, "efb_x_get"); }
int efb_x_set(efb_system *s, int32_t rhs, const double *x) { return vec_io((System *)s, s ? ((System *)s)->d_x : nullptr, rhs, nullptr, x, "efb_x_set"); }

}  // extern "C"
