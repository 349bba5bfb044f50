// Shared internals of libedgefem_b200 (sm_100a only).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "edgefem_b200.h"

namespace efb {

typedef double2 c128;  // (re, im)

// ------------------------------------------------------------------ complex helpers
__host__ __device__ __forceinline__ c128 cmake(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ c128 cadd(c128 a, c128 b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ c128 csub(c128 a, c128 b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ c128 cmul(c128 a, c128 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ c128 cmulconj(c128 a, c128 b) {  // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ c128 cscale(double s, c128 a) { return make_double2(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ c128 cconj(c128 a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ c128 cneg(c128 a) { return make_double2(-a.x, -a.y); }
__host__ __device__ __forceinline__ double cabs2(c128 a) { return a.x * a.x + a.y * a.y; }
// Smith's algorithm (what std::complex<double> division does without -ffast-math)
__host__ __device__ __forceinline__ c128 cdiv(c128 a, c128 b) {
  if (fabs(b.x) >= fabs(b.y)) {
    double r = b.y / b.x, den = b.x + b.y * r;
    return make_double2((a.x + a.y * r) / den, (a.y - a.x * r) / den);
  } else {
    double r = b.x / b.y, den = b.x * r + b.y;
    return make_double2((a.x * r + a.y) / den, (a.y * r - a.x) / den);
  }
}
__host__ __device__ __forceinline__ c128 cfma(c128 a, c128 b, c128 acc) {  // acc + a*b
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
  return acc;
}

// ------------------------------------------------------------------ limits
// Row-gather volume assembly.  A CTA owns a chunk of consecutive rows with <= ASM_CHUNK_NNZ entries and accumulates it in
// a shared-memory image.  Default: k_assemble_volume_s (rank-major schedule, one row per thread in every rank);
// fallback for meshes with an edge in more than ASM_MAX_RANK tets and EDGEFEM_B200_ASM_KERNEL=batch:
// k_assemble_volume_b (lanes = 32 consecutive incidences, adds in rounds).  The first-generation thread-per-row kernel
// (k_assemble_volume, EDGEFEM_B200_ASM_KERNEL=row) uses the larger ASMR_* chunks.  Measured on the 1.57 M-tet cube:
// 128-thread CTAs x 6 per SM 0.335 ms, 256 x 3 0.364 ms, 64 x 12 0.346 ms.
constexpr int ASM_CHUNK_NNZ = 1920;   // matrix entries per assembly CTA (30 KB of c128 / 15 KB of double accumulators)
constexpr int ASM_CHUNK_ROWS = 128;   // rows per assembly CTA = threads per CTA: one row per thread in every rank
constexpr int ASM_MAX_RANK = 62;      // scheduled kernel: most incident tets of an edge (else the batched kernel runs)
constexpr int ASM_SEC_STRIDE = ASM_MAX_RANK + 2;
constexpr int ASMB_THREADS = 128;
static_assert(ASM_CHUNK_ROWS <= ASMB_THREADS, "the scheduled kernel gives every row of a rank its own thread");
constexpr int ASMB_CTAS_PER_SM = 6;
constexpr int ASMR_CHUNK_NNZ = 5376;  // thread-per-row kernel: entries per CTA
constexpr int ASMR_CHUNK_ROWS = 768;  // thread-per-row kernel: rows per CTA (also the bank-skew padding)
constexpr int ASM_ACC_ENTRIES = ASMR_CHUNK_NNZ + ASMR_CHUNK_ROWS;  // 96 KB of c128 accumulators
constexpr int ASM_THREADS = 384;
constexpr int MAX_SLOTS = 256;        // distinct physical tags
constexpr int NSCAL = 16;             // per-system device scalars (c128)
constexpr int RED_MAX_BLOCKS = 1184;  // 148 SMs x 8
constexpr int SPMV_STREAM_W = 256;    // entries per warp in the CSR-stream SpMV

// cached per-tet geometry record (144 B, 16-byte aligned rows): see k_tet_geometry
struct alignas(16) TetGeom {
  double gg[4][4];     // Gram matrix of the barycentric gradients
  double V;            // volume
  uint32_t sign_slot;  // bits 0-5: edge k stored with orientation -1 ; bits 8-15: material slot
  uint32_t pad;
};
static_assert(sizeof(TetGeom) == 144, "TetGeom layout");

// the same information in 96 B = three 32-byte sectors (one 256-bit load each): the 10 distinct Gram products
// (row-major upper triangle: 00 01 02 03 11 12 13 22 23 33), the volume and V/10; sign|slot lives in Mesh::d_e2t_ss
struct alignas(32) TetRec {
  double g[10];
  double V;
  double Ieq;          // V / 10 (the division of src/edge_basis.cpp:28-30 done once per tet); V / 20 = Ieq / 2 exactly
};
static_assert(sizeof(TetRec) == 96, "TetRec layout");

// ------------------------------------------------------------------ handles
struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // user stopwatch
  bool ev_valid = false;
  int64_t launches = 0;
  int sm_count = 148;
  std::string err;
  void *dist = nullptr;  // efb::Dist (rank, world, NCCL communicator) once efb_dist_init ran
  void *d_flush = nullptr;  // 256 MB scratch written by efb_l2_flush
};

struct Mesh {
  Ctx *ctx = nullptr;
  int n_node = 0, n_tet = 0, m = 0, n_slots = 0;
  std::vector<int32_t> slot_tags;
  // host copies needed to build per-system maps
  std::vector<int32_t> h_tet_edges;  // [6*n_tet]
  std::vector<int32_t> h_e2t_ptr;    // [m+1]
  std::vector<int32_t> h_e2t_item;   // [6*n_tet]  tet<<3 | local
  std::vector<int32_t> h_edge_nodes; // [2*m]
  // device
  double4 *d_xyz = nullptr;          // [n_node] (x,y,z,0)
  int4 *d_tet_nodes = nullptr;       // [n_tet]
  uint8_t *d_tet_sign = nullptr;     // [n_tet] bit k set => orient[k] == -1
  uint8_t *d_tet_slot = nullptr;     // [n_tet]
  int32_t *d_e2t_ptr = nullptr;      // [m+1]
  int32_t *d_e2t_item = nullptr;     // [6*n_tet]
  int32_t *d_tet_edges = nullptr;    // [6*n_tet] (large meshes only: input of the device pattern builder)
  double *d_slot_bbox = nullptr;     // [n_slots*6] min xyz, max xyz over the slot's tets
  TetGeom *d_geom = nullptr;         // [n_tet] frequency-independent element geometry (thread-per-row kernel only)
  TetRec *d_rec = nullptr;           // [n_tet] the same, packed (batched kernel)
  uint16_t *d_e2t_ss = nullptr;      // [6*n_tet] sign | slot << 8 of the incidence's tet
  // host copy of the last CSR pattern + position map built on this mesh, keyed by the row range and the extras
  struct PatternCache {
    bool valid = false;
    int row0 = 0, row1 = 0;
    int64_t n_extra = 0;
    uint64_t hash = 0;
    std::vector<int32_t> rowptr, colidx;
    std::vector<uint16_t> pos;
  } pat_cache;
};

struct Port;

struct System {
  Ctx *ctx = nullptr;
  Mesh *mesh = nullptr;  // may be null (generic CSR system)
  int m = 0, n_matrix = 0, n_rhs = 0, n_sys = 0, n_node = 0;  // m = LOCAL rows
  int m_global = 0, row0 = 0;  // row-partitioned systems own rows [row0, row0+m) of m_global; columns are global ids
  int64_t nnz = 0;
  std::vector<int32_t> h_rowptr, h_colidx;
  int32_t *d_rowptr = nullptr, *d_colidx = nullptr, *d_diag_pos = nullptr;
  c128 *d_vals = nullptr;      // [n_matrix][nnz]
  c128 *d_b = nullptr;         // [n_sys][m]
  c128 *d_x = nullptr;         // [n_sys][m]
  uint8_t *d_dir_all = nullptr; // [m_global] Dirichlet flags of every edge (column lookups)
  uint8_t *d_dir = nullptr;    // = d_dir_all + row0: flags of the local rows
  bool has_dir = false;
  // assembly maps (mesh-born systems)
  uint16_t *d_e2t_pos = nullptr;   // [6*n_tet*6] column offsets inside the row (bit 15: Dirichlet column)
  int32_t *d_chunk_row = nullptr;  // [n_chunks+1]
  int n_chunks = 0;
  bool asm_row_kernel = false;     // chunks were cut for the thread-per-row kernel (ASMR_* limits)
  // rank-major assembly schedule (k_assemble_volume_s): per chunk, the r-th incidence of every non-Dirichlet row with
  // more than r incident tets, rows ascending, ranks ascending.  Built lazily; rebuilt when the Dirichlet flags change.
  bool sched_dirty = true, sched_ok = false;
  int32_t *d_sch_item = nullptr;   // [n_inc] tet << 3 | local edge, schedule order (chunk regions = incidence regions)
  uint16_t *d_sch_ss = nullptr;    // [n_inc] sign | slot << 8
  uint16_t *d_sch_row = nullptr;   // [n_inc] row inside the chunk
  uint16_t *d_sch_pos = nullptr;   // [n_inc*6] row-local positions (copy of d_e2t_pos in schedule order)
  int32_t *d_sch_sec = nullptr;    // [n_chunks][ASM_SEC_STRIDE]: number of ranks, then the section offsets
  int32_t *d_sch_flag = nullptr;   // [1] a chunk has a row with more than ASM_MAX_RANK incident tets
  // CSR-stream SpMV: row-aligned chunks of <= SPMV_STREAM_W entries (one warp each); null if a row is longer
  int32_t *d_sp_chunk = nullptr;   // [n_sp_chunks+1]
  int n_sp_chunks = 0;
  // Structures of the persistent small-system solver (built lazily by build_small_structs, rebuilt when the
  // Dirichlet flags or the gradient change).  The solver works on the FREE unknowns only: Dirichlet rows are identity
  // rows (x_e = b_e / A_ee) and Dirichlet columns hold explicit zeros, so both are dropped -- "compact" ids number
  // the free edges in ascending original order.  SELL-32 copy of the compact pattern: rows sorted by length (desc),
  // slices of 32 rows stored column-major and padded to the slice's longest row.
  bool small_dirty = true;
  // cluster-split persistent solver (cluster.cu): plan of the split, rebuilt with the small structures
  bool cl_dirty = true;
  void *cl_plan = nullptr;
  int last_cluster_c = 0, last_cluster_nr = 0, last_cluster_n = 0;  // shape of the last cluster launch (0: not used)
  int m_c = 0;                     // free unknowns
  int32_t *d_c_orig = nullptr;     // [m_c] compact id -> original edge id
  int2 *d_c_edge_nodes = nullptr;  // [m_c] (tail, head) node of the compact edge
  int32_t *d_c_n2e_ptr = nullptr;  // [n_node+1]
  int32_t *d_c_n2e_item = nullptr; // [<=2 m_c]  compact edge << 1 | (1 if the node is the head (+1) else 0 (-1))
  int32_t *d_sell_ptr = nullptr;   // [n_slices+1] entry offsets (multiples of 32)
  int32_t *d_sell_col = nullptr;   // [sell_total] compact column of every slot (0 for padding)
  int32_t *d_sell_src = nullptr;   // [sell_total] CSR position the slot's value comes from (-1 for padding)
  int32_t *d_sell_perm = nullptr;  // [n_slices*32] compact row of every SELL row (-1 for padding rows)
  c128 *d_sell_vals = nullptr;     // [n_matrix][sell_total] (lazy)
  int n_slices = 0;
  long long sell_total = 0;
  std::vector<uint8_t> h_dir;         // host copy of the Dirichlet flags of the local rows (empty: none set)
  std::vector<uint8_t> h_dir_all;     // row-partitioned systems: host copy of ALL flags (nodal lists of the distributed preconditioner)
  std::vector<int32_t> h_edge_nodes;  // host copy of the gradient (2 per edge; empty: none set)
  cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;  // around the persistent solver kernel
  bool small_timed = false;
  // gradient (aux preconditioner)
  int2 *d_edge_nodes = nullptr;    // [m]
  int32_t *d_n2e_ptr = nullptr;    // [n_node+1]
  int32_t *d_n2e_item = nullptr;   // [2m]  edge<<2 | (Dirichlet edge)<<1 | (1 if node is n1 (head, +1) else 0 (tail, -1))
  uint8_t *d_node_dir = nullptr;   // [n_node] node touches a Dirichlet edge
  // solver workspace (lazy)
  c128 *d_work = nullptr;      // [n_vec][n_sys][m]
  int n_work_vec = 0;
  c128 *d_dinv = nullptr;      // [n_matrix][m]
  c128 *d_linv = nullptr;      // [n_matrix][n_node]
  c128 *d_w = nullptr;         // [n_sys][n_node]
  c128 *d_scal = nullptr;      // [n_sys][NSCAL]
  double *d_partial = nullptr; // [n_sys][RED_MAX_BLOCKS][4]
  unsigned *d_counter = nullptr; // [n_sys]
  int32_t *d_state = nullptr;  // [n_sys][4]: active, iters, flag, pad
  int32_t *d_flag = nullptr;   // [1] device error flag (entry missing from the pattern)
  int32_t *d_job = nullptr;    // [1 + n_sys] work-queue counter of the persistent small-system solver, then the job order
  std::vector<int32_t> h_job_order;  // longest-expected-first order of the queue (kept alive for the async upload)
  std::vector<int32_t> last_iters;   // iterations per system of the previous solve (best predictor of the next one)
  // last assembly inputs (for efb_bench_kernel which=3)
  std::vector<double> last_omega;
  bool assembled = false;
  // materials scratch
  void *d_mat_blob = nullptr;
  size_t mat_blob_bytes = 0;
  std::vector<uint8_t> last_mat_blob;
  int last_mode = 0;
  void *dist_state = nullptr;  // efb::DistState of a row-partitioned system (dist.cu)
};

struct Port {
  System *sys = nullptr;
  int n_edges = 0;
  int64_t n_ms = 0;
  int32_t *d_edges = nullptr;
  c128 *d_w = nullptr;        // weights, zero on Dirichlet edges
  int32_t *d_ms_row = nullptr, *d_ms_col = nullptr, *d_ms_pos = nullptr;  // pos in the CSR pattern (-1: absent)
  double *d_ms_val = nullptr;
  c128 *d_e = nullptr;        // [m] dense port vector (weights scattered)
  c128 *d_tmp = nullptr;      // [m] M_s * v scratch
  int32_t *d_blk_pos = nullptr; // [n_edges*n_edges] CSR positions of the dense block (-1 absent)
};

// ------------------------------------------------------------------ error handling
void set_global_error(const std::string &s);
int fail(Ctx *ctx, int code, const char *fmt, ...);

#define EFB_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::efb::fail((ctx), EFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

// entry points that index vectors by global edge id refuse row-partitioned systems
#define EFB_WHOLE_ONLY(S, name)                                                                              \
  do {                                                                                                       \
    if ((S) && (S)->m != (S)->m_global)                                                                      \
      return ::efb::fail((S)->ctx, EFB_ERR_STATE, name ": not available on a row-partitioned system (efb_system_create_rows); use the efb_dist_* calls"); \
  } while (0)

#define EFB_CHECK_LAUNCH(ctx)                                                            \
  do {                                                                                   \
    (ctx)->launches++;                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return ::efb::fail((ctx), EFB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",        \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

// Caching device allocator (per device): large buffers released by destroyed handles are kept and
// handed to the next request of similar size, so repeated driver calls (one system per sweep call)
// do not pay cudaMalloc/cudaFree -- which cost milliseconds and occasionally tens of milliseconds.
size_t pool_round(size_t bytes);                   // allocation size for a request (small ones -> power of two)
void *pool_get(int device, size_t bytes);          // nullptr if no cached block fits
void pool_register(void *p, int device, size_t bytes);
void dfree(void *p);                               // returns the block to the pool or cudaFree()s it
void pool_trim();                                  // cudaFree every cached block

template <typename T>
int dev_alloc(Ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  const size_t bytes = pool_round(n * sizeof(T) + 16);  // 16 B of slack: 16-byte bulk copies of aligned supersets may read past the last element
  if (void *q = pool_get(ctx->device, bytes)) {
    *p = (T *)q;
    return EFB_OK;
  }
  cudaError_t e = cudaMalloc((void **)p, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give cached blocks back to the driver and retry once
    cudaGetLastError();
    pool_trim();
    e = cudaMalloc((void **)p, bytes);
  }
  if (e == cudaSuccess) pool_register(*p, ctx->device, bytes);
  if (e != cudaSuccess)
    return fail(ctx, e == cudaErrorMemoryAllocation ? EFB_ERR_NOMEM : EFB_ERR_CUDA,
                "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  return EFB_OK;
}
template <typename T>
int dev_upload(Ctx *ctx, T **p, const T *h, size_t n) {
  int rc = dev_alloc(ctx, p, n);
  if (rc) return rc;
  if (n) EFB_CUDA(ctx, cudaMemcpyAsync(*p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return EFB_OK;
}

struct Timed {  // records CUDA events around a compute call on the ctx stream
  Ctx *c;
  explicit Timed(Ctx *ctx) : c(ctx) { cudaEventRecord(c->ev0, c->stream); }
  ~Timed() {
    cudaEventRecord(c->ev1, c->stream);
    c->ev_valid = true;
  }
};

// EDGEFEM_B200_TRACE=2: host-side stage times inside the library calls (stderr)
struct SubTrace {
  bool on;
  std::chrono::steady_clock::time_point t;
  SubTrace() : on(false) {
    const char *e = getenv("EDGEFEM_B200_TRACE");
    on = e && atoi(e) >= 2;
    t = std::chrono::steady_clock::now();
  }
  void mark(const char *what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[edgefem-b200 trace]     . %s: %.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// internal entry points shared across translation units
int solver_free(System *s);
void cluster_plan_free(System *s);  // cluster.cu
int build_small_structs(System *s);  // abi.cu
// numbering.cu: large-mesh set-up on the device
bool device_setup_enabled(long long n_tet);
int device_e2t(Mesh *M, const int32_t *h_tet_edges, long long nt);
int mesh_host_e2t_item(Mesh *M);
int device_pattern(System *S, bool *done);
void dist_free(System *s);
int assemble_launch(System *s, int first, int count, int mode);
int launch_tet_geometry(Mesh *m);
int assemble_build_schedule(System *s);  // assemble.cu
bool asm_use_row_kernel();   // EDGEFEM_B200_ASM_KERNEL=row selects the older thread-per-row assembly kernel
void asm_chunk_limits(int *max_nnz, int *max_rows);

}  // namespace efb
