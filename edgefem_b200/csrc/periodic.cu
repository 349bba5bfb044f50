// a16: Bloch-phase periodic elimination, sparse and in place on the device.
// Reference (src/assemble_maxwell.cpp:496-574) densifies the whole matrix and applies, pair by
// pair,  A[m,:] += phase*A[s,:] ; A[:,m] += conj(phi)*o*A[:,s] ; b[m] += phase*b[s],  then turns
// slave rows/cols into identity rows.  For pairs with pairwise distinct masters and slaves (what
// build_periodic_pairs produces) the per-pair transforms commute and equal
//     A' = T A T^H ,  b' = T b ,   T = I + sum_k phase_k e_{m_k} e_{s_k}^T
// Entries with a slave index are pure SOURCES, entries without one are the only TARGETS, so the
// scatter below can run in place: sources are read, targets receive atomic adds.
#include <algorithm>
#include <unordered_map>

#include "common.cuh"

namespace efb {

__device__ __forceinline__ int csr_find_p(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, int r, int c) {
  int lo = rowptr[r], hi = rowptr[r + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = colidx[mid];
    if (v == c) return mid;
    if (v < c) lo = mid + 1; else hi = mid;
  }
  return -1;
}

__device__ __forceinline__ int row_of(const int32_t *__restrict__ rowptr, int m, long long k) {
  int lo = 0, hi = m;  // last r with rowptr[r] <= k
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= k) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_periodic_scatter(c128 *__restrict__ vals, long long nnz, int first, int m, const int32_t *__restrict__ rowptr,
                                   const int32_t *__restrict__ colidx, const int32_t *__restrict__ slave_of,
                                   const int32_t *__restrict__ pair_master, const c128 *__restrict__ pair_phase, int32_t *flag) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int q = colidx[k];
  const int p = row_of(rowptr, m, k);
  const int sp = slave_of[p], sq = slave_of[q];
  if (sp < 0 && sq < 0) return;
  c128 *A = vals + (size_t)(first + blockIdx.y) * nnz;
  c128 v = A[k];
  if (v.x == 0.0 && v.y == 0.0) return;
  int tr = p, tc = q;
  if (sp >= 0) {
    v = cmul(pair_phase[sp], v);
    tr = pair_master[sp];
  }
  if (sq >= 0) {
    v = cmul(cconj(pair_phase[sq]), v);
    tc = pair_master[sq];
  }
  const int pos = csr_find_p(rowptr, colidx, tr, tc);
  if (pos < 0) {
    atomicExch(flag, 1);
    return;
  }
  atomicAdd(&A[pos].x, v.x);
  atomicAdd(&A[pos].y, v.y);
}

__global__ void k_periodic_identity(c128 *__restrict__ vals, long long nnz, int first, int m, const int32_t *__restrict__ rowptr,
                                    const int32_t *__restrict__ colidx, const int32_t *__restrict__ slave_of) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  const int q = colidx[k];
  const int p = row_of(rowptr, m, k);
  if (slave_of[p] < 0 && slave_of[q] < 0) return;
  vals[(size_t)(first + blockIdx.y) * nnz + k] = cmake(p == q ? 1.0 : 0.0, 0.0);
}

__global__ void k_periodic_rhs(c128 *__restrict__ b, int m, int first_sys, const int32_t *__restrict__ master,
                               const int32_t *__restrict__ slave, const c128 *__restrict__ phase, int n_pairs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  c128 *bv = b + (size_t)(first_sys + blockIdx.y) * m;
  const int s = slave[k], ms = master[k];
  bv[ms] = cfma(phase[k], bv[s], bv[ms]);  // masters are pairwise distinct: no race
  bv[s] = cmake(0.0, 0.0);
}

}  // namespace efb

using namespace efb;

extern "C" {

int efb_periodic_extra(const efb_mesh *mesh_, int64_t n_base, const int32_t *base_rows, const int32_t *base_cols, int32_t n_pairs,
                       const int32_t *master, const int32_t *slave, int64_t *n_out, int32_t *rows_out, int32_t *cols_out) {
  const Mesh *M = (const Mesh *)mesh_;
  if (!M || !n_out || n_pairs < 0 || (n_pairs > 0 && (!master || !slave)) || n_base < 0 || (n_base > 0 && (!base_rows || !base_cols)))
    return fail(M ? M->ctx : nullptr, EFB_ERR_INVALID, "efb_periodic_extra: bad arguments");
  std::unordered_map<int32_t, int32_t> s2m;
  for (int k = 0; k < n_pairs; ++k) {
    if (master[k] < 0 || master[k] >= M->m || slave[k] < 0 || slave[k] >= M->m) return fail(M->ctx, EFB_ERR_INVALID, "efb_periodic_extra: edge out of range");
    s2m[slave[k]] = master[k];
  }
  int64_t cnt = 0;
  auto emit = [&](int32_t r, int32_t c) {
    if (rows_out) {
      rows_out[cnt] = r;
      cols_out[cnt] = c;
    }
    ++cnt;
  };
  auto mapped = [&](int32_t e, int32_t &out) {
    auto it = s2m.find(e);
    if (it == s2m.end()) return false;
    out = it->second;
    return true;
  };
  const int64_t nt = M->n_tet;
  for (int64_t t = 0; t < nt; ++t) {
    const int32_t *e6 = M->h_tet_edges.data() + 6 * t;
    int32_t me[6];
    bool any = false, is[6];
    for (int i = 0; i < 6; ++i) {
      is[i] = mapped(e6[i], me[i]);
      any |= is[i];
    }
    if (!any) continue;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        if (!is[i] && !is[j]) continue;
        emit(is[i] ? me[i] : e6[i], is[j] ? me[j] : e6[j]);
      }
  }
  for (int64_t k = 0; k < n_base; ++k) {
    int32_t mr, mc;
    const bool ir = mapped(base_rows[k], mr), ic = mapped(base_cols[k], mc);
    if (ir || ic) emit(ir ? mr : base_rows[k], ic ? mc : base_cols[k]);
  }
  for (int k = 0; k < n_pairs; ++k) emit(slave[k], slave[k]);
  *n_out = cnt;
  return EFB_OK;
}

int efb_apply_periodic(efb_system *sys_, int32_t first, int32_t count, int32_t n_pairs, const int32_t *master, const int32_t *slave,
                       const double *phase) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_apply_periodic");
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_apply_periodic: NULL system");
  Ctx *c = S->ctx;
  if (first < 0 || count <= 0 || first + count > S->n_matrix || n_pairs < 0 || (n_pairs > 0 && (!master || !slave || !phase)))
    return fail(c, EFB_ERR_INVALID, "efb_apply_periodic: bad arguments");
  if (n_pairs == 0) return EFB_OK;
  std::vector<int32_t> slave_of((size_t)S->m, -1);
  std::vector<uint8_t> is_master((size_t)S->m, 0);
  for (int k = 0; k < n_pairs; ++k) {
    if (master[k] < 0 || master[k] >= S->m || slave[k] < 0 || slave[k] >= S->m) return fail(c, EFB_ERR_INVALID, "efb_apply_periodic: edge out of range");
    if (slave_of[slave[k]] >= 0 || is_master[master[k]])
      return fail(c, EFB_ERR_INVALID, "efb_apply_periodic: an edge appears in two pairs (chained constraints are not supported)");
    slave_of[slave[k]] = k;
    is_master[master[k]] = 1;
  }
  for (int k = 0; k < n_pairs; ++k)
    if (slave_of[master[k]] >= 0) return fail(c, EFB_ERR_INVALID, "efb_apply_periodic: an edge is both master and slave (chained constraints are not supported)");
  EFB_CUDA(c, cudaSetDevice(c->device));
  int32_t *d_slave_of = nullptr, *d_master = nullptr, *d_slave = nullptr;
  c128 *d_phase = nullptr;
  int rc;
  if ((rc = dev_upload(c, &d_slave_of, slave_of.data(), slave_of.size()))) return rc;
  if ((rc = dev_upload(c, &d_master, master, (size_t)n_pairs))) return rc;
  if ((rc = dev_upload(c, &d_slave, slave, (size_t)n_pairs))) return rc;
  if ((rc = dev_upload(c, &d_phase, (const c128 *)phase, (size_t)n_pairs))) return rc;
  {
    Timed tm(c);
    dim3 g((unsigned)((S->nnz + 255) / 256), (unsigned)count);
    k_periodic_scatter<<<g, 256, 0, c->stream>>>(S->d_vals, (long long)S->nnz, first, S->m, S->d_rowptr, S->d_colidx, d_slave_of, d_master, d_phase, S->d_flag);
    c->launches++;
    k_periodic_identity<<<g, 256, 0, c->stream>>>(S->d_vals, (long long)S->nnz, first, S->m, S->d_rowptr, S->d_colidx, d_slave_of);
    c->launches++;
    dim3 gr((unsigned)((n_pairs + 127) / 128), (unsigned)(count * S->n_rhs));
    k_periodic_rhs<<<gr, 128, 0, c->stream>>>(S->d_b, S->m, first * S->n_rhs, d_master, d_slave, d_phase, n_pairs);
    c->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  int32_t h = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&h, S->d_flag, sizeof h, cudaMemcpyDeviceToHost);
  dfree(d_slave_of); dfree(d_master); dfree(d_slave); dfree(d_phase);
  if (e != cudaSuccess) return fail(c, EFB_ERR_CUDA, "efb_apply_periodic: %s", cudaGetErrorString(e));
  if (h) {
    cudaMemset(S->d_flag, 0, sizeof h);
    return fail(c, EFB_ERR_STATE, "efb_apply_periodic: a target entry is missing from the pattern (pass efb_periodic_extra() entries to efb_system_create)");
  }
  return EFB_OK;
}

}  // extern "C"
