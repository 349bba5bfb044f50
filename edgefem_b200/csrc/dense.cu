// Direct solve of SMALL systems on the device: dense complex128 LU with partial pivoting over the free unknowns.
// The robust last resort behind solve_linear's contract (src/solver.cpp:11-33,55-80: `use_direct` -> SparseLU, and
// "<method>->SparseLU" when the Krylov solver fails): the reference always gets a solution from its direct fallback;
// here a Krylov failure on a system of up to EFB_DENSE_MAX free unknowns (16 384 = 4.3 GB of dense storage) is
// re-solved by this factorisation, on the GPU, with no CPU fallback.  Non-symmetric systems (non-real Bloch phase)
// that BiCGSTAB cannot handle take this path.
//
// Blocked right-looking LU, column-major, block width 32:
//   k_lu_panel   one CTA: pivot search (block arg-max), row swap inside the panel, scale, rank-1 updates of the panel
//   k_laswp      the panel's row interchanges applied to every other column
//   k_trsm       U12 = L11^-1 A12 (thread per column, L11 in shared memory)
//   k_gemm       A22 -= L21 U12 (64 x 64 tiles, 4 x 4 complex per thread, operands staged in shared memory)
// then blocked forward / backward substitution for all right-hand sides of the matrix at once.
#include <algorithm>
#include <cmath>

#include "solve_internal.cuh"

namespace efb {

constexpr int LU_NB = 32;

__global__ void k_dense_fill(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const c128 *__restrict__ vals,
                             const int32_t *__restrict__ comp, const int32_t *__restrict__ orig, int n, c128 *A, size_t lda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = orig[i];
  for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
    const int j = comp[colidx[k]];
    if (j >= 0) A[(size_t)i + (size_t)j * lda] = vals[k];
  }
}

__global__ void __launch_bounds__(1024) k_lu_panel(c128 *A, size_t lda, int n, int k0, int nb, int32_t *ipiv, int32_t *info) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ c128 s_row[LU_NB];
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nth >> 5;
  for (int j = 0; j < nb; ++j) {
    c128 *col = A + (size_t)(k0 + j) * lda;
    double bv = -1.0;
    int bi = k0 + j;
    for (int i = k0 + j + tid; i < n; i += nth) {
      const double v = cabs2(col[i]);
      if (v > bv) { bv = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
      bv = lane < nw ? s_val[lane] : -1.0;
      bi = lane < nw ? s_idx[lane] : k0 + j;
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        s_idx[0] = bi;
        ipiv[k0 + j] = bi;
        if (!(bv > 0.0) || !isfinite(bv)) atomicExch(info, k0 + j + 1);  // singular / non-finite pivot
      }
    }
    __syncthreads();
    const int p = s_idx[0];
    if (tid < nb && p != k0 + j) {
      c128 *a = A + (size_t)(k0 + tid) * lda;
      const c128 t = a[k0 + j];
      a[k0 + j] = a[p];
      a[p] = t;
    }
    __syncthreads();
    if (tid < nb) s_row[tid] = A[(size_t)(k0 + j) + (size_t)(k0 + tid) * lda];
    __syncthreads();
    const c128 piv = s_row[j];
    const bool ok = piv.x != 0.0 || piv.y != 0.0;
    const c128 inv = ok ? cdiv(cmake(1.0, 0.0), piv) : cmake(0.0, 0.0);
    for (int i = k0 + j + 1 + tid; i < n; i += nth) {
      const c128 l = cmul(col[i], inv);
      col[i] = l;
      for (int jj = j + 1; jj < nb; ++jj) {
        c128 *c2 = A + (size_t)(k0 + jj) * lda;
        c2[i] = cfma(cneg(l), s_row[jj], c2[i]);
      }
    }
    __syncthreads();
  }
}

// row interchanges of the panel [k0, k0+nb) applied to the columns outside it
__global__ void k_laswp(c128 *A, size_t lda, int n, int k0, int nb, const int32_t *__restrict__ ipiv) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n - nb) return;
  if (c >= k0) c += nb;
  c128 *a = A + (size_t)c * lda;
  for (int j = 0; j < nb; ++j) {
    const int p = ipiv[k0 + j];
    if (p != k0 + j) {
      const c128 t = a[k0 + j];
      a[k0 + j] = a[p];
      a[p] = t;
    }
  }
}

// U12 = L11^-1 A12: thread per column right of the panel
__global__ void __launch_bounds__(128) k_trsm(c128 *A, size_t lda, int n, int k0, int nb) {
  __shared__ c128 L[LU_NB][LU_NB + 1];
  for (int i = threadIdx.x; i < LU_NB * LU_NB; i += blockDim.x) {
    const int r = i % LU_NB, cc = i / LU_NB;
    L[r][cc] = (r < nb && cc < nb) ? A[(size_t)(k0 + r) + (size_t)(k0 + cc) * lda] : cmake(0.0, 0.0);
  }
  __syncthreads();
  const int c = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  c128 *a = A + (size_t)c * lda + k0;
  c128 v[LU_NB];
#pragma unroll
  for (int j = 0; j < LU_NB; ++j) v[j] = j < nb ? a[j] : cmake(0.0, 0.0);
#pragma unroll
  for (int j = 1; j < LU_NB; ++j)
#pragma unroll
    for (int jj = 0; jj < j; ++jj) v[j] = cfma(cneg(L[j][jj]), v[jj], v[j]);
#pragma unroll
  for (int j = 0; j < LU_NB; ++j)
    if (j < nb) a[j] = v[j];
}

// A22 -= L21 U12 : tile 64 x 64 per CTA, 256 threads, 4 x 4 complex per thread
__global__ void __launch_bounds__(256) k_gemm(c128 *A, size_t lda, int n, int k0, int nb) {
  constexpr int KH = LU_NB / 2;  // the 32-deep product is staged in two halves (32 KB of static shared memory)
  __shared__ c128 Ls[KH][64];
  __shared__ c128 Us[KH][64];
  const int r0 = k0 + nb + blockIdx.x * 64, c0 = k0 + nb + blockIdx.y * 64;
  const int tid = threadIdx.x;
  const int rx = tid & 15, cx = tid >> 4;
  c128 acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = cmake(0.0, 0.0);
  for (int h = 0; h < 2; ++h) {
    if (h) __syncthreads();
    for (int i = tid; i < KH * 64; i += 256) {
      const int rr = i & 63, j = h * KH + (i >> 6);
      Ls[i >> 6][rr] = (j < nb && r0 + rr < n) ? A[(size_t)(r0 + rr) + (size_t)(k0 + j) * lda] : cmake(0.0, 0.0);
    }
    for (int i = tid; i < KH * 64; i += 256) {
      const int jl = i & (KH - 1), cc = i / KH, j = h * KH + jl;
      Us[jl][cc] = (j < nb && c0 + cc < n) ? A[(size_t)(k0 + j) + (size_t)(c0 + cc) * lda] : cmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < KH; ++j) {
      c128 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = Ls[j][rx * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = Us[j][cx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = cfma(a[u], b[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int c = c0 + cx * 4 + v;
    if (c >= n) continue;
    c128 *dst = A + (size_t)c * lda;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + rx * 4 + u;
      if (r < n) dst[r] = csub(dst[r], acc[u][v]);
    }
  }
}

// W[i][s] = b[s][orig[perm[i]]]   (row interchanges of the factorisation folded into one gather)
__global__ void k_rhs_gather(const c128 *__restrict__ b, int m, const int32_t *__restrict__ orig, const int32_t *__restrict__ perm, int n, int nrhs,
                             c128 *W, size_t ldw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = orig[perm[i]];
  for (int s = 0; s < nrhs; ++s) W[(size_t)i + (size_t)s * ldw] = b[(size_t)s * m + e];
}

// One block step of the substitutions for all right-hand sides.  LOWER: V[blk] = L11^-1 W[blk] (unit diagonal), then
// W[i] -= L[i, blk] V[blk] for the rows below; UPPER: V[blk] = U11^-1 W[blk], then W[i] -= U[i, blk] V[blk] above.
// Every CTA solves the 32 x 32 block redundantly (it only reads W[blk], final since the previous launch) and updates
// its own slice of rows; CTA 0 stores V[blk].
template <bool LOWER>
__global__ void __launch_bounds__(256) k_subst_step(const c128 *__restrict__ A, size_t lda, int n, int k0, int nb, c128 *W, c128 *V, size_t ldw,
                                                    int nrhs) {
  __shared__ c128 T[LU_NB][LU_NB + 1];
  extern __shared__ __align__(16) unsigned char dyn[];
  c128 *y = (c128 *)dyn;  // [nrhs][LU_NB]
  const int tid = threadIdx.x;
  for (int i = tid; i < LU_NB * LU_NB; i += blockDim.x) {
    const int r = i % LU_NB, cc = i / LU_NB;
    T[r][cc] = (r < nb && cc < nb) ? A[(size_t)(k0 + r) + (size_t)(k0 + cc) * lda] : cmake(0.0, 0.0);
  }
  for (int i = tid; i < nrhs * LU_NB; i += blockDim.x) {
    const int j = i % LU_NB, s = i / LU_NB;
    y[i] = j < nb ? W[(size_t)(k0 + j) + (size_t)s * ldw] : cmake(0.0, 0.0);
  }
  __syncthreads();
  for (int s = tid; s < nrhs; s += blockDim.x) {  // thread per right-hand side: serial substitution inside the block
    c128 *v = y + (size_t)s * LU_NB;
    if (LOWER) {
      for (int j = 1; j < nb; ++j) {
        c128 a = v[j];
        for (int jj = 0; jj < j; ++jj) a = cfma(cneg(T[j][jj]), v[jj], a);
        v[j] = a;
      }
    } else {
      for (int j = nb - 1; j >= 0; --j) {
        c128 a = v[j];
        for (int jj = j + 1; jj < nb; ++jj) a = cfma(cneg(T[j][jj]), v[jj], a);
        v[j] = cdiv(a, T[j][j]);
      }
    }
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int i = tid; i < nrhs * LU_NB; i += blockDim.x) {
      const int j = i % LU_NB, s = i / LU_NB;
      if (j < nb) V[(size_t)(k0 + j) + (size_t)s * ldw] = y[i];
    }
  const int lo = LOWER ? k0 + nb : 0, hi = LOWER ? n : k0;
  for (int i = lo + blockIdx.x * blockDim.x + tid; i < hi; i += gridDim.x * blockDim.x) {
    for (int s = 0; s < nrhs; ++s) {
      c128 a = W[(size_t)i + (size_t)s * ldw];
      const c128 *v = y + (size_t)s * LU_NB;
      for (int j = 0; j < nb; ++j) a = cfma(cneg(A[(size_t)i + (size_t)(k0 + j) * lda]), v[j], a);
      W[(size_t)i + (size_t)s * ldw] = a;
    }
  }
}

// x[s][orig[i]] = V[i][s] for the free unknowns; Dirichlet rows: x_e = b_e / A_ee
__global__ void k_x_scatter(const c128 *__restrict__ V, size_t ldw, const int32_t *__restrict__ orig, int n, int nrhs, c128 *x, int m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int s = 0; s < nrhs; ++s) x[(size_t)s * m + orig[i]] = V[(size_t)i + (size_t)s * ldw];
}
__global__ void k_x_dirichlet(const uint8_t *__restrict__ dir, const int32_t *__restrict__ diag_pos, const c128 *__restrict__ vals, const c128 *__restrict__ b,
                              c128 *x, int m, int nrhs) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m || !dir[e]) return;
  const int p = diag_pos[e];
  c128 d = cmake(1.0, 0.0);
  if (p >= 0 && (vals[p].x != 0.0 || vals[p].y != 0.0)) d = vals[p];
  for (int s = 0; s < nrhs; ++s) x[(size_t)s * m + e] = cdiv(b[(size_t)s * m + e], d);
}

static int dense_max_unknowns() {
  if (const char *e = getenv("EDGEFEM_B200_DENSE_MAX")) return std::max(0, atoi(e));
  return 16384;
}

}  // namespace efb

using namespace efb;

extern "C" {

int efb_solve_direct(efb_system *sys_, int32_t first_matrix, int32_t n_matrix, efb_solve_result *results) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_solve_direct");
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_solve_direct: NULL system");
  Ctx *c = S->ctx;
  if (!results || first_matrix < 0 || n_matrix <= 0 || first_matrix + n_matrix > S->n_matrix) return fail(c, EFB_ERR_INVALID, "efb_solve_direct: bad arguments");
  if (!S->assembled) return fail(c, EFB_ERR_STATE, "efb_solve_direct: matrix values were never assembled or set");
  if ((int)S->h_rowptr.size() != S->m + 1) return fail(c, EFB_ERR_STATE, "efb_solve_direct: the host copy of the pattern is missing");
  EFB_CUDA(c, cudaSetDevice(c->device));
  const int m = S->m, nrhs = S->n_rhs;
  const bool have_dir = (int)S->h_dir.size() == m;
  std::vector<int32_t> orig, comp((size_t)m, -1);
  for (int r = 0; r < m; ++r)
    if (!(have_dir && S->h_dir[r])) {
      comp[r] = (int32_t)orig.size();
      orig.push_back(r);
    }
  const int n = (int)orig.size();
  if (n > dense_max_unknowns())
    return fail(c, EFB_ERR_LIMIT, "efb_solve_direct: %d free unknowns exceed the dense-LU limit %d (EDGEFEM_B200_DENSE_MAX)", n, dense_max_unknowns());
  int rc = solver_alloc_public(S);
  if (rc) return rc;
  Timed tm(c);
  const size_t lda = (size_t)std::max(n, 1), ldw = lda;
  c128 *A = nullptr, *W = nullptr, *V = nullptr;
  int32_t *d_orig = nullptr, *d_comp = nullptr, *d_ipiv = nullptr, *d_perm = nullptr, *d_info = nullptr;
  auto cleanup = [&]() {
    cudaStreamSynchronize(c->stream);
    dfree(A); dfree(W); dfree(V); dfree(d_orig); dfree(d_comp); dfree(d_ipiv); dfree(d_perm); dfree(d_info);
  };
#define EFB_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc) {                 \
      cleanup();               \
      return _rc;              \
    }                          \
  } while (0)
  EFB_TRY(dev_alloc(c, &A, lda * (size_t)std::max(n, 1)));
  EFB_TRY(dev_alloc(c, &W, ldw * (size_t)nrhs));
  EFB_TRY(dev_alloc(c, &V, ldw * (size_t)nrhs));
  EFB_TRY(dev_upload(c, &d_orig, orig.data(), (size_t)std::max(n, 1)));
  EFB_TRY(dev_upload(c, &d_comp, comp.data(), (size_t)std::max(m, 1)));
  EFB_TRY(dev_alloc(c, &d_ipiv, (size_t)std::max(n, 1)));
  EFB_TRY(dev_alloc(c, &d_perm, (size_t)std::max(n, 1)));
  EFB_TRY(dev_alloc(c, &d_info, (size_t)1));
  const size_t subst_smem = (size_t)nrhs * LU_NB * sizeof(c128);
  if (subst_smem > 160 * 1024) {
    cleanup();
    return fail(c, EFB_ERR_LIMIT, "efb_solve_direct: too many right-hand sides per matrix (%d)", nrhs);
  }
  if (subst_smem > 40 * 1024) {
    cudaFuncSetAttribute(k_subst_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)subst_smem);
    cudaFuncSetAttribute(k_subst_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)subst_smem);
  }
  std::vector<int32_t> h_ipiv((size_t)std::max(n, 1)), perm((size_t)std::max(n, 1));
  for (int f = first_matrix; f < first_matrix + n_matrix; ++f) {
    const c128 *vals = S->d_vals + (size_t)f * S->nnz;
    const c128 *bsys = S->d_b + (size_t)f * nrhs * m;
    c128 *xsys = S->d_x + (size_t)f * nrhs * m;
    int32_t h_info = 0;
    if (n > 0) {
      cudaMemsetAsync(A, 0, lda * (size_t)n * sizeof(c128), c->stream);
      cudaMemsetAsync(d_info, 0, sizeof(int32_t), c->stream);
      k_dense_fill<<<(n + 127) / 128, 128, 0, c->stream>>>(S->d_rowptr, S->d_colidx, vals, d_comp, d_orig, n, A, lda);
      c->launches++;
      for (int k0 = 0; k0 < n; k0 += LU_NB) {
        const int nb = std::min(LU_NB, n - k0);
        const int rows = n - k0;
        const int pt = rows >= 4096 ? 1024 : rows >= 1024 ? 512 : 256;
        k_lu_panel<<<1, pt, 0, c->stream>>>(A, lda, n, k0, nb, d_ipiv, d_info);
        if (n - nb > 0) k_laswp<<<(n - nb + 127) / 128, 128, 0, c->stream>>>(A, lda, n, k0, nb, d_ipiv);
        const int rest = n - k0 - nb;
        if (rest > 0) {
          k_trsm<<<(rest + 127) / 128, 128, 0, c->stream>>>(A, lda, n, k0, nb);
          dim3 g((unsigned)((rest + 63) / 64), (unsigned)((rest + 63) / 64));
          k_gemm<<<g, 256, 0, c->stream>>>(A, lda, n, k0, nb);
          c->launches += 2;
        }
        c->launches += 2;
      }
      cudaMemcpyAsync(h_ipiv.data(), d_ipiv, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
      cudaMemcpyAsync(&h_info, d_info, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
      cudaError_t e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) {
        cleanup();
        return fail(c, EFB_ERR_CUDA, "efb_solve_direct: factorisation failed: %s", cudaGetErrorString(e));
      }
      for (int i = 0; i < n; ++i) perm[i] = i;
      for (int i = 0; i < n; ++i) std::swap(perm[i], perm[h_ipiv[i]]);
      cudaMemcpyAsync(d_perm, perm.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
      k_rhs_gather<<<(n + 127) / 128, 128, 0, c->stream>>>(bsys, m, d_orig, d_perm, n, nrhs, W, ldw);
      const int gb = std::max(1, std::min(c->sm_count * 2, (n + 255) / 256));
      for (int k0 = 0; k0 < n; k0 += LU_NB) {
        k_subst_step<true><<<gb, 256, subst_smem, c->stream>>>(A, lda, n, k0, std::min(LU_NB, n - k0), W, V, ldw, nrhs);
        c->launches++;
      }
      // backward substitution: the running vector is V (= y), the solution goes to W
      const int last = ((n - 1) / LU_NB) * LU_NB;
      for (int k0 = last; k0 >= 0; k0 -= LU_NB) {
        k_subst_step<false><<<gb, 256, subst_smem, c->stream>>>(A, lda, n, k0, std::min(LU_NB, n - k0), V, W, ldw, nrhs);
        c->launches++;
      }
      k_x_scatter<<<(n + 127) / 128, 128, 0, c->stream>>>(W, ldw, d_orig, n, nrhs, xsys, m);
      c->launches += 2;
    }
    if (n < m) {
      k_x_dirichlet<<<(m + 127) / 128, 128, 0, c->stream>>>(S->d_dir, S->d_diag_pos, vals, bsys, xsys, m, nrhs);
      c->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      cleanup();
      return fail(c, EFB_ERR_CUDA, "efb_solve_direct: kernel launch failed: %s", cudaGetErrorString(e));
    }
    for (int s = 0; s < nrhs; ++s) results[(size_t)(f - first_matrix) * nrhs + s].iters = h_info ? -h_info : 1;
  }
  cleanup();
#undef EFB_TRY
  // true residuals ||b - A x|| / ||b|| of every system
  rc = true_residuals(S, first_matrix, n_matrix, 1e-8, results);
  if (rc) return rc;
  for (int i = 0; i < n_matrix * nrhs; ++i) {
    results[i].method = EFB_METHOD_DIRECT;
    results[i].precond = EFB_PRECOND_NONE;
    if (results[i].iters < 0) {  // singular pivot reported by the factorisation
      results[i].converged = 0;
      results[i].iters = 0;
    } else {
      results[i].iters = 1;
    }
  }
  return EFB_OK;
}

int efb_solve_direct_limit(void) { return dense_max_unknowns(); }

}  // extern "C"
