// Internals shared by the solver translation units (solve.cu, cluster.cu).  Not part of the public ABI.
#pragma once
#include "common.cuh"

namespace efb {

enum Scal { S_RHO0 = 0, S_RHO1 = 1, S_PQ = 2, S_RR = 3, S_BB = 4, S_R0V = 5, S_TS = 6, S_TT = 7, S_ALPHA = 8, S_OMEGA = 9 };
enum State { ST_ACTIVE = 0, ST_ITERS = 1, ST_CONV = 2, ST_REC = 3 };
enum Vecs { V_R = 0, V_P = 1, V_Q = 2, V_Z = 3, V_R0 = 4, V_T = 5, V_Y = 6, V_NUM = 7 };


struct SolveDev {  // kernel-side view of the solver workspace
  int m, n_rhs, n_node;
  long long nnz;
  const int32_t *rowptr, *colidx;
  const c128 *vals;
  const uint8_t *dir, *node_dir;
  const int2 *edge_nodes;
  const int32_t *n2e_ptr, *n2e_item;
  c128 *dinv, *linv, *w;
  c128 *scal;
  double *partial;
  unsigned *counter;
  int32_t *state;
  double tol2;
  int max_it;
};


template <int N>
__device__ __forceinline__ void block_allreduce(double (&v)[N], double *red /* smem [33*N] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double a = v[k];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[wid * N + k] = a;
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double a = (lane < nw) ? red[lane * N + k] : 0.0;
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) red[32 * N + k] = a;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = red[32 * N + k];
  __syncthreads();
}


struct SolvePlan {
  System *S;
  SolveDev D;
  int first_matrix, n_matrix, first_sys, n_sys;
  int method, precond;
  bool aux;
  c128 *vec[V_NUM];
  dim3 vgrid, ngrid;
};


// cluster.cu: persistent Krylov solver with one thread-block CLUSTER per matrix (matrix slice resident in shared memory)
int run_krylov_cluster(SolvePlan &P, const efb_solve_opts *o, bool zero_x, bool *ran);
void cluster_plan_free(System *S);
void cluster_plan_cache_clear();
// solve.cu: workspace allocation and the final true-residual pass, shared with dense.cu
int solver_alloc_public(System *S);
int true_residuals(System *S, int first_matrix, int n_matrix, double tol, efb_solve_result *results);

}  // namespace efb
