/*
 * edgefem_b200.h -- C-ABI of the B200-native EdgeFEM frequency-domain solve hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++
 * or torch types.  The reference (jman4162/EdgeFEM) has no FFI of its own; the entry
 * points below are what its C++ functions for this path bind to when the Eigen CPU
 * arithmetic is replaced by the GPU.  Each group cites the reference code it replaces
 * (paths relative to the reference root).  The C++ host layer in include/edgefem/ (same
 * names/signatures as the reference headers) is the only intended caller; INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every call returns EFB_OK (0) or a negative efb_status; efb_last_error() gives text
 *   - complex numbers are interleaved (re, im) doubles ("c128")
 *   - indices are 0-based int32; node references are node *indices* (not Gmsh ids)
 *   - host buffers are caller-owned; handles are opaque and owned by the library
 *   - one efb_ctx per GPU; a ctx and its children are used from one thread at a time
 *   - no CPU fallback: every compute entry point fails with EFB_ERR_CUDA without a GPU
 */
#ifndef EDGEFEM_B200_H
#define EDGEFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFB_ABI_VERSION 1

typedef enum {
  EFB_OK = 0,
  EFB_ERR_INVALID = -1, /* bad argument */
  EFB_ERR_CUDA = -2,    /* CUDA runtime error / no device */
  EFB_ERR_NOMEM = -3,
  EFB_ERR_LIMIT = -4,   /* a structural limit was exceeded (row too long, too many tags) */
  EFB_ERR_STATE = -5    /* call order (e.g. solve before assemble) */
} efb_status;

typedef struct efb_ctx efb_ctx;
typedef struct efb_mesh efb_mesh;
typedef struct efb_system efb_system;
typedef struct efb_port efb_port;

/* ------------------------------------------------------------------ context */
int efb_abi_version(void);
int efb_device_count(void);
int efb_ctx_create(int device, efb_ctx **out);
void efb_ctx_destroy(efb_ctx *ctx);
const char *efb_last_error(const efb_ctx *ctx); /* ctx may be NULL: last global error */
int efb_ctx_sync(efb_ctx *ctx);
/* measurement helper: overwrite a 256 MB scratch buffer on the context's stream (evicts the 126 MB L2 between timed steps) */
int efb_l2_flush(efb_ctx *ctx);
/* time of the most recent efb_* compute call's kernels, ms (CUDA events on the ctx stream) */
double efb_last_kernel_ms(const efb_ctx *ctx);
/* number of kernels this ctx has launched since creation */
int64_t efb_launch_count(const efb_ctx *ctx);
/* CUDA-event stopwatch on the ctx stream (the stream every kernel of this ctx is launched on):
 * start records an event; stop records a second one, synchronises and returns the elapsed ms */
int efb_timer_start(efb_ctx *ctx);
int efb_timer_stop(efb_ctx *ctx, double *ms);

/* ------------------------------------------------------------------ mesh
 * Replaces the in-memory traversal of `Mesh` (include/edgefem/mesh.hpp:43-52) done by
 * assemble_maxwell (src/assemble_maxwell.cpp:114-204).  Edge numbering (a1,
 * src/mesh_gmsh.cpp:104-146) is an INPUT: tet_edges/tet_orient carry it bit-exactly.   */
typedef struct {
  int32_t n_node;
  const double *xyz;          /* [3*n_node] */
  int32_t n_tet;
  const int32_t *tet_nodes;   /* [4*n_tet] node indices */
  const int32_t *tet_edges;   /* [6*n_tet] global edge ids, local order (01,02,03,12,13,23) */
  const int8_t *tet_orient;   /* [6*n_tet] +1/-1 */
  const int32_t *tet_phys;    /* [n_tet] physical tag */
  int32_t n_edge;             /* m */
  const int32_t *edge_nodes;  /* [2*n_edge] node indices (n0,n1) of each global edge */
} efb_mesh_desc;

int efb_mesh_create(efb_ctx *ctx, const efb_mesh_desc *desc, efb_mesh **out);
void efb_mesh_destroy(efb_mesh *mesh);
/* distinct physical tags of the tets, ascending; the material "slots" used below */
int efb_mesh_num_slots(const efb_mesh *mesh);
int efb_mesh_get_slot_tags(const efb_mesh *mesh, int32_t *tags /* [n_slots] */);

/* ------------------------------------------------------------------ system
 * A batch of n_matrix matrices sharing one CSR pattern, each with n_rhs right-hand sides
 * and solutions.  Pattern = union of the 6x6 edge cliques of all tets (setFromTriplets,
 * src/assemble_maxwell.cpp:205) plus `extra` entries (coeffRef insertions: port blocks
 * :229, ABC / Dirichlet diagonals :263,:315,:345), row-major CSR, sorted columns.      */
int efb_system_create(efb_mesh *mesh, int64_t n_extra, const int32_t *extra_rows,
                      const int32_t *extra_cols, int32_t n_matrix, int32_t n_rhs,
                      efb_system **out);
/* generic system from a caller-supplied CSR matrix (solve_linear, src/solver.cpp:35);
 * vals may be NULL (set later with efb_system_set_values)                              */
int efb_system_create_csr(efb_ctx *ctx, int32_t m, int64_t nnz, const int32_t *rowptr,
                          const int32_t *colidx, const double *vals_c128, int32_t n_matrix,
                          int32_t n_rhs, efb_system **out);
void efb_system_destroy(efb_system *sys);
int efb_system_dims(const efb_system *sys, int32_t *m, int64_t *nnz, int32_t *n_matrix,
                    int32_t *n_rhs);
int efb_system_get_pattern(const efb_system *sys, int32_t *rowptr /* [m+1] */,
                           int32_t *colidx /* [nnz] */);
int efb_system_get_values(efb_system *sys, int32_t matrix, double *vals_c128 /* [nnz] */);
int efb_system_set_values(efb_system *sys, int32_t matrix, const double *vals_c128);
/* Dirichlet (PEC) edge flags, src/assemble_maxwell.cpp:322-347: rows+cols zeroed with the
 * entries kept, diagonal = 1, b = 0.  flags[m] (0/1).  Applied by efb_assemble_volume. */
int efb_system_set_dirichlet(efb_system *sys, const uint8_t *flags);
/* optional: discrete gradient for the auxiliary nodal-space preconditioner of a system
 * created with efb_system_create_csr (mesh-born systems have it already)               */
int efb_system_set_gradient(efb_system *sys, int32_t n_node, const int32_t *edge_nodes);

/* ------------------------------------------------------------------ materials / PML
 * Per-slot table (slot order = efb_mesh_get_slot_tags).  Static values are the
 * region-or-global resolution of MaxwellParams::get_eps_r(tag) (maxwell.hpp:100-103);
 * a dispersive model on a slot overrides the static value and is evaluated ON DEVICE
 * at each omega (materials/dispersive.hpp:62-66,127-138,182-194,251-273).              */
typedef enum { EFB_MODEL_NONE = 0, EFB_MODEL_DEBYE = 1, EFB_MODEL_LORENTZ = 2,
               EFB_MODEL_DRUDE = 3, EFB_MODEL_DRUDE_LORENTZ = 4 } efb_model_kind;

typedef struct {
  int32_t kind;       /* efb_model_kind */
  int32_t n_poles;    /* Lorentz poles (LORENTZ, DRUDE_LORENTZ) */
  int32_t pole_begin; /* index into the poles array */
  int32_t _pad;
  double p0;          /* DEBYE eps_s | LORENTZ eps_inf | DRUDE omega_p | DL eps_inf */
  double p1;          /* DEBYE eps_inf |               | DRUDE gamma   | DL omega_p */
  double p2;          /* DEBYE tau    |                |               | DL gamma_d */
} efb_model;

typedef struct { double delta_eps, omega0, gamma; } efb_pole;

typedef enum { EFB_PML_NONE = 0, EFB_PML_UNIFORM = 1, EFB_PML_TENSOR = 2 } efb_pml_kind;

typedef struct {
  int32_t kind;           /* efb_pml_kind; UNIFORM: s = 1 + j*sigma[0]/omega (assemble_maxwell.cpp:168-170) */
  int32_t enforce_heuristics;
  double sigma[3];        /* TENSOR: sigma_max per axis (assemble_maxwell.cpp:140-164) */
  double thickness[3];
  double grading_order;
} efb_pml;

typedef struct {
  int32_t n_slots;
  const double *eps_static_c128; /* [n_slots] */
  const double *mu_static_c128;  /* [n_slots] */
  const efb_model *eps_models;   /* [n_slots] or NULL */
  const efb_model *mu_models;    /* [n_slots] or NULL (eval_mu == 1 for every shipped model) */
  int32_t n_poles;
  const efb_pole *poles;
  const efb_pml *pml;            /* [n_slots] or NULL */
} efb_materials;

/* ------------------------------------------------------------------ assembly (K1, K6)
 * Volume term for matrices [first, first+count): per tet 6x6 K/(mu s) - k0^2 eps s M with
 * orientation signs, summed into the CSR pattern, Dirichlet mask applied
 * (src/assemble_maxwell.cpp:114-205,322-347; element matrices src/edge_basis.cpp:14-86).
 * mode: 0 = A(omega) ; 1 = K only (1/mu, static materials) ; 2 = M only (eps, static)
 * (src/sweep.cpp:90-172: K(e,e)=1 and M(e,e)=0 on Dirichlet edges).  omega[count].
 * Also zeroes the right-hand sides of those matrices.                                    */
int efb_assemble_volume(efb_system *sys, int32_t first, int32_t count, const double *omega,
                        const efb_materials *mat, int32_t mode);
/* K6: vals[dst] = vals[srcK] - k0sq * vals[srcM] (KMMatrices::combine, src/sweep.cpp:82-88) */
int efb_combine_km(efb_system *sys, int32_t dst_first, int32_t count, const double *k0sq,
                   int32_t src_k, int32_t src_m);
/* A[first+i](e,e) += coef[i] for e in edges (ABC src/assemble_maxwell.cpp:244-265; port ABC
 * :274-320).  Dirichlet edges are skipped on device.                                     */
int efb_add_diag(efb_system *sys, int32_t first, int32_t count, int32_t n,
                 const int32_t *edges, const double *coef_c128);

/* ------------------------------------------------------------------ ports (K2, K5)
 * Device-resident port: weights w on `edges` (zeroed on Dirichlet edges) and, optionally,
 * its surface mass matrix M_s as COO triplets (assemble_port_surface_mass,
 * src/ports/wave_port.cpp:549-585; duplicates are summed).                               */
int efb_port_create(efb_system *sys, int32_t n_edges, const int32_t *edges,
                    const double *weights_c128, int64_t n_ms, const int32_t *ms_rows,
                    const int32_t *ms_cols, const double *ms_vals, efb_port **out);
void efb_port_destroy(efb_port *port);
/* e <- e / sqrt(Re(e^H M_s e)) if that is > 1e-30 (src/assemble_maxwell.cpp:714-728);
 * returns the pre-normalisation norm_sq                                                  */
int efb_port_normalize_mass(efb_port *port, double *norm_sq);
/* A[first+i] += coef[i] * w w^H over non-Dirichlet port edges (src/assemble_maxwell.cpp:219-231) */
int efb_port_add_block(efb_system *sys, efb_port *port, int32_t first, int32_t count,
                       const double *coef_c128);
/* A[first+i] += coef[i] * M_s (src/assemble_maxwell.cpp:738-746) */
int efb_port_add_mass(efb_system *sys, efb_port *port, int32_t first, int32_t count,
                      const double *coef_c128);
/* b[rhs] += coef * w   (src/assemble_maxwell.cpp:233-241);  rhs = matrix*n_rhs + k */
int efb_port_rhs_weights(efb_system *sys, efb_port *port, int32_t rhs, const double *coef_c128);
/* b[rhs] += coef * M_s e (src/assemble_maxwell.cpp:749-752) */
int efb_port_rhs_mass(efb_system *sys, efb_port *port, int32_t rhs, const double *coef_c128);
/* V = sum conj(w_k) x(edge_k)  (src/assemble_maxwell.cpp:376-384) */
int efb_port_project_weights(efb_system *sys, efb_port *port, int32_t rhs, double *v_c128);
/* V = e^H M_s x (src/assemble_maxwell.cpp:774) */
int efb_port_project_mass(efb_system *sys, efb_port *port, int32_t rhs, double *v_c128);
/* batched forms for sweeps (one launch for `count` right-hand sides):
 * b[rhs_idx[i]] += coef[i] * (use_mass ? M_s e : w)   and   out[i] = use_mass ? e^H M_s x : w^H x */
int efb_port_rhs_batch(efb_system *sys, efb_port *port, int32_t count, const int32_t *rhs_idx,
                       const double *coef_c128, int32_t use_mass);
int efb_port_project_batch(efb_system *sys, efb_port *port, int32_t count, const int32_t *rhs_idx,
                           double *out_c128, int32_t use_mass);

/* ------------------------------------------------------------------ periodic (a16)
 * In-place Bloch elimination A <- T A T^H, b <- T b with T = I + sum phase_k e_m e_s^T,
 * then slave rows/cols -> identity, b[s] = 0 (src/assemble_maxwell.cpp:527-566).  The pair
 * targets must be in the pattern: pass efb_periodic_extra() entries to efb_system_create. */
int efb_periodic_extra(const efb_mesh *mesh, int64_t n_base_extra, const int32_t *base_rows,
                       const int32_t *base_cols, int32_t n_pairs, const int32_t *master,
                       const int32_t *slave, int64_t *n_out, int32_t *rows_out,
                       int32_t *cols_out); /* call with rows_out==NULL to size */
int efb_apply_periodic(efb_system *sys, int32_t first, int32_t count, int32_t n_pairs,
                       const int32_t *master, const int32_t *slave,
                       const double *phase_c128 /* [n_pairs]: phi*o_m*o_s */);

/* ------------------------------------------------------------------ rhs / solution */
int efb_rhs_zero(efb_system *sys, int32_t rhs);
int efb_rhs_set(efb_system *sys, int32_t rhs, const double *b_c128 /* [m] */);
int efb_rhs_get(efb_system *sys, int32_t rhs, double *b_c128);
int efb_x_get(efb_system *sys, int32_t rhs, double *x_c128);
int efb_x_set(efb_system *sys, int32_t rhs, const double *x_c128);
/* x[rhs][dst_k] = phase_k * x[rhs][src_k] (slave recovery, src/assemble_maxwell.cpp:607-613) */
int efb_x_recover(efb_system *sys, int32_t rhs, int32_t n, const int32_t *dst,
                  const int32_t *src, const double *phase_c128);

/* ------------------------------------------------------------------ solve (K3, K4)
 * Replaces solve_linear (src/solver.cpp:35-193).  Eigen's BiCGSTAB/IncompleteLUT/SparseLU
 * are not reproduced; the contract kept is SolveResult's: iterations, relative residual
 * ||b-Ax||/||b|| (TRUE residual, recomputed), converged flag.                           */
typedef enum { EFB_METHOD_AUTO = 0, EFB_METHOD_BICGSTAB = 1, EFB_METHOD_COCG = 2,
               EFB_METHOD_DIRECT = 3 /* reported by efb_solve_direct only */ } efb_method;
typedef enum { EFB_PRECOND_JACOBI = 0, EFB_PRECOND_AUX = 1 /* Jacobi + nodal gradient-space Jacobi */,
               EFB_PRECOND_NONE = 2 } efb_precond;

typedef struct {
  int32_t method;        /* efb_method; AUTO = COCG when symmetric_hint else BiCGSTAB */
  int32_t precond;       /* efb_precond */
  double tolerance;      /* relative residual (SolveOptions::tolerance, solver.hpp:20) */
  int32_t max_iterations;
  int32_t check_every;   /* host convergence poll interval in iterations (0 = default 32) */
  int32_t symmetric_hint;/* 1: A == A^T (complex symmetric) */
  int32_t zero_initial_guess; /* 1 (reference behaviour) or 0 to start from x */
  int32_t max_restarts;  /* true-residual restarts (default 3) */
  int32_t _pad;
} efb_solve_opts;

typedef struct {
  int32_t iters;
  int32_t converged;
  int32_t method;   /* efb_method actually used */
  int32_t precond;
  double residual;  /* true relative residual */
} efb_solve_result;

int efb_solve(efb_system *sys, int32_t first_matrix, int32_t n_matrix,
              const efb_solve_opts *opts, efb_solve_result *results /* [n_matrix*n_rhs] */);
/* Direct solve on the device: dense complex LU with partial pivoting over the free unknowns, all right-hand sides
 * of matrices [first, first+count).  The robust last resort behind solve_linear's contract (src/solver.cpp:11-33
 * `use_direct`, :55-80 "<method>->SparseLU" when the Krylov solver fails): used by the host layer when a Krylov solve
 * does not converge and the system has at most efb_solve_direct_limit() free unknowns (16 384 by default,
 * EDGEFEM_B200_DENSE_MAX); larger systems return EFB_ERR_LIMIT.  results: iters = 1, true relative residual. */
int efb_solve_direct(efb_system *sys, int32_t first_matrix, int32_t n_matrix, efb_solve_result *results);
int efb_solve_direct_limit(void);
/* drop every cached derived structure of the library (cluster-split plans shared between systems) */
void efb_clear_caches(void);
/* device time (ms, CUDA events on the ctx stream) of the persistent one-CTA-per-matrix COCG kernel of the
 * most recent efb_solve on this system, or -1 if that solve used the multi-kernel path */
int efb_system_last_solve_kernel_ms(efb_system *sys, double *ms);
/* shape of the cluster-split persistent solver launch of the most recent efb_solve on this system: CTAs per cluster
 * (0: that solve took another path), right-hand sides per job, resident clusters */
int efb_system_last_solve_shape(efb_system *sys, int32_t *cluster_ctas, int32_t *rhs_per_job, int32_t *n_clusters);

/* ------------------------------------------------------------------ diagnostics of the cluster split (host only, no GPU)
 * The plan that splits one small system over the CTAs of a thread-block cluster (free unknowns in reverse
 * Cuthill-McKee order, contiguous row bands, window/halo and nodal exchange lists; csrc/cluster_plan.hpp), built from
 * a CSR pattern, Dirichlet flags (or NULL), the edge end nodes (or NULL: no auxiliary space) and the flags of the rows
 * whose values are not all real (their ELL blocks store complex128, the others doubles).  `get` copies the
 * named index array, widened to int64, and returns its length (-1: unknown name); tests/test_cluster_plan.py runs
 * the kernel's algorithm on these arrays on the CPU. */
int efb_debug_cluster_plan_build(int32_t m, const int32_t *rowptr, const int32_t *colidx, const uint8_t *dir, int32_t n_node,
                                 const int32_t *edge_nodes, const uint8_t *row_complex /* [m] or NULL = every row */,
                                 int32_t cluster_ctas, void **plan);
void efb_debug_cluster_plan_free(void *plan);
int64_t efb_debug_cluster_plan_get(void *plan, const char *name, int64_t *buf, int64_t capacity);
/* y = A[matrix] x on device, host in/out (test + diagnostics) */
int efb_spmv_host(efb_system *sys, int32_t matrix, const double *x_c128, double *y_c128);

/* ------------------------------------------------------------------ benchmarking hooks
 * Run a kernel `reps` times on resident data and return the average ms (CUDA events on the
 * ctx stream).  which: 0 = SpMV (all matrices, all rhs), 1 = one BiCGSTAB iteration,
 * 2 = one COCG iteration, 3 = volume assembly (mode 0, last used materials/omegas),
 * 4 = FP64 FMA throughput probe (sm_count*8 CTAs x 256 threads x 8 chains x 4096 FMAs per launch). */
int efb_bench_kernel(efb_system *sys, int32_t which, int32_t reps, double *avg_ms);

/* ------------------------------------------------------------------ field post-processing
 * efb_huygens_eval: per surface triangle i (node INDICES tri_nodes[3i..], parent tet tri_tet[i], that tet's six global
 * edge ids tri_tet_edges[6i..]) the centroid r, the unit normal n pointing away from the parent tet, the area, and the
 * tangential E and H = curl E / (j omega mu0 mu_r) of the solution (x[rhs] of `sys`, or a host vector) at the centroid -- the body of the triangle loop
 * of extract_huygens_surface (src/post/huygens_surface.cpp:59-135) incl. its projection E - conj(E.n) n.
 * efb_stratton_chu: far field of the Love currents J = n x H, M = -n x E,
 *   E_far = (j k0 / 4 pi) sum_s [ Z0 (rhat x J) x rhat - rhat x M ] exp(-j k0 rhat.r_s) area_s
 * projected on theta_hat / phi_hat for n_dir directions (theta[i], phi[i]) -- src/post/ntf.cpp:86-203.           */
int efb_huygens_eval(efb_mesh *mesh, efb_system *sys /* or NULL */, int32_t rhs, const double *x_host_c128 /* [m] or NULL */,
                     int32_t n_tri, const int32_t *tri_nodes, const int32_t *tri_tet, const int32_t *tri_tet_edges, double omega, const double *mu_r_c128, double *r_out /* [3n] */,
                     double *n_out /* [3n] */, double *E_tan_c128 /* [3n] */, double *H_tan_c128 /* [3n] */, double *area_out /* [n] */);
int efb_stratton_chu(efb_ctx *ctx, int32_t n_s, const double *r, const double *n, const double *E_c128, const double *H_c128,
                     const double *area, int32_t n_dir, const double *theta, const double *phi, double k0,
                     double *e_theta_c128 /* [n_dir] */, double *e_phi_c128 /* [n_dir] */);

/* ------------------------------------------------------------------ edge numbering on the device
 * Replaces the sequential unordered_map walk of build_edges (src/mesh_gmsh.cpp:104-146) for large meshes, bit-exact:
 * first-seen ids over tets x (01,02,03,12,13,23) then tris x (01,12,20), orient = +1 iff conn[a] < conn[b],
 * edges[id] = (min, max) node ids.  conn arrays hold node IDS (as in the .msh file), 0 <= id < 2^32.  Call once to
 * size (`edges` NULL), or pass an `edges` buffer of edges_capacity pairs (6*n_tet + 3*n_tri always suffices).
 * Two radix sorts + a few passes over 6*n_tet + 3*n_tri keys. */
int efb_build_edges(efb_ctx *ctx, int64_t n_tet, const int64_t *tet_conn /* [4t] */, int64_t n_tri, const int64_t *tri_conn /* [3k] */,
                    int32_t *tet_edges /* [6t] */, int8_t *tet_orient /* [6t] */, int32_t *tri_edges /* [3k] */,
                    int8_t *tri_orient /* [3k] */, int64_t *n_edges, int64_t *edges /* [2*capacity] or NULL */, int64_t edges_capacity);

/* ------------------------------------------------------------------ row-partitioned single large system
 * One process per GPU.  Rank r owns rows [r*chunk, min(m,(r+1)*chunk)), chunk = ceil(m/world), of the global edge
 * space; every rank holds the whole mesh (efb_mesh_create) and creates its block with efb_system_create_rows, sets
 * the GLOBAL Dirichlet flags (efb_system_set_dirichlet, m_global bytes), assembles it with efb_assemble_volume (no
 * communication: the row-gather kernel only needs the tets incident to its rows) and fills its slice of the
 * right-hand side with efb_rhs_set (local length).  Replaces: the reference has no multi-GPU path; this is the
 * SpMV + Krylov part of solve_linear (src/solver.cpp:11-193) for matrices larger than one GPU should carry.
 *
 * efb_dist_unique_id / efb_dist_init wrap ncclGetUniqueId / ncclCommInitRank (libnccl.so.2 is resolved with dlopen at
 * the first call): rank 0 creates the 128-byte id, the caller ships it to the other ranks by any means
 * (torch.distributed broadcast, MPI, a file) and every rank calls efb_dist_init(ctx, rank, world, id).
 * efb_dist_solve: COCG + Jacobi or + the auxiliary-space preconditioner (EFB_PRECOND_AUX: every rank gathers G^T r for the nodes of its
 * own edges over the peers' exported residual, one more scalar all-reduce per iteration).  halo_mode 0: the SpMV kernel loads off-rank vector entries directly from the peers'
 * memory (CUDA IPC mappings over NVLink) -- no exchange collective; halo_mode 1: ncclAllGather of the vector + local
 * SpMV (the library baseline).  Two scalar all-reduces per iteration in both modes.  Every rank must make the same
 * sequence of efb_dist_* calls.  efb_x_get returns the local slice of the solution. */
int efb_system_create_rows(efb_mesh *mesh, int32_t row_begin, int32_t row_end, int32_t n_matrix, int32_t n_rhs, efb_system **out);
int efb_dist_unique_id(uint8_t *id128);
int efb_dist_init(efb_ctx *ctx, int32_t rank, int32_t world, const uint8_t *id128);
void efb_dist_finalize(efb_ctx *ctx);
int efb_dist_row_range(int32_t m, int32_t rank, int32_t world, int32_t *row_begin, int32_t *row_end);
int efb_dist_solve(efb_system *sys, const efb_solve_opts *opts, efb_solve_result *result, int32_t halo_mode);
/* which 0: distributed SpMV + fused epilogue alone; 1: one full COCG + Jacobi iteration (2 kernels + 2 all-reduces);
 * 2: one COCG + auxiliary-space iteration (4 kernels + 3 all-reduces) */
int efb_dist_bench(efb_system *sys, int32_t which, int32_t reps, int32_t halo_mode, double *avg_ms);

#ifdef __cplusplus
}
#endif
#endif /* EDGEFEM_B200_H */
