// K1 volume assembly (row-gather, no atomics), K6 K/M combine, K2 boundary terms, K5 port
// projections.  All FP64 / complex128 on the CUDA cores (nothing here is a dense
// contraction).  Reference semantics: src/assemble_maxwell.cpp:114-347,
// src/edge_basis.cpp:14-86, include/edgefem/materials/dispersive.hpp, src/sweep.cpp:82-172.
#include <algorithm>
#include <type_traits>

#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace efb {

constexpr double C0 = 299792458.0;  // src/assemble_maxwell.cpp:39

struct SlotMat {  // one per physical-tag slot, device copy of efb_materials
  c128 eps_s, mu_s;
  efb_model em, mm;
  efb_pml pml;
};

// ---------------------------------------------------------------- dispersive models
// include/edgefem/materials/dispersive.hpp:62-66 (Debye), :127-138 (Lorentz),
// :182-194 (Drude), :251-273 (Drude-Lorentz).  e^{+jwt} convention: loss => Im eps < 0.
__device__ c128 eval_model_eps(const efb_model &md, const efb_pole *__restrict__ poles, double w) {
  c128 eps;
  switch (md.kind) {
    case EFB_MODEL_DEBYE:
      eps = cdiv(cmake(md.p0 - md.p1, 0.0), cmake(1.0, w * md.p2));
      eps.x += md.p1;
      return eps;
    case EFB_MODEL_LORENTZ:
      eps = cmake(md.p0, 0.0);
      break;
    case EFB_MODEL_DRUDE:
      if (w == 0.0) return cmake(-1e30, 0.0);
      eps = cdiv(cmake(md.p0 * md.p0, 0.0), cmake(w * w, md.p1 * w));
      return cmake(1.0 - eps.x, -eps.y);
    case EFB_MODEL_DRUDE_LORENTZ:
      eps = cmake(md.p0, 0.0);
      if (w != 0.0) {
        c128 d = cdiv(cmake(md.p1 * md.p1, 0.0), cmake(w * w, md.p2 * w));
        eps = csub(eps, d);
      } else {
        eps = cmake(-1e30, 0.0);
      }
      break;
    default:
      return cmake(1.0, 0.0);
  }
  const double w2 = w * w;
  for (int k = 0; k < md.n_poles; ++k) {
    const efb_pole p = poles[md.pole_begin + k];
    const double w02 = p.omega0 * p.omega0;
    eps = cadd(eps, cdiv(cmake(p.delta_eps * w02, 0.0), cmake(w02 - w2, p.gamma * w)));
  }
  return eps;
}

// ---------------------------------------------------------------- element geometry
struct V3 {
  double x, y, z;
};
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 vcross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 vscale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }

// Frequency-independent geometry of one tet, cached on the device when the mesh is uploaded:
// Gram matrix gg[i][j] = grad(lambda_i).grad(lambda_j) and the volume V
// (src/edge_basis.cpp:14-26: gradients = columns of B^-T = cofactors/det, V = |det B|/6).
// Both element matrices are functions of (gg, V) only:
//   K(i,j) = V c_i.c_j, c_i = 2 g_a x g_b  ==  4V [ gg(a,c) gg(b,d) - gg(a,d) gg(b,c) ]   (Binet-Cauchy)
//   M(i,j) = gg(b,d) I(a,c) - gg(b,c) I(a,d) - gg(a,d) I(b,c) + gg(a,c) I(b,d),  I = V/10 (equal) | V/20
// so the per-frequency assembly never touches coordinates (except for tensor-PML tets).
__global__ void k_tet_geometry(const double4 *__restrict__ xyz, const int4 *__restrict__ tet_nodes,
                               const uint8_t *__restrict__ tet_sign, const uint8_t *__restrict__ tet_slot, int n_tet,
                               TetGeom *__restrict__ geom, TetRec *__restrict__ rec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tet) return;
  const int4 nd = tet_nodes[t];
  const double4 q0 = xyz[nd.x], q1 = xyz[nd.y], q2 = xyz[nd.z], q3 = xyz[nd.w];
  const V3 v0{q0.x, q0.y, q0.z}, v1{q1.x, q1.y, q1.z}, v2{q2.x, q2.y, q2.z}, v3{q3.x, q3.y, q3.z};
  const V3 b0 = vsub(v0, v3), b1 = vsub(v1, v3), b2 = vsub(v2, v3);
  const V3 c12 = vcross(b1, b2), c20 = vcross(b2, b0), c01 = vcross(b0, b1);
  const double det = vdot(b0, c12);
  const double inv = 1.0 / det;
  V3 g[4];
  g[0] = vscale(inv, c12);
  g[1] = vscale(inv, c20);
  g[2] = vscale(inv, c01);
  g[3] = {-g[0].x - g[1].x - g[2].x, -g[0].y - g[1].y - g[2].y, -g[0].z - g[1].z - g[2].z};
  TetGeom G;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      const double d = vdot(g[i], g[j]);
      G.gg[i][j] = d;
      G.gg[j][i] = d;
    }
  G.V = fabs(det) / 6.0;
  G.sign_slot = (uint32_t)tet_sign[t] | ((uint32_t)tet_slot[t] << 8);
  G.pad = 0;
  if (geom) geom[t] = G;
  TetRec R;
  int q = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) R.g[q++] = G.gg[i][j];
  R.V = G.V;
  R.Ieq = G.V / 10.0;
  rec[t] = R;
}

// sign|slot word of every (edge, tet) incidence, in the order of the incidence list (coalesced for the assembly)
__global__ void k_incidence_ss(const int32_t *__restrict__ e2t_item, const uint8_t *__restrict__ tet_sign,
                               const uint8_t *__restrict__ tet_slot, long long n_inc, uint16_t *__restrict__ ss) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inc) return;
  const int t = e2t_item[i] >> 3;
  ss[i] = (uint16_t)((unsigned)tet_sign[t] | ((unsigned)tet_slot[t] << 8));
}

int launch_tet_geometry(Mesh *M) {
  Ctx *c = M->ctx;
  if (M->n_tet == 0) return EFB_OK;
  k_tet_geometry<<<(M->n_tet + 127) / 128, 128, 0, c->stream>>>(M->d_xyz, M->d_tet_nodes, M->d_tet_sign, M->d_tet_slot, M->n_tet, M->d_geom, M->d_rec);
  EFB_CHECK_LAUNCH(c);
  const long long n_inc = 6ll * M->n_tet;
  k_incidence_ss<<<(unsigned)((n_inc + 255) / 256), 256, 0, c->stream>>>(M->d_e2t_item, M->d_tet_sign, M->d_tet_slot, n_inc, M->d_e2t_ss);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

__device__ __forceinline__ void prefetch_l2(const void *p) {  // a TetGeom record is 144 B: at most 2 lines
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)p + 128));
}

// PML scalar stretch of one tet (src/assemble_maxwell.cpp:121-172)
__device__ c128 pml_stretch(const efb_pml &pm, const double *__restrict__ bbox, V3 cen, double omega) {
  if (omega == 0.0 || pm.kind == EFB_PML_NONE) return cmake(1.0, 0.0);
  if (pm.kind == EFB_PML_UNIFORM) return cmake(1.0, pm.sigma[0] / omega);
  const double cc[3] = {cen.x, cen.y, cen.z};
  double im = 0.0;
#pragma unroll
  for (int ax = 0; ax < 3; ++ax) {
    const double smax = pm.sigma[ax], th = pm.thickness[ax];
    double sig = 0.0;
    if (smax > 0.0 && th > 0.0) {
      const double dmin = cc[ax] - bbox[ax], dmax = bbox[3 + ax] - cc[ax];
      const double md = fmin(dmin, dmax);
      double xi = 0.0;
      if (md < th) xi = 1.0 - md / th;
      if (pm.enforce_heuristics && xi > 0.0) xi = fmax(xi, 1e-3);
      sig = smax * pow(xi, pm.grading_order);
      if (pm.enforce_heuristics && sig < smax * 1e-3) sig = smax * 1e-3;
      im += sig / omega;
    }
  }
  return cmake(1.0, im / 3.0);  // mean of (1 + j sigma_ax / omega)
}

__device__ __noinline__ c128 pml_stretch_of_tet(const efb_pml &pm, const double *__restrict__ bbox, const double4 *__restrict__ xyz,
                                                const int4 *__restrict__ tet_nodes, int t, double omega) {
  const int4 nd = tet_nodes[t];
  const double4 q0 = xyz[nd.x], q1 = xyz[nd.y], q2 = xyz[nd.z], q3 = xyz[nd.w];
  const V3 cen{(q0.x + q1.x + q2.x + q3.x) / 4.0, (q0.y + q1.y + q2.y + q3.y) / 4.0, (q0.z + q1.z + q2.z + q3.z) / 4.0};
  return pml_stretch(pm, bbox, cen, omega);
}

// ---------------------------------------------------------------- K1: volume assembly
// (thread-per-row variant, EDGEFEM_B200_ASM_KERNEL=row; superseded by k_assemble_volume_g below)
// grid (n_chunks, count).  One CTA owns a contiguous row chunk (<= ASMR_CHUNK_NNZ entries): every
// thread walks the incident tets of its rows (ascending tet index => deterministic sums), forms
// the needed element-matrix row from the cached Gram record and accumulates into shared memory;
// the chunk is then written ONCE, fully coalesced, with the Dirichlet mask applied.  No atomics.
__global__ void __launch_bounds__(ASM_THREADS, 2)
k_assemble_volume(const TetGeom *__restrict__ geom, const double4 *__restrict__ xyz, const int4 *__restrict__ tet_nodes,
                  const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ e2t_item,
                  const uint16_t *__restrict__ e2t_pos, const int32_t *__restrict__ chunk_row,
                  const int32_t *__restrict__ rowptr, const int32_t *__restrict__ diag_pos,
                  const uint8_t *__restrict__ dir, const SlotMat *__restrict__ slots,
                  const efb_pole *__restrict__ poles, const double *__restrict__ slot_bbox,
                  const double *__restrict__ omegas, int n_slots, int mode, int first, long long nnz,
                  c128 *__restrict__ vals, long long pos_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // entry i of local row lr lives at acc[i + lr]: the +lr skew spreads the row starts of the 32
  // lanes of a warp (consecutive rows, ~16 entries = 64 words apart) over the shared-memory banks
  c128 *acc = (c128 *)smem_raw;                                   // [ASM_ACC_ENTRIES]
  c128 *s_kf = acc + ASM_ACC_ENTRIES;                             // [n_slots]
  c128 *s_mf = s_kf + n_slots;                                    // [n_slots]
  int32_t *s_rowptr = (int32_t *)(s_mf + n_slots);                // [ASMR_CHUNK_ROWS+1]
  uint16_t *s_rowid = (uint16_t *)(s_rowptr + ASMR_CHUNK_ROWS + 1); // [ASMR_CHUNK_NNZ] local row of every entry
  uint8_t *s_pml = (uint8_t *)(s_rowid + ASMR_CHUNK_NNZ);          // [n_slots]

  const int chunk = blockIdx.x, fi = blockIdx.y;
  const int r0 = chunk_row[chunk], r1 = chunk_row[chunk + 1];
  const int base = rowptr[r0];
  const int cnt = rowptr[r1] - base;
  const int nrow = r1 - r0;
  const double omega = omegas[fi];
  const double k0 = omega / C0;
  const double k0sq = k0 * k0;

  // get the first Gram record of this thread's row moving towards L2 while the CTA sets up
  if ((int)threadIdx.x < nrow) {
    const int r = r0 + threadIdx.x;
    const int kb = e2t_ptr[r];
    if (kb < e2t_ptr[r + 1]) prefetch_l2(geom + (e2t_item[kb] >> 3));
  }
  for (int i = threadIdx.x; i < cnt + nrow; i += blockDim.x) acc[i] = cmake(0.0, 0.0);
  for (int i = threadIdx.x; i <= nrow; i += blockDim.x) s_rowptr[i] = rowptr[r0 + i] - base;
  for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
    const SlotMat sm = slots[s];
    c128 eps = sm.eps_s, mu = sm.mu_s;
    if (mode == 0) {
      if (sm.em.kind != EFB_MODEL_NONE) eps = eval_model_eps(sm.em, poles, omega);
      // every shipped model's eval_mu is the base-class 1.0 (dispersive.hpp:28-31)
      if (sm.mm.kind != EFB_MODEL_NONE) mu = cmake(1.0, 0.0);
    }
    c128 kf, mf;
    if (mode == 0) {
      kf = cdiv(cmake(1.0, 0.0), mu);
      mf = cscale(-k0sq, eps);
    } else if (mode == 1) {
      kf = cdiv(cmake(1.0, 0.0), mu);
      mf = cmake(0.0, 0.0);
    } else {
      kf = cmake(0.0, 0.0);
      mf = eps;
    }
    s_kf[s] = kf;
    s_mf[s] = mf;
    s_pml[s] = (mode == 0) ? (uint8_t)sm.pml.kind : (uint8_t)0;
  }
  __syncthreads();

  const double diag_one = (mode == 2) ? 0.0 : 1.0;
  for (int lr = threadIdx.x; lr < nrow; lr += blockDim.x) {
    const int r = r0 + lr;
    const int rs = s_rowptr[lr], re = s_rowptr[lr + 1];
    for (int i = rs; i < re; ++i) s_rowid[i] = (uint16_t)lr;
    c128 *arow = acc + rs + lr;
    if (dir[r]) {  // Dirichlet row: zeros (kept in the pattern) and the unit diagonal
      const int dp = diag_pos[r];
      if (dp >= 0) arow[dp - (base + rs)] = cmake(diag_one, 0.0);
      continue;
    }
    const int kb = e2t_ptr[r], ke = e2t_ptr[r + 1];
    if (kb >= ke) continue;
    // the id of incidence k+2 is in flight and the Gram record of incidence k+1 is being pulled into L2
    // while incidence k is accumulated (no extra registers: prefetch, not load)
    int it_cur = e2t_item[kb];
    int it_nxt = (kb + 1 < ke) ? e2t_item[kb + 1] : -1;
    for (int k = kb; k < ke; ++k) {
      const int it_n2 = (k + 2 < ke) ? e2t_item[k + 2] : -1;
      if (it_nxt >= 0) prefetch_l2(geom + (it_nxt >> 3));
      const int item = it_cur;
      const int t = item >> 3, li = item & 7;
      const TetGeom *__restrict__ G = geom + t;
      const int a = (li < 3) ? 0 : (li < 5 ? 1 : 2);
      const int b = (li == 0) ? 1 : ((li == 1 || li == 3) ? 2 : 3);
      // rows a and b of the Gram matrix (32 B each), then (V, sign|slot)
      const double2 *ra2 = (const double2 *)G->gg[a], *rb2 = (const double2 *)G->gg[b];
      const double2 a01 = ra2[0], a23 = ra2[1], b01 = rb2[0], b23 = rb2[1];
      const double2 tail = *(const double2 *)&G->V;
      const uint16_t *pp = e2t_pos + ((size_t)k * 6 - pos_off);  // 6 x uint16 = 12 bytes, 4-byte aligned
      const uint32_t p01 = *(const uint32_t *)(pp), p23 = *(const uint32_t *)(pp + 2), p45 = *(const uint32_t *)(pp + 4);
      // bit 15 of a position: the column is a Dirichlet edge (entry stays an explicit zero)
      const int pos[6] = {(int)(p01 & 0xffff), (int)(p01 >> 16), (int)(p23 & 0xffff), (int)(p23 >> 16), (int)(p45 & 0xffff), (int)(p45 >> 16)};
      const double ra[4] = {a01.x, a01.y, a23.x, a23.y}, rb[4] = {b01.x, b01.y, b23.x, b23.y};
      const double V = tail.x;
      const unsigned packed = (unsigned)__double2loint(tail.y);
      const unsigned sg = packed & 0xffu;
      const int slot = (int)((packed >> 8) & 0xffu);
      c128 kf = s_kf[slot], mf = s_mf[slot];
      if (s_pml[slot]) {
        const c128 st = pml_stretch_of_tet(slots[slot].pml, slot_bbox + slot * 6, xyz, tet_nodes, t, omega);
        kf = cdiv(kf, st);
        mf = cmul(mf, st);
      }
      const double V4 = 4.0 * V, Ieq = V / 10.0, Ine = V / 20.0;
      const unsigned si = (sg >> li) & 1u;
      constexpr int PA[6] = {0, 0, 0, 1, 1, 2}, PB[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (pos[j] & 0x8000) continue;
        const int c = PA[j], d = PB[j];
        const double Kj = V4 * (ra[c] * rb[d] - ra[d] * rb[c]);
        double Mj = 0.0;
        Mj += rb[d] * ((a == c) ? Ieq : Ine);
        Mj -= rb[c] * ((a == d) ? Ieq : Ine);
        Mj -= ra[d] * ((b == c) ? Ieq : Ine);
        Mj += ra[c] * ((b == d) ? Ieq : Ine);
        const double sgn = (((sg >> j) & 1u) ^ si) ? -1.0 : 1.0;
        const double kk = Kj * sgn, mm = Mj * sgn;
        c128 v = arow[pos[j]];
        v.x += kk * kf.x + mm * mf.x;
        v.y += kk * kf.y + mm * mf.y;
        arow[pos[j]] = v;
      }
      it_cur = it_nxt;
      it_nxt = it_n2;
    }
  }
  __syncthreads();

  // write-out: a pure shared -> global copy, fully coalesced 16-byte stores (the Dirichlet mask was
  // applied while accumulating)
  c128 *__restrict__ out = vals + (size_t)(first + fi) * (size_t)nnz + base;
#pragma unroll 4
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) out[i] = acc[i + s_rowid[i]];
}

size_t assemble_smem_bytes(int n_slots) {
  return (size_t)ASM_ACC_ENTRIES * sizeof(c128) + 2 * (size_t)n_slots * sizeof(c128) + (ASMR_CHUNK_ROWS + 1) * sizeof(int32_t) +
         (size_t)ASMR_CHUNK_NNZ * sizeof(uint16_t) + (size_t)n_slots + 16;
}

// ---------------------------------------------------------------- K1 (batched): volume assembly
// Same row-gather scheme and the same summation order as k_assemble_volume (bit-identical sums), with the lanes of a
// warp mapped to 32 CONSECUTIVE (edge, tet) incidences of the row-sorted incidence list instead of to 32 rows:
//   * a warp owns a contiguous range of the chunk's rows (ranges balanced by incidence count) and walks the flat
//     incidence stream of those rows in batches of 32 -- every lane has work in every batch (no valence imbalance),
//     the tet ids, sign|slot words and position blocks of a batch are contiguous in memory (coalesced), and only the
//     96-byte tet records are gathered: three 256-bit loads per lane, each one whole 32-byte sector;
//   * every lane forms its 6 entries in registers; the adds into the shared-memory chunk image are then done in
//     rounds: in round r the r-th incidence (of this batch) of every row goes, so no two lanes of a round share a row
//     and the adds into an entry happen in ascending tet order -- deterministic, no atomics.
__host__ __device__ constexpr int asm_sym(int p, int q) {  // index of g(p,q) in TetRec::g (packed upper triangle)
  return (p <= q) ? (p * (9 - p)) / 2 + (q - p) : (q * (9 - q)) / 2 + (p - q);
}
__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ void load_record(const TetRec *__restrict__ rec, int item, double (&g)[12]) {
  const double *rp = (const double *)(rec + (item >> 3));
  ldg256(rp, g[0], g[1], g[2], g[3]);
  ldg256(rp + 4, g[4], g[5], g[6], g[7]);
  ldg256(rp + 8, g[8], g[9], g[10], g[11]);
}
template <bool PML, bool REAL, typename acc_t>
__device__ __forceinline__ void record_entries(const double (&g)[12], int item, unsigned ss, const c128 *s_kf, const c128 *s_mf,
                                               const uint8_t *s_pml, const SlotMat *__restrict__ slots,
                                               const double *__restrict__ slot_bbox, const double4 *__restrict__ xyz,
                                               const int4 *__restrict__ tet_nodes, double omega, acc_t (&v)[6]);

// The 6 entries of local row li of tet t (columns in local-edge order), combined with the slot's factors:
// v[j] = s_li s_j (K(li,j) kf + M(li,j) mf), K and M from the packed Gram record (see k_tet_geometry).
template <bool PML, bool REAL, typename acc_t>
__device__ __forceinline__ void incidence_entries(const TetRec *__restrict__ rec, int item, unsigned ss, const c128 *s_kf,
                                                  const c128 *s_mf, const uint8_t *s_pml, const SlotMat *__restrict__ slots,
                                                  const double *__restrict__ slot_bbox, const double4 *__restrict__ xyz,
                                                  const int4 *__restrict__ tet_nodes, double omega, acc_t (&v)[6]) {
  double g[12];
  load_record(rec, item, g);
  record_entries<PML, REAL>(g, item, ss, s_kf, s_mf, s_pml, slots, slot_bbox, xyz, tet_nodes, omega, v);
}

template <bool PML, bool REAL, typename acc_t>
__device__ __forceinline__ void record_entries(const double (&g)[12], int item, unsigned ss, const c128 *s_kf, const c128 *s_mf,
                                               const uint8_t *s_pml, const SlotMat *__restrict__ slots,
                                               const double *__restrict__ slot_bbox, const double4 *__restrict__ xyz,
                                               const int4 *__restrict__ tet_nodes, double omega, acc_t (&v)[6]) {
  constexpr int PA[6] = {0, 0, 0, 1, 1, 2}, PB[6] = {1, 2, 3, 2, 3, 3};
  const int t = item >> 3, li = item & 7;
  const double V = g[10], Ieq = g[11];
  const unsigned sg = ss & 0xffu;
  const int slot = (int)((ss >> 8) & 0xffu);
  c128 kf = s_kf[slot], mf = s_mf[slot];
  if (PML) {
    if (s_pml[slot]) {
      const c128 sv = pml_stretch_of_tet(slots[slot].pml, slot_bbox + slot * 6, xyz, tet_nodes, t, omega);
      kf = cdiv(kf, sv);
      mf = cmul(mf, sv);
    }
  }
  const int a = (li < 3) ? 0 : (li < 5 ? 1 : 2);
  const int b = (li == 0) ? 1 : ((li == 1 || li == 3) ? 2 : 3);
  // rows a and b of the Gram matrix out of the packed upper triangle
  double ra[4], rb4[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double xa = 0.0, xb = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const double gv = g[asm_sym(p, q)];
      if (a == p) xa = gv;
      if (b == p) xb = gv;
    }
    ra[q] = xa;
    rb4[q] = xb;
  }
  const double V4 = 4.0 * V, Ine = 0.5 * Ieq;
  const unsigned si = (sg >> li) & 1u;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = PA[j], d = PB[j];
    const double Kj = V4 * (ra[c] * rb4[d] - ra[d] * rb4[c]);
    double Mj = 0.0;
    Mj += rb4[d] * ((a == c) ? Ieq : Ine);
    Mj -= rb4[c] * ((a == d) ? Ieq : Ine);
    Mj -= ra[d] * ((b == c) ? Ieq : Ine);
    Mj += ra[c] * ((b == d) ? Ieq : Ine);
    const double sgn = (((sg >> j) & 1u) ^ si) ? -1.0 : 1.0;
    const double kk = Kj * sgn, mm = Mj * sgn;
    // explicit fma: the real and the complex image, and every kernel variant, round the same way
    if constexpr (REAL) v[j] = fma(kk, kf.x, mm * mf.x);
    else v[j] = cmake(fma(kk, kf.x, mm * mf.x), fma(kk, kf.y, mm * mf.y));
  }
}

// REAL: every slot's 1/mu and eps are real and no PML is present (lossless media, the common case): the chunk image
// holds doubles, the shared-memory adds move half the bytes, the imaginary parts are written as +0.0 -- exactly what
// the complex path produces for such materials.
template <bool PML, bool REAL>
__global__ void __launch_bounds__(ASMB_THREADS, ASMB_CTAS_PER_SM)
k_assemble_volume_b(const TetRec *__restrict__ rec, const double4 *__restrict__ xyz, const int4 *__restrict__ tet_nodes,
                    const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ e2t_item,
                    const uint16_t *__restrict__ e2t_ss, const uint16_t *__restrict__ e2t_pos,
                    const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowptr,
                    const int32_t *__restrict__ diag_pos, const uint8_t *__restrict__ dir,
                    const SlotMat *__restrict__ slots, const efb_pole *__restrict__ poles,
                    const double *__restrict__ slot_bbox, const double *__restrict__ omegas, int n_slots, int mode,
                    int first, long long nnz, c128 *__restrict__ vals, long long pos_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using acc_t = typename std::conditional<REAL, double, c128>::type;
  acc_t *acc = (acc_t *)smem_raw;                               // [ASM_CHUNK_NNZ]
  c128 *s_kf = (c128 *)(acc + ASM_CHUNK_NNZ);                   // [n_slots]
  c128 *s_mf = s_kf + n_slots;                                  // [n_slots]
  int32_t *s_rowptr = (int32_t *)(s_mf + n_slots);              // [ASM_CHUNK_ROWS+1]
  int32_t *s_inc = s_rowptr + ASM_CHUNK_ROWS + 1;               // [ASM_CHUNK_ROWS+1]
  uint8_t *s_dir = (uint8_t *)(s_inc + ASM_CHUNK_ROWS + 1);     // [ASM_CHUNK_ROWS]
  uint8_t *s_pml = s_dir + ASM_CHUNK_ROWS;                      // [n_slots]
  auto make_acc = [](double re) -> acc_t {
    if constexpr (REAL) return re; else return cmake(re, 0.0);
  };

  const int chunk = blockIdx.x, fi = blockIdx.y;
  const int r0 = chunk_row[chunk], r1 = chunk_row[chunk + 1];
  const int base = rowptr[r0];
  const int cnt = rowptr[r1] - base;
  const int nrow = r1 - r0;
  const double omega = omegas[fi];
  const double k0 = omega / C0;
  const double k0sq = k0 * k0;
  const int tid = threadIdx.x;

  for (int i = tid; i < cnt; i += ASMB_THREADS) acc[i] = make_acc(0.0);
  for (int i = tid; i <= nrow; i += ASMB_THREADS) {
    s_rowptr[i] = rowptr[r0 + i] - base;
    s_inc[i] = e2t_ptr[r0 + i];
  }
  for (int i = tid; i < nrow; i += ASMB_THREADS) s_dir[i] = dir[r0 + i];
  for (int s = tid; s < n_slots; s += ASMB_THREADS) {
    const SlotMat sm = slots[s];
    c128 eps = sm.eps_s, mu = sm.mu_s;
    if (mode == 0) {
      if (sm.em.kind != EFB_MODEL_NONE) eps = eval_model_eps(sm.em, poles, omega);
      // every shipped model's eval_mu is the base-class 1.0 (dispersive.hpp:28-31)
      if (sm.mm.kind != EFB_MODEL_NONE) mu = cmake(1.0, 0.0);
    }
    c128 kf, mf;
    if (mode == 0) {
      kf = cdiv(cmake(1.0, 0.0), mu);
      mf = cscale(-k0sq, eps);
    } else if (mode == 1) {
      kf = cdiv(cmake(1.0, 0.0), mu);
      mf = cmake(0.0, 0.0);
    } else {
      kf = cmake(0.0, 0.0);
      mf = eps;
    }
    s_kf[s] = kf;
    s_mf[s] = mf;
    if (PML) s_pml[s] = (mode == 0) ? (uint8_t)sm.pml.kind : (uint8_t)0;
  }
  __syncthreads();

  // Dirichlet rows: zeros (kept in the pattern) and the unit diagonal; no incidence of theirs is accumulated
  const double diag_one = (mode == 2) ? 0.0 : 1.0;
  for (int lr = tid; lr < nrow; lr += ASMB_THREADS) {
    if (s_dir[lr]) {
      const int dp = diag_pos[r0 + lr];
      if (dp >= 0) acc[dp - base] = make_acc(diag_one);
    }
  }

  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = ASMB_THREADS / 32;
  // row range of this warp: boundaries at equal shares of the chunk's incidences (whole rows)
  const int inc0 = s_inc[0], inc_total = s_inc[nrow] - inc0;
  int lr_a, lr_b;
  {
    int bound[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int wi = warp + q;
      if (wi >= NW) {
        bound[q] = nrow;
      } else {
        const int target = inc0 + (int)(((long long)inc_total * wi) / NW);
        int lo = 0, hi = nrow;  // first row with s_inc[row] >= target
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (s_inc[mid] < target) lo = mid + 1; else hi = mid;
        }
        bound[q] = lo;
      }
    }
    lr_a = bound[0];
    lr_b = bound[1];
  }
  const int k_begin = s_inc[lr_a], k_end = s_inc[lr_b];
  // does a row of this warp's range own no incidence (an edge outside every tet)?  Then two rows start at the same
  // incidence and the ballot look-up below cannot tell them apart: such warps use the binary search.
  bool any_empty = false;
  for (int r = lr_a + lane; r < lr_b; r += 32) any_empty |= (s_inc[r + 1] == s_inc[r]);
  any_empty = __any_sync(0xffffffffu, any_empty);
  int lr_prev = lr_a - 1;  // row of the last incidence before the current batch

  for (int kb = k_begin; kb < k_end; kb += 32) {
    const int k = kb + lane;
    const bool in = k < k_end;
    int item = 0;
    unsigned ss = 0, p01 = 0x80008000u, p23 = 0x80008000u, p45 = 0x80008000u;
    if (in) {
      item = __ldg(e2t_item + k);
      ss = __ldg(e2t_ss + k);
      const uint32_t *pp = (const uint32_t *)(e2t_pos + ((size_t)k * 6 - pos_off));  // 12 bytes, 4-byte aligned
      p01 = __ldg(pp);
      p23 = __ldg(pp + 1);
      p45 = __ldg(pp + 2);
    }
    // row of this incidence: last row of the warp's range whose first incidence is <= k
    int lr;
    if (!any_empty) {
      // rows starting inside this batch, as a bit mask over the lanes: lane t looks at row lr_prev + 1 + t
      const int cand = lr_prev + 1 + lane;
      unsigned bit = 0;
      if (cand < lr_b) {
        const int st = s_inc[cand] - kb;
        if (st < 32) bit = 1u << st;  // st >= 0: cand starts after the last incidence of the previous batch
      }
      const unsigned starts = __reduce_or_sync(0xffffffffu, bit);
      lr = lr_prev + __popc(starts & (0xffffffffu >> (31 - lane)));
      lr_prev += __popc(starts);
    } else {
      int lo = lr_a, hi = lr_b - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_inc[mid] <= k) lo = mid; else hi = mid - 1;
      }
      lr = lo;
    }
    const bool live = in && !s_dir[lr];
    // rank of this incidence among the ones of its row inside this batch (rows are contiguous runs of lanes)
    const int rb = lane - max(0, s_inc[lr] - kb);
    acc_t v[6];
    int pos[6] = {(int)(p01 & 0xffff), (int)(p01 >> 16), (int)(p23 & 0xffff), (int)(p23 >> 16), (int)(p45 & 0xffff), (int)(p45 >> 16)};
    if (live) {
      incidence_entries<PML, REAL>(rec, item, ss, s_kf, s_mf, s_pml, slots, slot_bbox, xyz, tet_nodes, omega, v);
    } else {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        v[j] = make_acc(0.0);
        pos[j] = 0x8000;
      }
    }
    // rounds: bit 15 of a position = the column is a Dirichlet edge (entry stays an explicit zero)
    acc_t *arow = acc + s_rowptr[lr];
    const int n_round = __reduce_max_sync(0xffffffffu, live ? rb + 1 : 0);
    for (int r = 0; r < n_round; ++r) {
      if (live && rb == r) {
        acc_t o[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) o[j] = arow[pos[j] & 0x7fff];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          if (!(pos[j] & 0x8000)) {
            if constexpr (REAL) arow[pos[j]] = o[j] + v[j];
            else arow[pos[j]] = cadd(o[j], v[j]);
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // write-out: a pure shared -> global copy, fully coalesced 16-byte streaming stores
  c128 *__restrict__ out = vals + (size_t)(first + fi) * (size_t)nnz + base;
#pragma unroll 4
  for (int i = tid; i < cnt; i += ASMB_THREADS) {
    if constexpr (REAL) __stcs(&out[i], cmake(acc[i], 0.0));
    else __stcs(&out[i], acc[i]);
  }
}

// ---------------------------------------------------------------- K1 (scheduled): volume assembly
// The batched kernel pays for its in-order adds with rounds in which only one lane per row is active, and a 16-byte
// shared-memory access costs one wavefront per quarter warp however few lanes are active -- the L1 data pipe, not HBM,
// bounds it (ncu: l1tex__data_pipe_lsu_wavefronts 85 % of peak).  Here the order of work is fixed once per system by
// k_build_schedule: inside a chunk, rank r holds the r-th incident tet of every non-Dirichlet row that has more than r
// of them, rows ascending.  The CTA walks the ranks with a barrier in between; inside a rank every thread takes one
// incidence and all of them belong to DIFFERENT rows, so all lanes add at once, conflict-free, and an entry still
// receives its tets in ascending order (bit-identical sums).  The schedule stream (item, sign|slot, row, positions:
// 20 B per incidence) is contiguous; only the 96-byte tet records are gathered.
__global__ void __launch_bounds__(512)
k_build_schedule(const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ e2t_item, const uint16_t *__restrict__ e2t_ss,
                 const uint16_t *__restrict__ e2t_pos, const int32_t *__restrict__ chunk_row, const uint8_t *__restrict__ dir,
                 long long k_off, int32_t *__restrict__ sch_item, uint16_t *__restrict__ sch_ss, uint16_t *__restrict__ sch_row,
                 uint16_t *__restrict__ sch_pos, int32_t *__restrict__ sch_sec, int32_t *__restrict__ flag) {
  typedef cub::BlockScan<int, 512> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int s_max;
  const int chunk = blockIdx.x, lr = threadIdx.x;
  const int r0 = chunk_row[chunk], nrow = chunk_row[chunk + 1] - r0;
  if (lr == 0) s_max = 0;
  __syncthreads();
  int val = 0, kbeg = 0;
  if (lr < nrow) {
    kbeg = e2t_ptr[r0 + lr];
    val = dir[r0 + lr] ? 0 : e2t_ptr[r0 + lr + 1] - kbeg;
  }
  atomicMax(&s_max, val);
  __syncthreads();
  const int maxv = s_max;
  int32_t *sec = sch_sec + (size_t)chunk * ASM_SEC_STRIDE;
  if (maxv > ASM_MAX_RANK) {
    if (lr == 0) {
      *flag = 1;
      sec[0] = 0;
    }
    return;
  }
  const long long gbase = (long long)e2t_ptr[r0] - k_off;  // the chunk's region of the schedule = its incidence region
  // Rows are taken in order of DESCENDING valence (stable): the rows with more than r incident tets are then a prefix
  // of that order for every r, so the thread that owns sorted row t in rank 0 owns it in every rank -- a row's adds all
  // come from one thread and the rank loop of the assembly kernel needs no barrier.
  // First touches: an off-diagonal entry of a row receives at most two tets (the two that share the face spanned by
  // the edge pair; one for opposite edges), so most adds are the FIRST contribution to their entry: bit 14 of a
  // position marks them and the kernel stores instead of load-add-store (60 % of the shared-memory loads of the adds).
  __shared__ int s_val[ASM_CHUNK_ROWS];
  if (lr < ASM_CHUNK_ROWS) s_val[lr] = lr < nrow ? val : -1;
  __syncthreads();
  int spos = 0;
  if (lr < nrow)
    for (int o = 0; o < nrow; ++o) spos += (s_val[o] > val) || (s_val[o] == val && o < lr);
  unsigned long long seen0 = 0ull, seen1 = 0ull;  // row-local positions < 128 already touched by an earlier rank
  int running = 0;
  for (int r = 0; r < maxv; ++r) {
    const int has = val > r ? 1 : 0;
    const int total = __syncthreads_count(has);
    if (has) {
      const long long k = (long long)kbeg + r, q = gbase + running + spos;
      sch_item[q] = e2t_item[k];
      sch_ss[q] = e2t_ss[k];
      sch_row[q] = (uint16_t)lr;
      const uint16_t *src = e2t_pos + (k - k_off) * 6;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        uint16_t pj = src[j];
        const unsigned pp = pj & 0x3fffu;
        if (!(pj & 0x8000u) && pp < 128u) {
          unsigned long long &sm = pp < 64u ? seen0 : seen1;
          const unsigned long long bit = 1ull << (pp & 63u);
          if (!(sm & bit)) pj |= 0x4000u;
          sm |= bit;
        }
        sch_pos[q * 6 + j] = pj;
      }
    }
    if (lr == 0) sec[1 + r] = running;
    running += total;
  }
  if (lr == 0) {
    sec[0] = maxv;
    sec[1 + maxv] = running;
  }
}

int assemble_build_schedule(System *S) {
  Ctx *c = S->ctx;
  Mesh *M = S->mesh;
  S->sched_ok = false;
  S->sched_dirty = false;
  static const bool off = getenv("EDGEFEM_B200_ASM_KERNEL") && strcmp(getenv("EDGEFEM_B200_ASM_KERNEL"), "batch") == 0;
  if (off || S->asm_row_kernel || !M || !S->d_e2t_pos || S->n_chunks == 0) return EFB_OK;
  const long long k_off = (long long)M->h_e2t_ptr[S->row0];
  const long long n_inc = (long long)M->h_e2t_ptr[S->row0 + S->m] - k_off;
  int rc;
  if (!S->d_sch_item) {
    if ((rc = dev_alloc(c, &S->d_sch_item, (size_t)std::max<long long>(n_inc, 1)))) return rc;
    if ((rc = dev_alloc(c, &S->d_sch_ss, (size_t)std::max<long long>(n_inc, 1)))) return rc;
    if ((rc = dev_alloc(c, &S->d_sch_row, (size_t)std::max<long long>(n_inc, 1)))) return rc;
    if ((rc = dev_alloc(c, &S->d_sch_pos, (size_t)std::max<long long>(n_inc * 6, 1)))) return rc;
    if ((rc = dev_alloc(c, &S->d_sch_sec, (size_t)S->n_chunks * ASM_SEC_STRIDE))) return rc;
    if ((rc = dev_alloc(c, &S->d_sch_flag, (size_t)1))) return rc;
  }
  EFB_CUDA(c, cudaMemsetAsync(S->d_sch_flag, 0, sizeof(int32_t), c->stream));
  k_build_schedule<<<(unsigned)S->n_chunks, 512, 0, c->stream>>>(M->d_e2t_ptr + S->row0, M->d_e2t_item, M->d_e2t_ss, S->d_e2t_pos,
                                                                  S->d_chunk_row, S->d_dir, k_off, S->d_sch_item, S->d_sch_ss,
                                                                  S->d_sch_row, S->d_sch_pos, S->d_sch_sec, S->d_sch_flag);
  EFB_CHECK_LAUNCH(c);
  int32_t flag = 0;
  EFB_CUDA(c, cudaMemcpyAsync(&flag, S->d_sch_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  S->sched_ok = flag == 0;
  return EFB_OK;
}

template <bool PML, bool REAL, int CTAS>
__global__ void __launch_bounds__(ASMB_THREADS, CTAS)
k_assemble_volume_s(const TetRec *__restrict__ rec, const double4 *__restrict__ xyz, const int4 *__restrict__ tet_nodes,
                    const int32_t *__restrict__ e2t_ptr, const int32_t *__restrict__ sch_item,
                    const uint16_t *__restrict__ sch_ss, const uint16_t *__restrict__ sch_row,
                    const uint16_t *__restrict__ sch_pos, const int32_t *__restrict__ sch_sec,
                    const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowptr,
                    const int32_t *__restrict__ diag_pos, const uint8_t *__restrict__ dir,
                    const SlotMat *__restrict__ slots, const efb_pole *__restrict__ poles,
                    const double *__restrict__ slot_bbox, const double *__restrict__ omegas, int n_slots, int mode,
                    int first, long long nnz, c128 *__restrict__ vals, long long k_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using acc_t = typename std::conditional<REAL, double, c128>::type;
  acc_t *acc = (acc_t *)smem_raw;                               // [ASM_CHUNK_NNZ]
  c128 *s_kf = (c128 *)(acc + ASM_CHUNK_NNZ);                   // [n_slots]
  c128 *s_mf = s_kf + n_slots;                                  // [n_slots]
  int32_t *s_rowptr = (int32_t *)(s_mf + n_slots);              // [ASM_CHUNK_ROWS+1]
  int32_t *s_sec = s_rowptr + ASM_CHUNK_ROWS + 1;               // [ASM_SEC_STRIDE]
  uint8_t *s_pml = (uint8_t *)(s_sec + ASM_SEC_STRIDE);         // [n_slots]
  auto make_acc = [](double re) -> acc_t {
    if constexpr (REAL) return re; else return cmake(re, 0.0);
  };

  const int chunk = blockIdx.x, fi = blockIdx.y;
  const int r0 = chunk_row[chunk], r1 = chunk_row[chunk + 1];
  const int base = rowptr[r0];
  const int cnt = rowptr[r1] - base;
  const int nrow = r1 - r0;
  const double omega = omegas[fi];
  const double k0 = omega / C0;
  const double k0sq = k0 * k0;
  const int tid = threadIdx.x;
  const long long gbase = (long long)e2t_ptr[r0] - k_off;

  for (int i = tid; i < cnt; i += ASMB_THREADS) acc[i] = make_acc(0.0);
  for (int i = tid; i <= nrow; i += ASMB_THREADS) s_rowptr[i] = rowptr[r0 + i] - base;
  if (tid < ASM_SEC_STRIDE) s_sec[tid] = sch_sec[(size_t)chunk * ASM_SEC_STRIDE + tid];
  for (int s = tid; s < n_slots; s += ASMB_THREADS) {
    const SlotMat sm = slots[s];
    c128 eps = sm.eps_s, mu = sm.mu_s;
    if (mode == 0) {
      if (sm.em.kind != EFB_MODEL_NONE) eps = eval_model_eps(sm.em, poles, omega);
      // every shipped model's eval_mu is the base-class 1.0 (dispersive.hpp:28-31)
      if (sm.mm.kind != EFB_MODEL_NONE) mu = cmake(1.0, 0.0);
    }
    c128 kf, mf;
    if (mode == 0) {
      kf = cdiv(cmake(1.0, 0.0), mu);
      mf = cscale(-k0sq, eps);
    } else if (mode == 1) {
      kf = cdiv(cmake(1.0, 0.0), mu);
      mf = cmake(0.0, 0.0);
    } else {
      kf = cmake(0.0, 0.0);
      mf = eps;
    }
    s_kf[s] = kf;
    s_mf[s] = mf;
    if (PML) s_pml[s] = (mode == 0) ? (uint8_t)sm.pml.kind : (uint8_t)0;
  }
  __syncthreads();

  // Dirichlet rows: zeros (kept in the pattern) and the unit diagonal; they are not in the schedule
  const double diag_one = (mode == 2) ? 0.0 : 1.0;
  for (int lr = tid; lr < nrow; lr += ASMB_THREADS) {
    if (dir[r0 + lr]) {
      const int dp = diag_pos[r0 + lr];
      if (dp >= 0) acc[dp - base] = make_acc(diag_one);
    }
  }

  // Rank loop, software-pipelined: a rank's section holds at most one incidence per thread (ASM_CHUNK_ROWS <=
  // ASMB_THREADS).  The schedule words run two ranks ahead and the record gather one rank ahead of the adds, so the
  // only thing between two barriers is arithmetic on data that is already in registers plus the shared-memory adds.
  struct Sched {
    int item;        // -1: this thread has no incidence in the rank
    unsigned ss_row; // sign|slot | row << 16
    unsigned p01, p23, p45;
  };
  const int n_rank = s_sec[0];
  auto load_sched = [&](int r) -> Sched {
    Sched w{-1, 0u, 0x80008000u, 0x80008000u, 0x80008000u};
    if (r < n_rank) {
      const int q = s_sec[1 + r] + tid;
      if (q < s_sec[2 + r]) {
        const long long gq = gbase + q;
        w.item = __ldg(sch_item + gq);
        w.ss_row = (unsigned)__ldg(sch_ss + gq) | ((unsigned)__ldg(sch_row + gq) << 16);
        const uint32_t *pp = (const uint32_t *)(sch_pos + gq * 6);  // 12 bytes, 4-byte aligned
        w.p01 = __ldg(pp);
        w.p23 = __ldg(pp + 1);
        w.p45 = __ldg(pp + 2);
      }
    }
    return w;
  };
  auto add_row = [&](const Sched &w, const acc_t (&v)[6]) {
    // every thread owns its row in all ranks; bit 15 of a position = Dirichlet column (stays zero), bit 14 = first
    // contribution to the entry (plain store: the image is zero there)
    const int pos[6] = {(int)(w.p01 & 0xffff), (int)(w.p01 >> 16), (int)(w.p23 & 0xffff), (int)(w.p23 >> 16),
                        (int)(w.p45 & 0xffff), (int)(w.p45 >> 16)};
    acc_t *arow = acc + s_rowptr[w.ss_row >> 16];
#pragma unroll
    for (int h = 0; h < 6; h += 3) {
      acc_t o[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // predicated loads (no branches: the lanes of a warp differ in their flags)
        o[j] = make_acc(0.0);
        if (!(pos[h + j] & 0xc000)) o[j] = arow[pos[h + j]];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!(pos[h + j] & 0x8000)) {
          if constexpr (REAL) arow[pos[h + j] & 0x3fff] = o[j] + v[h + j];
          else arow[pos[h + j] & 0x3fff] = cadd(o[j], v[h + j]);
        }
      }
    }
  };
  if constexpr (REAL) {
    Sched w0 = load_sched(0);
    Sched w1 = load_sched(1);
    double g[12];
    if (w0.item >= 0) load_record(rec, w0.item, g);
    for (int r = 0; r < n_rank; ++r) {
      const bool on = w0.item >= 0;
      acc_t v[6];
      if (on) record_entries<PML, REAL>(g, w0.item, w0.ss_row & 0xffffu, s_kf, s_mf, s_pml, slots, slot_bbox, xyz, tet_nodes, omega, v);
      // the record registers are free again: start the next rank's gather and the schedule words of the one after
      if (w1.item >= 0) load_record(rec, w1.item, g);
      const Sched w2 = load_sched(r + 2);
      if (on) add_row(w0, v);  // no barrier: the schedule gives a row to the same thread in every rank
      w0 = w1;
      w1 = w2;
    }
    __syncthreads();
  } else {
    // complex image: the pipelined form needs more registers than three resident CTAs leave (measured slower)
    for (int r = 0; r < n_rank; ++r) {
      const Sched w = load_sched(r);
      if (w.item >= 0) {
        acc_t v[6];
        incidence_entries<PML, REAL>(rec, w.item, w.ss_row & 0xffffu, s_kf, s_mf, s_pml, slots, slot_bbox, xyz, tet_nodes, omega, v);
        add_row(w, v);
      }
    }
    __syncthreads();
  }

  // write-out: a pure shared -> global copy, fully coalesced 16-byte streaming stores
  c128 *__restrict__ out = vals + (size_t)(first + fi) * (size_t)nnz + base;
#pragma unroll 4
  for (int i = tid; i < cnt; i += ASMB_THREADS) {
    if constexpr (REAL) __stcs(&out[i], cmake(acc[i], 0.0));
    else __stcs(&out[i], acc[i]);
  }
}

size_t assemble_b_smem_bytes(int n_slots) {
  return (size_t)ASM_CHUNK_NNZ * sizeof(c128) + 2 * (size_t)n_slots * sizeof(c128) + 2 * (size_t)(ASM_CHUNK_ROWS + 1) * sizeof(int32_t) +
         ASM_CHUNK_ROWS + (size_t)n_slots + 16;
}

bool asm_use_row_kernel() {
  static const bool v = [] {
    const char *e = getenv("EDGEFEM_B200_ASM_KERNEL");
    return e && strcmp(e, "row") == 0;
  }();
  return v;
}
void asm_chunk_limits(int *max_nnz, int *max_rows) {
  if (asm_use_row_kernel()) {
    *max_nnz = ASMR_CHUNK_NNZ;
    *max_rows = ASMR_CHUNK_ROWS;
  } else {
    *max_nnz = ASM_CHUNK_NNZ;
    *max_rows = ASM_CHUNK_ROWS;
  }
}

template <bool PML, bool REAL>
static int launch_assemble_b2(System *S, int first, int count, int mode, const unsigned char *blob, size_t off_poles, size_t off_om) {
  Ctx *c = S->ctx;
  Mesh *M = S->mesh;
  const int ns = M->n_slots;
  const size_t smem = assemble_b_smem_bytes(ns);
  if (S->sched_dirty) {
    int rcs = assemble_build_schedule(S);
    if (rcs) return rcs;
  }
  if (S->sched_ok) {
    dim3 grid((unsigned)S->n_chunks, (unsigned)count);
#define EFB_ASM_S(CTAS)                                                                                                   \
  EFB_CUDA(c, cudaFuncSetAttribute(k_assemble_volume_s<PML, REAL, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                   (int)assemble_b_smem_bytes(MAX_SLOTS)));                                               \
  k_assemble_volume_s<PML, REAL, CTAS><<<grid, ASMB_THREADS, smem, c->stream>>>(                                          \
      M->d_rec, M->d_xyz, M->d_tet_nodes, M->d_e2t_ptr + S->row0, S->d_sch_item, S->d_sch_ss, S->d_sch_row, S->d_sch_pos, \
      S->d_sch_sec, S->d_chunk_row, S->d_rowptr, S->d_diag_pos, S->d_dir, (const SlotMat *)blob,                          \
      (const efb_pole *)(blob + off_poles), M->d_slot_bbox, (const double *)(blob + off_om), ns, mode, first,             \
      (long long)S->nnz, S->d_vals, (long long)M->h_e2t_ptr[S->row0])
    // (4 resident CTAs at 64 registers spill in the rank loop: measured 0.54 ms against 0.36 ms on the 1.57 M-tet cube)
    EFB_ASM_S(ASMB_CTAS_PER_SM);
#undef EFB_ASM_S
    EFB_CHECK_LAUNCH(c);
    return EFB_OK;
  }
  EFB_CUDA(c, cudaFuncSetAttribute(k_assemble_volume_b<PML, REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assemble_b_smem_bytes(MAX_SLOTS)));
  dim3 grid((unsigned)S->n_chunks, (unsigned)count);
  const long long k_off = (long long)M->h_e2t_ptr[S->row0];
  k_assemble_volume_b<PML, REAL><<<grid, ASMB_THREADS, smem, c->stream>>>(
      M->d_rec, M->d_xyz, M->d_tet_nodes, M->d_e2t_ptr + S->row0, M->d_e2t_item, M->d_e2t_ss, S->d_e2t_pos, S->d_chunk_row,
      S->d_rowptr, S->d_diag_pos, S->d_dir, (const SlotMat *)blob, (const efb_pole *)(blob + off_poles), M->d_slot_bbox,
      (const double *)(blob + off_om), ns, mode, first, (long long)S->nnz, S->d_vals, k_off * 6);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}
static int launch_assemble_b(System *S, int first, int count, int mode, const unsigned char *blob, size_t off_poles, size_t off_om) {
  // the PML variant (per-tet stretch from the coordinates) only when a slot of this call carries a PML
  // and the real-image variant when every factor 1/mu, eps of this call is real
  bool any_pml = false, all_real = true;
  const SlotMat *hs = (const SlotMat *)S->last_mat_blob.data();
  for (int i = 0; i < S->mesh->n_slots; ++i) {
    if (mode == 0) any_pml |= hs[i].pml.kind != EFB_PML_NONE;
    all_real &= hs[i].eps_s.y == 0.0 && hs[i].mu_s.y == 0.0;
    if (mode == 0) all_real &= hs[i].em.kind == EFB_MODEL_NONE;  // a dispersive model evaluates to a complex eps
  }
  static const bool no_real = getenv("EDGEFEM_B200_ASM_NO_REAL") != nullptr;
  if (any_pml) return launch_assemble_b2<true, false>(S, first, count, mode, blob, off_poles, off_om);
  if (all_real && !no_real) return launch_assemble_b2<false, true>(S, first, count, mode, blob, off_poles, off_om);
  return launch_assemble_b2<false, false>(S, first, count, mode, blob, off_poles, off_om);
}

// blob layout: [SlotMat x n_slots][efb_pole x n_poles][double omega x count]
int assemble_launch(System *S, int first, int count, int mode) {
  Ctx *c = S->ctx;
  Mesh *M = S->mesh;
  const int ns = M->n_slots;
  const size_t off_poles = (size_t)ns * sizeof(SlotMat);
  const size_t total = S->last_mat_blob.size();
  const size_t off_om = total - (size_t)count * sizeof(double);
  if (S->mat_blob_bytes < total) {
    cudaStreamSynchronize(c->stream);
    dfree(S->d_mat_blob);
    S->d_mat_blob = nullptr;
    unsigned char *p = nullptr;
    int rc = dev_alloc(c, &p, total);
    if (rc) return rc;
    S->d_mat_blob = p;
    S->mat_blob_bytes = total;
  }
  EFB_CUDA(c, cudaMemcpyAsync(S->d_mat_blob, S->last_mat_blob.data(), total, cudaMemcpyHostToDevice, c->stream));
  const unsigned char *blob = (const unsigned char *)S->d_mat_blob;
  if (!S->asm_row_kernel) return launch_assemble_b(S, first, count, mode, blob, off_poles, off_om);
  if (!M->d_geom) return fail(c, EFB_ERR_STATE, "assembly: the mesh was uploaded without the thread-per-row geometry cache");
  const size_t smem = assemble_smem_bytes(ns);
  EFB_CUDA(c, cudaFuncSetAttribute(k_assemble_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assemble_smem_bytes(MAX_SLOTS)));
  dim3 grid((unsigned)S->n_chunks, (unsigned)count);
  k_assemble_volume<<<grid, ASM_THREADS, smem, c->stream>>>(
      M->d_geom, M->d_xyz, M->d_tet_nodes, M->d_e2t_ptr + S->row0, M->d_e2t_item, S->d_e2t_pos,
      S->d_chunk_row, S->d_rowptr, S->d_diag_pos, S->d_dir, (const SlotMat *)blob, (const efb_pole *)(blob + off_poles),
      M->d_slot_bbox, (const double *)(blob + off_om), ns, mode, first, (long long)S->nnz, S->d_vals,
      (long long)M->h_e2t_ptr[S->row0] * 6);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

// ---------------------------------------------------------------- small kernels
__device__ __forceinline__ int csr_find(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, int r, int c) {
  int lo = rowptr[r], hi = rowptr[r + 1];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = colidx[mid];
    if (v == c) return mid;
    if (v < c) lo = mid + 1; else hi = mid;
  }
  return -1;
}

__device__ __forceinline__ void atomic_cadd(c128 *p, c128 v) {
  atomicAdd(&p->x, v.x);
  atomicAdd(&p->y, v.y);
}

__global__ void k_combine_km(c128 *__restrict__ vals, long long nnz, int dst_first, int count,
                             const double *__restrict__ k0sq, int src_k, int src_m) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const c128 K = vals[(size_t)src_k * nnz + i], Mv = vals[(size_t)src_m * nnz + i];
  for (int f = 0; f < count; ++f) {
    const double q = k0sq[f];
    vals[(size_t)(dst_first + f) * nnz + i] = cmake(K.x - q * Mv.x, K.y - q * Mv.y);
  }
}

__global__ void k_add_diag(c128 *__restrict__ vals, long long nnz, int first, const int32_t *__restrict__ edges, int n,
                           const c128 *__restrict__ coef, const int32_t *__restrict__ diag_pos,
                           const uint8_t *__restrict__ dir, int32_t *flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
  if (k >= n) return;
  const int e = edges[k];
  if (dir[e]) return;
  const int p = diag_pos[e];
  if (p < 0) {
    atomicExch(flag, 1);
    return;
  }
  atomic_cadd(&vals[(size_t)(first + f) * nnz + p], coef[f]);
}

__global__ void k_port_prepare(const int32_t *__restrict__ edges, c128 *w, int n, const uint8_t *__restrict__ dir, c128 *e_dense) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = edges[k];
  c128 v = w[k];
  if (dir[e]) v = cmake(0.0, 0.0);
  w[k] = v;
  e_dense[e] = v;
}

__global__ void k_ms_pos(const int32_t *__restrict__ rows, const int32_t *__restrict__ cols, long long n,
                         const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, int32_t *pos) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  pos[i] = csr_find(rowptr, colidx, rows[i], cols[i]);
}

__global__ void k_port_block(c128 *__restrict__ vals, long long nnz, int first, const int32_t *__restrict__ edges,
                             const c128 *__restrict__ w, int n, const c128 *__restrict__ coef,
                             const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                             const uint8_t *__restrict__ dir, int32_t *flag) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)n * n) return;
  const int a = (int)(idx / n), b = (int)(idx % n);
  const int ea = edges[a], eb = edges[b];
  if (dir[ea] | dir[eb]) return;
  const int p = csr_find(rowptr, colidx, ea, eb);
  if (p < 0) {
    atomicExch(flag, 1);
    return;
  }
  const c128 wab = cmul(w[a], cconj(w[b]));
  atomic_cadd(&vals[(size_t)(first + blockIdx.y) * nnz + p], cmul(coef[blockIdx.y], wab));
}

__global__ void k_port_mass(c128 *__restrict__ vals, long long nnz, int first, const int32_t *__restrict__ pos,
                            const double *__restrict__ mv, long long n, const c128 *__restrict__ coef, int32_t *flag) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = pos[i];
  if (p < 0) {
    atomicExch(flag, 1);
    return;
  }
  atomic_cadd(&vals[(size_t)(first + blockIdx.y) * nnz + p], cscale(mv[i], coef[blockIdx.y]));
}

__global__ void k_rhs_weights(c128 *b, const int32_t *__restrict__ edges, const c128 *__restrict__ w, int n, c128 coef,
                              const uint8_t *__restrict__ dir) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = edges[k];
  if (dir[e]) return;
  atomic_cadd(&b[e], cmul(coef, w[k]));
}

__global__ void k_rhs_mass(c128 *b, const int32_t *__restrict__ rows, const int32_t *__restrict__ cols,
                           const double *__restrict__ mv, long long n, const c128 *__restrict__ e_dense, c128 coef) {
  // the entries are sorted by (row, column) (efb_port_create): the first entry of a row sums the row in order and is the only
  // writer of b[row] in this launch -- a fixed summation order instead of fp64 atomics, so b is bit-reproducible
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  if (i > 0 && rows[i - 1] == r) return;
  c128 acc = cmake(0.0, 0.0);
  for (long long k = i; k < n && rows[k] == r; ++k) acc = cadd(acc, cscale(mv[k], e_dense[cols[k]]));
  b[r] = cadd(b[r], cmul(coef, acc));
}

// single-CTA deterministic reductions (ports have O(100) edges)
__device__ c128 block_sum(c128 v) {
  __shared__ c128 sh[32];
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, o);
    v.y += __shfl_down_sync(0xffffffffu, v.y, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  c128 r = cmake(0.0, 0.0);
  if (wid == 0) {
    r = (lane < (blockDim.x + 31) / 32) ? sh[lane] : cmake(0.0, 0.0);
    for (int o = 16; o > 0; o >>= 1) {
      r.x += __shfl_down_sync(0xffffffffu, r.x, o);
      r.y += __shfl_down_sync(0xffffffffu, r.y, o);
    }
  }
  return r;  // valid in thread 0
}

__global__ void k_project_weights(const c128 *__restrict__ x, const int32_t *__restrict__ edges, const c128 *__restrict__ w,
                                  int n, const uint8_t *__restrict__ dir, c128 *out) {
  c128 acc = cmake(0.0, 0.0);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int e = edges[k];
    if (!dir[e]) acc = cadd(acc, cmulconj(w[k], x[e]));
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) *out = acc;
}

// out = sum conj(u[row]) * val * v[col]
__global__ void k_bilinear_mass(const c128 *__restrict__ u, const c128 *__restrict__ v, const int32_t *__restrict__ rows,
                                const int32_t *__restrict__ cols, const double *__restrict__ mv, long long n, c128 *out) {
  c128 acc = cmake(0.0, 0.0);
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = cadd(acc, cmulconj(u[rows[i]], cscale(mv[i], v[cols[i]])));
  acc = block_sum(acc);
  if (threadIdx.x == 0) *out = acc;
}

__global__ void k_port_scale(c128 *w, int n, c128 *e_dense, const int32_t *__restrict__ edges, const c128 *__restrict__ nsq) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double ns = nsq->x;
  if (!(ns > 1e-30)) return;
  const double s = 1.0 / sqrt(ns);
  const c128 v = cscale(s, w[k]);
  w[k] = v;
  e_dense[edges[k]] = v;
}

__global__ void k_x_recover(c128 *x, const int32_t *__restrict__ dst, const int32_t *__restrict__ src,
                            const c128 *__restrict__ phase, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  x[dst[k]] = cmul(phase[k], x[src[k]]);
}

__global__ void k_rhs_weights_batch(c128 *b, int m, const int32_t *__restrict__ rhs_idx, const c128 *__restrict__ coef,
                                    const int32_t *__restrict__ edges, const c128 *__restrict__ w, int n, const uint8_t *__restrict__ dir) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int e = edges[k];
  if (dir[e]) return;
  atomic_cadd(&b[(size_t)rhs_idx[blockIdx.y] * m + e], cmul(coef[blockIdx.y], w[k]));
}

__global__ void k_rhs_mass_batch(c128 *b, int m, const int32_t *__restrict__ rhs_idx, const c128 *__restrict__ coef,
                                 const int32_t *__restrict__ rows, const int32_t *__restrict__ cols, const double *__restrict__ mv,
                                 long long n, const c128 *__restrict__ e_dense) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rows[i];
  if (i > 0 && rows[i - 1] == r) return;  // see k_rhs_mass
  c128 acc = cmake(0.0, 0.0);
  for (long long k = i; k < n && rows[k] == r; ++k) acc = cadd(acc, cscale(mv[k], e_dense[cols[k]]));
  c128 *dst = &b[(size_t)rhs_idx[blockIdx.y] * m + r];
  *dst = cadd(*dst, cmul(coef[blockIdx.y], acc));
}

// one CTA per right-hand side: V = sum conj(w) x (weights) or e^H M_s x (mass)
__global__ void k_project_batch(const c128 *__restrict__ x, int m, const int32_t *__restrict__ rhs_idx, c128 *out, int use_mass,
                                const int32_t *__restrict__ edges, const c128 *__restrict__ w, int n_edges,
                                const uint8_t *__restrict__ dir, const c128 *__restrict__ e_dense, const int32_t *__restrict__ rows,
                                const int32_t *__restrict__ cols, const double *__restrict__ mv, long long n_ms) {
  const c128 *xv = x + (size_t)rhs_idx[blockIdx.x] * m;
  c128 acc = cmake(0.0, 0.0);
  if (use_mass) {
    for (long long i = threadIdx.x; i < n_ms; i += blockDim.x) acc = cadd(acc, cmulconj(e_dense[rows[i]], cscale(mv[i], xv[cols[i]])));
  } else {
    for (int k = threadIdx.x; k < n_edges; k += blockDim.x) {
      const int e = edges[k];
      if (!dir[e]) acc = cadd(acc, cmulconj(w[k], xv[e]));
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

static int check_flag(System *S, const char *who) {
  Ctx *c = S->ctx;
  int32_t h = 0;
  EFB_CUDA(c, cudaMemcpyAsync(&h, S->d_flag, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h) {
    EFB_CUDA(c, cudaMemsetAsync(S->d_flag, 0, sizeof h, c->stream));
    return fail(c, EFB_ERR_STATE, "%s: an entry is missing from the CSR pattern (pass it as an extra entry to efb_system_create)", who);
  }
  return EFB_OK;
}

// uploads `n` c128 to a temporary device buffer
struct TmpBuf {  // per-call scratch from the caching allocator; the kernels reading it have finished when it goes back
  void *p = nullptr;
  Ctx *c = nullptr;
  int alloc(Ctx *ctx, size_t bytes) {
    c = ctx;
    unsigned char *q = nullptr;
    int rc = dev_alloc(ctx, &q, bytes);
    p = q;
    return rc;
  }
  ~TmpBuf() {
    if (!p) return;
    cudaStreamSynchronize(c->stream);
    dfree(p);
  }
};

}  // namespace efb

using namespace efb;

extern "C" {

int efb_assemble_volume(efb_system *sys_, int32_t first, int32_t count, const double *omega, const efb_materials *mat, int32_t mode) {
  System *S = (System *)sys_;
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_assemble_volume: NULL system");
  Ctx *c = S->ctx;
  if (!S->mesh) return fail(c, EFB_ERR_STATE, "efb_assemble_volume: system was not created from a mesh");
  Mesh *M = S->mesh;
  if (!omega || !mat || first < 0 || count <= 0 || first + count > S->n_matrix || mode < 0 || mode > 2)
    return fail(c, EFB_ERR_INVALID, "efb_assemble_volume: bad arguments");
  if (mat->n_slots != M->n_slots || !mat->eps_static_c128 || !mat->mu_static_c128 || (mat->n_poles > 0 && !mat->poles))
    return fail(c, EFB_ERR_INVALID, "efb_assemble_volume: material table must have one entry per mesh slot (%d)", M->n_slots);
  EFB_CUDA(c, cudaSetDevice(c->device));
  const int ns = M->n_slots;
  std::vector<uint8_t> blob((size_t)ns * sizeof(SlotMat) + (size_t)std::max(0, mat->n_poles) * sizeof(efb_pole) + (size_t)count * sizeof(double));
  SlotMat *sm = (SlotMat *)blob.data();
  for (int s = 0; s < ns; ++s) {
    memset(&sm[s], 0, sizeof(SlotMat));
    sm[s].eps_s = cmake(mat->eps_static_c128[2 * s], mat->eps_static_c128[2 * s + 1]);
    sm[s].mu_s = cmake(mat->mu_static_c128[2 * s], mat->mu_static_c128[2 * s + 1]);
    if (mat->eps_models) sm[s].em = mat->eps_models[s];
    if (mat->mu_models) sm[s].mm = mat->mu_models[s];
    if (mat->pml) sm[s].pml = mat->pml[s];
    for (const efb_model *md : {&sm[s].em, &sm[s].mm})
      if (md->n_poles < 0 || md->pole_begin < 0 || md->pole_begin + md->n_poles > std::max(0, mat->n_poles))
        return fail(c, EFB_ERR_INVALID, "efb_assemble_volume: pole range of slot %d out of bounds", s);
  }
  if (mat->n_poles > 0) memcpy(blob.data() + (size_t)ns * sizeof(SlotMat), mat->poles, (size_t)mat->n_poles * sizeof(efb_pole));
  memcpy(blob.data() + blob.size() - (size_t)count * sizeof(double), omega, (size_t)count * sizeof(double));
  S->last_mat_blob.swap(blob);
  S->last_omega.assign(omega, omega + count);
  S->last_mode = mode;
  Timed tm(c);
  int rc = assemble_launch(S, first, count, mode);
  if (rc) return rc;
  EFB_CUDA(c, cudaMemsetAsync(S->d_b + (size_t)first * S->n_rhs * S->m, 0, (size_t)count * S->n_rhs * S->m * sizeof(c128), c->stream));
  S->assembled = true;
  return EFB_OK;
}

int efb_combine_km(efb_system *sys_, int32_t dst_first, int32_t count, const double *k0sq, int32_t src_k, int32_t src_m) {
  System *S = (System *)sys_;
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_combine_km: NULL system");
  Ctx *c = S->ctx;
  if (!k0sq || count <= 0 || dst_first < 0 || dst_first + count > S->n_matrix || src_k < 0 || src_k >= S->n_matrix || src_m < 0 ||
      src_m >= S->n_matrix || (src_k >= dst_first && src_k < dst_first + count) || (src_m >= dst_first && src_m < dst_first + count))
    return fail(c, EFB_ERR_INVALID, "efb_combine_km: bad arguments (sources must lie outside the destination range)");
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf tb;
  if (int rc_ = tb.alloc(c, (size_t)count * sizeof(double))) return rc_;
  EFB_CUDA(c, cudaMemcpyAsync(tb.p, k0sq, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  {
    Timed tm(c);
    k_combine_km<<<(unsigned)((S->nnz + 255) / 256), 256, 0, c->stream>>>(S->d_vals, (long long)S->nnz, dst_first, count, (const double *)tb.p, src_k, src_m);
    EFB_CHECK_LAUNCH(c);
    EFB_CUDA(c, cudaMemsetAsync(S->d_b + (size_t)dst_first * S->n_rhs * S->m, 0, (size_t)count * S->n_rhs * S->m * sizeof(c128), c->stream));
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  S->assembled = true;
  return EFB_OK;
}

int efb_add_diag(efb_system *sys_, int32_t first, int32_t count, int32_t n, const int32_t *edges, const double *coef) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_add_diag");
  if (!S) return fail(nullptr, EFB_ERR_INVALID, "efb_add_diag: NULL system");
  Ctx *c = S->ctx;
  if (n < 0 || (n > 0 && !edges) || !coef || first < 0 || count <= 0 || first + count > S->n_matrix)
    return fail(c, EFB_ERR_INVALID, "efb_add_diag: bad arguments");
  if (n == 0) return EFB_OK;
  for (int k = 0; k < n; ++k)
    if (edges[k] < 0 || edges[k] >= S->m) return fail(c, EFB_ERR_INVALID, "efb_add_diag: edge out of range");
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf te, tc;
  if (int rc_ = te.alloc(c, (size_t)n * sizeof(int32_t))) return rc_;
  if (int rc_ = tc.alloc(c, (size_t)count * sizeof(c128))) return rc_;
  EFB_CUDA(c, cudaMemcpyAsync(te.p, edges, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaMemcpyAsync(tc.p, coef, (size_t)count * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)count);
  k_add_diag<<<grid, 128, 0, c->stream>>>(S->d_vals, (long long)S->nnz, first, (const int32_t *)te.p, n, (const c128 *)tc.p, S->d_diag_pos, S->d_dir, S->d_flag);
  EFB_CHECK_LAUNCH(c);
  return check_flag(S, "efb_add_diag");
}

// ------------------------------------------------------------------ ports
int efb_port_create(efb_system *sys_, int32_t n_edges, const int32_t *edges, const double *weights, int64_t n_ms,
                    const int32_t *ms_rows, const int32_t *ms_cols, const double *ms_vals, efb_port **out) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_port_create");
  if (!S || !out) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_create: NULL argument");
  Ctx *c = S->ctx;
  *out = nullptr;
  if (n_edges <= 0 || !edges || !weights || n_ms < 0 || (n_ms > 0 && (!ms_rows || !ms_cols || !ms_vals)))
    return fail(c, EFB_ERR_INVALID, "efb_port_create: bad arguments");
  for (int k = 0; k < n_edges; ++k)
    if (edges[k] < 0 || edges[k] >= S->m) return fail(c, EFB_ERR_INVALID, "efb_port_create: edge out of range");
  for (int64_t i = 0; i < n_ms; ++i)
    if (ms_rows[i] < 0 || ms_rows[i] >= S->m || ms_cols[i] < 0 || ms_cols[i] >= S->m)
      return fail(c, EFB_ERR_INVALID, "efb_port_create: M_s entry out of range");
  EFB_CUDA(c, cudaSetDevice(c->device));
  Port *P = new Port();
  P->sys = S;
  P->n_edges = n_edges;
  P->n_ms = n_ms;
  int rc;
  if ((rc = dev_upload(c, &P->d_edges, edges, (size_t)n_edges))) return rc;
  if ((rc = dev_upload(c, &P->d_w, (const c128 *)weights, (size_t)n_edges))) return rc;
  {
    // entries sorted by (row, column): the right-hand-side kernels sum a row in this order (k_rhs_mass)
    std::vector<int64_t> ord((size_t)n_ms);
    for (int64_t i = 0; i < n_ms; ++i) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b2) {
      return ms_rows[a] != ms_rows[b2] ? ms_rows[a] < ms_rows[b2] : ms_cols[a] < ms_cols[b2];
    });
    std::vector<int32_t> sr((size_t)n_ms), sc((size_t)n_ms);
    std::vector<double> sv((size_t)n_ms);
    for (int64_t i = 0; i < n_ms; ++i) {
      sr[i] = ms_rows[ord[i]];
      sc[i] = ms_cols[ord[i]];
      sv[i] = ms_vals[ord[i]];
    }
    if ((rc = dev_upload(c, &P->d_ms_row, sr.data(), (size_t)n_ms))) return rc;
    if ((rc = dev_upload(c, &P->d_ms_col, sc.data(), (size_t)n_ms))) return rc;
    if ((rc = dev_upload(c, &P->d_ms_val, sv.data(), (size_t)n_ms))) return rc;
    EFB_CUDA(c, cudaStreamSynchronize(c->stream));  // the uploads read the temporaries
  }
  if ((rc = dev_alloc(c, &P->d_ms_pos, (size_t)n_ms))) return rc;
  if ((rc = dev_alloc(c, &P->d_e, (size_t)S->m))) return rc;
  if ((rc = dev_alloc(c, &P->d_tmp, (size_t)4))) return rc;
  EFB_CUDA(c, cudaMemsetAsync(P->d_e, 0, (size_t)S->m * sizeof(c128), c->stream));
  k_port_prepare<<<(n_edges + 127) / 128, 128, 0, c->stream>>>(P->d_edges, P->d_w, n_edges, S->d_dir, P->d_e);
  EFB_CHECK_LAUNCH(c);
  if (n_ms > 0) {
    k_ms_pos<<<(unsigned)((n_ms + 127) / 128), 128, 0, c->stream>>>(P->d_ms_row, P->d_ms_col, (long long)n_ms, S->d_rowptr, S->d_colidx, P->d_ms_pos);
    EFB_CHECK_LAUNCH(c);
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  *out = (efb_port *)P;
  return EFB_OK;
}

void efb_port_destroy(efb_port *port_) {
  Port *P = (Port *)port_;
  if (!P) return;
  cudaSetDevice(P->sys->ctx->device);
  dfree(P->d_edges); dfree(P->d_w); dfree(P->d_ms_row); dfree(P->d_ms_col); dfree(P->d_ms_pos);
  dfree(P->d_ms_val); dfree(P->d_e); dfree(P->d_tmp); dfree(P->d_blk_pos);
  delete P;
}

int efb_port_normalize_mass(efb_port *port_, double *norm_sq) {
  Port *P = (Port *)port_;
  if (!P) return fail(nullptr, EFB_ERR_INVALID, "efb_port_normalize_mass: NULL port");
  System *S = P->sys;
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  k_bilinear_mass<<<1, 256, 0, c->stream>>>(P->d_e, P->d_e, P->d_ms_row, P->d_ms_col, P->d_ms_val, (long long)P->n_ms, P->d_tmp);
  EFB_CHECK_LAUNCH(c);
  k_port_scale<<<(P->n_edges + 127) / 128, 128, 0, c->stream>>>(P->d_w, P->n_edges, P->d_e, P->d_edges, P->d_tmp);
  EFB_CHECK_LAUNCH(c);
  c128 h;
  EFB_CUDA(c, cudaMemcpyAsync(&h, P->d_tmp, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (norm_sq) *norm_sq = h.x;
  return EFB_OK;
}

static int upload_coef(Ctx *c, TmpBuf &tb, const double *coef, int count) {
  if (int rc_ = tb.alloc(c, (size_t)count * sizeof(c128))) return rc_;
  EFB_CUDA(c, cudaMemcpyAsync(tb.p, coef, (size_t)count * sizeof(c128), cudaMemcpyHostToDevice, c->stream));
  return EFB_OK;
}

int efb_port_add_block(efb_system *sys_, efb_port *port_, int32_t first, int32_t count, const double *coef) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  if (!S || !P || P->sys != S || !coef || first < 0 || count <= 0 || first + count > S->n_matrix)
    return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_add_block: bad arguments");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf tc;
  int rc = upload_coef(c, tc, coef, count);
  if (rc) return rc;
  const long long nn = (long long)P->n_edges * P->n_edges;
  dim3 grid((unsigned)((nn + 127) / 128), (unsigned)count);
  k_port_block<<<grid, 128, 0, c->stream>>>(S->d_vals, (long long)S->nnz, first, P->d_edges, P->d_w, P->n_edges, (const c128 *)tc.p, S->d_rowptr, S->d_colidx, S->d_dir, S->d_flag);
  EFB_CHECK_LAUNCH(c);
  return check_flag(S, "efb_port_add_block");
}

int efb_port_add_mass(efb_system *sys_, efb_port *port_, int32_t first, int32_t count, const double *coef) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  if (!S || !P || P->sys != S || !coef || first < 0 || count <= 0 || first + count > S->n_matrix)
    return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_add_mass: bad arguments");
  Ctx *c = S->ctx;
  if (P->n_ms == 0) return EFB_OK;
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf tc;
  int rc = upload_coef(c, tc, coef, count);
  if (rc) return rc;
  dim3 grid((unsigned)((P->n_ms + 127) / 128), (unsigned)count);
  k_port_mass<<<grid, 128, 0, c->stream>>>(S->d_vals, (long long)S->nnz, first, P->d_ms_pos, P->d_ms_val, (long long)P->n_ms, (const c128 *)tc.p, S->d_flag);
  EFB_CHECK_LAUNCH(c);
  return check_flag(S, "efb_port_add_mass");
}

int efb_port_rhs_weights(efb_system *sys_, efb_port *port_, int32_t rhs, const double *coef) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  if (!S || !P || P->sys != S || !coef || rhs < 0 || rhs >= S->n_sys) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_rhs_weights: bad arguments");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  k_rhs_weights<<<(P->n_edges + 127) / 128, 128, 0, c->stream>>>(S->d_b + (size_t)rhs * S->m, P->d_edges, P->d_w, P->n_edges, cmake(coef[0], coef[1]), S->d_dir);
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

int efb_port_rhs_mass(efb_system *sys_, efb_port *port_, int32_t rhs, const double *coef) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  if (!S || !P || P->sys != S || !coef || rhs < 0 || rhs >= S->n_sys) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_rhs_mass: bad arguments");
  Ctx *c = S->ctx;
  if (P->n_ms == 0) return EFB_OK;
  EFB_CUDA(c, cudaSetDevice(c->device));
  k_rhs_mass<<<(unsigned)((P->n_ms + 127) / 128), 128, 0, c->stream>>>(S->d_b + (size_t)rhs * S->m, P->d_ms_row, P->d_ms_col, P->d_ms_val, (long long)P->n_ms, P->d_e, cmake(coef[0], coef[1]));
  EFB_CHECK_LAUNCH(c);
  return EFB_OK;
}

int efb_port_project_weights(efb_system *sys_, efb_port *port_, int32_t rhs, double *v) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  if (!S || !P || P->sys != S || !v || rhs < 0 || rhs >= S->n_sys) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_project_weights: bad arguments");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  k_project_weights<<<1, 256, 0, c->stream>>>(S->d_x + (size_t)rhs * S->m, P->d_edges, P->d_w, P->n_edges, S->d_dir, P->d_tmp + 1);
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cudaMemcpyAsync(v, P->d_tmp + 1, sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_port_project_mass(efb_system *sys_, efb_port *port_, int32_t rhs, double *v) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  if (!S || !P || P->sys != S || !v || rhs < 0 || rhs >= S->n_sys) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_port_project_mass: bad arguments");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  k_bilinear_mass<<<1, 256, 0, c->stream>>>(P->d_e, S->d_x + (size_t)rhs * S->m, P->d_ms_row, P->d_ms_col, P->d_ms_val, (long long)P->n_ms, P->d_tmp + 2);
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cudaMemcpyAsync(v, P->d_tmp + 2, sizeof(c128), cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

// ---- batched variants: one launch for many right-hand sides (sweeps: F frequencies x P ports)
static int batch_args(System *S, Port *P, int32_t count, const int32_t *rhs_idx, const char *who) {
  if (!S || !P || P->sys != S || count <= 0 || !rhs_idx) return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "%s: bad arguments", who);
  for (int i = 0; i < count; ++i)
    if (rhs_idx[i] < 0 || rhs_idx[i] >= S->n_sys) return fail(S->ctx, EFB_ERR_INVALID, "%s: rhs index out of range", who);
  return EFB_OK;
}

int efb_port_rhs_batch(efb_system *sys_, efb_port *port_, int32_t count, const int32_t *rhs_idx, const double *coef, int32_t use_mass) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  int rc = batch_args(S, P, count, rhs_idx, "efb_port_rhs_batch");
  if (rc) return rc;
  if (!coef) return fail(S->ctx, EFB_ERR_INVALID, "efb_port_rhs_batch: coef is NULL");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf ti, tc;
  if (int rc_ = ti.alloc(c, (size_t)count * 4)) return rc_;
  if (int rc_ = tc.alloc(c, (size_t)count * 16)) return rc_;
  EFB_CUDA(c, cudaMemcpyAsync(ti.p, rhs_idx, (size_t)count * 4, cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaMemcpyAsync(tc.p, coef, (size_t)count * 16, cudaMemcpyHostToDevice, c->stream));
  if (use_mass) {
    if (P->n_ms > 0) {
      dim3 g((unsigned)((P->n_ms + 127) / 128), (unsigned)count);
      k_rhs_mass_batch<<<g, 128, 0, c->stream>>>(S->d_b, S->m, (const int32_t *)ti.p, (const c128 *)tc.p, P->d_ms_row, P->d_ms_col, P->d_ms_val, (long long)P->n_ms, P->d_e);
      EFB_CHECK_LAUNCH(c);
    }
  } else {
    dim3 g((unsigned)((P->n_edges + 127) / 128), (unsigned)count);
    k_rhs_weights_batch<<<g, 128, 0, c->stream>>>(S->d_b, S->m, (const int32_t *)ti.p, (const c128 *)tc.p, P->d_edges, P->d_w, P->n_edges, S->d_dir);
    EFB_CHECK_LAUNCH(c);
  }
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_port_project_batch(efb_system *sys_, efb_port *port_, int32_t count, const int32_t *rhs_idx, double *out, int32_t use_mass) {
  System *S = (System *)sys_;
  Port *P = (Port *)port_;
  int rc = batch_args(S, P, count, rhs_idx, "efb_port_project_batch");
  if (rc) return rc;
  if (!out) return fail(S->ctx, EFB_ERR_INVALID, "efb_port_project_batch: out is NULL");
  Ctx *c = S->ctx;
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf ti, to;
  if (int rc_ = ti.alloc(c, (size_t)count * 4)) return rc_;
  if (int rc_ = to.alloc(c, (size_t)count * 16)) return rc_;
  EFB_CUDA(c, cudaMemcpyAsync(ti.p, rhs_idx, (size_t)count * 4, cudaMemcpyHostToDevice, c->stream));
  k_project_batch<<<(unsigned)count, 128, 0, c->stream>>>(S->d_x, S->m, (const int32_t *)ti.p, (c128 *)to.p, use_mass, P->d_edges, P->d_w, P->n_edges,
                                                          S->d_dir, P->d_e, P->d_ms_row, P->d_ms_col, P->d_ms_val, (long long)P->n_ms);
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cudaMemcpyAsync(out, to.p, (size_t)count * 16, cudaMemcpyDeviceToHost, c->stream));
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

int efb_x_recover(efb_system *sys_, int32_t rhs, int32_t n, const int32_t *dst, const int32_t *src, const double *phase) {
  System *S = (System *)sys_;
  EFB_WHOLE_ONLY(S, "efb_x_recover");
  if (!S || rhs < 0 || rhs >= S->n_sys || n < 0 || (n > 0 && (!dst || !src || !phase)))
    return fail(S ? S->ctx : nullptr, EFB_ERR_INVALID, "efb_x_recover: bad arguments");
  if (n == 0) return EFB_OK;
  Ctx *c = S->ctx;
  for (int k = 0; k < n; ++k)
    if (dst[k] < 0 || dst[k] >= S->m || src[k] < 0 || src[k] >= S->m) return fail(c, EFB_ERR_INVALID, "efb_x_recover: index out of range");
  EFB_CUDA(c, cudaSetDevice(c->device));
  TmpBuf td, ts, tp;
  if (int rc_ = td.alloc(c, (size_t)n * 4)) return rc_;
  if (int rc_ = ts.alloc(c, (size_t)n * 4)) return rc_;
  if (int rc_ = tp.alloc(c, (size_t)n * 16)) return rc_;
  EFB_CUDA(c, cudaMemcpyAsync(td.p, dst, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaMemcpyAsync(ts.p, src, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  EFB_CUDA(c, cudaMemcpyAsync(tp.p, phase, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
  k_x_recover<<<(n + 127) / 128, 128, 0, c->stream>>>(S->d_x + (size_t)rhs * S->m, (const int32_t *)td.p, (const int32_t *)ts.p, (const c128 *)tp.p, n);
  EFB_CHECK_LAUNCH(c);
  EFB_CUDA(c, cudaStreamSynchronize(c->stream));
  return EFB_OK;
}

}  // extern "C"
