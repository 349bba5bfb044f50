// Host-side plan of the cluster-split persistent Krylov solver (cluster.cu): how ONE small system is split over the
// C CTAs of a thread-block cluster.  Pure host code (no CUDA calls) so that the index logic can be exercised without
// a GPU (efb_debug_cluster_plan_*, tests/test_cluster_plan.py emulates the kernel's algorithm on these arrays).
//
//  * free unknowns only (Dirichlet rows/columns dropped), renumbered by reverse Cuthill-McKee so that the matrix is
//    banded: "position" = RCM rank of a free edge;
//  * CTA c owns the contiguous positions [lo_c, hi_c) (balanced by entries) and keeps the values of those rows in
//    shared memory for the whole solve, in a thread-per-row ELL layout (rows sorted by length, 32-row blocks stored
//    column-major);
//  * the SpMV input of CTA c is the contiguous WINDOW [wlo_c, wlo_c + Wn_c) of positions its rows reference: own
//    entries plus a halo that is pulled from the owners' shared memory (DSMEM) once per iteration;
//  * auxiliary-space preconditioner: every CTA keeps G^T r for the nodes of its own edges ("my nodes"); partial
//    sums over own edges are exchanged through DSMEM (nsrc lists name every CTA that touches a node).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <numeric>
#include <queue>
#include <string>
#include <thread>
#include <vector>

namespace efb {

constexpr int CL_MAX_C = 8;       // portable cluster size
constexpr int CL_INFO_STRIDE = 20;
// cta_info[c][...]
enum ClInfo {
  CI_N_OWN = 0, CI_LO = 1, CI_WLO = 2, CI_WN = 3, CI_N_MY = 4, CI_N_HALO = 5, CI_N_BLK = 6, CI_N_SLOTS = 7,
  CI_OFF_ROW = 8, CI_OFF_SLOT = 9, CI_OFF_BLK = 10, CI_OFF_HALO = 11, CI_OFF_NODE = 12, CI_OFF_N2E = 13, CI_OFF_NSRC = 14,
  CI_N_PUSH = 15, CI_OFF_PUSH = 16
};

struct ClusterPlanHost {
  int C = 0, mc = 0, m = 0;
  bool aux = false;
  int max_own = 0, max_w = 0, max_my = 0, max_slots = 0, max_halo = 0, max_n2e = 0, max_nsrc = 0;
  int max_push = 0;    // most own rows (with multiplicity) a CTA sends into other CTAs' halos
  int max_vunits = 0;  // largest value image of a CTA in 8-byte units (real blocks 1 per slot, complex blocks 2)
  std::vector<int32_t> c_orig;     // [mc] position -> original edge id
  std::vector<int32_t> cta_info;   // [C][CL_INFO_STRIDE]
  // per own row, local order (sorted by length, descending), concatenated over CTAs at CI_OFF_ROW
  std::vector<int32_t> row_edge;   // original edge id
  std::vector<uint16_t> row_ws;    // window slot of the row itself
  std::vector<uint16_t> row_n0, row_n1;  // my-node slots of the tail / head node (aux only)
  std::vector<int32_t> blk_off;    // per CTA n_blk+1 slot offsets (multiples of 32), at CI_OFF_BLK
  std::vector<int32_t> blk_voff;   // same indexing: offsets of the blocks' VALUES in 8-byte units.  A block whose rows are all
                                   // real (row_complex == 0) stores one double per slot, the others a complex128:
                                   // voff[b+1] - voff[b] == 2 * (off[b+1] - off[b]) marks a complex block
  std::vector<int32_t> slot_src;   // CSR position of the slot's value (-1: padding), at CI_OFF_SLOT
  std::vector<uint16_t> slot_col;  // window slot of the slot's column
  std::vector<uint16_t> halo_ws;   // window slot of every halo entry, at CI_OFF_HALO
  std::vector<uint32_t> halo_src;  // owner CTA << 16 | owner local row
  // the transpose of the halo lists: what a CTA SENDS (the iteration pushes z into the halo staging of the reader before the
  // barrier instead of the reader pulling it over DSMEM after it), at CI_OFF_PUSH, CI_N_PUSH entries
  std::vector<uint16_t> push_row;  // own local row
  std::vector<uint32_t> push_dst;  // reader CTA << 16 | index of the entry in the reader's halo list
  std::vector<int32_t> node_id;    // [n_my] global node id, at CI_OFF_NODE
  std::vector<int32_t> n2e_ptr;    // [n_my+1] relative item offsets, at CI_OFF_NODE + c (one extra per CTA)
  std::vector<uint32_t> n2e_item;  // own local row << 1 | head, at CI_OFF_N2E
  std::vector<int32_t> nsrc_ptr;   // [n_my+1] relative, same placement as n2e_ptr
  std::vector<uint32_t> nsrc_item; // CTA << 16 | that CTA's my-node slot (ascending CTA, self included), at CI_OFF_NSRC
  std::string error;
  std::vector<std::pair<const char *, double>> timing;  // host milliseconds per build stage (EDGEFEM_B200_TRACE=2 prints them)
};

// reverse Cuthill-McKee over a symmetric pattern given as CSR adjacency without self loops
inline std::vector<int32_t> rcm_order(int n, const std::vector<int32_t> &ptr, const std::vector<int32_t> &adj) {
  std::vector<int32_t> order;
  order.reserve(n);
  std::vector<uint8_t> seen(n, 0);
  std::vector<int32_t> deg(n), level(n), q;
  for (int i = 0; i < n; ++i) deg[i] = ptr[i + 1] - ptr[i];
  std::vector<int32_t> by_deg(n);
  std::iota(by_deg.begin(), by_deg.end(), 0);
  std::stable_sort(by_deg.begin(), by_deg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
  auto bfs_far = [&](int start, std::vector<uint8_t> &mark) {  // farthest node of minimum degree from `start` (same component)
    q.clear();
    q.push_back(start);
    mark[start] = 1;
    level[start] = 0;
    size_t head = 0;
    while (head < q.size()) {
      const int v = q[head++];
      for (int k = ptr[v]; k < ptr[v + 1]; ++k) {
        const int w = adj[k];
        if (!mark[w]) {
          mark[w] = 1;
          level[w] = level[v] + 1;
          q.push_back(w);
        }
      }
    }
    const int last = level[q.back()];
    int best = q.back();
    for (size_t i = q.size(); i-- > 0 && level[q[i]] == last;)
      if (deg[q[i]] < deg[best]) best = q[i];
    for (int v : q) mark[v] = 0;
    return std::make_pair(best, last);
  };
  std::vector<uint8_t> mark(n, 0);
  std::vector<int32_t> nb;
  for (int s0 : by_deg) {
    if (seen[s0]) continue;
    // pseudo-peripheral start node: two sweeps of "go to the farthest node" (more sweeps rarely move it and each costs a
    // full BFS of host time on the solve path)
    int start = s0, ecc = -1;
    for (int it = 0; it < 2; ++it) {
      auto fr = bfs_far(start, mark);
      if (fr.second <= ecc) break;
      ecc = fr.second;
      start = fr.first;
    }
    const size_t first = order.size();
    order.push_back(start);
    seen[start] = 1;
    size_t head = first;
    while (head < order.size()) {
      const int v = order[head++];
      nb.clear();
      for (int k = ptr[v]; k < ptr[v + 1]; ++k)
        if (!seen[adj[k]]) {
          seen[adj[k]] = 1;
          nb.push_back(adj[k]);
        }
      std::stable_sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] < deg[b]; });
      order.insert(order.end(), nb.begin(), nb.end());
    }
  }
  std::reverse(order.begin(), order.end());
  return order;
}

// rowptr/colidx: m x m CSR pattern (sorted columns); dir: m flags or nullptr; edge_nodes: 2m node indices or nullptr
// row_complex (or NULL = every row): rows whose values are not all real.  A lossless system is real except for the rows of
// the port faces; storing the other rows as doubles halves the matrix slice, which is what decides how small a cluster can
// hold the system in shared memory (WR-90: 8 CTAs with complex values, 6 with mixed ones -> 24 instead of 15 resident).
inline bool build_cluster_plan(int m, const int32_t *rowptr, const int32_t *colidx, const uint8_t *dir, int n_node, const int32_t *edge_nodes,
                               int C, ClusterPlanHost &P, const uint8_t *row_complex = nullptr) {
  P = ClusterPlanHost();
  auto t_last = std::chrono::steady_clock::now();
  auto stage = [&](const char *name) {
    const auto t = std::chrono::steady_clock::now();
    P.timing.push_back({name, std::chrono::duration<double, std::milli>(t - t_last).count()});
    t_last = t;
  };
  P.C = C;
  P.m = m;
  P.aux = edge_nodes != nullptr && n_node > 0;
  if (C < 1 || C > CL_MAX_C) {
    P.error = "cluster size out of range";
    return false;
  }
  auto is_dir = [&](int e) { return dir && dir[e] != 0; };
  std::vector<int32_t> orig0, comp0((size_t)m, -1);
  for (int r = 0; r < m; ++r)
    if (!is_dir(r)) {
      comp0[r] = (int32_t)orig0.size();
      orig0.push_back(r);
    }
  const int mc = (int)orig0.size();
  P.mc = mc;
  if (mc == 0) {
    P.error = "no free unknowns";
    return false;
  }
  // free-free adjacency (no self loops) straight from the rows: FEM patterns are structurally symmetric; for a
  // non-symmetric pattern the order is merely less good (windows are computed from the actual rows below)
  std::vector<int32_t> aptr((size_t)mc + 1, 0), adj;
  adj.reserve((size_t)rowptr[m]);
  for (int i = 0; i < mc; ++i) {
    const int r = orig0[i];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int j = comp0[colidx[k]];
      if (j >= 0 && j != i) adj.push_back(j);
    }
    aptr[i + 1] = (int32_t)adj.size();
  }
  stage("adjacency");
  const std::vector<int32_t> rcm = rcm_order(mc, aptr, adj);  // position -> compact0 id
  stage("RCM order");
  std::vector<int32_t> pos_of((size_t)mc);
  P.c_orig.resize(mc);
  for (int p = 0; p < mc; ++p) {
    pos_of[rcm[p]] = p;
    P.c_orig[p] = orig0[rcm[p]];
  }
  // rows in position order: sorted free columns as positions (+ the CSR position of each)
  std::vector<int32_t> rptr((size_t)mc + 1, 0);
  for (int p = 0; p < mc; ++p) {
    const int r = P.c_orig[p];
    int n = 0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) n += comp0[colidx[k]] >= 0;
    rptr[p + 1] = rptr[p] + n;
  }
  std::vector<int32_t> rcol((size_t)rptr[mc]), rsrc((size_t)rptr[mc]);
  for (int p = 0; p < mc; ++p) {
    const int r = P.c_orig[p];
    int at = rptr[p];
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
      const int j = comp0[colidx[k]];
      if (j < 0) continue;
      rcol[at] = pos_of[j];
      rsrc[at] = k;
      ++at;
    }
  }
  stage("rows in position order");
  // contiguous partition balanced by (entries + 8) per row
  std::vector<int32_t> lo(C + 1, mc);
  {
    long long tot = 0;
    for (int p = 0; p < mc; ++p) tot += rptr[p + 1] - rptr[p] + 8;
    long long acc = 0;
    int c = 0;
    lo[0] = 0;
    for (int p = 0; p < mc; ++p) {
      while (c + 1 < C && acc * C >= tot * (c + 1)) lo[++c] = p;
      acc += rptr[p + 1] - rptr[p] + 8;
    }
    while (c + 1 < C) lo[++c] = mc;
    lo[C] = mc;
  }
  std::vector<int32_t> part_of((size_t)mc), local_of((size_t)mc);
  std::vector<std::vector<int32_t>> local_rows(C);  // local order -> position
  for (int c = 0; c < C; ++c) {
    auto &L = local_rows[c];
    for (int p = lo[c]; p < lo[c + 1]; ++p) {
      L.push_back(p);
      part_of[p] = c;
    }
    // complex rows first (their blocks come first: 16-byte alignment of the complex values), each group by length
    auto is_cplx = [&](int pos) { return row_complex ? (row_complex[P.c_orig[pos]] != 0) : true; };
    std::stable_sort(L.begin(), L.end(), [&](int a, int b) {
      const bool ca = is_cplx(a), cb = is_cplx(b);
      if (ca != cb) return ca;
      return rptr[a + 1] - rptr[a] > rptr[b + 1] - rptr[b];
    });
    for (size_t t = 0; t < L.size(); ++t) local_of[L[t]] = (int32_t)t;
  }
  // my nodes per CTA
  std::vector<std::vector<int32_t>> my_nodes(C);
  std::vector<std::vector<int32_t>> node_slot_of(C);  // sparse: via binary search in my_nodes
  if (P.aux)
    for (int c = 0; c < C; ++c) {
      auto &N = my_nodes[c];
      for (int p : local_rows[c]) {
        N.push_back(edge_nodes[2 * (size_t)P.c_orig[p]]);
        N.push_back(edge_nodes[2 * (size_t)P.c_orig[p] + 1]);
      }
      std::sort(N.begin(), N.end());
      N.erase(std::unique(N.begin(), N.end()), N.end());
    }
  auto slot_of_node = [&](int c, int node) {
    const auto &N = my_nodes[c];
    return (int)(std::lower_bound(N.begin(), N.end(), node) - N.begin());
  };
  // which CTAs touch a node (ascending)
  std::vector<std::vector<uint32_t>> touch;  // per node: list of cta<<16|slot
  if (P.aux) {
    touch.resize(n_node);
    for (int c = 0; c < C; ++c)
      for (size_t s = 0; s < my_nodes[c].size(); ++s) touch[my_nodes[c][s]].push_back((uint32_t)c << 16 | (uint32_t)s);
  }
  stage("partition, nodes, touch lists");
  std::vector<ClusterPlanHost> parts(C);
  auto build_cta = [&](int c) {
    ClusterPlanHost &Q = parts[c];
    Q.cta_info.assign(CL_INFO_STRIDE, 0);
    int32_t *I = Q.cta_info.data();
    const auto &L = local_rows[c];
    const int n_own = (int)L.size();
    int wlo = lo[c], whi = lo[c + 1];
    for (int p : L)
      for (int k = rptr[p]; k < rptr[p + 1]; ++k) {
        wlo = std::min(wlo, rcol[k]);
        whi = std::max(whi, rcol[k] + 1);
      }
    if (n_own == 0) wlo = whi = lo[c];
    const int Wn = whi - wlo;
    const int n_my = (int)my_nodes[c].size();
    if (n_own > 65535 || Wn > 65535 || n_my > 65535) {
      Q.error = "slice too large for 16-bit slots";
      return;
    }
    I[CI_N_OWN] = n_own; I[CI_LO] = lo[c]; I[CI_WLO] = wlo; I[CI_WN] = Wn; I[CI_N_MY] = n_my;
    I[CI_OFF_ROW] = (int32_t)Q.row_edge.size();
    I[CI_OFF_SLOT] = (int32_t)Q.slot_src.size();
    I[CI_OFF_BLK] = (int32_t)Q.blk_off.size();
    I[CI_OFF_HALO] = (int32_t)Q.halo_ws.size();
    I[CI_OFF_NODE] = (int32_t)Q.node_id.size();
    I[CI_OFF_N2E] = (int32_t)Q.n2e_item.size();
    I[CI_OFF_NSRC] = (int32_t)Q.nsrc_item.size();
    // rows
    for (int t = 0; t < n_own; ++t) {
      const int p = L[t], e = P.c_orig[p];
      Q.row_edge.push_back(e);
      Q.row_ws.push_back((uint16_t)(p - wlo));
      if (P.aux) {
        Q.row_n0.push_back((uint16_t)slot_of_node(c, edge_nodes[2 * (size_t)e]));
        Q.row_n1.push_back((uint16_t)slot_of_node(c, edge_nodes[2 * (size_t)e + 1]));
      } else {
        Q.row_n0.push_back(0);
        Q.row_n1.push_back(0);
      }
    }
    // ELL blocks of 32 local rows
    const int n_blk = (n_own + 31) / 32;
    I[CI_N_BLK] = n_blk;
    int slots = 0, vunits = 0;
    for (int b = 0; b < n_blk; ++b) {
      Q.blk_off.push_back(slots);
      Q.blk_voff.push_back(vunits);
      bool blk_cplx = false;
      for (int l = 0; l < 32 && b * 32 + l < n_own; ++l)
        blk_cplx |= row_complex ? (row_complex[P.c_orig[L[b * 32 + l]]] != 0) : true;
      int width = 0;
      for (int l = 0; l < 32 && b * 32 + l < n_own; ++l) width = std::max(width, rptr[L[b * 32 + l] + 1] - rptr[L[b * 32 + l]]);
      const size_t base = Q.slot_src.size();
      Q.slot_src.resize(base + (size_t)width * 32, -1);
      Q.slot_col.resize(base + (size_t)width * 32, 0);
      for (int l = 0; l < 32; ++l) {
        const int t = b * 32 + l;
        const uint16_t self = t < n_own ? (uint16_t)(L[t] - wlo) : (uint16_t)0;
        for (int k = 0; k < width; ++k) Q.slot_col[base + (size_t)k * 32 + l] = self;  // padding reads a valid slot
      }
      // The order of the entries inside a row is free, and the kernel gathers p from shared memory with one 16-byte
      // load per entry: the 8 lanes of a quarter warp are served in one wavefront only if their window slots fall in
      // 8 different bank groups (slot mod 8).  Per quarter and step, the rows pick -- fewest choices first -- an unused
      // bank group among the entries they have left (their fullest one), else their fullest group.
      constexpr int FILL_MAX = 96;  // rows longer than this keep the column order
      for (int q = 0; q < 4 && width <= FILL_MAX; ++q) {
        int32_t ecs[8][FILL_MAX], esrc[8][FILL_MAX];  // entries bucketed by bank group (lane-local arrays)
        int beg[8][8], cnt[8][8], left[8], ng[8];
        for (int u = 0; u < 8; ++u) {
          left[u] = 0;
          ng[u] = 0;
          for (int g = 0; g < 8; ++g) cnt[u][g] = 0;
          const int t = b * 32 + 8 * q + u;
          if (t >= n_own) continue;
          const int p = L[t];
          for (int k = rptr[p]; k < rptr[p + 1]; ++k) cnt[u][(rcol[k] - wlo) & 7]++;
          int run = 0;
          for (int g = 0; g < 8; ++g) {
            beg[u][g] = run;
            run += cnt[u][g];
            ng[u] += cnt[u][g] > 0;
            cnt[u][g] = 0;
          }
          for (int k = rptr[p + 1] - 1; k >= rptr[p]; --k) {  // descending: taking from the back yields ascending columns
            const int cs = rcol[k] - wlo, g = cs & 7, at = beg[u][g] + cnt[u][g]++;
            ecs[u][at] = cs;
            esrc[u][at] = rsrc[k];
            left[u]++;
          }
        }
        for (int k = 0; k < width; ++k) {
          unsigned used = 0, served = 0;
          for (int pick = 0; pick < 8; ++pick) {
            int u = -1;
            for (int v = 0; v < 8; ++v) {
              if ((served >> v & 1u) || left[v] == 0) continue;
              if (u < 0 || ng[v] < ng[u]) u = v;
            }
            if (u < 0) break;
            int best = -1, any = -1, bc = 0, ac = 0;
            for (int g = 0; g < 8; ++g) {
              const int cg = cnt[u][g];
              if (cg > ac) { ac = cg; any = g; }
              if (!(used >> g & 1u) && cg > bc) { bc = cg; best = g; }
            }
            const int g = best >= 0 ? best : any;
            used |= 1u << g;
            served |= 1u << u;
            const int at = beg[u][g] + --cnt[u][g];
            if (cnt[u][g] == 0) ng[u]--;
            left[u]--;
            Q.slot_col[base + (size_t)k * 32 + 8 * q + u] = (uint16_t)ecs[u][at];
            Q.slot_src[base + (size_t)k * 32 + 8 * q + u] = esrc[u][at];
          }
        }
      }
      if (width > FILL_MAX)
        for (int l = 0; l < 32 && b * 32 + l < n_own; ++l) {
          const int p = L[b * 32 + l];
          for (int k = rptr[p]; k < rptr[p + 1]; ++k) {
            Q.slot_src[base + (size_t)(k - rptr[p]) * 32 + l] = rsrc[k];
            Q.slot_col[base + (size_t)(k - rptr[p]) * 32 + l] = (uint16_t)(rcol[k] - wlo);
          }
        }
      slots += width * 32;
      vunits += width * 32 * (blk_cplx ? 2 : 1);
    }
    Q.blk_off.push_back(slots);
    Q.blk_voff.push_back(vunits);
    I[CI_N_SLOTS] = slots;
    // halo
    int n_halo = 0;
    for (int p = wlo; p < whi; ++p) {
      if (p >= lo[c] && p < lo[c + 1]) continue;
      Q.halo_ws.push_back((uint16_t)(p - wlo));
      Q.halo_src.push_back((uint32_t)part_of[p] << 16 | (uint32_t)local_of[p]);
      ++n_halo;
    }
    I[CI_N_HALO] = n_halo;
    // nodal lists
    {
      std::vector<std::vector<uint32_t>> items(n_my);
      if (P.aux)
        for (int t = 0; t < n_own; ++t) {
          const int e = P.c_orig[L[t]];
          items[slot_of_node(c, edge_nodes[2 * (size_t)e])].push_back((uint32_t)t << 1);
          items[slot_of_node(c, edge_nodes[2 * (size_t)e + 1])].push_back((uint32_t)t << 1 | 1u);
        }
      int acc = 0, acc2 = 0;
      for (int s = 0; s < n_my; ++s) {
        Q.node_id.push_back(my_nodes[c][s]);
        Q.n2e_ptr.push_back(acc);
        Q.nsrc_ptr.push_back(acc2);
        for (uint32_t it : items[s]) Q.n2e_item.push_back(it);
        acc += (int)items[s].size();
        for (uint32_t it : touch[my_nodes[c][s]]) Q.nsrc_item.push_back(it);
        acc2 += (int)touch[my_nodes[c][s]].size();
      }
      Q.n2e_ptr.push_back(acc);   // one extra entry per CTA: pointer arrays live at CI_OFF_NODE + c
      Q.nsrc_ptr.push_back(acc2);
    }
    Q.max_own = std::max(Q.max_own, n_own);
    Q.max_w = std::max(Q.max_w, Wn);
    Q.max_my = std::max(Q.max_my, n_my);
    Q.max_slots = std::max(Q.max_slots, slots);
    Q.max_vunits = std::max(Q.max_vunits, vunits);
    Q.max_halo = std::max(Q.max_halo, n_halo);
    Q.max_n2e = std::max(Q.max_n2e, Q.n2e_ptr.back());
    Q.max_nsrc = std::max(Q.max_nsrc, Q.nsrc_ptr.back());
  };
  if (C > 1) {  // the CTAs are independent (the bank-aware fill dominates): one host thread each
    std::vector<std::thread> th;
    for (int c = 0; c < C; ++c) th.emplace_back(build_cta, c);
    for (auto &t : th) t.join();
  } else {
    build_cta(0);
  }
  stage("per-CTA lists (one thread each)");
  P.cta_info.assign((size_t)C * CL_INFO_STRIDE, 0);
  auto app = [](auto &dst, const auto &src) { dst.insert(dst.end(), src.begin(), src.end()); };
  for (int c = 0; c < C; ++c) {
    const ClusterPlanHost &Q = parts[c];
    if (!Q.error.empty()) {
      P.error = Q.error;
      return false;
    }
    int32_t *I = &P.cta_info[(size_t)c * CL_INFO_STRIDE];
    for (int k = 0; k < CL_INFO_STRIDE; ++k) I[k] = Q.cta_info[k];
    I[CI_OFF_ROW] = (int32_t)P.row_edge.size();
    I[CI_OFF_SLOT] = (int32_t)P.slot_src.size();
    I[CI_OFF_BLK] = (int32_t)P.blk_off.size();
    I[CI_OFF_HALO] = (int32_t)P.halo_ws.size();
    I[CI_OFF_NODE] = (int32_t)P.node_id.size();
    I[CI_OFF_N2E] = (int32_t)P.n2e_item.size();
    I[CI_OFF_NSRC] = (int32_t)P.nsrc_item.size();
    app(P.row_edge, Q.row_edge); app(P.row_ws, Q.row_ws); app(P.row_n0, Q.row_n0); app(P.row_n1, Q.row_n1);
    app(P.blk_off, Q.blk_off); app(P.blk_voff, Q.blk_voff); app(P.slot_src, Q.slot_src); app(P.slot_col, Q.slot_col);
    app(P.halo_ws, Q.halo_ws); app(P.halo_src, Q.halo_src); app(P.node_id, Q.node_id);
    app(P.n2e_ptr, Q.n2e_ptr); app(P.n2e_item, Q.n2e_item); app(P.nsrc_ptr, Q.nsrc_ptr); app(P.nsrc_item, Q.nsrc_item);
    P.max_own = std::max(P.max_own, Q.max_own); P.max_w = std::max(P.max_w, Q.max_w); P.max_my = std::max(P.max_my, Q.max_my);
    P.max_slots = std::max(P.max_slots, Q.max_slots); P.max_halo = std::max(P.max_halo, Q.max_halo);
    P.max_vunits = std::max(P.max_vunits, Q.max_vunits);
    P.max_n2e = std::max(P.max_n2e, Q.max_n2e); P.max_nsrc = std::max(P.max_nsrc, Q.max_nsrc);
  }
  {  // push lists = halo lists transposed, ordered by (own row, reader)
    std::vector<std::vector<std::pair<uint16_t, uint32_t>>> out(C);
    for (int c = 0; c < C; ++c) {
      const int32_t *I = &P.cta_info[(size_t)c * CL_INFO_STRIDE];
      for (int h = 0; h < I[CI_N_HALO]; ++h) {
        const uint32_t src = P.halo_src[(size_t)I[CI_OFF_HALO] + h];
        out[src >> 16].push_back({(uint16_t)(src & 0xffffu), (uint32_t)c << 16 | (uint32_t)h});
      }
    }
    for (int c = 0; c < C; ++c) {
      std::sort(out[c].begin(), out[c].end());
      int32_t *I = &P.cta_info[(size_t)c * CL_INFO_STRIDE];
      I[CI_OFF_PUSH] = (int32_t)P.push_row.size();
      I[CI_N_PUSH] = (int32_t)out[c].size();
      for (auto &e : out[c]) {
        P.push_row.push_back(e.first);
        P.push_dst.push_back(e.second);
      }
      P.max_push = std::max(P.max_push, (int)out[c].size());
    }
  }
  stage("merge + push lists");
  return true;
}

}  // namespace efb
