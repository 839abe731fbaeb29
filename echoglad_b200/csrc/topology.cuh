// Closed-form topology of the EchoGLAD hierarchical graph (no networkx, no edge lists).
//
// Reference: create_graphs / add_inter_aux_task_edges / add_inter_main_task_edges
// (src/core/datasets.py:375-521) followed by PyG from_networkx (:258).  Node index inside a frame is
// [connection nodes][aux level 1 .. aux level n][main grid][coordinate nodes], each lattice row-major.
// The order in which the reference emits the neighbours of a node is
//     sorted(lower-index neighbours) ++ higher-index neighbours in networkx insertion order
// (SURVEY.md §3.3); `for_each_neighbor(.., sorted=false, ..)` reproduces it, `sorted=true` yields the
// ascending order used by the compute CSR.
#pragma once
#include "common.cuh"

namespace eg {

struct Topo {
  int S, naux, main_only, coord, conn, mdiag, adiag;
  int nlev;
  int lsize[EG_MAX_LEVELS];
  int loff[EG_MAX_LEVELS];
  int nconn, ncoord, N, N0, crop, half;
};

// returns 0 on success, <0 on an unsupported spec
inline int topo_init(Topo& t, const eg_graph_spec& s) {
  t.S = s.frame_size;
  t.main_only = s.use_main_graph_only != 0;
  t.naux = t.main_only ? 0 : s.num_aux_graphs;
  t.coord = (!t.main_only) && s.use_coordinate_graph != 0;
  t.conn = (!t.main_only) && s.use_connection_nodes != 0;
  t.mdiag = s.main_diagonal != 0;
  t.adiag = s.aux_diagonal != 0;
  if (t.S < 2 || t.S > 4096) return -1;
  if (!t.main_only && (t.naux < 1 || t.naux > 12)) return -1;
  t.nlev = t.naux + 1;
  t.nconn = t.conn ? t.naux + 1 : 0;
  t.ncoord = t.coord ? 4 : 0;
  t.half = t.S / 2;
  t.crop = 0;
  if (!t.main_only) {
    int P = 1 << t.naux;
    if (P < t.half) return -2;  // reference builds a malformed crop (negative python slice), Appendix A
    t.crop = (P - t.half) / 2;
  }
  int off = t.nconn;
  for (int l = 0; l < t.nlev; ++l) {
    int p = (l == t.nlev - 1) ? t.S : (1 << (l + 1));
    t.lsize[l] = p;
    t.loff[l] = off;
    off += p * p;
  }
  t.N0 = off - t.nconn;
  t.N = off + t.ncoord;
  return 0;
}

template <class F>
__host__ __device__ __forceinline__ void for_each_neighbor(const Topo& t, int u, bool sorted, F&& f) {
  if (u < t.nconn) {  // connection hub k: K(naux+1) + every node of aux level k+1 (k+1 <= naux-1)
    for (int j = 0; j < t.nconn; ++j)
      if (j != u) f(j);
    int g = u + 1;
    if (g <= t.naux - 1) {
      int off = t.loff[g - 1], cnt = t.lsize[g - 1] * t.lsize[g - 1];
      for (int v = 0; v < cnt; ++v) f(off + v);
    }
    return;
  }
  if (u >= t.N - t.ncoord) {  // coordinate K4, isolated from the lattices
    int base = t.N - t.ncoord;
    for (int j = 0; j < t.ncoord; ++j)
      if (base + j != u) f(base + j);
    return;
  }
  int l = t.nlev - 1;
  while (u < t.loff[l]) --l;
  const int p = t.lsize[l], off = t.loff[l];
  const int a = (u - off) / p, b = (u - off) % p;
  const bool is_main = (l == t.nlev - 1);
  const bool diag = is_main ? t.mdiag : t.adiag;
  // ---- lower-index neighbours, ascending ----
  if (t.conn && !is_main && (l + 1) <= t.naux - 1) f(l);  // hub of this aux level
  if (is_main) {
    if (!t.main_only && a < 2 * t.half && b < 2 * t.half) {
      int P = t.lsize[t.naux - 1];
      f(t.loff[t.naux - 1] + (t.crop + a / 2) * P + t.crop + b / 2);
    }
  } else if (l >= 1) {
    f(t.loff[l - 1] + (a / 2) * (p / 2) + b / 2);
  }
  if (diag && a > 0 && b > 0) f(u - p - 1);
  if (a > 0) f(u - p);
  if (diag && a > 0 && b < p - 1) f(u - p + 1);
  if (b > 0) f(u - 1);
  // ---- higher-index neighbours ----
  if (sorted) {
    if (b < p - 1) f(u + 1);
    if (diag && a < p - 1 && b > 0) f(u + p - 1);
    if (a < p - 1) f(u + p);
    if (diag && a < p - 1 && b < p - 1) f(u + p + 1);
  } else {  // networkx insertion order: grid down, right; then the two diagonal edge lists
    if (a < p - 1) f(u + p);
    if (b < p - 1) f(u + 1);
    if (diag && a < p - 1 && b < p - 1) f(u + p + 1);
    if (diag && a < p - 1 && b > 0) f(u + p - 1);
  }
  if (!is_main) {
    if (l < t.naux - 1) {
      int off2 = t.loff[l + 1], p2 = 2 * p;
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) f(off2 + (2 * a + i) * p2 + 2 * b + j);
    } else if (a >= t.crop && a < t.crop + t.half && b >= t.crop && b < t.crop + t.half) {
      int off2 = t.loff[t.nlev - 1];
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) f(off2 + (2 * (a - t.crop) + i) * t.S + 2 * (b - t.crop) + j);
    }
  }
}

__host__ __device__ __forceinline__ int degree_of(const Topo& t, int u) {
  int d = 0;
  for_each_neighbor(t, u, true, [&](int) { ++d; });
  return d;
}

}  // namespace eg
