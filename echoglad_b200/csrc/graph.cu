// Static hierarchical graph: closed-form CSR built once on the device, plus the exports that prove
// bit-exactness against the reference's networkx/PyG edge_index (src/core/datasets.py:375-521, :258).
#include <cub/device/device_scan.cuh>

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <cmath>
#include <utility>
#include <vector>

#include "topology.cuh"

namespace eg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace eg

struct eg_graph {
  eg_graph_spec spec;
  eg::Topo topo;
  eg_graph_info info;
  int device;
  int nnz;
  int32_t* rowptr;  // [N+1]
  int32_t* col;     // [nnz]   sources, ascending, self loop last
  float* w;         // [nnz]   dis[src]*dis[dst]
  float* dis;       // [N]
  // Output-tile table of the tensor-core kernels: tiles of 128 node ids (-1 = padding).  Lattice levels
  // whose side is a multiple of 16 are cut into 8 x 16 patches so that the stencil neighbours of a tile
  // (up/down/left/right rows, the 2x2 children, the parent) are shared inside the tile and hit L1; all
  // other nodes are packed in index order.  Every node of the frame appears exactly once.
  int32_t* tile_nodes;  // [tiles_per_frame][128]
  int tiles_per_frame;
  int32_t* tile_groups;  // [tiles_per_frame][16]: first node / row count of each 16-row group (build_groups)
  // Gather plan of the fused kernel, per tile (see TilePlan in common.cuh): the unique source rows a tile
  // reads most (its own rows, the lattice halo, the parents) are staged in shared memory; every edge of a
  // tile row is either a slot of that stage or a direct global read (e.g. the 4 children of an aux node).
  eg::TilePlan plan;
  eg::PatchPlan patch;  // TMA path of the fused kernel (regular lattices only)
  // pool scratch of the patch path ([SMs][2][kPoolRows][128] floats, zero at allocation), one per stream that has
  // launched on this graph (launches on one stream are ordered; two streams must not share it)
  struct PoolScratch {
    cudaStream_t stream;
    float* pool;
  };
  std::vector<PoolScratch> pool;
  std::mutex pool_mu;
};

using namespace eg;

namespace {

__global__ void degree_kernel(Topo t, int32_t* __restrict__ deg1, float* __restrict__ dis) {
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= t.N) return;
  int d = degree_of(t, u) + 1;  // + GCN self loop (PyG add_remaining_self_loops)
  deg1[u] = d;
  // PyG: deg.pow(-0.5) in fp32; evaluate in double and round once (deg is a small integer)
  dis[u] = (float)(1.0 / sqrt((double)d));
}

__global__ void fill_kernel(Topo t, const int32_t* __restrict__ rowptr, const float* __restrict__ dis,
                            int32_t* __restrict__ col, float* __restrict__ w) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= t.N) return;
  int e = rowptr[v];
  const float dv = dis[v];
  for_each_neighbor(t, v, true, [&](int u) {
    col[e] = u;
    w[e] = dis[u] * dv;  // gcn_norm: deg_inv_sqrt[row] * 1 * deg_inv_sqrt[col]
    ++e;
  });
  col[e] = v;
  w[e] = dv * dv;
}

__global__ void export_kernel(Topo t, const int32_t* __restrict__ rowptr, int E, int batch,
                              int64_t* __restrict__ out) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)batch * t.N) return;
  int b = (int)(gid / t.N), v = (int)(gid % t.N);
  long long cols = (long long)batch * E;
  long long e = (long long)b * E + (rowptr[v] - v);
  long long base = (long long)b * t.N;
  for_each_neighbor(t, v, false, [&](int u) {
    out[e] = base + v;
    out[cols + e] = base + u;
    ++e;
  });
}

__global__ void check_kernel(Topo t, const int32_t* __restrict__ rowptr, int E, int batch,
                             const int64_t* __restrict__ ei, int32_t* __restrict__ mismatch) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)batch * t.N) return;
  int b = (int)(gid / t.N), v = (int)(gid % t.N);
  long long cols = (long long)batch * E;
  long long e = (long long)b * E + (rowptr[v] - v);
  long long base = (long long)b * t.N;
  int bad = 0;
  for_each_neighbor(t, v, false, [&](int u) {
    bad += (ei[e] != base + v) + (ei[cols + e] != base + u);
    ++e;
  });
  if (bad) atomicAdd(mismatch, bad);
}

__global__ void set_int_kernel(int32_t* p, int32_t v) { *p = v; }

void fill_info(const Topo& t, eg_graph_info* info, int num_edges, int max_degree) {
  memset(info, 0, sizeof(*info));
  info->num_nodes = t.N;
  info->num_edges = num_edges;
  info->num_pixel_nodes = t.N0;
  info->first_pixel_node = t.nconn;
  info->num_coord_nodes = t.ncoord;
  info->num_levels = t.nlev;
  for (int l = 0; l < t.nlev; ++l) {
    info->level_size[l] = t.lsize[l];
    info->level_offset[l] = t.loff[l];
  }
  info->max_degree = max_degree;
  info->crop_offset = t.crop;
}

// Tile table contract: every aligned group of 16 tile rows is a run of CONSECUTIVE node ids, possibly
// shorter than 16 (then padded with -1 at its end); the epilogue of the tensor-core kernels addresses a
// group as base row + immediate offsets.
std::vector<int32_t> build_tiles(const Topo& t) {
  std::vector<int32_t> tiles;
  int in_group = 0, last = -2;
  auto close_group = [&]() {
    while (in_group % 16) {
      tiles.push_back(-1);
      ++in_group;
    }
    in_group = 0;
  };
  auto push_misc = [&](int v) {
    if (in_group == 16 || (in_group && v != last + 1)) close_group();
    tiles.push_back(v);
    ++in_group;
    last = v;
  };
  // nodes outside the 8 x 16 patches first: connection hubs, lattices whose side is not a multiple of 16
  // (in index order), coordinate nodes
  for (int v = 0; v < t.nconn; ++v) push_misc(v);
  for (int l = 0; l < t.nlev; ++l) {
    const int p = t.lsize[l], off = t.loff[l];
    if (p % 16 != 0)
      for (int v = 0; v < p * p; ++v) push_misc(off + v);
  }
  for (int v = t.N - t.ncoord; v < t.N; ++v) push_misc(v);
  close_group();
  while (tiles.size() % 128) tiles.push_back(-1);
  // patches: (level, a0, b0), level by level ...
  struct Patch { int l, a0, b0; };
  std::vector<Patch> coarse, fine;  // patches of the aux levels / of the main level
  for (int l = 0; l < t.nlev; ++l) {
    const int p = t.lsize[l];
    if (p % 16 != 0) continue;
    for (int a0 = 0; a0 < p; a0 += 8)
      for (int b0 = 0; b0 < p; b0 += 16) (l == t.nlev - 1 ? fine : coarse).push_back({l, a0, b0});
  }
  // ... level by level, coarse to fine.  Three other orders were measured on B200 (r02z, forward at batch 64) and
  // dropped: the aux patches spread evenly among the main patches (1.36 ms against 1.30), depth-first over the patch
  // pyramid (a patch followed by the patches of its children: 1.31 ms, DRAM reads 3.66 against 3.54 GB) and fine to
  // coarse with the coarse patch one round behind its children (1.36 ms, DRAM reads 4.16 GB).
  std::vector<Patch> order = coarse;
  order.insert(order.end(), fine.begin(), fine.end());
  for (const Patch& q : order) {
    const int p = t.lsize[q.l], off = t.loff[q.l];
    for (int a = q.a0; a < q.a0 + 8; ++a)
      for (int b = q.b0; b < q.b0 + 16; ++b) tiles.push_back(off + a * p + b);
  }
  return tiles;
}

// per tile: [0..7] first node of each 16-row group (-1 = empty), [8..15] rows in the group
std::vector<int32_t> build_groups(const std::vector<int32_t>& tiles) {
  const size_t T = tiles.size() / 128;
  std::vector<int32_t> grp(T * 16, 0);
  for (size_t ti = 0; ti < T; ++ti)
    for (int gi = 0; gi < 8; ++gi) {
      const int32_t* q = &tiles[ti * 128 + gi * 16];
      int cnt = 0;
      while (cnt < 16 && q[cnt] >= 0) ++cnt;
      grp[ti * 16 + gi] = cnt ? q[0] : -1;
      grp[ti * 16 + 8 + gi] = cnt;
    }
  return grp;
}

struct HostPlan {
  std::vector<int4> hdr;
  std::vector<int32_t> src;
  std::vector<PlanRow> rows;
};

// dis / weights exactly as degree_kernel / fill_kernel compute them on the device
HostPlan build_plan(const Topo& t, const std::vector<int32_t>& tiles) {
  const int T = (int)(tiles.size() / 128);
  std::vector<float> dis(t.N);
  std::vector<int32_t> rowbeg(t.N + 1, 0);
  for (int u = 0; u < t.N; ++u) {
    const int d = degree_of(t, u) + 1;
    dis[u] = (float)(1.0 / sqrt((double)d));
    rowbeg[u + 1] = rowbeg[u] + d;
  }
  HostPlan hp;
  hp.hdr.assign(T, make_int4(0, 0, 0, 0));
  hp.src.assign((size_t)T * kPlanSrc, -1);
  PlanRow zero;
  memset(&zero, 0, sizeof(zero));
  hp.rows.assign((size_t)T * 128, zero);
  std::vector<std::pair<int, int>> uses;  // (source, count)
  std::vector<int> nb, st, far;
  for (int ti = 0; ti < T; ++ti) {
    uses.clear();
    // hub rows have thousands of neighbours: they are CSR rows and do not vote for staged sources
    constexpr int kVoteDegree = kPlanStaged + kPlanFar;
    for (int r = 0; r < 128; ++r) {
      const int v = tiles[(size_t)ti * 128 + r];
      if (v < 0) continue;
      uses.emplace_back(v, 1 << 20);  // own rows always staged (the self loop is read from the stage)
      if (degree_of(t, v) > kVoteDegree) continue;
      for_each_neighbor(t, v, true, [&](int u) { uses.emplace_back(u, 1); });
    }
    std::sort(uses.begin(), uses.end());
    {  // merge duplicates
      size_t o = 0;
      for (size_t i = 0; i < uses.size(); ++i) {
        if (o && uses[o - 1].first == uses[i].first) uses[o - 1].second += uses[i].second;
        else uses[o++] = uses[i];
      }
      uses.resize(o);
    }
    // keep the most used sources (ties: lower index), then number the slots in ascending source order
    std::stable_sort(uses.begin(), uses.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
      return a.second != b.second ? a.second > b.second : a.first < b.first;
    });
    if ((int)uses.size() > kPlanSrc) uses.resize(kPlanSrc);
    std::sort(uses.begin(), uses.end());
    for (int s2 = 0; s2 < (int)uses.size(); ++s2) hp.src[(size_t)ti * kPlanSrc + s2] = uses[s2].first;
    auto slot_of = [&](int u) {
      auto it = std::lower_bound(uses.begin(), uses.end(), std::make_pair(u, 0),
                                 [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
      return (it != uses.end() && it->first == u) ? (int)(it - uses.begin()) : -1;
    };
    int ks = 0, nfar = 0, has_csr = 0;
    for (int r = 0; r < 128; ++r) {
      const int v = tiles[(size_t)ti * 128 + r];
      if (v < 0) continue;
      PlanRow& pr = hp.rows[(size_t)ti * 128 + r];
      nb.clear();
      st.clear();
      far.clear();
      bool fits = degree_of(t, v) <= 4096;
      if (fits) for_each_neighbor(t, v, true, [&](int u) { nb.push_back(u); });
      for (int u : nb) {
        if (slot_of(u) >= 0 && (int)st.size() < kPlanStaged) st.push_back(u);
        else far.push_back(u);
      }
      const int self_slot = slot_of(v);
      fits = fits && (int)far.size() <= kPlanFar && self_slot >= 0;
      if (!fits) {
        pr.csr_beg = rowbeg[v];
        pr.csr_deg = rowbeg[v + 1] - rowbeg[v];
        has_csr = 1;
        continue;
      }
      for (int k = 0; k < (int)st.size(); ++k) {
        pr.slot[k] = (uint8_t)slot_of(st[k]);
        pr.w[k] = dis[st[k]] * dis[v];
      }
      pr.slot[7] = (uint8_t)self_slot;
      pr.w[7] = dis[v] * dis[v];
      for (int k = 0; k < (int)far.size(); ++k) {
        pr.far_node[k] = far[k];
        pr.far_w[k] = dis[far[k]] * dis[v];
      }
      ks = std::max(ks, (int)st.size());
      if (!far.empty()) nfar = kPlanFar;
    }
    hp.hdr[ti] = make_int4((int)uses.size(), ks, nfar, has_csr);
  }
  return hp;
}

struct HostPatchPlan {
  int ok = 0;
  std::vector<PatchTile> tiles;
  std::vector<PatchBlockW> blocks;
  std::vector<int32_t> seq, unit_off;  // processing sequence of a frame (PatchPlan)
};

// Patch plan (see PatchTile / PatchBlockW in common.cuh).  Weights are dis[v] * dis[u] with dis exactly as
// degree_kernel computes it, so the patch path and the CSR agree bit for bit on every coefficient.
HostPatchPlan build_patch_plan(const Topo& t, const std::vector<int32_t>& tiles) {
  HostPatchPlan pp;
  const int T = (int)(tiles.size() / 128);
  pp.ok = !t.mdiag && !t.adiag && !t.conn;
  if (!pp.ok) return pp;
  PatchTile zt;
  memset(&zt, 0, sizeof(zt));
  zt.cls = 2;
  zt.qlevel = zt.clevel = -1;
  zt.pool_rel = -1;
  pp.tiles.assign(T, zt);
  PatchBlockW zb;
  memset(&zb, 0, sizeof(zb));
  pp.blocks.assign((size_t)T * 32, zb);
  auto dis = [&](int u) { return (float)(1.0 / sqrt((double)(degree_of(t, u) + 1))); };
  const int main_l = t.nlev - 1;
  int npatch = 0;
  for (int l = 0; l < t.nlev; ++l)
    if (t.lsize[l] % 16 == 0) npatch += (t.lsize[l] / 8) * (t.lsize[l] / 16);
  int ti = T - npatch;  // build_tiles: the misc tiles come first (they stay CSR tiles), then the patches level by level
  if (ti < 0) {
    pp.ok = 0;
    return pp;
  }
  for (; ti < T; ++ti) {
    {
      // the patch of this tile, from its first node (the order of the patches in the table is free)
      const int v0 = tiles[(size_t)ti * 128];
      int l = t.nlev - 1;
      while (l > 0 && v0 < t.loff[l]) --l;
      const int p = t.lsize[l], off = t.loff[l];
      const int y0 = v0 < off ? -1 : (v0 - off) / p, x0 = v0 < off ? -1 : (v0 - off) % p;
      if (v0 < off || p % 16 != 0 || y0 % 8 != 0 || x0 % 16 != 0 || y0 + 8 > p) {
        pp.ok = 0;
        return pp;
      }
      const bool is_main = l == main_l;
      {
        PatchTile& pt = pp.tiles[ti];
        pt.level = l;
        pt.y0 = y0;
        pt.x0 = x0;
        pt.node0 = off;
        pt.side = p;
        // parents (for_each_neighbor: main -> crop window of the last aux level; aux l >= 1 -> level l - 1)
        if (is_main) {
          if (!t.main_only) {
            pt.qlevel = t.naux - 1;
            pt.qy = t.crop + y0 / 2;
            pt.qx = t.crop + x0 / 2;
          }
        } else if (l >= 1) {
          pt.qlevel = l - 1;
          pt.qy = y0 / 2;
          pt.qx = x0 / 2;
        }
        // children
        if (!is_main) {
          if (l < t.naux - 1) {
            pt.clevel = l + 1;
            pt.cy = 2 * y0;
            pt.cx = 2 * x0;
          } else {
            pt.clevel = main_l;
            pt.cy = 2 * (y0 - t.crop);
            pt.cx = 2 * (x0 - t.crop);
          }
          pt.cnode0 = t.loff[pt.clevel];
          pt.cside = t.lsize[pt.clevel];
        }
        bool any_child = false;
        for (int q = 0; q < 32; ++q) {
          PatchBlockW& bw = pp.blocks[(size_t)ti * 32 + q];
          const int by = q >> 3, bx = q & 7;
          for (int n = 0; n < 4; ++n) {
            const int a = y0 + 2 * by + (n >> 1), b = x0 + 2 * bx + (n & 1);
            const int v = off + a * p + b;
            const float dv = dis(v);
            if (a > 0) bw.wl[n][0] = dv * dis(v - p);
            if (b > 0) bw.wl[n][1] = dv * dis(v - 1);
            if (b < p - 1) bw.wl[n][2] = dv * dis(v + 1);
            if (a < p - 1) bw.wl[n][3] = dv * dis(v + p);
            if (pt.qlevel >= 0) {
              const int P = t.lsize[pt.qlevel];
              const int pa = is_main ? t.crop + a / 2 : a / 2, pb = is_main ? t.crop + b / 2 : b / 2;
              if (!is_main || (a < 2 * t.half && b < 2 * t.half)) bw.wl[n][4] = dv * dis(t.loff[pt.qlevel] + pa * P + pb);
            }
            bw.wl[n][5] = dv * dv;
            bw.dv[n] = dv;
            if (pt.clevel >= 0) {
              const int P2 = t.lsize[pt.clevel], off2 = t.loff[pt.clevel];
              const bool window = l < t.naux - 1 || (a >= t.crop && a < t.crop + t.half && b >= t.crop && b < t.crop + t.half);
              if (window) {
                const int ca = l < t.naux - 1 ? 2 * a : 2 * (a - t.crop), cb = l < t.naux - 1 ? 2 * b : 2 * (b - t.crop);
                for (int i = 0; i < 2; ++i)
                  for (int j = 0; j < 2; ++j) bw.wc[n][i * 2 + j] = dv * dis(off2 + (ca + i) * P2 + cb + j);
                bw.wp[n] = dv;
                any_child = true;
              }
            }
          }
        }
        pt.cls = any_child ? 1 : 0;
        if (!any_child) pt.clevel = -1;
      }
    }
  }
  if (ti != T) pp.ok = 0;
  // ---- units.  A family = a patch of the LAST aux level that has children + the main patches that hold them: the main
  // patches write the pooled child sums of their 2 x 2 blocks (pool_rel), the aux patch reads them (cls 3).  Needs the
  // parents of a main patch inside ONE aux patch: crop a multiple of 8 (224 / 7: 8, 448 / 8: 16); otherwise every patch
  // with children keeps the direct loads.  EG_PATCH_FAMILIES=0 switches the families off (A/B timing).
  const char* fam_env = getenv("EG_PATCH_FAMILIES");
  const bool families = !(fam_env && fam_env[0] == '0') && !t.main_only && t.lsize[main_l] % 16 == 0 &&
                        t.naux >= 1 && t.lsize[t.naux - 1] % 16 == 0 && t.crop % 8 == 0;
  std::vector<int> tile_at_main, tile_at_aux;  // tile id of the patch at (row, column) of the main / last aux level
  std::vector<char> in_family(T, 0);
  if (families) {
    const int pm = t.lsize[main_l], pa = t.lsize[t.naux - 1];
    tile_at_main.assign((pm / 8) * (pm / 16), -1);
    tile_at_aux.assign((pa / 8) * (pa / 16), -1);
    for (int i = 0; i < T; ++i) {
      const PatchTile& pt = pp.tiles[i];
      if (pt.cls == 2) continue;
      if (pt.level == main_l) tile_at_main[(pt.y0 / 8) * (pm / 16) + pt.x0 / 16] = i;
      if (pt.level == t.naux - 1) tile_at_aux[(pt.y0 / 8) * (pa / 16) + pt.x0 / 16] = i;
    }
  }
  auto push_unit = [&](const std::vector<int>& members) {
    pp.unit_off.push_back((int)pp.seq.size());
    for (int m : members) pp.seq.push_back(m);
  };
  std::vector<std::vector<int>> fam_units;
  if (families) {
    const int pm = t.lsize[main_l], pa = t.lsize[t.naux - 1];
    for (int ty = 0; ty < pa / 8; ++ty)
      for (int tx = 0; tx < pa / 16; ++tx) {
        const int ai = tile_at_aux[ty * (pa / 16) + tx];
        if (ai < 0 || pp.tiles[ai].cls != 1) continue;
        std::vector<int> members;
        for (int my = 0; my < pm / 8; ++my)
          for (int mx = 0; mx < pm / 16; ++mx) {
            const int qy = t.crop + (my * 8) / 2, qx = t.crop + (mx * 16) / 2;  // parents of the main patch: [qy, qy+4) x [qx, qx+8)
            if (qy / 8 != ty || qx / 16 != tx) continue;
            const int mi = tile_at_main[my * (pm / 16) + mx];
            if (mi < 0) continue;
            pp.tiles[mi].pool_rel = (qy - ty * 8) * 16 + (qx - tx * 16);
            in_family[mi] = 1;
            members.push_back(mi);
          }
        members.push_back(ai);
        pp.tiles[ai].cls = 3;
        in_family[ai] = 1;
        fam_units.push_back(members);
      }
  }
  for (int i = 0; i < T; ++i)
    if (!in_family[i]) push_unit({i});  // singles first, in table order (coarse levels, childless aux patches, ...)
  for (const auto& m : fam_units) push_unit(m);
  pp.unit_off.push_back((int)pp.seq.size());
  if ((int)pp.seq.size() != T) pp.ok = 0;
  return pp;
}

int init_topo(const eg_graph_spec* spec, Topo& t) {
  if (!spec) {
    set_error("spec is NULL");
    return EG_ERR_INVALID;
  }
  int rc = topo_init(t, *spec);
  if (rc == -2) {
    set_error("unsupported spec: 2^num_aux_graphs (%d) < frame_size/2 (%d); the reference builds a "
              "malformed centre crop for it (src/core/datasets.py:502)", 1 << spec->num_aux_graphs,
              spec->frame_size / 2);
    return EG_ERR_UNSUPPORTED;
  }
  if (rc != 0) {
    set_error("invalid graph spec (frame_size=%d, num_aux_graphs=%d)", spec->frame_size, spec->num_aux_graphs);
    return EG_ERR_INVALID;
  }
  return EG_OK;
}

}  // namespace

extern "C" {

const char* eg_version(void) { return "echoglad_b200 0.1 (sm_100a)"; }
const char* eg_last_error(void) { return eg::g_err; }
size_t eg_workspace_bytes(void) { return eg::kWorkspaceBytes; }

int eg_graph_spec_info(const eg_graph_spec* spec, eg_graph_info* info) {
  Topo t;
  int rc = init_topo(spec, t);
  if (rc) return rc;
  EG_CHECK_ARG(info, "info is NULL");
  long long e = 0;
  int maxd = 0;
  for (int u = 0; u < t.N; ++u) {
    int d = degree_of(t, u);
    e += d;
    if (d + 1 > maxd) maxd = d + 1;
  }
  fill_info(t, info, (int)e, maxd);
  return EG_OK;
}

int eg_graph_host_edge_index(const eg_graph_spec* spec, int batch, int64_t* out) {
  Topo t;
  int rc = init_topo(spec, t);
  if (rc) return rc;
  EG_CHECK_ARG(out && batch >= 1, "bad arguments");
  long long E = 0;
  for (int u = 0; u < t.N; ++u) E += degree_of(t, u);
  long long cols = E * batch, e = 0;
  for (int b = 0; b < batch; ++b) {
    long long base = (long long)b * t.N;
    for (int v = 0; v < t.N; ++v)
      for_each_neighbor(t, v, false, [&](int u) {
        out[e] = base + v;
        out[cols + e] = base + u;
        ++e;
      });
  }
  return EG_OK;
}

int eg_graph_host_node_type(const eg_graph_spec* spec, int batch, double* out) {
  Topo t;
  int rc = init_topo(spec, t);
  if (rc) return rc;
  EG_CHECK_ARG(out && batch >= 1, "bad arguments");
  for (int b = 0; b < batch; ++b)
    for (int v = 0; v < t.N; ++v)
      out[(size_t)b * t.N + v] = v < t.nconn ? 2.0 : (v >= t.N - t.ncoord ? 1.0 : 0.0);
  return EG_OK;
}

int eg_graph_create(const eg_graph_spec* spec, int device, eg_graph** out) {
  Topo t;
  int rc = init_topo(spec, t);
  if (rc) return rc;
  EG_CHECK_ARG(out, "out is NULL");
  int prev = 0;
  EG_CUDA(cudaGetDevice(&prev));
  EG_CUDA(cudaSetDevice(device));
  eg_graph* g = new eg_graph();
  g->spec = *spec;
  g->topo = t;
  g->device = device;
  int32_t* deg = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  auto fail = [&](int code) {
    cudaFree(deg);
    cudaFree(tmp);
    eg_graph_destroy(g);
    cudaSetDevice(prev);
    return code;
  };
#define EG_TRY(call)                                                                         \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      set_error("%s failed: %s", #call, cudaGetErrorString(_e));                             \
      return fail(EG_ERR_CUDA);                                                              \
    }                                                                                        \
  } while (0)
  EG_TRY(cudaMalloc(&g->rowptr, sizeof(int32_t) * (t.N + 1)));
  EG_TRY(cudaMalloc(&g->dis, sizeof(float) * t.N));
  EG_TRY(cudaMalloc(&deg, sizeof(int32_t) * (t.N + 1)));
  EG_TRY(cudaMemset(deg, 0, sizeof(int32_t) * (t.N + 1)));
  const int threads = 128, blocks = (t.N + threads - 1) / threads;
  degree_kernel<<<blocks, threads>>>(t, deg, g->dis);
  EG_TRY(cudaGetLastError());
  EG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg, g->rowptr, t.N + 1));
  EG_TRY(cudaMalloc(&tmp, tmp_bytes));
  EG_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, g->rowptr, t.N + 1));
  int32_t nnz = 0;
  EG_TRY(cudaMemcpy(&nnz, g->rowptr + t.N, sizeof(int32_t), cudaMemcpyDeviceToHost));
  g->nnz = nnz;
  EG_TRY(cudaMalloc(&g->col, sizeof(int32_t) * nnz));
  EG_TRY(cudaMalloc(&g->w, sizeof(float) * nnz));
  fill_kernel<<<blocks, threads>>>(t, g->rowptr, g->dis, g->col, g->w);
  EG_TRY(cudaGetLastError());
  {
    std::vector<int32_t> tiles = build_tiles(t);
    g->tiles_per_frame = (int)(tiles.size() / 128);
    EG_TRY(cudaMalloc(&g->tile_nodes, sizeof(int32_t) * tiles.size()));
    EG_TRY(cudaMemcpy(g->tile_nodes, tiles.data(), sizeof(int32_t) * tiles.size(), cudaMemcpyHostToDevice));
    std::vector<int32_t> grp = build_groups(tiles);
    HostPlan hp = build_plan(t, tiles);
    auto up = [&](const void* h, size_t bytes, void** d) {
      cudaError_t e = cudaMalloc(d, bytes);
      return e != cudaSuccess ? e : cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    EG_TRY(up(grp.data(), grp.size() * sizeof(int32_t), (void**)&g->tile_groups));
    EG_TRY(up(hp.hdr.data(), hp.hdr.size() * sizeof(int4), (void**)&g->plan.hdr));
    EG_TRY(up(hp.src.data(), hp.src.size() * sizeof(int32_t), (void**)&g->plan.src));
    EG_TRY(up(hp.rows.data(), hp.rows.size() * sizeof(PlanRow), (void**)&g->plan.rows));
    HostPatchPlan pp = build_patch_plan(t, tiles);
    g->patch.ok = pp.ok;
    if (pp.ok) {
      EG_TRY(up(pp.tiles.data(), pp.tiles.size() * sizeof(PatchTile), (void**)&g->patch.tiles));
      EG_TRY(up(pp.blocks.data(), pp.blocks.size() * sizeof(PatchBlockW), (void**)&g->patch.blocks));
      EG_TRY(up(pp.seq.data(), pp.seq.size() * sizeof(int32_t), (void**)&g->patch.seq));
      EG_TRY(up(pp.unit_off.data(), pp.unit_off.size() * sizeof(int32_t), (void**)&g->patch.unit_off));
      g->patch.units_per_frame = (int)pp.unit_off.size() - 1;
    }
  }
  std::vector<int32_t> hdeg(t.N);
  EG_TRY(cudaMemcpy(hdeg.data(), deg, sizeof(int32_t) * t.N, cudaMemcpyDeviceToHost));
  int maxd = 0;
  for (int d : hdeg) maxd = d > maxd ? d : maxd;
  fill_info(t, &g->info, nnz - t.N, maxd);
  EG_TRY(cudaDeviceSynchronize());
#undef EG_TRY
  cudaFree(deg);
  cudaFree(tmp);
  cudaSetDevice(prev);
  *out = g;
  return EG_OK;
}

void eg_graph_destroy(eg_graph* g) {
  if (!g) return;
  cudaFree(g->rowptr);
  cudaFree(g->col);
  cudaFree(g->w);
  cudaFree(g->dis);
  cudaFree(g->tile_nodes);
  cudaFree(g->tile_groups);
  cudaFree((void*)g->plan.hdr);
  cudaFree((void*)g->plan.src);
  cudaFree((void*)g->plan.rows);
  cudaFree((void*)g->patch.tiles);
  cudaFree((void*)g->patch.blocks);
  cudaFree((void*)g->patch.seq);
  cudaFree((void*)g->patch.unit_off);
  for (auto& ps : g->pool) cudaFree(ps.pool);
  delete g;
}

int eg_graph_get_info(const eg_graph* g, eg_graph_info* info) {
  EG_CHECK_ARG(g && info, "NULL argument");
  *info = g->info;
  return EG_OK;
}

int eg_graph_csr(const eg_graph* g, const int32_t** rowptr, const int32_t** col, const float** w,
                 const float** dis) {
  EG_CHECK_ARG(g, "graph is NULL");
  if (rowptr) *rowptr = g->rowptr;
  if (col) *col = g->col;
  if (w) *w = g->w;
  if (dis) *dis = g->dis;
  return EG_OK;
}

int eg_graph_tiles(const eg_graph* g, const int32_t** tile_nodes, int32_t* tiles_per_frame) {
  EG_CHECK_ARG(g, "graph is NULL");
  if (tile_nodes) *tile_nodes = g->tile_nodes;
  if (tiles_per_frame) *tiles_per_frame = g->tiles_per_frame;
  return EG_OK;
}

// Host-only self check of the tile table and the gather plan (no GPU needed): every node in exactly one
// tile, 16-row groups consecutive, and for every row the plan (staged slots -> node ids, far nodes, self
// loop) or its CSR flag reproduces the neighbour list and the gcn_norm weights.  stats (optional, int64[6]):
// tiles, plan rows, CSR rows, far edges, staged edges, largest staged-source count.  Returns the number of
// violations (0 = consistent) or a negative error code.
int eg_graph_plan_check(const eg_graph_spec* spec, int64_t* stats) {
  Topo t;
  int rc = init_topo(spec, t);
  if (rc) return rc;
  const std::vector<int32_t> tiles = build_tiles(t);
  const std::vector<int32_t> grp = build_groups(tiles);
  const HostPlan hp = build_plan(t, tiles);
  const int T = (int)(tiles.size() / 128);
  long long bad = 0, plan_rows = 0, csr_rows = 0, far_edges = 0, staged_edges = 0, max_src = 0;
  long long cls[2] = {0, 0};
  std::vector<int> seen(t.N, 0);
  for (int32_t v : tiles)
    if (v >= 0) {
      if (v >= t.N) ++bad;
      else ++seen[v];
    }
  for (int v = 0; v < t.N; ++v) bad += seen[v] != 1;
  for (int ti = 0; ti < T; ++ti)
    for (int gi = 0; gi < 8; ++gi) {
      const int base = grp[ti * 16 + gi], cnt = grp[ti * 16 + 8 + gi];
      for (int k = 0; k < 16; ++k) {
        const int v = tiles[(size_t)ti * 128 + gi * 16 + k];
        bad += (k < cnt) ? (v != base + k) : (v != -1);
      }
    }
  std::vector<float> dis(t.N);
  for (int u = 0; u < t.N; ++u) dis[u] = (float)(1.0 / sqrt((double)(degree_of(t, u) + 1)));
  std::vector<std::pair<int, float>> want, got;
  for (int ti = 0; ti < T; ++ti) {
    const int4 h = hp.hdr[ti];
    max_src = std::max<long long>(max_src, h.x);
    // tile classes of the fused kernel (gcn_tc.cu): lattice (<= 5 staged neighbours, nothing else) / general
    if (h.y <= 5 && !h.z && !h.w) ++cls[0];
    else ++cls[1];
    bad += h.x < 1 || h.x > kPlanSrc || h.y < 0 || h.y > kPlanStaged || (h.z != 0 && h.z != kPlanFar);
    const int32_t* src = &hp.src[(size_t)ti * kPlanSrc];
    for (int k = 0; k < kPlanSrc; ++k) bad += (k < h.x) ? (src[k] < 0 || src[k] >= t.N) : (src[k] != -1);
    for (int r = 0; r < 128; ++r) {
      const int v = tiles[(size_t)ti * 128 + r];
      const PlanRow& pr = hp.rows[(size_t)ti * 128 + r];
      got.clear();
      for (int k = 0; k < 8; ++k)
        if (pr.w[k] != 0.f) {
          if (pr.slot[k] >= h.x || (k < 7 && k >= h.y)) ++bad;
          else got.emplace_back(src[pr.slot[k]], pr.w[k]);
          staged_edges += k < 7;
        } else if (pr.slot[k] != 0) {
          ++bad;
        }
      for (int k = 0; k < 4; ++k)
        if (pr.far_w[k] != 0.f) {
          bad += h.z == 0;
          got.emplace_back(pr.far_node[k], pr.far_w[k]);
          ++far_edges;
        } else if (pr.far_node[k] != 0) {
          ++bad;
        }
      if (v < 0) {
        bad += !got.empty() || pr.csr_deg != 0;
        continue;
      }
      if (pr.csr_deg) {
        ++csr_rows;
        bad += !got.empty() || !h.w || pr.csr_deg != degree_of(t, v) + 1;
        int beg = 0;
        for (int u = 0; u < v; ++u) beg += degree_of(t, u) + 1;
        bad += pr.csr_beg != beg;
        continue;
      }
      ++plan_rows;
      want.clear();
      for_each_neighbor(t, v, true, [&](int u) { want.emplace_back(u, dis[u] * dis[v]); });
      want.emplace_back(v, dis[v] * dis[v]);
      bad += pr.w[7] == 0.f || src[pr.slot[7]] != v;  // entry 7 (summed last) is the self loop
      std::sort(want.begin(), want.end());
      std::sort(got.begin(), got.end());
      bad += want != got;
    }
  }
  if (stats) {
    stats[0] = T, stats[1] = plan_rows, stats[2] = csr_rows, stats[3] = far_edges, stats[4] = staged_edges;
    stats[5] = max_src;
    stats[6] = cls[0], stats[7] = cls[1];
  }
  return (int)std::min<long long>(bad, 1 << 30);
}

// Host-only self check of the patch plan (TMA path of the fused kernel): replays, for every patch tile and node, the
// sources the kernel reads -- box position -> node id, exactly as the tensor-map coordinates address them; children
// either through the direct 4 x 4 window or through the pool rows the members of the tile's unit write -- and compares
// the (source, weight) set with the CSR row.  Also: the sequence is a permutation of the tiles, and a unit of more than
// one tile is a family (pool writers, then their reader).  stats (optional, int64[6]): plan usable (0/1), plain patch
// tiles, patch tiles with children (direct loads), CSR tiles, patch tiles reading the pool, units per frame.
// Returns the number of violations or a negative error code.
int eg_graph_patch_check(const eg_graph_spec* spec, int64_t* stats) {
  Topo t;
  int rc = init_topo(spec, t);
  if (rc) return rc;
  const std::vector<int32_t> tiles = build_tiles(t);
  const HostPatchPlan pp = build_patch_plan(t, tiles);
  const int T = (int)(tiles.size() / 128);
  long long bad = 0, cls[4] = {0, 0, 0, 0};
  if (stats) stats[0] = pp.ok, stats[1] = stats[2] = stats[3] = stats[4] = stats[5] = 0;
  if (!pp.ok) return 0;
  std::vector<float> dis(t.N);
  for (int u = 0; u < t.N; ++u) dis[u] = (float)(1.0 / sqrt((double)(degree_of(t, u) + 1)));
  auto at = [&](int level, int y, int x) {  // node at a box position, -1 = outside the lattice (TMA fills zeros)
    const int p = t.lsize[level];
    return (y < 0 || x < 0 || y >= p || x >= p) ? -1 : t.loff[level] + y * p + x;
  };
  // sequence: permutation; units; pool contents per family reader
  struct PoolTerm { int node; float w; };
  std::vector<std::vector<std::vector<PoolTerm>>> pool_of(T);  // reader tile -> pool row -> terms
  {
    std::vector<int> seen(T, 0);
    bad += (int)pp.seq.size() != T || pp.unit_off.empty() || pp.unit_off.front() != 0 || pp.unit_off.back() != T;
    for (int32_t v : pp.seq)
      if (v < 0 || v >= T) ++bad;
      else ++seen[v];
    for (int i = 0; i < T; ++i) bad += seen[i] != 1;
    for (size_t u = 0; u + 1 < pp.unit_off.size() && !bad; ++u) {
      const int i0 = pp.unit_off[u], i1 = pp.unit_off[u + 1];
      bad += i1 <= i0;
      if (i1 - i0 == 1) {
        const PatchTile& pt = pp.tiles[pp.seq[i0]];
        bad += pt.cls == 3 || pt.pool_rel >= 0;  // pool traffic only inside a family
        continue;
      }
      const int reader = pp.seq[i1 - 1];
      bad += pp.tiles[reader].cls != 3;
      pool_of[reader].assign(kPoolRows, {});
      for (int i = i0; i < i1 - 1; ++i) {
        const int w = pp.seq[i];
        const PatchTile& pt = pp.tiles[w];
        bad += pt.cls != 0 || pt.pool_rel < 0;
        for (int q = 0; q < 32; ++q) {
          const int row = pt.pool_rel + 16 * (q >> 3) + (q & 7);
          if (row < 0 || row >= kPoolRows) {
            ++bad;
            continue;
          }
          const PatchBlockW& bw = pp.blocks[(size_t)w * 32 + q];
          for (int n = 0; n < 4; ++n)
            pool_of[reader][row].push_back({at(pt.level, pt.y0 + 2 * (q >> 3) + (n >> 1), pt.x0 + 2 * (q & 7) + (n & 1)), bw.dv[n]});
        }
      }
    }
  }
  std::vector<std::pair<int, float>> want, got;
  for (int ti = 0; ti < T; ++ti) {
    const PatchTile& pt = pp.tiles[ti];
    if (pt.cls < 0 || pt.cls > 3) {
      ++bad;
      continue;
    }
    ++cls[pt.cls];
    if (pt.cls == 2) continue;
    bad += pt.node0 != t.loff[pt.level] || pt.side != t.lsize[pt.level] || (pt.cls == 1 || pt.cls == 3) != (pt.clevel >= 0);
    if (pt.clevel >= 0) bad += pt.cnode0 != t.loff[pt.clevel] || pt.cside != t.lsize[pt.clevel];
    for (int q = 0; q < 32; ++q) {
      const PatchBlockW& bw = pp.blocks[(size_t)ti * 32 + q];
      const int by = q >> 3, bx = q & 7;
      for (int n = 0; n < 4; ++n) {
        const int ny = n >> 1, nx = n & 1;
        const int y = pt.y0 + 2 * by + ny, x = pt.x0 + 2 * bx + nx;
        const int v = at(pt.level, y, x);
        bad += v < 0 || v != tiles[(size_t)ti * 128 + (2 * by + ny) * 16 + 2 * bx + nx];
        got.clear();
        auto add = [&](float w, int u) {
          if (w == 0.f) return;
          if (u < 0) ++bad;  // a weight on a position the copy fills with zeros
          else got.emplace_back(u, w);
        };
        add(bw.wl[n][0], at(pt.level, y - 1, x));
        add(bw.wl[n][1], at(pt.level, y, x - 1));
        add(bw.wl[n][2], at(pt.level, y, x + 1));
        add(bw.wl[n][3], at(pt.level, y + 1, x));
        add(bw.wl[n][4], pt.qlevel >= 0 ? at(pt.qlevel, pt.qy + by, pt.qx + bx) : -1);
        add(bw.wl[n][5], v);
        if (v >= 0) bad += bw.dv[n] != dis[v];
        if (pt.cls == 3) {  // children through the pool row of the node itself: row (2 by + ny) * 16 + 2 bx + nx
          if (bw.wp[n] != 0.f) {
            bad += v < 0 || bw.wp[n] != dis[v];
            const auto& terms = pool_of[ti].empty() ? std::vector<PoolTerm>() : pool_of[ti][(2 * by + ny) * 16 + 2 * bx + nx];
            bad += terms.size() != 4;
            for (const PoolTerm& tm : terms) add(tm.w * bw.wp[n], tm.node);
          }
        } else {
          bad += bw.wp[n] != 0.f && pt.cls != 1;
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
              add(bw.wc[n][i * 2 + j],
                  pt.clevel >= 0 ? at(pt.clevel, pt.cy + 4 * by + 2 * ny + i, pt.cx + 4 * bx + 2 * nx + j) : -1);
        }
        if (v < 0) continue;
        want.clear();
        for_each_neighbor(t, v, true, [&](int u) { want.emplace_back(u, dis[u] * dis[v]); });
        want.emplace_back(v, dis[v] * dis[v]);
        std::sort(want.begin(), want.end());
        std::sort(got.begin(), got.end());
        bad += want != got;
      }
    }
  }
  if (stats) {
    stats[1] = cls[0], stats[2] = cls[1], stats[3] = cls[2], stats[4] = cls[3];
    stats[5] = (long long)pp.unit_off.size() - 1;
  }
  return (int)std::min<long long>(bad, 1 << 30);
}

int eg_graph_export_edge_index(const eg_graph* g, int batch, int64_t* out, void* stream) {
  EG_CHECK_ARG(g && out && batch >= 1, "bad arguments");
  long long total = (long long)batch * g->topo.N;
  const int threads = 128;
  export_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
      g->topo, g->rowptr, g->info.num_edges, batch, out);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_graph_check_edge_index(const eg_graph* g, int batch, const int64_t* edge_index, int64_t num_cols,
                              int32_t* mismatch, void* stream) {
  EG_CHECK_ARG(g && edge_index && mismatch && batch >= 1, "bad arguments");
  cudaStream_t s = as_stream(stream);
  if (num_cols != (int64_t)batch * g->info.num_edges) {
    set_int_kernel<<<1, 1, 0, s>>>(mismatch, -1);
    EG_LAUNCH_CHECK();
    return EG_OK;
  }
  set_int_kernel<<<1, 1, 0, s>>>(mismatch, 0);
  long long total = (long long)batch * g->topo.N;
  const int threads = 128;
  check_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
      g->topo, g->rowptr, g->info.num_edges, batch, edge_index, mismatch);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // extern "C"

// accessors for the other translation units
namespace eg {
const Topo& graph_topo(const eg_graph* g) { return g->topo; }
const eg_graph_info& graph_info(const eg_graph* g) { return g->info; }
const int32_t* graph_rowptr(const eg_graph* g) { return g->rowptr; }
const int32_t* graph_col(const eg_graph* g) { return g->col; }
const float* graph_w(const eg_graph* g) { return g->w; }
const float* graph_dis(const eg_graph* g) { return g->dis; }
int graph_nnz(const eg_graph* g) { return g->nnz; }
const int32_t* graph_tile_nodes(const eg_graph* g) { return g->tile_nodes; }
int graph_tiles_per_frame(const eg_graph* g) { return g->tiles_per_frame; }
const int32_t* graph_tile_groups(const eg_graph* g) { return g->tile_groups; }
const TilePlan& graph_plan(const eg_graph* g) { return g->plan; }
const PatchPlan& graph_patch_plan(const eg_graph* g) { return g->patch; }

// Pool scratch of the patch path for launches on stream s (allocated on first use, zero-filled: rows of nodes without
// children are read -- with weight 0 -- but never written).
int graph_pool_scratch(const eg_graph* gc, cudaStream_t s, float** pool) {
  eg_graph* g = const_cast<eg_graph*>(gc);
  std::lock_guard<std::mutex> lock(g->pool_mu);
  for (auto& e : g->pool)
    if (e.stream == s) {
      *pool = e.pool;
      return EG_OK;
    }
  const size_t bytes = (size_t)kNumSMs * 2 * kPoolRows * 128 * sizeof(float);
  float* ptr = nullptr;
  EG_CUDA(cudaMalloc(&ptr, bytes));
  EG_CUDA(cudaMemsetAsync(ptr, 0, bytes, s));
  g->pool.push_back({s, ptr});
  *pool = ptr;
  return EG_OK;
}
}  // namespace eg
