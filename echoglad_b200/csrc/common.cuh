// Shared helpers for the echoglad_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/echoglad_b200.h"

#include <atomic>

namespace eg {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// brackets the kernels of one launch helper with CUDA events when eg_profile_enable(1) is active
class ProfileScope {
 public:
  ProfileScope(const char* name, cudaStream_t s);
  ~ProfileScope();
 private:
  const char* name_;
  cudaStream_t stream_;
  bool on_;
  void* beg_ = nullptr;
  void* end_ = nullptr;
};

#define EG_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      eg::set_error(__VA_ARGS__);               \
      return EG_ERR_INVALID;                    \
    }                                           \
  } while (0)

#define EG_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      eg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return EG_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

// after every kernel launch: counts the launch (eg_launch_count) and surfaces launch errors
#define EG_LAUNCH_CHECK()                          \
  do {                                             \
    eg::g_launches.fetch_add(1, std::memory_order_relaxed); \
    EG_CUDA(cudaGetLastError());                   \
  } while (0)

constexpr int kNumSMs = 148;            // B200: UPPER BOUND used to size workspaces / partial-slot tables at compile time
constexpr int kMaxParts = 4 * kNumSMs;  // upper bound on per-CTA partial slots of any reduction
// SM count of the CURRENT device (cudaDevAttrMultiProcessorCount, cached per device), capped at kNumSMs so that
// every grid sized from it fits the workspace layout: the launch helpers size persistent grids with this.
int num_sms();
// true exactly once per (mask, current device): guards per-device one-time setup such as
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which is a per-device attribute.  Thread-safe.
bool first_use_on_current_device(std::atomic<unsigned long long>& mask);
// workspace: [kMaxParts][2][128] doubles for column statistics + [kNumSMs][128*128] floats for the
// weight-gradient partials + slack
constexpr size_t kStatsBytes = (size_t)kMaxParts * 2 * 128 * sizeof(double);
constexpr size_t kWgradBytes = (size_t)kNumSMs * 128 * 128 * sizeof(float);
constexpr size_t kWorkspaceBytes = kStatsBytes + kWgradBytes + 65536;

// Per-tile gather plan of the fused tensor-core kernel (built on the host in graph.cu, device arrays).
// The unique source rows a tile reads most (own rows, lattice halo, parents) are STAGED in shared memory
// once per K chunk; a tile row then sums up to 7 staged neighbours, up to 4 FAR neighbours read straight
// from global (the 2x2 children of an aux node) and its self loop (staged, always last).  Rows that do
// not fit this shape (hubs of use_connection_nodes, 'grid-diagonal' aux rows) are CSR rows: the plan
// weights are zero and the row is summed from the device CSR instead.
constexpr int kPlanSrc = 216;   // staged source rows per tile (216 x 128 B = 27 KB per K chunk)
constexpr int kPlanStaged = 7;  // staged non-self edges per row
constexpr int kPlanFar = 4;     // far edges per row
struct alignas(16) PlanRow {    // 80 bytes
  uint8_t slot[8];     // [0..6] staged neighbours in summation order, [7] the row itself (self loop)
  int32_t csr_beg;     // CSR rows: first CSR entry, else 0
  int32_t csr_deg;     // CSR rows: entries incl. the self loop, else 0
  float w[8];          // deg^-1/2[u] deg^-1/2[v] of slot[k]; 0 = unused (slot 0 is always a valid row)
  int32_t far_node[4]; // frame-local node ids read from global; unused = node 0 with weight 0
  float far_w[4];
};
static_assert(sizeof(PlanRow) == 80, "PlanRow layout");
struct TilePlan {
  const int4* hdr;      // [tiles] {staged source rows, max staged non-self edges, far edges (0 / 4), has CSR rows}
  const int32_t* src;   // [tiles][kPlanSrc] staged node ids, ascending (-1 = unused)
  const PlanRow* rows;  // [tiles][128]
};

// ---- patch plan of the fused kernel's TMA path (gcn_tc.cu, MODE = patch; built in graph.cu) -----------------
// A regular 4-neighbour lattice level whose side is a multiple of 16 is cut into 8 x 16 patches (the tile table).
// Per K chunk a patch tile stages, with one TMA box copy each,
//   P: the haloed patch, [10][18] lattice positions x 128 B (out-of-lattice positions arrive as zeros),
//   Q: the [4][8] parents of the patch (coarser level),
// The 2 x 2 children of a coarse node are NOT staged.  Two routes:
//   * families (the last aux level -> the main level, where 90 % of the child bytes are): an aux patch and the (up to)
//     2 x 2 main patches that hold its children form a UNIT that one SM processes back to back.  The main patches have
//     the four children of a coarse node in one half-warp's registers anyway: they write the pooled sum
//     S[parent] = sum_c dis[c] x[c] (one row per 2 x 2 block) to a per-SM scratch buffer in global memory (128 rows,
//     L2-resident, double-buffered by unit), and the aux patch adds dis[v] * S[v] -- 16 KB per chunk instead of 64 KB of
//     child rows, and the main-level rows cross HBM once.  Same SM, so a CTA barrier orders the writes and the reads;
//   * every other patch with children: the half-warp that owns a 2 x 2 block reads its 16 child rows straight from
//     global, issued one chunk ahead (16 LDG.64 per lane in flight across the operand stores / barrier round trip).
// Either way a tile costs one slot use per chunk and the ring stays four chunks ahead.
// A compute half-warp owns a 2 x 2 block of the patch; its 4 x 6 lattice weights (up, left, right, down, parent, self;
// 0 = no such edge) and 4 x 4 child weights are gcn_norm's dis[v] * dis[u], precomputed per block.
constexpr int kPatchPRows = 10 * 18, kPatchQRows = 4 * 8;
constexpr int kPoolRows = 128;  // pooled rows of one aux patch
struct alignas(16) PatchTile {  // 64 bytes
  int32_t cls;              // 0 = patch, 1 = patch with children (direct loads), 2 = CSR tile (rows summed from the device
                            // CSR), 3 = patch with children through the pool of its unit
  int32_t level;            // lattice level of the patch
  int32_t y0, x0;           // patch origin inside the level
  int32_t qlevel, qy, qx;   // parents: level (-1 = none) and origin of the 4 x 8 box
  int32_t clevel, cy, cx;   // children: level (-1 = none) and origin of the 16 x 32 box (may lie partly outside)
  int32_t node0, side;      // frame-local node id of lattice position (0, 0) and side of the level
  int32_t cnode0, cside;    // same for the children level
  int32_t pool_rel;         // >= 0: member of a unit -- block (by, bx) writes its pooled sum to pool row pool_rel + 16 by + bx; -1: no
  int32_t pad_;
};
static_assert(sizeof(PatchTile) == 64, "PatchTile layout");
struct alignas(16) PatchBlockW {  // 192 bytes
  float wl[4][6];  // node (a, b, c, d) = ((0,0), (0,1), (1,0), (1,1)) of the block x (up, left, right, down, parent, self)
  float wc[4][4];  // node x child (2 ny + i, 2 nx + j) -> [i * 2 + j]
  float dv[4];     // dis[node]: its coefficient in the pooled sum of its parent
  float wp[4];     // dis[node] if the node has children (coefficient of its pooled child sum), else 0
};
static_assert(sizeof(PatchBlockW) == 192, "PatchBlockW layout");
struct PatchPlan {
  int ok;                      // 0: the graph has tiles this path cannot run (diagonal lattices, hubs): use the gather plan
  const PatchTile* tiles;      // [tiles_per_frame]
  const PatchBlockW* blocks;   // [tiles_per_frame][32]
  // processing sequence of a frame: units (runs of tiles one SM processes back to back: a family, or a single tile);
  // the persistent grid deals UNITS round-robin
  const int32_t* seq;          // [tiles_per_frame] tile ids, unit by unit
  const int32_t* unit_off;     // [units_per_frame + 1] first position of each unit in seq
  int units_per_frame;
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- counter-based dropout RNG: one 64-bit mix per group of 4 consecutive elements -----------------
// keep(element e) <=> 16-bit lane of mix(seed, e/4) >= p * 65536.  Stateless, so backward recomputes it.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  float t = p * 65536.0f;
  return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : (uint32_t)t);
}
// returns 4 keep-flags (bit i = element 4*group+i kept)
__host__ __device__ __forceinline__ uint32_t drop_keep4(uint64_t seed, uint64_t group, uint32_t thr) {
  uint64_t r = mix64(seed ^ (group * 0x9e3779b97f4a7c15ULL));
  uint32_t k = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) k |= (((uint32_t)(r >> (16 * i)) & 0xffffu) >= thr ? 1u : 0u) << i;
  return k;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming (evict-first) 128-bit store for outputs that are not re-read by this kernel
__device__ __forceinline__ void st4_stream(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

}  // namespace eg
