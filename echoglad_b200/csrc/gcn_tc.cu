// Fused message passing + dense transform on the 5th-generation tensor cores (tcgen05 / TMEM):
//
//     Out[tile rows] = (A_hat X)[tile rows] * op(W) (+ bias) (+ addend)        one persistent kernel
//
// Replaces PyG GCNConv = propagate(scatter_add) + nn.Linear (src/core/models.py:330,431) with the
// re-association A_hat (X W^T) = (A_hat X) W^T (4e-7 rms), so the aggregated rows never
// touch HBM: compute warps gather / weight / sum the neighbour rows of a 128-node tile (atomic-free, fixed order),
// split the sums into a tf32 part and a bf16 correction part and write them straight into swizzled shared-memory
// operand tiles; one elected thread issues tcgen05.mma (tf32 main term + one bf16 MMA for both correction terms) into a
// double-buffered TMEM accumulator; epilogue warps drain it (tcgen05.ld), add bias / residual gradient, accumulate the
// BatchNorm column statistics and store rows.
//
// One body, three MODES that share the MMA issuer, the epilogue and the operand / accumulator rings:
//   kLinear  plain per-node transform (classifier layer 0 and its input gradient): the tile's own rows are staged;
//   kGather  any graph (round 1): three LOADER warps copy the tile's unique source rows (own rows + lattice halo +
//            parents, <= 216 rows x 128 B, eg::TilePlan) with cp.async into a 3-stage raw ring; a lane group of 8 owns a
//            row (6 LDS.128 + 12 FFMA2 on lattice tiles); the 2x2 children of an aux node are read from global one row
//            group ahead; hub rows through the device CSR.  Runs graphs with hubs / diagonal lattices;
//   kPatch   regular 4-neighbour lattices (round 2, see PatchTile in common.cuh): ONE thread stages the haloed 8x16 patch
//            and the parents with two TMA box copies per chunk (4-deep ring); a half-warp owns a 2x2 node block (13 LDS.64
//            + 24 FFMA2 per 4 rows); children arrive pooled through the tile's unit (families) or by direct loads issued
//            one chunk ahead; tiles are walked unit by unit.
//
// The product is computed TRANSPOSED, D^T[f][r] = sum_k Wop[f][k] * A[r][k]:
//   * the weight (tf32 part and bf16 correction part, 2 x 128 TMEM columns) is the M-side operand and lives in TENSOR
//     MEMORY for the whole kernel, so each MMA reads only the 4 KB node-tile slice from shared memory (half the
//     shared-memory traffic of an SS-mode MMA) and all of shared memory is a ring of operand stages;
//   * the accumulator has one output feature per TMEM lane and one tile row per column, so an epilogue
//     thread owns a feature: bias and the column statistics are per-thread scalars, and for a fixed row
//     the 32 lanes of a warp hold 32 consecutive features = one coalesced 128-byte store.  No staging.
//
// Tile = 128 output rows (8x16 lattice patches, see eg_graph::tile_nodes); K is consumed in 4 chunks of 32 features; the
// split sums fill a 3-stage OPERAND ring ([128 rows x 128 B] tf32 + correction = 32 KB per stage).
// TMEM: 256 columns of weight + 2 x 128 of accumulator.
// Warps: 0-3 epilogue (TMEM lane quadrant = warp id), 4 MMA issuer, 5-7 loaders (patch mode: 5 = the TMA thread), 8-23
// compute; register budgets per warpgroup with setmaxnreg (RegBudget below).
//
// Development switches (never defined in the shipped build; `EG_NVCC_EXTRA=-D... python echoglad_b200/build.py`, prebuilt
// variants tools/build_variants.sh + tools/gpu_pd.sh): EG_TC_TIMING adds per-role wait-cycle counters (eg_tc_debug_read,
// printed by tools/kernel_bench.py); EG_DBG_NOGATHER / NOEMIT / NOFENCE / NOCOMPUTE / NOLOAD / NOMMA / NOSTORE / NOEPI /
// SMALLOUT and EG_PD_NOGATHER / NOCHILD / PLAINAUX / NOPOOLOUT / NOTMA each remove one piece of work (WRONG results,
// timing only: the knock-out tables of DESIGN.md 4.1); EG_TF32X3 restores the three-MMA tf32 split.
#include <cuda.h>  // CUtensorMap (the encode function is fetched through cudaGetDriverEntryPoint: no -lcuda)

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc05.cuh"

struct eg_graph;
namespace eg {
const eg_graph_info& graph_info(const eg_graph* g);
const int32_t* graph_rowptr(const eg_graph* g);
const int32_t* graph_col(const eg_graph* g);
const float* graph_w(const eg_graph* g);
const int32_t* graph_tile_nodes(const eg_graph* g);
const int32_t* graph_tile_groups(const eg_graph* g);
int graph_tiles_per_frame(const eg_graph* g);
const TilePlan& graph_plan(const eg_graph* g);
const PatchPlan& graph_patch_plan(const eg_graph* g);
int graph_pool_scratch(const eg_graph* g, cudaStream_t s, float** pool);
int launch_stats_finalize(int nparts, int cols, int stride, long long rows, const double* parts, float* mean,
                          float* var, cudaStream_t s);
}  // namespace eg

using namespace eg;
using namespace eg::tc;

// Forward output rows are written once and re-read only by a later kernel: streaming (evict-first) stores keep
// the L2 for the frame's input rows, which the halo / parent / child gathers re-read (+1..3 %).  The backward's
// outputs (dX, A_hat dH) are consumed immediately by the weight-gradient / BatchNorm kernels and stay plain
// (measured: streaming them costs 15 % of the backward).
#define EG_ST_OUT(ptr, v) __stcs((ptr), (v))
#define EG_ST_AGG(ptr, v) st4((ptr), (v))

// L2 prefetch distance of the loader warps, in tile rounds (0 = off).  r02e on B200, batch 64: 1 -> forward 1.442 ms
// (off: 1.478), backward 3.17 (3.24); 2 / 3 / 5 -> 1.50 / 1.56 / 1.58 ms (the prefetched rows compete with the
// frame's halo rows for the L2).
#ifndef EG_PF_TILES
#define EG_PF_TILES 1
#endif

namespace {

constexpr int kStages = 3;                 // operand ring (hi + lo tiles)
constexpr int kProdWarps = 16;             // compute warps: 8 tile rows per warp and stage
constexpr int kRowsPerProd = 128 / kProdWarps;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kLoadWarps = 3;
constexpr int kLoadWarp0 = kMmaWarp + 1;
constexpr int kProdWarp0 = kLoadWarp0 + kLoadWarps;
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;  // 768 threads launched with 80 registers each
static_assert(kProdWarp0 == 8, "warpgroup layout of setmaxnreg");
// Register budget by warpgroup (setmaxnreg): epilogue 72, MMA + loaders 56, the 16 compute warps 88
// (each SM sub-partition hosts 1 + 1 + 4 of them: (72 + 56 + 4 x 88) x 32 lanes = 15360 <= 16384 registers).
#ifndef EG_REGS_PATCH
#define EG_REGS_PATCH 88, 80, 48
#endif
struct RegBudget {
  int compute, epi, load;
};
constexpr RegBudget kRegsDefault{88, 72, 56};
// patch mode: the loader warpgroup only hosts the MMA issuer and the one TMA thread (48 registers are plenty), which
// lets the epilogue keep the launch's 80 (at 72 it spilled inside its slab loop as
// soon as the compute code grew: r02x)
constexpr RegBudget kRegsPatch{EG_REGS_PATCH};
constexpr bool regs_fit(RegBudget r) { return 4 * 32 * (r.epi + r.load) + kProdWarps * 32 * r.compute <= 65536 && r.epi <= 80 && r.load <= 80; }
static_assert(regs_fit(kRegsDefault) && regs_fit(kRegsPatch), "register pool of the SM (64 K registers, launched at 80 per thread)");
constexpr int kTileBytes = 128 * 128;      // one [128 rows x 32 tf32] operand tile
constexpr uint32_t kABytes = kStages * 2 * kTileBytes;
// Kernel modes: plain per-node transform / gather plan (any graph: cp.async row copies, per-row slot plan) / patch plan
// (regular lattices: TMA box copies, one 2 x 2 node block per half-warp; see PatchTile in common.cuh).
constexpr int kLinear = 0, kGather = 1, kPatch = 2;
template <int MODE>
struct Lay {
  static constexpr int kRawStages = MODE == kPatch ? 4 : 3;
  // gather / linear: kPlanSrc staged rows (linear mode stages the tile's own 128 rows); patch: P + Q boxes (a C sub-stage
  // of 128 child rows reuses the front of a slot)
  static constexpr int kRawRows = MODE == kPatch ? kPatchPRows + kPatchQRows : kPlanSrc;
  static constexpr uint32_t kRawBytes = kRawRows * 128;
  // per compute warp and tile: gather: 8 plan rows + tile header; patch: its two blocks' weights + the tile descriptor
  static constexpr uint32_t kPlanWarpBytes =
      MODE == kPatch ? 2 * sizeof(PatchBlockW) + sizeof(PatchTile) : kRowsPerProd * sizeof(PlanRow) + 16;
  static constexpr uint32_t kPlanBytes = kProdWarps * 2 * kPlanWarpBytes;  // double-buffered per warp
  static constexpr uint32_t kSmemBytes = kABytes + kRawStages * kRawBytes + kPlanBytes + 256 /*barriers*/ + 1024 /*align*/;
  // shared-memory map, byte offsets from the 1024-byte aligned base
  static constexpr uint32_t kOffRaw = kABytes;
  static constexpr uint32_t kOffPlan = kOffRaw + kRawStages * kRawBytes;
  static constexpr uint32_t kOffBars = kOffPlan + kPlanBytes;
  static constexpr uint32_t kOffFull = kOffBars, kOffEmpty = kOffFull + 8 * kStages, kOffRawFull = kOffEmpty + 8 * kStages,
                            kOffRawEmpty = kOffRawFull + 8 * kRawStages;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory of one CTA");
};
constexpr int kTmemCols = 512;             // [0,128) W hi, [128,256) W lo, [256,384) / [384,512) accumulators
constexpr uint32_t kTmemAcc = 256;
static_assert(kPlanSrc % (kLoadWarps * 4) == 0 && kPlanSrc >= 128, "loader mapping");

struct PatchMaps {  // tensor maps of the node tensor, per lattice level: [3 l + 0 / 1] = P / Q box (see PatchTile)
  CUtensorMap m[3 * EG_MAX_LEVELS];
};

struct TcParams {
  const int32_t* tile_nodes;   // GATHER: [tiles_per_frame][128]
  const int32_t* tile_groups;  // GATHER: [tiles_per_frame][16] first node / rows of each 16-row group
  TilePlan plan;               // gather mode: per-tile staged sources and per-row edges
  PatchPlan patch;             // patch mode: tile descriptors, per-block weights, processing sequence
  float* pool;                 // patch mode: pooled child sums, [SMs][2][kPoolRows][128]
  int batch;                   // patch mode: frames
  int tiles_per_frame;
  int nodes_per_frame;
  int num_tiles;              // < 2^31 / 128 (rows < 2^31, checked by the launchers)
  long long rows;             // total rows of X / Out
  const int32_t* rowptr;
  const int32_t* col;
  const float* w;
  const float* X;
  const float* W;
  int trans_w;                // 1: Out = A W^T (nn.Linear forward), 0: Out = A W (its input gradient)
  const float* bias;
  const float* addend;
  float* Out;
  float* AggOut;              // optional: the aggregated rows A_hat X themselves
  double* stat_parts;         // optional: [gridDim.x][2][128] column sum / sum of squares partials
  // EPI instances only (inference layer): Out = act(acc * ep_scale[f] + ep_shift[f]) + addend
  const float* ep_scale;
  const float* ep_shift;
  int ep_relu;
};

#ifdef EG_TC_TIMING
__device__ long long g_tc_dbg[kNumSMs][20];  // per CTA: cycles spent waiting, by role (see eg_tc_debug_read)
#define TC_TIMED_WAIT(slot, bar, par)            \
  do {                                           \
    const long long _t = clock64();              \
    mbar_wait(bar, par);                         \
    dbg_acc[slot] += clock64() - _t;             \
  } while (0)
#else
#define TC_TIMED_WAIT(slot, bar, par) mbar_wait(bar, par)
#endif

__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// packed fp32 pairs (one FADD2 / FFMA2 issue slot for two elements): used by the epilogue
struct F2 {
  unsigned long long u;
};
__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
  return F2{((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo)};
}
__device__ __forceinline__ float f2_lo(F2 a) { return __uint_as_float((uint32_t)a.u); }
__device__ __forceinline__ float f2_hi(F2 a) { return __uint_as_float((uint32_t)(a.u >> 32)); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
  return r;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u));
  return r;
}

// acc += w * x on a packed pair (FFMA2 with the scalar weight broadcast)
__device__ __forceinline__ void fma2(F2& acc, float w, F2 x) {
  const unsigned long long w2 = ((unsigned long long)__float_as_uint(w) << 32) | __float_as_uint(w);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.u) : "l"(w2), "l"(x.u));
}
__device__ __forceinline__ F2 lds_f2(uint32_t addr) {
  F2 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v.u) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts2(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ F2 ldg_f2(const float* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return f2_pack(v.x, v.y);
}
__device__ __forceinline__ void st_f2(float* p, F2 v) { *reinterpret_cast<float2*>(p) = make_float2(f2_lo(v), f2_hi(v)); }
// Pool rows (per-SM scratch, rewritten every unit): kept in L2 (evict_last) in the forward launch -- with the default
// policy the streaming traffic of the kernel evicts every dirty pool line before it is rewritten, 0.39 GB of DRAM writes
// per launch.
__device__ __forceinline__ uint64_t l2_policy(bool evict_last) {
  uint64_t pol;
  if (evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_f2_keep(float* p, F2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(f2_lo(v)), "f"(f2_hi(v)), "l"(pol)
               : "memory");
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& x) {
  acc.x = fmaf(w, x.x, acc.x);
  acc.y = fmaf(w, x.y, acc.y);
  acc.z = fmaf(w, x.z, acc.z);
  acc.w = fmaf(w, x.w, acc.w);
}
// Same arithmetic (round-to-nearest fp32 FMA per element) as two packed FFMA2 instructions: sm_100 issues
// fma.rn.f32x2 with the scalar weight broadcast to both halves, which halves the FMA issue slots of the gather.
__device__ __forceinline__ void fma4_x2(float4& acc, float w, const float4& x) {
  unsigned long long a0, a1;
  const unsigned long long w2 = ((unsigned long long)__float_as_uint(w) << 32) | __float_as_uint(w);
  const unsigned long long x0 = ((unsigned long long)__float_as_uint(x.y) << 32) | __float_as_uint(x.x);
  const unsigned long long x1 = ((unsigned long long)__float_as_uint(x.w) << 32) | __float_as_uint(x.z);
  const unsigned long long c0 = ((unsigned long long)__float_as_uint(acc.y) << 32) | __float_as_uint(acc.x);
  const unsigned long long c1 = ((unsigned long long)__float_as_uint(acc.w) << 32) | __float_as_uint(acc.z);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a0) : "l"(w2), "l"(x0), "l"(c0));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a1) : "l"(w2), "l"(x1), "l"(c1));
  acc.x = __uint_as_float((uint32_t)a0);
  acc.y = __uint_as_float((uint32_t)(a0 >> 32));
  acc.z = __uint_as_float((uint32_t)a1);
  acc.w = __uint_as_float((uint32_t)(a1 >> 32));
}

// 128-bit load of a far (not staged) source row slice: read once per SM, so it bypasses the ~28 KB of L1 left beside
// the 203 KB of shared memory (ld.global.cg; measured -1 % against ld.global.nc, L1::no_allocate +4 %).
__device__ __forceinline__ float4 ld_far(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ int lane_id() {
  int l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// child rows: read once per SM, straight from L2 (no L1 allocation)
__device__ __forceinline__ F2 ldcg_f2(const float* p) {
  const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
  return f2_pack(v.x, v.y);
}
// 3xTF32 operand split of a packed pair -> operand tiles (see emit in tc_body for the layout of the correction tile)
__device__ __forceinline__ void emit_pair(uint32_t a_hi, uint32_t off, F2 v) {
#ifdef EG_DBG_NOEMIT
  if (f2_lo(v) == 123.456f) sts2(a_hi + off, make_uint2(0, 0));
  return;
#endif
  uint2 hi, lo;
  split_tf32_op(f2_lo(v), hi.x, lo.x);
  split_tf32_op(f2_hi(v), hi.y, lo.y);
  sts2(a_hi + off, hi);
#ifdef EG_TF32X3
  sts2(a_hi + kTileBytes + off, lo);
#else
  sts2(a_hi + kTileBytes + off, make_uint2(pack_bf16x2(__uint_as_float(lo.x), __uint_as_float(lo.y)),
                                           pack_bf16x2(f2_lo(v), f2_hi(v))));
#endif
}

// ---- patch mode, one tile WITH CHILDREN, as a function of its own --------------------------------------------------
// The 2 x 2 children of the block's four nodes (a 4 x 4 window of the finer level) are not staged: 16 LDG.64 per lane,
// issued one chunk AHEAD -- right after the previous chunk's sums, when the 13 staged rows are dead -- so the operand
// stores / fence / barrier round trip cover most of their L2 latency; they are summed first, before the staged rows are
// loaded.  Children exist per NODE (all four or none): a zero weight marks a node without (crop border).
// __noinline__ on purpose: inlined into the tile loop, its 32 registers of loads in flight made the compiler spill
// kernel-wide values (the lane id, ring addresses) that the plain tiles and the MMA issuer then re-loaded from local
// memory on their critical paths (plain tiles +25 %, r02r-w).
struct AuxTileArgs {
  const float* cwin;   // direct: first child row of the block's window; pooled: pool row of node a; this lane's 2 features, chunk 0
  float* agg_a;        // A_hat dH side output (backward) of node a, or nullptr
  uint32_t wlu;        // shared address of the block's weights: wl[4][6], wc[4][4], dv[4], wp[4]
  uint32_t pb, qb;     // shared offsets of P[0][0] / the parent row of the block inside a slot (slot base not added)
  uint32_t so_a, so_b; // operand-tile offsets of tile rows a and b (c, d: + 2048)
  int cside, side;     // sides of the children level / the patch level
  uint32_t use, chunk; // ring counters (in / out)
};
// POOLED: the children arrive as ONE pooled row per node (written by the main patches of the tile's unit, see
// PatchPlan): 4 loads per lane and chunk instead of 16, weights wp[4].
template <bool POOLED, uint32_t kRawStagesT, uint32_t kRawBytesT, uint32_t kOffRawFullT, uint32_t kOffRawEmptyT,
          uint32_t kOffFullT, uint32_t kOffEmptyT>
#ifndef EG_AUX_INLINE
#define EG_AUX_INLINE __forceinline__
#endif
__device__ EG_AUX_INLINE void patch_aux_tile(AuxTileArgs& a) {
  constexpr uint32_t sm = 0x400;
  const int lane = lane_id();
  uint32_t use = a.use, chunk = a.chunk;
  const uint32_t wlu = a.wlu;
  const float* cwin = a.cwin;
  const int cside = a.cside;
  uint32_t has = 0;
#pragma unroll
  for (int n = 0; n < 4; ++n)
    has |= (__uint_as_float(lds_u32(POOLED ? wlu + 176 + n * 4 : wlu + 96 + n * 16)) != 0.f ? 1u : 0u) << n;
  constexpr int NCH = POOLED ? 4 : 16;
  F2 ch[NCH];
  auto load_children = [&](int kc) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      if constexpr (POOLED) {  // pool row of node n = (ny, nx): row of node a + 16 ny + nx
        ch[n] = (has >> n & 1u) ? ldcg_f2(cwin + ((n >> 1) * 16 + (n & 1)) * 128 + kc * 32) : f2_pack(0.f, 0.f);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = 2 * (n >> 1) + (k >> 1), c = 2 * (n & 1) + (k & 1);
#ifdef EG_PD_NOCHILD
          ch[n * 4 + k] = f2_pack(0.f, 0.f);
#else
          ch[n * 4 + k] = (has >> n & 1u) ? ldcg_f2(cwin + ((long long)r * cside + c) * 128 + kc * 32) : f2_pack(0.f, 0.f);
#endif
        }
      }
    }
  };
  load_children(0);
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
    const uint32_t rs = use % kRawStagesT;
    mbar_wait_a(sm + kOffRawFullT + rs * 8, (use / kRawStagesT) & 1u);
    const uint32_t ro = rs * kRawBytesT;
    F2 aa = f2_pack(0.f, 0.f), ab = aa, ac = aa, ad = aa;
    if constexpr (POOLED) {
      const float4 wp = lds4(wlu + 176);
      fma2(aa, wp.x, ch[0]), fma2(ab, wp.y, ch[1]), fma2(ac, wp.z, ch[2]), fma2(ad, wp.w, ch[3]);
    } else {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float4 wc = lds4(wlu + 96 + n * 16);
        F2& acc = n == 0 ? aa : n == 1 ? ab : n == 2 ? ac : ad;
        fma2(acc, wc.x, ch[n * 4]), fma2(acc, wc.y, ch[n * 4 + 1]), fma2(acc, wc.z, ch[n * 4 + 2]), fma2(acc, wc.w, ch[n * 4 + 3]);
      }
    }
    {
      const uint32_t pa = a.pb + ro;
      const F2 p01 = lds_f2(pa + (0 * 18 + 1) * 128), p02 = lds_f2(pa + (0 * 18 + 2) * 128);
      const F2 p10 = lds_f2(pa + (1 * 18 + 0) * 128), p11 = lds_f2(pa + (1 * 18 + 1) * 128);
      const F2 p12 = lds_f2(pa + (1 * 18 + 2) * 128), p13 = lds_f2(pa + (1 * 18 + 3) * 128);
      const F2 p20 = lds_f2(pa + (2 * 18 + 0) * 128), p21 = lds_f2(pa + (2 * 18 + 1) * 128);
      const F2 p22 = lds_f2(pa + (2 * 18 + 2) * 128), p23 = lds_f2(pa + (2 * 18 + 3) * 128);
      const F2 p31 = lds_f2(pa + (3 * 18 + 1) * 128), p32 = lds_f2(pa + (3 * 18 + 2) * 128);
      const F2 pq = lds_f2(a.qb + ro);
      {
        const float4 w0 = lds4(wlu), w1 = lds4(wlu + 16), w2 = lds4(wlu + 32);  // wl[0][0..5], wl[1][0..5]
        fma2(aa, w0.x, p01), fma2(aa, w0.y, p10), fma2(aa, w0.z, p12), fma2(aa, w0.w, p21), fma2(aa, w1.x, pq), fma2(aa, w1.y, p11);
        fma2(ab, w1.z, p02), fma2(ab, w1.w, p11), fma2(ab, w2.x, p13), fma2(ab, w2.y, p22), fma2(ab, w2.z, pq), fma2(ab, w2.w, p12);
      }
      {
        const float4 w3 = lds4(wlu + 48), w4 = lds4(wlu + 64), w5 = lds4(wlu + 80);  // wl[2][0..5], wl[3][0..5]
        fma2(ac, w3.x, p11), fma2(ac, w3.y, p20), fma2(ac, w3.z, p22), fma2(ac, w3.w, p31), fma2(ac, w4.x, pq), fma2(ac, w4.y, p21);
        fma2(ad, w4.z, p12), fma2(ad, w4.w, p21), fma2(ad, w5.x, p23), fma2(ad, w5.y, p32), fma2(ad, w5.z, pq), fma2(ad, w5.w, p22);
      }
    }
    asm volatile("" : "+l"(aa.u), "+l"(ab.u), "+l"(ac.u), "+l"(ad.u) : : "memory");  // the slot's loads have landed
    __syncwarp();
    if (lane == 0) mbar_arrive_a(sm + kOffRawEmptyT + rs * 8);
    ++use;
    if (kc < 3) load_children(kc + 1);
    const uint32_t stage = chunk % kStages;
    mbar_wait_a(sm + kOffEmptyT + stage * 8, ((chunk / kStages) & 1u) ^ 1u);
    const uint32_t a_hi = sm + stage * 2 * kTileBytes;
    emit_pair(a_hi, a.so_a, aa);
    emit_pair(a_hi, a.so_b, ab);
    emit_pair(a_hi, a.so_a + 2048, ac);
    emit_pair(a_hi, a.so_b + 2048, ad);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive_a(sm + kOffFullT + stage * 8);
    ++chunk;
    if (a.agg_a) {
      float* o = a.agg_a + kc * 32;
      st_f2(o, aa);
      st_f2(o + 128, ab);
      st_f2(o + (long long)a.side * 128, ac);
      st_f2(o + (long long)a.side * 128 + 128, ad);
    }
  }
  a.use = use;
  a.chunk = chunk;
}

// EPI: the epilogue applies a per-feature affine map and an optional ReLU to the accumulator before the addend -- an
// eval-mode GNN layer (conv bias + BatchNorm with running statistics + activation + residual) in ONE launch.  A separate
// instance, so the training kernels' code is exactly what it was.
template <int MODE, bool EPI = false>
__device__ __forceinline__ void tc_body(const TcParams& p, const PatchMaps* pm) {
  constexpr bool GATHER = MODE != kLinear;
  using L = Lay<MODE>;
  constexpr int kRawStages = L::kRawStages, kRawRows = L::kRawRows;
  constexpr uint32_t kRawBytes = L::kRawBytes, kPlanWarpBytes = L::kPlanWarpBytes, kPlanBytes = L::kPlanBytes;
  constexpr uint32_t kOffRaw = L::kOffRaw, kOffPlan = L::kOffPlan, kOffFull = L::kOffFull, kOffEmpty = L::kOffEmpty,
                     kOffRawFull = L::kOffRawFull, kOffRawEmpty = L::kOffRawEmpty;
  (void)kRawRows, (void)kOffPlan, (void)kPlanWarpBytes, (void)pm;
  // Patch mode walks UNITS (PatchPlan: runs of tiles one SM processes back to back), dealt round-robin; every role walks
  // the same sequence.  `t` is read one step ahead of its use wherever a role needs it early.
  struct PIter {
    int u, i, i1, b, t;
  };
  const int units_total = MODE == kPatch ? p.batch * p.patch.units_per_frame : 0;
  // (iterators are passed and returned BY VALUE: taken by reference they lived in local memory, and the MMA issuer and
  // the epilogue re-loaded them between tiles)
  auto pit_at = [&](int u) -> PIter {
    PIter it{u, 0, 0, 0, 0};
    if (u < units_total) {
      it.b = u / p.patch.units_per_frame;
      const int uf = u - it.b * p.patch.units_per_frame;
      it.i = __ldg(p.patch.unit_off + uf);
      it.i1 = __ldg(p.patch.unit_off + uf + 1);
      it.t = __ldg(p.patch.seq + it.i);
    }
    return it;
  };
  auto pit_begin = [&]() -> PIter { return pit_at((int)blockIdx.x); };
  auto pit_valid = [&](PIter it) { return it.u < units_total; };
  auto pit_next = [&](PIter it) -> PIter {
    if (it.i + 1 < it.i1) {
      ++it.i;
      it.t = __ldg(p.patch.seq + it.i);
      return it;
    }
    return pit_at(it.u + (int)gridDim.x);
  };
  (void)units_total, (void)pit_begin, (void)pit_valid, (void)pit_next;
  // The MMA issuer and the epilogue walk "virtual tile indices" b * tiles_per_frame + t (-1 = done): the running tile
  // index in linear / gather mode, the unit sequence in patch mode (state `wit`, owned by the role).
  auto walk_tile = [&](PIter wit, int v) -> int {
    if constexpr (MODE == kPatch) return pit_valid(wit) ? 0 : -1;  // (patch mode: only validity; frame / tile come from wit)
    return v < p.num_tiles ? v : -1;
  };
  (void)walk_tile;

#ifdef EG_TC_TIMING
  long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long dbg_n = 0;
  const long long dbg_t0 = clock64();
#endif
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned base in the shared address space (the swizzle patterns are functions of the shared address)
  // The kernel has no static shared memory and is never launched in a cluster, so its dynamic shared memory starts
  // right after the 1 KB the system reserves per CTA: shared-window address 0x400, already 1024-byte aligned.
  // Using the literal turns every ring / barrier address below into an immediate (the compiler otherwise keeps the
  // base in a register that the register-tight lattice loop spilled to local memory: 2 % of the kernel, r01j).
  // Checked once per thread; a different layout traps (launch failure) instead of corrupting memory.
  constexpr uint32_t sm = 0x400;
  if (((smem_u32(smem_raw) + 1023u) & ~1023u) != sm) __trap();
  uint8_t* smem = smem_raw + (sm - smem_u32(smem_raw));
  uint8_t* sA = smem;                     // [stage][hi|lo][16 KB]
  uint8_t* sRaw = sA + kABytes;           // [raw stage][kRawRows][128 B]
  uint8_t* sPlan = sRaw + kRawStages * kRawBytes;  // [compute warp][2][kPlanWarpBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPlan + kPlanBytes);
  uint64_t* full = bars;                  // [kStages]     compute -> MMA
  uint64_t* empty = full + kStages;       // [kStages]     MMA -> compute
  uint64_t* raw_full = empty + kStages;   // [kRawStages]  loaders (cp.async completion) -> compute
  uint64_t* raw_empty = raw_full + kRawStages;   // [kRawStages]  compute -> loaders
  uint64_t* acc_full = raw_empty + kRawStages;   // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;     // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Tiles are dealt round-robin: at any time the 148 CTAs work on ~148 consecutive tiles of the SAME frame,
  // whose 37 MB of rows stay in the 126 MB L2 for the halo / parent / children re-reads (measured: contiguous
  // per-CTA ranges raise the DRAM reads of the forward kernel from 3.5 GB to 5.6 GB at batch 64).

  // ---- one-time setup -----------------------------------------------------------------------------------
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], kProdWarps);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kRawStages; ++s) {
      // gather / linear: one cp.async.mbarrier.arrive.noinc per loader thread; patch: the producer's expect_tx arrival
      mbar_init(&raw_full[s], MODE == kPatch ? 1 : kLoadWarps * 32);
      mbar_init(&raw_empty[s], kProdWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
    }
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < kEpiWarps) {
    // weight -> TMEM: thread (warp q, lane t) owns output feature f = 32 q + t = TMEM lane f and writes
    // Wop[f][k] = trans_w ? W[f][k] : W[k][f] for k = 0..127 as tf32 hi (columns k) and lo (columns 128 + k)
    const int f = warp * 32 + lane;
#pragma unroll 1
    for (int slab = 0; slab < 4; ++slab) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int kk = slab * 32 + k;
        split_tf32(__ldg(p.W + (p.trans_w ? f * 128 + kk : kk * 128 + f)), hi[k], lo[k]);
      }
      const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + slab * 32;
      tmem_st32(ta, hi);
#ifdef EG_TF32X3
      tmem_st32(ta + 128, lo);
#else
      // correction operand, bf16, K = 64 per chunk in the order the compute warps emit (see emit / emit2): a group of
      // G features (G = 4: float4 lanes, G = 2: patch mode) occupies 2 G consecutive K slots, first the slots that meet
      // the node tile's LOW parts (they carry W_hi), then the slots that meet its HIGH parts (they carry W_lo).
      // Column c holds K slots 2 c (low half) and 2 c + 1.
      constexpr int G = MODE == kPatch ? 2 : 4;
      uint32_t cr[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kap = 2 * c + e, grp = kap / (2 * G), i = kap % (2 * G);
          const int feat = grp * G + (i % G);
          v[e] = i < G ? __uint_as_float(hi[feat]) : __uint_as_float(lo[feat]);
        }
        cr[c] = pack_bf16x2(v[0], v[1]);
      }
      tmem_st32(ta + 128, cr);
#endif
    }
    tmem_st_wait();
  }
  if constexpr (MODE == kPatch) {
    // The parent region of a slot is only written by tiles that have parents; its weights are zero otherwise, but
    // 0 x (uninitialised shared memory) may be NaN: start from zeros (from then on it holds zeros or finite rows).
    if (warp >= kProdWarp0) {
      for (int i = tid - kProdWarp0 * 32; i < kRawStages * kPatchQRows * 8; i += kProdWarps * 32) {
        const int slot = i / (kPatchQRows * 8), o = i % (kPatchQRows * 8);
        sts4(sm + kOffRaw + slot * kRawBytes + kPatchPRows * 128 + o * 16, make_uint4(0, 0, 0, 0));
      }
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // one setmaxnreg per warpgroup: warps 0-7 release registers, warps 8-23 take them
  if (warp >= kProdWarp0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MODE == kPatch ? kRegsPatch.compute : kRegsDefault.compute));
  } else if (warp < kEpiWarps) {
    if constexpr ((MODE == kPatch ? kRegsPatch.epi : kRegsDefault.epi) < 80)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MODE == kPatch ? kRegsPatch.epi : kRegsDefault.epi));
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MODE == kPatch ? kRegsPatch.load : kRegsDefault.load));
  }
#ifdef EG_TC_TIMING
  dbg_acc[6] = clock64() - dbg_t0;  // one-time setup (TMEM allocation, weight -> TMEM, barriers)
#endif
  if (warp >= kProdWarp0) {
    // ===== compute warps: copy raw rows two chunks ahead, gather -> split -> swizzled operand tile ==========
    const int pw = warp - kProdWarp0;
    const int g = lane >> 3, j = lane & 7;
    constexpr int kIters = kRowsPerProd / 4;
    const uint32_t plan_u = sm + kOffPlan + pw * 2 * kPlanWarpBytes;  // this warp's two plan buffers
    const uint32_t lane_raw = kOffRaw + j * 16;  // this lane's 16 bytes of a staged 128-byte row slice
    uint32_t soff[kIters];  // swizzled position of this lane's 16 bytes inside an operand tile
#pragma unroll
    for (int i = 0; i < kIters; ++i) {
      soff[i] = sw128_off(pw * kRowsPerProd + i * 4 + g, j);
      asm volatile("" : "+r"(soff[i]));  // opaque: kept in a register instead of being recomputed from tid per chunk
    }
    // The warp's plan rows (8 x 80 B, contiguous) + the tile header travel global -> shared with cp.async one
    // tile ahead, so a tile starts with LDS instead of an exposed L2 round trip.
    auto prefetch_plan = [&](int tile, uint32_t buf) {
      if (tile >= p.num_tiles) return;
      const int t = tile % p.tiles_per_frame;
      const uint32_t dst = plan_u + buf * kPlanWarpBytes;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.plan.rows + (size_t)t * 128 + pw * kRowsPerProd);
      cp_async16(dst + lane * 16, src + lane * 16);
      if (lane < 8) cp_async16(dst + (32 + lane) * 16, src + (32 + lane) * 16);
      if (lane == 8) cp_async16(dst + kRowsPerProd * sizeof(PlanRow), p.plan.hdr + t);
    };

    uint32_t chunk = 0;
    // waits for the raw stage and a free operand stage of `chunk`; returns the shared addresses of both
    auto acquire = [&](uint32_t& raw, uint32_t& a_hi) {
      const uint32_t stage = chunk % kStages, phase = (chunk / kStages) & 1u;
      const uint32_t rs = chunk % kRawStages, rphase = (chunk / kRawStages) & 1u;
#ifdef EG_TC_TIMING
      const long long _t = clock64();
      mbar_wait_a(sm + kOffRawFull + rs * 8, rphase);
      dbg_acc[1] += clock64() - _t;
      const long long _t2 = clock64();
      mbar_wait_a(sm + kOffEmpty + stage * 8, phase ^ 1u);
      dbg_acc[0] += clock64() - _t2;
#else
      mbar_wait_a(sm + kOffRawFull + rs * 8, rphase);
      mbar_wait_a(sm + kOffEmpty + stage * 8, phase ^ 1u);
#endif
      raw = sm + rs * kRawBytes;
      a_hi = sm + stage * 2 * kTileBytes;
    };
    auto release = [&]() {
#ifdef EG_TC_TIMING
      const long long _tr = clock64();
      fence_proxy_async_smem();
      dbg_acc[2] += clock64() - _tr;
#elif !defined(EG_DBG_NOFENCE)
      fence_proxy_async_smem();
#endif
#ifdef EG_TC_TIMING
      const long long _ta = clock64();
#endif
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(sm + kOffFull + (chunk % kStages) * 8);
        mbar_arrive_a(sm + kOffRawEmpty + (chunk % kRawStages) * 8);
      }
      ++chunk;
#ifdef EG_TC_TIMING
      dbg_acc[5] += clock64() - _ta;
#endif
    };
    auto emit = [&](uint32_t a_hi, int i, const float4& acc) {  // 3xTF32 split -> operand tiles
#ifdef EG_DBG_NOEMIT
      if (acc.x == 123.456f) sts4(a_hi + soff[i], make_uint4(0, 0, 0, 0));
      return;
#endif
      uint4 hi, lo;
      split_tf32_op(acc.x, hi.x, lo.x);
      split_tf32_op(acc.y, hi.y, lo.y);
      split_tf32_op(acc.z, hi.z, lo.z);
      split_tf32_op(acc.w, hi.w, lo.w);
      sts4(a_hi + soff[i], hi);
#ifdef EG_TF32X3
      sts4(a_hi + kTileBytes + soff[i], lo);
#else
      // correction tile (bf16, same 16 bytes): the four low parts, then the four values themselves
      sts4(a_hi + kTileBytes + soff[i],
           make_uint4(pack_bf16x2(__uint_as_float(lo.x), __uint_as_float(lo.y)),
                      pack_bf16x2(__uint_as_float(lo.z), __uint_as_float(lo.w)), pack_bf16x2(acc.x, acc.y),
                      pack_bf16x2(acc.z, acc.w)));
#endif
    };

    if constexpr (MODE == kLinear) {
      // linear mode: raw slot r = tile row r (rows past the end were not copied: zero them)
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < 4; ++kc) {
          uint32_t raw, a_hi;
          acquire(raw, a_hi);
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const int r = pw * kRowsPerProd + i * 4 + g;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((long long)tile * 128 + r < p.rows) acc = lds4(raw + lane_raw + r * 128);
            emit(a_hi, i, acc);
          }
          release();
        }
      }
    } else if constexpr (MODE == kGather) {
      const bool agg_out = p.AggOut != nullptr;
      uint32_t pbuf = 0;
      prefetch_plan(blockIdx.x, 0);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, pbuf ^= 1u) {
        const int b = tile / p.tiles_per_frame;
        const int t = tile - b * p.tiles_per_frame;
        const uint32_t prow = plan_u + pbuf * kPlanWarpBytes + g * sizeof(PlanRow);  // + i * 4 rows
        const long long frow0 = (long long)b * p.nodes_per_frame;
        const float* fbase = p.X + frow0 * 128;
        const int32_t* tnode = p.tile_nodes + t * 128 + pw * kRowsPerProd + g;
        auto store_agg = [&](int i, int coff, const float4& acc) {
          const int node = __ldg(tnode + i * 4);
          if (node >= 0) EG_ST_AGG(p.AggOut + (frow0 + node) * 128 + coff, acc);
        };
#ifdef EG_TC_TIMING
        const long long _tp = clock64();
#endif
        cp_async_wait_all();  // this tile's plan rows (prefetched one tile ahead)
        __syncwarp();
        uint32_t raw, a_hi;
        const float4 hdr = lds4(plan_u + pbuf * kPlanWarpBytes + kRowsPerProd * sizeof(PlanRow));
        const int ks = __float_as_int(hdr.y), nfar = __float_as_int(hdr.z), has_csr = __float_as_int(hdr.w);
        if (ks <= 5 && !nfar && !has_csr) {
          // ---- lattice tile (every main-level tile): <= 5 staged neighbours + the row itself, all in shared
          // memory.  Slot byte offsets and weights sit in registers for the 4 chunks; the loop body is
          // 6 LDS.128 + 12 FFMA2 + split + 2 STS.128 per row.
          uint32_t so[kIters][6];
          float wv[kIters][6];
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const float4 a = lds4(prow + i * 4 * sizeof(PlanRow));
            const float4 wa = lds4(prow + i * 4 * sizeof(PlanRow) + 16);
            const float4 wb = lds4(prow + i * 4 * sizeof(PlanRow) + 32);
            const uint32_t s_lo = __float_as_uint(a.x), s_hi = __float_as_uint(a.y);
            so[i][0] = (__byte_perm(s_lo, 0, 0x4440) << 7) + lane_raw;
            so[i][1] = (__byte_perm(s_lo, 0, 0x4441) << 7) + lane_raw;
            so[i][2] = (__byte_perm(s_lo, 0, 0x4442) << 7) + lane_raw;
            so[i][3] = (__byte_perm(s_lo, 0, 0x4443) << 7) + lane_raw;
            so[i][4] = (__byte_perm(s_hi, 0, 0x4440) << 7) + lane_raw;
            so[i][5] = (__byte_perm(s_hi, 0, 0x4443) << 7) + lane_raw;  // the row itself
            wv[i][0] = wa.x, wv[i][1] = wa.y, wv[i][2] = wa.z, wv[i][3] = wa.w, wv[i][4] = wb.x, wv[i][5] = wb.w;
          }
          prefetch_plan(tile + gridDim.x, pbuf ^ 1u);
#ifdef EG_TC_TIMING
          dbg_acc[4] += clock64() - _tp;
#endif
#pragma unroll 1
          for (int kc = 0; kc < 4; ++kc) {
#ifdef EG_TC_TIMING
            const long long _tl = clock64();
#endif
            acquire(raw, a_hi);
#ifdef EG_DBG_NOCOMPUTE
            release();
            continue;
#endif
#ifdef EG_TC_TIMING
            const long long _tg = clock64();
#endif
            float4 x[kIters][6];
#pragma unroll
            for (int i = 0; i < kIters; ++i)
#pragma unroll
              for (int k = 0; k < 6; ++k) {
#ifdef EG_DBG_NOGATHER
                x[i][k] = k == 5 ? lds4(raw + so[i][k]) : make_float4(1.f, 2.f, 3.f, 4.f);
#else
                x[i][k] = lds4(raw + so[i][k]);
#endif
              }
            float4 aggv[kIters];
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int k = 0; k < 6; ++k) fma4_x2(acc, wv[i][k], x[i][k]);  // plan order, self loop last
              emit(a_hi, i, acc);
              aggv[i] = acc;
            }
#ifdef EG_TC_TIMING
            dbg_acc[3] += clock64() - _tg;  // lattice tiles: gather + split + operand stores of one chunk
            ++dbg_n;
#endif
            release();
            if (agg_out) {  // A_hat dH side output: stored after the chunk is handed to the MMA
#pragma unroll
              for (int i = 0; i < kIters; ++i) store_agg(i, kc * 32 + j * 4, aggv[i]);
            }
#ifdef EG_TC_TIMING
            dbg_acc[7] += clock64() - _tl;
#endif
          }
          continue;
        }
        // ---- general tile (aux levels, hubs, ragged lattices): up to 7 staged neighbours, 4 far neighbours read from
        // global (the 2x2 children of an aux node: 512 rows per tile, not staged), CSR rows (hubs).  A dedicated class
        // for aux lattice tiles (all 8 child loads of a chunk in flight, FFMA2, plan re-read from shared memory) was
        // measured equal in the forward and 1-3 % slower in the backward (r01h/i) and removed.
        uint2 slots[kIters];
        float4 w0[kIters], w1[kIters];
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const float4 a = lds4(prow + i * 4 * sizeof(PlanRow));
          slots[i] = make_uint2(__float_as_uint(a.x), __float_as_uint(a.y));
          w0[i] = lds4(prow + i * 4 * sizeof(PlanRow) + 16);
          w1[i] = lds4(prow + i * 4 * sizeof(PlanRow) + 32);
        }
        prefetch_plan(tile + gridDim.x, pbuf ^ 1u);
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const int coff = kc * 32 + j * 4;
          float4 fx[4];
          if (nfar) {  // far rows of the first row group: in flight across the barrier waits
            const float4 fn = lds4(prow + 48);
            fx[0] = ld_far(fbase + (long long)__float_as_int(fn.x) * 128 + coff);
            fx[1] = ld_far(fbase + (long long)__float_as_int(fn.y) * 128 + coff);
            fx[2] = ld_far(fbase + (long long)__float_as_int(fn.z) * 128 + coff);
            fx[3] = ld_far(fbase + (long long)__float_as_int(fn.w) * 128 + coff);
          }
          acquire(raw, a_hi);
          const uint32_t rawl = raw + lane_raw;
          float4 aggv[kIters];
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t s_lo = slots[i].x, s_hi = slots[i].y;
            asm volatile("" : "+r"(s_lo), "+r"(s_hi));  // keep the slot -> address arithmetic inside the chunk loop
            // staged neighbours 0..4 (+ 5, 6 on tiles that have them): loads first, then the sums in plan
            // order (unused entries of a row point at slot 0 with weight 0)
            float4 x0 = lds4(rawl + (__byte_perm(s_lo, 0, 0x4440) << 7));
            float4 x1 = lds4(rawl + (__byte_perm(s_lo, 0, 0x4441) << 7));
            float4 x2 = lds4(rawl + (__byte_perm(s_lo, 0, 0x4442) << 7));
            float4 x3 = lds4(rawl + (__byte_perm(s_lo, 0, 0x4443) << 7));
            float4 x4 = lds4(rawl + (__byte_perm(s_hi, 0, 0x4440) << 7));
            fma4(acc, w0[i].x, x0);
            fma4(acc, w0[i].y, x1);
            if (ks > 5) {
              x0 = lds4(rawl + (__byte_perm(s_hi, 0, 0x4441) << 7));
              x1 = lds4(rawl + (__byte_perm(s_hi, 0, 0x4442) << 7));
            }
            fma4(acc, w0[i].z, x2);
            fma4(acc, w0[i].w, x3);
            fma4(acc, w1[i].x, x4);
            x2 = lds4(rawl + (__byte_perm(s_hi, 0, 0x4443) << 7));  // the row itself
            if (ks > 5) {
              fma4(acc, w1[i].y, x0);
              fma4(acc, w1[i].z, x1);
            }
            if (nfar) {
              const float4 fw = lds4(prow + i * 4 * sizeof(PlanRow) + 64);
              fma4(acc, fw.x, fx[0]);
              fma4(acc, fw.y, fx[1]);
              fma4(acc, fw.z, fx[2]);
              fma4(acc, fw.w, fx[3]);
              if (i + 1 < kIters) {
                const float4 fn = lds4(prow + (i + 1) * 4 * sizeof(PlanRow) + 48);
                fx[0] = ld_far(fbase + (long long)__float_as_int(fn.x) * 128 + coff);
                fx[1] = ld_far(fbase + (long long)__float_as_int(fn.y) * 128 + coff);
                fx[2] = ld_far(fbase + (long long)__float_as_int(fn.z) * 128 + coff);
                fx[3] = ld_far(fbase + (long long)__float_as_int(fn.w) * 128 + coff);
              }
            }
            fma4(acc, w1[i].w, x2);  // self loop last
            if (has_csr) {  // hub rows etc.: summed from the device CSR (plan weights are zero)
              const float4 a = lds4(prow + i * 4 * sizeof(PlanRow));
              const int cbeg = __float_as_int(a.z), cdeg = __float_as_int(a.w);
              for (int e = cbeg; e < cbeg + cdeg; ++e)
                fma4(acc, __ldg(p.w + e), ldg4(fbase + (long long)__ldg(p.col + e) * 128 + coff));
            }
            emit(a_hi, i, acc);
            aggv[i] = acc;
          }
          release();
          if (agg_out) {
#pragma unroll
            for (int i = 0; i < kIters; ++i) store_agg(i, coff, aggv[i]);
          }
        }
      }
    } else {
      // ===== patch mode: a half-warp owns a 2 x 2 node block of the 8 x 16 patch, a lane 2 of the chunk's 32 features.
      // Per chunk: 12 haloed-patch rows + 1 parent row (13 LDS.64, immediate offsets from one base) feed the four sums
      // (24 FFMA2); patches with children add 16 child rows when the sub-stage of their block row arrives.
      const int h = lane >> 4, l16 = lane & 15;
      const int q = pw * 2 + h, by = q >> 3, bx = q & 7;
      auto op_off = [&](int r) {  // this lane's 8 bytes of tile row r inside a K-major SWIZZLE_128B operand tile
        return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((l16 >> 1) ^ (r & 7)) << 4) + (l16 & 1) * 8);
      };
      const int r_a = 32 * by + 2 * bx;  // tile rows of the block: a = r_a, b = r_a + 1, c = r_a + 16, d = r_a + 17
      uint32_t so_a = op_off(r_a), so_b = op_off(r_a + 1);
      asm volatile("" : "+r"(so_a), "+r"(so_b));
      const uint32_t pb = sm + kOffRaw + (2 * by * 18 + 2 * bx) * 128 + l16 * 8;  // P[0][0] of the block's 4 x 4 window
      const uint32_t qb = sm + kOffRaw + kPatchPRows * 128 + (by * 8 + bx) * 128 + l16 * 8;
      const bool agg_out = p.AggOut != nullptr;
      auto prefetch_patch = [&](int t, uint32_t buf) {
        if (t < 0) return;
        const uint32_t dst = plan_u + buf * kPlanWarpBytes;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.patch.blocks + (size_t)t * 32 + pw * 2);
        const uint8_t* tsrc = reinterpret_cast<const uint8_t*>(p.patch.tiles + t);
        if (lane < 24) cp_async16(dst + lane * 16, wsrc + lane * 16);  // 2 x 192 B of weights
        else if (lane < 28) cp_async16(dst + lane * 16, tsrc + (lane - 24) * 16);
      };
      uint32_t use = 0;  // raw-slot uses so far (the producer counts the same sequence)
      auto wait_raw = [&]() -> uint32_t {
        const uint32_t rs = use % kRawStages, rphase = (use / kRawStages) & 1u;
#ifdef EG_TC_TIMING
        const long long _t = clock64();
        mbar_wait_a(sm + kOffRawFull + rs * 8, rphase);
        dbg_acc[1] += clock64() - _t;
#else
        mbar_wait_a(sm + kOffRawFull + rs * 8, rphase);
#endif
        return rs * kRawBytes;
      };
      auto release_raw = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sm + kOffRawEmpty + (use % kRawStages) * 8);
        ++use;
      };
      auto wait_op = [&]() -> uint32_t {
        const uint32_t stage = chunk % kStages, phase = (chunk / kStages) & 1u;
#ifdef EG_TC_TIMING
        const long long _t = clock64();
        mbar_wait_a(sm + kOffEmpty + stage * 8, phase ^ 1u);
        dbg_acc[0] += clock64() - _t;
        ++dbg_n;
#else
        mbar_wait_a(sm + kOffEmpty + stage * 8, phase ^ 1u);
#endif
        return sm + stage * 2 * kTileBytes;
      };
      auto release_op = [&]() {
#ifdef EG_TC_TIMING
        const long long _t = clock64();
        fence_proxy_async_smem();
        dbg_acc[2] += clock64() - _t;
#else
        fence_proxy_async_smem();
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive_a(sm + kOffFull + (chunk % kStages) * 8);
        ++chunk;
      };
      auto emit2 = [&](uint32_t a_hi, uint32_t off, F2 v) { emit_pair(a_hi, off, v); };
      uint32_t pbuf = 0;
      // this SM's pool buffers: [2][kPoolRows][128] floats, alternating by unit (`par`), so the writers of the next unit
      // never meet a reader of the current one
      float* const pool_cta = p.pool + (size_t)blockIdx.x * 2 * kPoolRows * 128 + l16 * 2;
      // (forward launches only: measured -1.3 % there, +3..5 % on the backward launch, whose 4 U of streams need the L2)
      const uint64_t pool_policy = l2_policy(p.AggOut == nullptr);
      uint32_t par = 0;
      // (b, t): this tile; t_nxt: the next one (-1 = none), whose plan is prefetched during this tile; `ahead` runs
      // TWO tiles ahead, so that the table reads behind it are never waited for
      PIter ahead = pit_begin();
      int b = ahead.b, t = pit_valid(ahead) ? ahead.t : -1;
      ahead = pit_next(ahead);
      int b_nxt = ahead.b, t_nxt = pit_valid(ahead) ? ahead.t : -1;
      ahead = pit_next(ahead);
      prefetch_patch(t, 0);
      for (; t >= 0; pbuf ^= 1u, b = b_nxt, t = t_nxt, b_nxt = ahead.b, t_nxt = pit_valid(ahead) ? ahead.t : -1, ahead = pit_next(ahead)) {
        const long long frow0 = (long long)b * p.nodes_per_frame;
        cp_async_wait_all();  // this tile's weights and descriptor (prefetched one tile ahead)
        __syncwarp();
        const uint32_t pl = plan_u + pbuf * kPlanWarpBytes;
        const float4 d0 = lds4(pl + 384), d2 = lds4(pl + 416), d3 = lds4(pl + 432);
        const int cls = __float_as_int(d0.x), y0 = __float_as_int(d0.z), x0 = __float_as_int(d0.w);
        const int node0 = __float_as_int(d2.z), side = __float_as_int(d2.w);
#ifndef EG_PD_PLAINAUX
        if (cls == 1 || cls == 3) {  // patch with children: a function of its own (see patch_aux_tile)
          AuxTileArgs a;
          a.agg_a = agg_out ? p.AggOut + (frow0 + node0 + (long long)(y0 + 2 * by) * side + x0 + 2 * bx) * 128 + l16 * 2
                            : nullptr;
          a.wlu = pl + h * 192;
          a.pb = pb, a.qb = qb, a.so_a = so_a, a.so_b = so_b;
          a.side = side;
          a.use = use, a.chunk = chunk;
          prefetch_patch(t_nxt, pbuf ^ 1u);
          if (cls == 3) {
            // the pooled child sums of this patch were written by the main patches of its unit, by ALL compute warps of
            // this CTA: one CTA-scope fence + barrier among the 16 compute warps orders them before the reads below
            __threadfence_block();
            asm volatile("bar.sync 1, %0;" ::"n"(kProdWarps * 32) : "memory");
            a.cwin = pool_cta + ((size_t)par * kPoolRows + (2 * by) * 16 + 2 * bx) * 128;
            a.cside = 0;
            patch_aux_tile<true, kRawStages, kRawBytes, kOffRawFull, kOffRawEmpty, kOffFull, kOffEmpty>(a);
            par ^= 1u;
          } else {
            const int cy = __float_as_int(d2.x), cx = __float_as_int(d2.y);
            const int cnode0 = __float_as_int(d3.x), cside = __float_as_int(d3.y);
            a.cwin = p.X + (frow0 + cnode0 + (long long)(cy + 4 * by) * cside + cx + 4 * bx) * 128 + l16 * 2;
            a.cside = cside;
            patch_aux_tile<false, kRawStages, kRawBytes, kOffRawFull, kOffRawEmpty, kOffFull, kOffEmpty>(a);
          }
          use = a.use, chunk = a.chunk;
          continue;
        }
#endif
        if (cls == 2) {
          // ---- CSR tile (ragged small lattices, coordinate nodes): rows summed straight from the device CSR
          prefetch_patch(t_nxt, pbuf ^ 1u);
          const float* fbase = p.X + frow0 * 128 + l16 * 2;
#pragma unroll 1
          for (int kc = 0; kc < 4; ++kc) {
            const uint32_t a_hi = wait_op();
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int r = pw * 8 + h * 4 + i;
              const int node = __ldg(p.tile_nodes + t * 128 + r);
              F2 acc = f2_pack(0.f, 0.f);
              if (node >= 0) {
                const int e1 = __ldg(p.rowptr + node + 1);
                for (int e = __ldg(p.rowptr + node); e < e1; ++e)  // sources ascending, self loop last
                  fma2(acc, __ldg(p.w + e), ldg_f2(fbase + (long long)__ldg(p.col + e) * 128 + kc * 32));
                if (agg_out) st_f2(p.AggOut + (frow0 + node) * 128 + kc * 32 + l16 * 2, acc);
              }
              emit2(a_hi, op_off(r), acc);
            }
            release_op();
          }
          continue;
        }
        float wl[24];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const float4 v = lds4(pl + h * 192 + i * 16);
          wl[4 * i] = v.x, wl[4 * i + 1] = v.y, wl[4 * i + 2] = v.z, wl[4 * i + 3] = v.w;
        }
        prefetch_patch(t_nxt, pbuf ^ 1u);
        // member of a unit: the pooled sum of this block (= the child term of its parent) goes to the SM's pool buffer
        const int pool_rel = __float_as_int(d3.z);
        float* const pool_row = pool_rel >= 0 ? pool_cta + ((size_t)par * kPoolRows + pool_rel + by * 16 + bx) * 128 : nullptr;
        const uint32_t dvu = pl + h * 192 + 160;
        float* agg_a = agg_out ? p.AggOut + (frow0 + node0 + (long long)(y0 + 2 * by) * side + x0 + 2 * bx) * 128 + l16 * 2
                               : nullptr;
#pragma unroll 1
        for (int kc = 0; kc < 4; ++kc) {
          const uint32_t ro = wait_raw();
          const uint32_t pa = pb + ro;
#ifdef EG_PD_NOGATHER
#define lds_f2x(a) f2_pack(1.f, 2.f)
#else
#define lds_f2x(a) lds_f2(a)
#endif
          const F2 p01 = lds_f2x(pa + (0 * 18 + 1) * 128), p02 = lds_f2x(pa + (0 * 18 + 2) * 128);
          const F2 p10 = lds_f2x(pa + (1 * 18 + 0) * 128), p11 = lds_f2(pa + (1 * 18 + 1) * 128);
          const F2 p12 = lds_f2x(pa + (1 * 18 + 2) * 128), p13 = lds_f2x(pa + (1 * 18 + 3) * 128);
          const F2 p20 = lds_f2x(pa + (2 * 18 + 0) * 128), p21 = lds_f2x(pa + (2 * 18 + 1) * 128);
          const F2 p22 = lds_f2x(pa + (2 * 18 + 2) * 128), p23 = lds_f2x(pa + (2 * 18 + 3) * 128);
          const F2 p31 = lds_f2x(pa + (3 * 18 + 1) * 128), p32 = lds_f2x(pa + (3 * 18 + 2) * 128);
          const F2 pq = lds_f2x(qb + ro);
          F2 aa = f2_pack(0.f, 0.f), ab = aa, ac = aa, ad = aa;
          // order: up, left, right, down, parent, self (then the children)
          fma2(aa, wl[0], p01), fma2(aa, wl[1], p10), fma2(aa, wl[2], p12), fma2(aa, wl[3], p21), fma2(aa, wl[4], pq), fma2(aa, wl[5], p11);
          fma2(ab, wl[6], p02), fma2(ab, wl[7], p11), fma2(ab, wl[8], p13), fma2(ab, wl[9], p22), fma2(ab, wl[10], pq), fma2(ab, wl[11], p12);
          fma2(ac, wl[12], p11), fma2(ac, wl[13], p20), fma2(ac, wl[14], p22), fma2(ac, wl[15], p31), fma2(ac, wl[16], pq), fma2(ac, wl[17], p21);
          fma2(ad, wl[18], p12), fma2(ad, wl[19], p21), fma2(ad, wl[20], p23), fma2(ad, wl[21], p32), fma2(ad, wl[22], pq), fma2(ad, wl[23], p22);
#ifndef EG_PD_NOPOOLOUT
          if (pool_row) {  // sum_c dis[c] x[c] over the block's four nodes = the pooled child sum of their parent
            const float4 dv = lds4(dvu);
            F2 ps = f2_pack(0.f, 0.f);
            fma2(ps, dv.x, p11), fma2(ps, dv.y, p12), fma2(ps, dv.z, p21), fma2(ps, dv.w, p22);
            st_f2_keep(pool_row + kc * 32, ps, pool_policy);
          }
#endif
          asm volatile("" : "+l"(aa.u), "+l"(ab.u), "+l"(ac.u), "+l"(ad.u) : : "memory");  // the slot's loads have landed
          release_raw();
          const uint32_t a_hi = wait_op();
          emit2(a_hi, so_a, aa);
          emit2(a_hi, so_b, ab);
          emit2(a_hi, so_a + 2048, ac);
          emit2(a_hi, so_b + 2048, ad);
          release_op();
#ifndef EG_PD_NOAGG
          if (agg_out) {  // A_hat dH side output: stored after the chunk is handed to the MMA
            float* o = agg_a + kc * 32;
            st_f2(o, aa);
            st_f2(o + 128, ab);
            st_f2(o + (long long)side * 128, ac);
            st_f2(o + (long long)side * 128 + 128, ad);
          }
#endif
        }
      }
    }
  } else if (MODE == kPatch && warp >= kLoadWarp0) {
    // ===== patch mode producer: ONE thread issues the TMA box copies of every slot use (the other loader lanes idle)
    if constexpr (MODE == kPatch) {
      if (warp == kLoadWarp0 && lane == 0) {
        const uint32_t bar_raw_full = sm + kOffRawFull, bar_raw_empty = sm + kOffRawEmpty;
        uint32_t use = 0;
        auto slot_acquire = [&](uint32_t bytes, uint32_t& dst, uint32_t& bar) {
          const uint32_t rs = use % kRawStages, rphase = (use / kRawStages) & 1u;
#ifdef EG_TC_TIMING
          const long long _t = clock64();
          mbar_wait_a(bar_raw_empty + rs * 8, rphase ^ 1u);
          dbg_acc[0] += clock64() - _t;
#else
          mbar_wait_a(bar_raw_empty + rs * 8, rphase ^ 1u);
#endif
          bar = bar_raw_full + rs * 8;
          dst = sm + kOffRaw + rs * kRawBytes;
#ifdef EG_PD_NOTMA
          mbar_arrive_a(bar);
#else
          mbar_arrive_expect_tx_a(bar, bytes);
#endif
          ++use;
        };
        auto load_desc = [&](const PIter& it, int4& a, int4& b4) {
          if (!pit_valid(it)) return;
          const int4* d = reinterpret_cast<const int4*>(p.patch.tiles + it.t);
          a = __ldg(d), b4 = __ldg(d + 1);
        };
        int4 na = make_int4(2, 0, 0, 0), nb = na;
        PIter cur = pit_begin();
        load_desc(cur, na, nb);
        while (pit_valid(cur)) {
          const int4 ta = na, tb = nb;  // {cls, level, y0, x0} {qlevel, qy, qx, clevel}
          const int b = cur.b;
          cur = pit_next(cur);
          load_desc(cur, na, nb);  // the next tile's descriptor, in flight while this tile's copies are issued
          if (ta.x == 2) continue;
          const CUtensorMap* mp = &pm->m[3 * ta.y];
          const CUtensorMap* mq = tb.x >= 0 ? &pm->m[3 * tb.x + 1] : nullptr;
#pragma unroll 1
          for (int kc = 0; kc < 4; ++kc) {
            uint32_t dst, bar;
            slot_acquire(kPatchPRows * 128 + (mq ? kPatchQRows * 128 : 0), dst, bar);
#ifndef EG_PD_NOTMA
            tma_load_4d(dst, mp, kc * 32, ta.w - 1, ta.z - 1, b, bar);
            if (mq) tma_load_4d(dst + kPatchPRows * 128, mq, kc * 32, tb.z, tb.y, b, bar);
#endif
          }
        }
      }
    }
  } else if (warp >= kLoadWarp0) {
    // ===== loader warps: unique source rows of the tile, one K chunk per raw stage, cp.async ==================
    constexpr int kPer = kRawRows / (kLoadWarps * 4);  // source rows per thread (8 lanes x 16 B per row)
    const int q = (warp - kLoadWarp0) * 32 + lane;
    const int sl = q >> 3, j = q & 7;
    const uint32_t bar_raw_full = sm + kOffRawFull, bar_raw_empty = sm + kOffRawEmpty;
    const uint32_t dst0 = sm + kOffRaw + sl * 128 + j * 16;
    // frame-local node staged in slot sl + (kLoadWarps*4) * m (-1 = none), fetched one tile ahead
    auto load_src = [&](int tile, int (&node)[kPer]) {
      if (GATHER && tile < p.num_tiles) {
        const int t = tile % p.tiles_per_frame;
#pragma unroll
        for (int m = 0; m < kPer; ++m) node[m] = __ldg(p.plan.src + (size_t)t * kPlanSrc + sl + kLoadWarps * 4 * m);
      }
    };
    int nxt[kPer];
    load_src(blockIdx.x, nxt);
    uint32_t chunk = 0;
#if EG_PF_TILES > 0
    // L2 prefetch of the OWN rows of the tile EG_PF_TILES rounds ahead (its 16-row groups are contiguous 8 KB
    // segments: one bulk prefetch each, issued by 8 threads of the first loader warp).  The raw ring only holds
    // 3 chunks (0.75 tile) of loads in flight per SM, which does not cover the DRAM latency under load (ncu r02d:
    // 22 % of the compute warps' samples wait for a raw stage); with the rows already in L2 the ring only has to
    // cover the L2 latency.
    auto prefetch_tile = [&](int tile) {
      if (!GATHER || tile >= p.num_tiles || q >= 8) return;
      const int b = tile / p.tiles_per_frame, t = tile - b * p.tiles_per_frame;
      const int first = __ldg(p.tile_groups + t * 16 + q), cnt = __ldg(p.tile_groups + t * 16 + 8 + q);
      if (first < 0 || cnt <= 0) return;
      const float* src = p.X + ((long long)b * p.nodes_per_frame + first) * 128;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(cnt * 512) : "memory");
    };
#pragma unroll 1
    for (int d = 1; d < EG_PF_TILES; ++d) prefetch_tile(blockIdx.x + d * gridDim.x);
#endif
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
#if EG_PF_TILES > 0
      prefetch_tile(tile + EG_PF_TILES * gridDim.x);
#endif
      int srow[kPer];  // global row (< 2^31, checked by the launcher), or -1
      if (GATHER) {
        const int base = (tile / p.tiles_per_frame) * p.nodes_per_frame;
#pragma unroll
        for (int m = 0; m < kPer; ++m) srow[m] = nxt[m] >= 0 ? base + nxt[m] : -1;
        load_src(tile + gridDim.x, nxt);
      } else {
#pragma unroll
        for (int m = 0; m < kPer; ++m) {
          const int s2 = sl + kLoadWarps * 4 * m;
          const long long row = (long long)tile * 128 + s2;
          srow[m] = (s2 < 128 && row < p.rows) ? (int)row : -1;
        }
      }
      for (int kc = 0; kc < 4; ++kc, ++chunk) {
        const uint32_t rs = chunk % kRawStages, rphase = (chunk / kRawStages) & 1u;
#ifdef EG_TC_TIMING
        const long long _t = clock64();
        mbar_wait_a(bar_raw_empty + rs * 8, rphase ^ 1u);
        dbg_acc[0] += clock64() - _t;
#else
        mbar_wait_a(bar_raw_empty + rs * 8, rphase ^ 1u);
#endif
        const uint32_t dst = dst0 + rs * kRawBytes;
        const float* srcb = p.X + kc * 32 + j * 4;
#pragma unroll
        for (int m = 0; m < kPer; ++m) {
#ifdef EG_DBG_NOLOAD
          if (srow[m] >= 0 && m == 0) cp_async16(dst + kLoadWarps * 4 * m * 128, srcb + (long long)srow[m] * 128);
#else
          if (srow[m] >= 0) cp_async16(dst + kLoadWarps * 4 * m * 128, srcb + (long long)srow[m] * 128);
#endif
        }
        cp_async_mbar_arrive_a(bar_raw_full + rs * 8);
      }
    }
  } else if (warp == kMmaWarp) {
    const int lane = lane_id();  // fresh read: keeps the issuer's only use of the lane out of the kernel-wide live set
    // ===== MMA issuer ======================================================================================
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 0, 0);
    uint32_t chunk = 0, it = 0;
    // the issuer only needs the NUMBER of tiles this CTA processes (patch mode: summed over its units, once)
    uint32_t my_tiles = 0;
    if constexpr (MODE == kPatch) {
      // (the 32 lanes share the walk: done serially by one lane it took ~20 us during which nothing was issued)
      for (int u = blockIdx.x + lane * gridDim.x; u < units_total; u += 32 * gridDim.x) {
        const int uf = u % p.patch.units_per_frame;
        my_tiles += __ldg(p.patch.unit_off + uf + 1) - __ldg(p.patch.unit_off + uf);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) my_tiles += __shfl_xor_sync(0xffffffffu, my_tiles, o);
    } else {
      my_tiles = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    }
    for (; it < my_tiles; ++it) {
      const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
      TC_TIMED_WAIT(0, &acc_empty[buf], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + kTmemAcc + buf * 128;
      for (int kc = 0; kc < 4; ++kc, ++chunk) {
        const uint32_t stage = chunk % kStages, phase = (chunk / kStages) & 1u;
        TC_TIMED_WAIT(1, &full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b_hi = smem_u32(sA) + stage * 2 * kTileBytes, b_lo = b_hi + kTileBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {  // UMMA_K = 8 tf32: 8 TMEM columns of W, 32 B of the swizzle atom
            const uint32_t w_hi = tmem_base + kc * 32 + ks * 8, w_lo = w_hi + 128;
            const uint32_t o = ks * 32;
#ifndef EG_DBG_NOMMA
#ifdef EG_TF32X3
            umma_tf32_ts(d, w_hi, umma_desc_k128(b_lo + o), idesc, (kc | ks) != 0);
            umma_tf32_ts(d, w_lo, umma_desc_k128(b_hi + o), idesc, 1u);
            umma_tf32_ts(d, w_hi, umma_desc_k128(b_hi + o), idesc, 1u);
#else
            // W_hi x A_hi in tf32 (exact products), then both correction terms W_hi x A_lo + W_lo x A_hi as ONE bf16
            // MMA over 16 interleaved K slots: their operands only need 8 bits (each term is 2^-11 of the result)
            umma_bf16_ts(d, w_lo, umma_desc_k128(b_lo + o), umma_idesc_bf16(128, 128), (kc | ks) != 0);
            umma_tf32_ts(d, w_hi, umma_desc_k128(b_hi + o), idesc, 1u);
#endif
#endif
          }
          umma_commit(&empty[stage]);
          if (kc == 3) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp < kEpiWarps) {
    // ===== epilogue: thread <-> output feature; TMEM -> registers -> 128-byte row segments ===================
    // A 16-column slab of the accumulator = one 16-row group of the tile = consecutive rows of Out
    // (tile table contract), so a slab is addressed as one base pointer + immediate offsets.
    const int ew = warp;             // TMEM lanes [32 ew, 32 ew + 32)
    const int f = ew * 32 + lane;    // output feature owned by this thread
    const float bias = p.bias ? __ldg(p.bias + f) : 0.f;
    float ep_sc = 1.f, ep_sh = 0.f;
    if constexpr (EPI) ep_sc = __ldg(p.ep_scale + f), ep_sh = __ldg(p.ep_shift + f);
    auto ep_map = [&](float v) {
      const float y = fmaf(v, ep_sc, ep_sh);
      return p.ep_relu ? fmaxf(y, 0.f) : y;
    };
    (void)ep_sc, (void)ep_sh, (void)ep_map;
    double s_sum = 0.0, s_sq = 0.0;
    // row groups of a tile: lane l < 8 holds the first global row of group l (or -1), lane 8 + l its row count.
    // They are fetched ONE TILE AHEAD (the table read is a dependent global load that would otherwise be exposed
    // at the top of every tile).
    PIter wit = MODE == kPatch ? pit_begin() : PIter{0, 0, 0, 0, 0};  // patch mode: the tile `tile_rows` is asked about
    auto tile_rows = [&](int tile, long long& base, int& cnt) {
      base = -1;
      cnt = 0;
      if (tile < 0 || tile >= p.num_tiles) return;
      if (GATHER) {
        int b, t;
        if constexpr (MODE == kPatch) {  // (the walker knows frame and tile: no division)
          b = wit.b, t = wit.t;
        } else {
          b = tile / p.tiles_per_frame;
          t = tile - b * p.tiles_per_frame;
        }
        const int v = lane < 16 ? __ldg(p.tile_groups + t * 16 + lane) : 0;
        cnt = v;
        base = (lane < 8 && v >= 0) ? (long long)b * p.nodes_per_frame + v : -1;
      } else {
        const long long r0 = (long long)tile * 128 + (lane & 7) * 16;
        base = r0 < p.rows ? r0 : -1;
        cnt = (int)max(0LL, min(16LL, p.rows - r0));
      }
    };
    long long gbase, nbase;
    int gcnt, ncnt;
    int vt = (int)blockIdx.x;
    tile_rows(walk_tile(wit, vt), gbase, gcnt);
    uint32_t it = 0;
    for (int tile = walk_tile(wit, vt); tile >= 0; ++it) {
      const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
      if constexpr (MODE == kPatch) wit = pit_next(wit);
      vt += gridDim.x;
      tile = walk_tile(wit, vt);           // (from here on `tile` is the NEXT tile)
      tile_rows(tile, nbase, ncnt);        // consumed after the slab loop
      float adA[16], adB[16];
      auto slab_rows = [&](int sl, long long& base, int& cnt) {
        base = __shfl_sync(0xffffffffu, gbase, sl);
        cnt = __shfl_sync(0xffffffffu, gcnt, GATHER ? 8 + sl : sl);
      };
      auto load_res = [&](int sl, float (&ad)[16]) {
        long long base;
        int cnt;
        slab_rows(sl, base, cnt);
        const float* adp = p.addend + max(base, 0LL) * 128 + f;
#pragma unroll
#ifdef EG_PD_NORES
        for (int i = 0; i < 16; ++i) ad[i] = (float)i;
        (void)adp;
#else
        for (int i = 0; i < 16; ++i) ad[i] = i < cnt ? __ldg(adp + i * 128) : 0.f;
#endif
      };
      if (p.addend) load_res(0, adA);  // in flight across the wait for the accumulator
      TC_TIMED_WAIT(0, &acc_full[buf], acc_phase);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + kTmemAcc + buf * 128;
      if (p.addend) {
        // backward: Out = acc + bias + residual gradient.  The 16 residual loads of a slab are issued one slab
        // AHEAD (two register sets, ping-pong), so their L2 latency overlaps the previous slab's TMEM load and
        // stores instead of stalling the epilogue 8 times per tile (it was the bottleneck of the backward).
        auto finish = [&](int sl, const float (&ad)[16]) {
          uint32_t v[16];
          tmem_ld16(tacc + sl * 16, v);
          long long base;
          int cnt;
          slab_rows(sl, base, cnt);
          #ifdef EG_DBG_SMALLOUT
          float* out = p.Out + (max(base, 0LL) & 1023) * 128 + f;  // timing experiment: all stores hit the same 512 KB
#else
          float* out = p.Out + max(base, 0LL) * 128 + f;
#endif
          tmem_ld_wait();
          float s = 0.f, q = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < cnt) {  // warp-uniform predicate
              float o;
              if constexpr (EPI) o = ep_map(__uint_as_float(v[i])) + ad[i];
              else o = __uint_as_float(v[i]) + bias + ad[i];
              out[i * 128] = o;
              if (p.stat_parts) {  // (no caller asks for both today; kept for the ABI)
                s += o;
                q = fmaf(o, o, q);
              }
            }
          s_sum += (double)s;
          s_sq += (double)q;
        };
#pragma unroll 1
        for (int sl = 0; sl < 8; sl += 2) {
          load_res(sl + 1, adB);
          finish(sl, adA);
          if (sl + 2 < 8) load_res(sl + 2, adA);
          finish(sl + 1, adB);
        }
      } else {
#pragma unroll 1
#ifdef EG_DBG_NOEPI
        for (int sl = 0; sl < 0; ++sl) {
#else
        for (int sl = 0; sl < 8; ++sl) {  // 16 tile rows (accumulator columns) at a time
#endif
          uint32_t v[16];
          tmem_ld16(tacc + sl * 16, v);
          const long long base = __shfl_sync(0xffffffffu, gbase, sl);
          const int cnt = __shfl_sync(0xffffffffu, gcnt, GATHER ? 8 + sl : sl);
          #ifdef EG_DBG_SMALLOUT
          float* out = p.Out + (max(base, 0LL) & 1023) * 128 + f;  // timing experiment: all stores hit the same 512 KB
#else
          float* out = p.Out + max(base, 0LL) * 128 + f;
#endif
          float s = 0.f, q = 0.f;
          tmem_ld_wait();
          if constexpr (EPI) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < cnt) EG_ST_OUT(out + i * 128, ep_map(__uint_as_float(v[i])));  // warp-uniform predicate
          } else if (cnt == 16) {
            // two rows per packed instruction: o = v + bias, s += o, q += o * o  (24 instead of 48 issue slots)
            const F2 b2 = f2_pack(bias, bias);
            F2 s2 = f2_pack(0.f, 0.f), q2 = s2;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const F2 o2 = f2_add(F2{((unsigned long long)v[i + 1] << 32) | v[i]}, b2);
#ifndef EG_DBG_NOSTORE
              EG_ST_OUT(out + i * 128, f2_lo(o2));
              EG_ST_OUT(out + (i + 1) * 128, f2_hi(o2));
#endif
              s2 = f2_add(s2, o2);
              q2 = f2_fma(o2, o2, q2);
            }
            s = f2_lo(s2) + f2_hi(s2);
            q = f2_lo(q2) + f2_hi(q2);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float o = __uint_as_float(v[i]) + bias;
              if (i < cnt) {  // warp-uniform
                EG_ST_OUT(out + i * 128, o);
                s += o;
                q = fmaf(o, o, q);
              }
            }
          }
          s_sum += (double)s;
          s_sq += (double)q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (p.addend) {
        // residual rows of this CTA's NEXT tile -> L2 now, so that its slab loop reads them at L2 latency
#pragma unroll
        for (int m = 0; m < 4; ++m) {  // 8 groups x 64 lines of 128 B, 4 lines per epilogue thread
          const int line = f + 128 * m, grp = line >> 6, lo = line & 63;
          const long long base = __shfl_sync(0xffffffffu, nbase, grp);
          const int cnt = __shfl_sync(0xffffffffu, ncnt, GATHER ? 8 + grp : grp);
          if (base >= 0 && lo < cnt * 4) prefetch_l2(p.addend + base * 128 + lo * 32);
        }
      }
      gbase = nbase;
      gcnt = ncnt;
    }
    if (p.stat_parts) {
      p.stat_parts[(size_t)blockIdx.x * 256 + f] = s_sum;
      p.stat_parts[(size_t)blockIdx.x * 256 + 128 + f] = s_sq;
    }
  }

#ifdef EG_TC_TIMING
  if (lane == 0) {
    long long* d = g_tc_dbg[blockIdx.x];
    if (warp == kProdWarp0) { d[0] = dbg_acc[0]; d[5] = dbg_acc[1]; d[7] = dbg_acc[2]; d[8] = dbg_acc[3]; d[9] = dbg_n; d[10] = clock64() - dbg_t0; d[16] = dbg_acc[4]; d[17] = dbg_acc[5]; d[18] = dbg_acc[6]; d[19] = dbg_acc[7]; } // compute warp 0: wait operand stage, wait raw, fence; span
    if (warp == kProdWarp0 + kProdWarps - 1) { d[11] = dbg_acc[3]; d[12] = dbg_acc[0]; d[13] = dbg_acc[1]; d[14] = clock64() - dbg_t0; d[15] = dbg_acc[2]; }  // last compute warp
    if (warp == kLoadWarp0) d[6] = dbg_acc[0];                       // loader: wait for a free raw stage
    if (warp == kMmaWarp) { d[1] = dbg_acc[0]; d[2] = dbg_acc[1]; } // MMA: wait acc_empty, wait full
    if (warp == 0) { d[3] = dbg_acc[0]; d[4] = clock64() - dbg_t0; } // epilogue: wait acc_full; total cycles
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int MODE, bool EPI>
__global__ void __launch_bounds__(kThreads, 1) gcn_tc_kernel(const TcParams p) {
  tc_body<MODE, EPI>(p, nullptr);
}
template <bool EPI>
__global__ void __launch_bounds__(kThreads, 1) gcn_patch_kernel(const TcParams p, const __grid_constant__ PatchMaps pm) {
  tc_body<kPatch, EPI>(p, &pm);
}

template <int MODE, bool EPI = false>
int launch(const TcParams& p, const PatchMaps* pm, float* mean, float* var, void* ws, size_t ws_bytes, const char* name,
           cudaStream_t s) {
  const bool stats = mean && var;
  if (stats && (!ws || ws_bytes < kWorkspaceBytes)) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  static std::atomic<unsigned long long> attr_mask{0};  // per template instance, one bit per device
  if (first_use_on_current_device(attr_mask)) {
    if (MODE == kPatch)
      EG_CUDA(cudaFuncSetAttribute(gcn_patch_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<MODE>::kSmemBytes));
    else
      EG_CUDA(cudaFuncSetAttribute(gcn_tc_kernel<MODE == kPatch ? kGather : MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)Lay<MODE>::kSmemBytes));
  }
  const int sms = num_sms();
  const long long work = MODE == kPatch ? (long long)p.batch * p.patch.units_per_frame : (long long)p.num_tiles;
  const int grid = (int)(work < sms ? work : sms);
  TcParams q = p;
  q.stat_parts = stats ? reinterpret_cast<double*>(ws) : nullptr;
  {
    ProfileScope prof(name, s);
    if (MODE == kPatch)
      gcn_patch_kernel<EPI><<<grid, kThreads, Lay<MODE>::kSmemBytes, s>>>(q, *pm);
    else
      gcn_tc_kernel<MODE == kPatch ? kGather : MODE, EPI><<<grid, kThreads, Lay<MODE>::kSmemBytes, s>>>(q);
    EG_LAUNCH_CHECK();
  }
  if (stats) return launch_stats_finalize(grid, 128, 128, p.rows, q.stat_parts, mean, var, s);
  return EG_OK;
}

// ---- tensor maps of the node tensor (patch mode) ------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Level l of the node tensor X [batch * N, 128] as a 4-D tensor (feature, x, y, frame); boxes of 32 features.
int encode_patch_maps(const eg_graph_info& info, int batch, const float* X, PatchMaps& out) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return EG_ERR_CUDA;
  }
  static const cuuint32_t kBox[2][2] = {{18, 10}, {8, 4}};  // P, Q: (x, y) extent
  for (int l = 0; l < info.num_levels; ++l) {
    const cuuint64_t side = (cuuint64_t)info.level_size[l];
    const cuuint64_t dims[4] = {128, side, side, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {512, side * 512, (cuuint64_t)info.num_nodes * 512};
    void* base = const_cast<float*>(X) + (size_t)info.level_offset[l] * 128;
    for (int k = 0; k < 2; ++k) {
      const cuuint32_t box[4] = {32, kBox[k][0], kBox[k][1], 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      const CUresult r = enc(&out.m[3 * l + k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for level %d (side %d), box %d", (int)r, l, (int)side, k);
        return EG_ERR_CUDA;
      }
    }
  }
  return EG_OK;
}

// Plan selection: 0 = automatic (the patch plan wherever the graph allows it), 1 = always the gather plan.  Initialised
// from EG_GCN_PLAN=gather (development / A-B timing); eg_gcn_plan_select() switches it at run time (parity tests run
// both plans on the same inputs).
std::atomic<int> g_plan_mode{-1};
bool patch_plan_enabled() {
  int m = g_plan_mode.load(std::memory_order_relaxed);
  if (m < 0) {
    const char* e = getenv("EG_GCN_PLAN");
    m = (e && strcmp(e, "gather") == 0) ? 1 : 0;
    g_plan_mode.store(m, std::memory_order_relaxed);
  }
  return m == 0;
}

}  // namespace

extern "C" int eg_gcn_plan_select(int mode) {
  patch_plan_enabled();  // resolve the environment default first
  return g_plan_mode.exchange(mode ? 1 : 0, std::memory_order_relaxed);
}

#ifdef EG_TC_TIMING
extern "C" int eg_tc_debug_read(long long* out) {  // HOST buffer of kNumSMs * 16 counters
  return cudaMemcpyFromSymbol(out, g_tc_dbg, sizeof(long long) * kNumSMs * 20) == cudaSuccess ? 0 : -2;
}
#endif

namespace eg {

template <bool EPI>
static int launch_gcn_any(const eg_graph* g, int batch, const float* X, const float* W, int trans_w, const float* bias,
                          const float* addend, float* Out, float* AggOut, float* mean, float* var, void* ws,
                          size_t ws_bytes, const float* ep_scale, const float* ep_shift, int ep_relu, cudaStream_t s) {
  const eg_graph_info& info = graph_info(g);
  TcParams p{};
  p.ep_scale = ep_scale;
  p.ep_shift = ep_shift;
  p.ep_relu = ep_relu;
  p.tile_nodes = graph_tile_nodes(g);
  p.tile_groups = graph_tile_groups(g);
  p.plan = graph_plan(g);
  p.tiles_per_frame = graph_tiles_per_frame(g);
  p.nodes_per_frame = info.num_nodes;
  p.num_tiles = batch * p.tiles_per_frame;
  p.rows = (long long)batch * info.num_nodes;
  if (p.rows >= (1LL << 31)) {
    set_error("batch * num_nodes = %lld does not fit the 32-bit row index of the tensor-core kernels", p.rows);
    return EG_ERR_INVALID;
  }
  p.rowptr = graph_rowptr(g);
  p.col = graph_col(g);
  p.w = graph_w(g);
  p.X = X;
  p.W = W;
  p.trans_w = trans_w;
  p.bias = bias;
  p.addend = addend;
  p.Out = Out;
  p.AggOut = AggOut;
  const PatchPlan& pp = graph_patch_plan(g);
  if (pp.ok && patch_plan_enabled() && encode_tiled_fn() != nullptr) {  // (no tensor-map encoder in the driver: gather plan)
    p.patch = pp;
    p.batch = batch;
    if (int rc = graph_pool_scratch(g, s, &p.pool)) return rc;
    PatchMaps maps;
    if (int rc = encode_patch_maps(info, batch, X, maps)) return rc;
    return launch<kPatch, EPI>(p, &maps, mean, var, ws, ws_bytes, EPI ? "gcn_tc_eval" : AggOut ? "gcn_tc_bwd" : "gcn_tc_fwd", s);
  }
  return launch<kGather, EPI>(p, nullptr, mean, var, ws, ws_bytes, EPI ? "gcn_tc_eval" : AggOut ? "gcn_tc_bwd" : "gcn_tc_fwd", s);
}

// Out = (A_hat X) op(W) + bias + addend over the batched graph; AggOut (optional) receives A_hat X.
int launch_gcn_tc(const eg_graph* g, int batch, const float* X, const float* W, int trans_w, const float* bias,
                  const float* addend, float* Out, float* AggOut, float* mean, float* var, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
  return launch_gcn_any<false>(g, batch, X, W, trans_w, bias, addend, Out, AggOut, mean, var, ws, ws_bytes, nullptr,
                               nullptr, 0, s);
}

// Out = act((A_hat X) W^T * scale + shift) + addend: an eval-mode GNN layer in one launch (scale / shift: device
// float[128], the conv bias and the BatchNorm running statistics folded by the caller).
int launch_gcn_tc_eval(const eg_graph* g, int batch, const float* X, const float* W, const float* scale,
                       const float* shift, int relu, const float* addend, float* Out, cudaStream_t s) {
  return launch_gcn_any<true>(g, batch, X, W, 1, nullptr, addend, Out, nullptr, nullptr, nullptr, nullptr, 0, scale,
                              shift, relu, s);
}

// C = A op(W) + bias + addend (C may alias A: a tile is read completely before its epilogue writes it).
int launch_linear_tc(long long rows, const float* A, const float* W, int trans_w, const float* bias,
                     const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                     cudaStream_t s) {
  if (rows >= (1LL << 31)) {
    set_error("rows = %lld does not fit the 32-bit row index of the tensor-core kernels", rows);
    return EG_ERR_INVALID;
  }
  TcParams p{};
  p.num_tiles = (int)((rows + 127) / 128);
  p.rows = rows;
  p.X = A;
  p.W = W;
  p.trans_w = trans_w;
  p.bias = bias;
  p.addend = addend;
  p.Out = C;
  return launch<kLinear>(p, nullptr, mean, var, ws, ws_bytes, "linear_tc", s);
}

}  // namespace eg
