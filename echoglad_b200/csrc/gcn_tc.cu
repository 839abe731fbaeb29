// Fused message passing + dense transform on the 5th-generation tensor cores (tcgen05 / TMEM):
//
//     Out[tile rows] = (A_hat X)[tile rows] * op(W) (+ bias) (+ addend)        one persistent kernel
//
// Replaces PyG GCNConv = propagate(scatter_add) + nn.Linear (src/core/models.py:330,431) with the
// re-association A_hat (X W^T) = (A_hat X) W^T (4e-7 rms), so the aggregated rows never
// touch HBM: producer warps gather/weight/sum the neighbour rows of a 128-node tile (atomic-free,
// ascending source order, self loop last), split them into tf32 hi/lo parts and write them straight into
// swizzled shared-memory operand tiles; one elected thread issues 3xTF32 tcgen05.mma into a
// double-buffered TMEM accumulator; epilogue warps drain it (tcgen05.ld), add bias / residual gradient,
// accumulate the BatchNorm column statistics and store rows.  With GATHER = false the same pipeline is
// the plain per-node transform (classifier layer 0 and its input gradient).
//
// The product is computed TRANSPOSED, D^T[f][r] = sum_k Wop[f][k] * A[r][k]:
//   * the weight (hi and lo parts, 2 x 128 TMEM columns) is the M-side operand and lives in TENSOR MEMORY
//     for the whole kernel, so each MMA reads only the 4 KB node-tile slice from shared memory (half the
//     shared-memory traffic of an SS-mode MMA) and all of shared memory is a ring of operand stages;
//   * the accumulator has one output feature per TMEM lane and one tile row per column, so an epilogue
//     thread owns a feature: bias and the column statistics are per-thread scalars, and for a fixed row
//     the 32 lanes of a warp hold 32 consecutive features = one coalesced 128-byte store.  No staging.
//
// Work decomposition.  Tile = 128 output rows (8x16 lattice patches, see eg_graph::tile_nodes); K is
// consumed in 4 chunks of 32 features, one ring stage = [128 rows x 128 B] hi + lo = 32 KB, 6 stages, so
// the producers run up to 1.5 tiles ahead of the tensor core and the per-stage gather working set
// (~208 neighbour rows x 128 B) stays in L1.  TMEM: 256 columns of weight + 2 x 128 of accumulator.
// Warps: 0-3 epilogue (TMEM lane quadrant = warp id), 4 MMA issuer, 5.. producers.
#include "common.cuh"
#include "tc05.cuh"

struct eg_graph;
namespace eg {
const eg_graph_info& graph_info(const eg_graph* g);
const int32_t* graph_rowptr(const eg_graph* g);
const int32_t* graph_col(const eg_graph* g);
const float* graph_w(const eg_graph* g);
const int32_t* graph_tile_nodes(const eg_graph* g);
int graph_tiles_per_frame(const eg_graph* g);
int launch_stats_finalize(int nparts, int cols, int stride, long long rows, const double* parts, float* mean,
                          float* var, cudaStream_t s);
}  // namespace eg

using namespace eg;
using namespace eg::tc;

namespace {

constexpr int kStages = 4;  // 128 KB of operand ring; the rest of the 256 KB SM array stays L1 for the gather
constexpr int kProdWarps = 16;             // 8 tile rows per producer warp and stage
constexpr int kRowsPerProd = 128 / kProdWarps;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;
constexpr int kTileBytes = 128 * 128;      // one [128 rows x 32 tf32] operand tile
constexpr uint32_t kABytes = kStages * 2 * kTileBytes;
constexpr uint32_t kSmemBytes = kABytes + 256 /*barriers*/ + 1024 /*align*/;
constexpr int kTmemCols = 512;             // [0,128) W hi, [128,256) W lo, [256,384) / [384,512) accumulators
constexpr uint32_t kTmemAcc = 256;

struct TcParams {
  const int32_t* tile_nodes;  // GATHER: [tiles_per_frame][128]
  int tiles_per_frame;
  int nodes_per_frame;
  long long num_tiles;
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
};

template <bool GATHER>
__device__ __forceinline__ int tile_row(const TcParams& p, long long tile, int r) {
  // global row index of tile row r, or -1
  if (GATHER) {
    const long long b = tile / p.tiles_per_frame;
    const int t = (int)(tile - b * p.tiles_per_frame);
    const int node = __ldg(p.tile_nodes + t * 128 + r);
    return node < 0 ? -1 : (int)(b * p.nodes_per_frame + node);
  }
  const long long row = tile * 128 + r;
  return row < p.rows ? (int)row : -1;
}

#ifdef EG_TC_TIMING
__device__ long long g_tc_dbg[kNumSMs][8];  // per CTA: cycles spent waiting, by role (see eg_tc_debug_read)
#define TC_TIMED_WAIT(slot, bar, par)            \
  do {                                           \
    const long long _t = clock64();              \
    mbar_wait(bar, par);                         \
    dbg_acc[slot] += clock64() - _t;             \
  } while (0)
#else
#define TC_TIMED_WAIT(slot, bar, par) mbar_wait(bar, par)
#endif

template <bool GATHER>
__global__ void __launch_bounds__(kThreads, 1) gcn_tc_kernel(const TcParams p) {
#ifdef EG_TC_TIMING
  long long dbg_acc[2] = {0, 0};
  const long long dbg_t0 = clock64();
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                     // [stage][hi|lo][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + kABytes);
  uint64_t* full = bars;                  // [kStages]  producers -> MMA
  uint64_t* empty = bars + kStages;       // [kStages]  MMA -> producers
  uint64_t* acc_full = bars + 2 * kStages;       // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

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
      tmem_st32(ta + 128, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp > kMmaWarp) {
    // ===== producers: gather -> split -> swizzled operand tile =============================================
    const int pw = warp - (kMmaWarp + 1);
    const int g = lane >> 3, j = lane & 7;
    const uint32_t gmask = 0xFFu << (lane & 24);
    uint32_t chunk = 0;
    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      constexpr int kIters = kRowsPerProd / 4;
      // Per tile: this lane's rows and (GATHER) their first 16 CSR entries, one entry per lane of the
      // 8-lane group in two register sets -- the K chunks below then need a single round of feature loads.
      int grow[kIters], beg[kIters], deg[kIters], ec[kIters][2];
      float ewt[kIters][2];
      const float* fbase = p.X;
#pragma unroll
      for (int i = 0; i < kIters; ++i) {
        const int r = pw * kRowsPerProd + i * 4 + g;
        grow[i] = tile_row<GATHER>(p, tile, r);
        beg[i] = deg[i] = 0;
        ec[i][0] = ec[i][1] = 0;
        ewt[i][0] = ewt[i][1] = 0.f;
        if (GATHER && grow[i] >= 0) {
          const int node = __ldg(p.tile_nodes + (int)(tile % p.tiles_per_frame) * 128 + r);
          beg[i] = __ldg(p.rowptr + node);
          deg[i] = __ldg(p.rowptr + node + 1) - beg[i];
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if (8 * h + j < deg[i]) {
              ec[i][h] = __ldg(p.col + beg[i] + 8 * h + j);
              ewt[i][h] = __ldg(p.w + beg[i] + 8 * h + j);
            }
        }
      }
      if (GATHER) fbase = p.X + (tile / p.tiles_per_frame) * (long long)p.nodes_per_frame * 128;
      {  // pull the rows of this CTA's NEXT tile into L2 (4 x 128 B lines per row, lanes j < 4 of each group)
        const long long nt = tile + gridDim.x;
        if (nt < p.num_tiles && j < 4) {
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const int nr = tile_row<GATHER>(p, nt, pw * kRowsPerProd + i * 4 + g);
            if (nr >= 0) prefetch_l2(p.X + (long long)nr * 128 + j * 32);
          }
        }
      }
      for (int kc = 0; kc < 4; ++kc, ++chunk) {
        const uint32_t stage = chunk % kStages, phase = (chunk / kStages) & 1u;
        TC_TIMED_WAIT(0, &empty[stage], phase ^ 1u);
        uint8_t* a_hi = sA + stage * 2 * kTileBytes;
        uint8_t* a_lo = a_hi + kTileBytes;
        const int coff = kc * 32 + j * 4;
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int r = pw * kRowsPerProd + i * 4 + g;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          if (GATHER) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (8 * h < deg[i]) {  // uniform inside the 8-lane group
                const int n = deg[i] - 8 * h;
                float4 x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const int c = __shfl_sync(gmask, ec[i][h], k, 8);
                  x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (k < n) {
                    const float* src = fbase + (long long)c * 128 + coff;
                    x[k] = ldg4(src);
                    if (j == 0 && kc < 3) prefetch_l1(src + 32);  // next K chunk of this neighbour row -> L1
                  }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float wk = __shfl_sync(gmask, ewt[i][h], k, 8);  // 0 for k >= n
                  acc.x = fmaf(wk, x[k].x, acc.x);
                  acc.y = fmaf(wk, x[k].y, acc.y);
                  acc.z = fmaf(wk, x[k].z, acc.z);
                  acc.w = fmaf(wk, x[k].w, acc.w);
                }
              }
            }
            for (int e = beg[i] + 16; e < beg[i] + deg[i]; ++e) {  // hub rows (connection nodes) only
              const float wk = __ldg(p.w + e);
              const float4 x = ldg4(fbase + (long long)__ldg(p.col + e) * 128 + coff);
              acc.x = fmaf(wk, x.x, acc.x);
              acc.y = fmaf(wk, x.y, acc.y);
              acc.z = fmaf(wk, x.z, acc.z);
              acc.w = fmaf(wk, x.w, acc.w);
            }
            if (p.AggOut && grow[i] >= 0) st4(p.AggOut + (long long)grow[i] * 128 + coff, acc);
          } else if (grow[i] >= 0) {
            acc = ldg4(p.X + (long long)grow[i] * 128 + coff);
          }
          uint4 hi, lo;
          split_tf32(acc.x, hi.x, lo.x);
          split_tf32(acc.y, hi.y, lo.y);
          split_tf32(acc.z, hi.z, lo.z);
          split_tf32(acc.w, hi.w, lo.w);
          const uint32_t off = sw128_off(r, j);
          *reinterpret_cast<uint4*>(a_hi + off) = hi;
          *reinterpret_cast<uint4*>(a_lo + off) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer ======================================================================================
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 0, 0);
    uint32_t chunk = 0, it = 0;
    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
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
            umma_tf32_ts(d, w_hi, umma_desc_k128(b_lo + o), idesc, (kc | ks) != 0);
            umma_tf32_ts(d, w_lo, umma_desc_k128(b_hi + o), idesc, 1u);
            umma_tf32_ts(d, w_hi, umma_desc_k128(b_hi + o), idesc, 1u);
#endif
          }
          umma_commit(&empty[stage]);
          if (kc == 3) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: thread <-> output feature; TMEM -> registers -> 128-byte row segments ===================
    const int ew = warp;             // TMEM lanes [32 ew, 32 ew + 32)
    const int f = ew * 32 + lane;    // output feature owned by this thread
    const float bias = p.bias ? __ldg(p.bias + f) : 0.f;
    double s_sum = 0.0, s_sq = 0.0;
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, acc_phase = (it >> 1) & 1u;
      int myrow[4];  // lane i holds the global row of tile rows i, 32 + i, 64 + i, 96 + i
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) myrow[sl] = tile_row<GATHER>(p, tile, sl * 32 + lane);
      TC_TIMED_WAIT(0, &acc_full[buf], acc_phase);
      tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {  // 16 tile rows (accumulator columns) at a time
        uint32_t v[16];
#ifdef EG_DBG_NOLDTM
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = sl + i;
#else
        tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + kTmemAcc + buf * 128 + sl * 16, v);
#endif
        int row[16];
        float ad[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {  // all residual loads in flight before the accumulator is consumed
          row[i] = __shfl_sync(0xffffffffu, myrow[sl >> 1], (sl & 1) * 16 + i);
          ad[i] = 0.f;
          if (p.addend) ad[i] = __ldg(p.addend + (long long)max(row[i], 0) * 128 + f);
        }
        tmem_ld_wait();
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float o = __uint_as_float(v[i]) + bias + ad[i];
          if (row[i] >= 0) {  // warp-uniform
#ifndef EG_DBG_NOSTORE
            p.Out[(long long)row[i] * 128 + f] = o;
#endif
            s += o;
            q = fmaf(o, o, q);
          }
        }
        s_sum += (double)s;
        s_sq += (double)q;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (p.stat_parts) {
      p.stat_parts[(size_t)blockIdx.x * 256 + f] = s_sum;
      p.stat_parts[(size_t)blockIdx.x * 256 + 128 + f] = s_sq;
    }
  }

#ifdef EG_TC_TIMING
  if (lane == 0) {
    long long* d = g_tc_dbg[blockIdx.x];
    if (warp == kMmaWarp + 1) d[0] = dbg_acc[0];                    // producer warp 0: wait for a free stage
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

template <bool GATHER>
int launch(const TcParams& p, float* mean, float* var, void* ws, size_t ws_bytes, const char* name, cudaStream_t s) {
  const bool stats = mean && var;
  if (stats && (!ws || ws_bytes < kWorkspaceBytes)) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  static bool attr_done = false;
  if (!attr_done) {
    EG_CUDA(cudaFuncSetAttribute(gcn_tc_kernel<GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_done = true;
  }
  const int grid = (int)(p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs);
  TcParams q = p;
  q.stat_parts = stats ? reinterpret_cast<double*>(ws) : nullptr;
  {
    ProfileScope prof(name, s);
    gcn_tc_kernel<GATHER><<<grid, kThreads, kSmemBytes, s>>>(q);
    EG_LAUNCH_CHECK();
  }
  if (stats) return launch_stats_finalize(grid, 128, 128, p.rows, q.stat_parts, mean, var, s);
  return EG_OK;
}

}  // namespace

#ifdef EG_TC_TIMING
extern "C" int eg_tc_debug_read(long long* out) {  // HOST buffer of kNumSMs * 8 counters
  return cudaMemcpyFromSymbol(out, g_tc_dbg, sizeof(long long) * kNumSMs * 8) == cudaSuccess ? 0 : -2;
}
#endif

namespace eg {

// Out = (A_hat X) op(W) + bias + addend over the batched graph; AggOut (optional) receives A_hat X.
int launch_gcn_tc(const eg_graph* g, int batch, const float* X, const float* W, int trans_w, const float* bias,
                  const float* addend, float* Out, float* AggOut, float* mean, float* var, void* ws, size_t ws_bytes,
                  cudaStream_t s) {
  const eg_graph_info& info = graph_info(g);
  TcParams p{};
  p.tile_nodes = graph_tile_nodes(g);
  p.tiles_per_frame = graph_tiles_per_frame(g);
  p.nodes_per_frame = info.num_nodes;
  p.num_tiles = (long long)batch * p.tiles_per_frame;
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
  return launch<true>(p, mean, var, ws, ws_bytes, "gcn_tc", s);
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
  p.num_tiles = (rows + 127) / 128;
  p.rows = rows;
  p.X = A;
  p.W = W;
  p.trans_w = trans_w;
  p.bias = bias;
  p.addend = addend;
  p.Out = C;
  return launch<false>(p, mean, var, ws, ws_bytes, "linear_tc", s);
}

}  // namespace eg
