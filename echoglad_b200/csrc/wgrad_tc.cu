// Weight gradient of the per-node transforms on the 5th-generation tensor cores:
//
//     dW[o][i] = sum_n G[n][o] * X[n][i]      (G = A_hat dH of a GCNConv, or dZ of classifier layer 0)
//     dbias[o] = sum_n G[n][o]                 (optional)
//
// Replaces the autograd of nn.Linear inside PyG GCNConv.lin and node_classifiers[k][0]
// (src/core/models.py:330,364; loss.backward() at src/engine.py:272).  HBM-bound: both [rows,128] operands
// are read exactly once (2 U), the result is 64 KB.
//
// The reduction runs over the ROWS, so both operands are "MN-major" for the MMA (the 128 features of a row
// are contiguous, the K index is the row): a block of 32 rows of G and of X travels global -> shared with
// cp.async straight into the canonical MN-major layout of 32-bit operands (SWIZZLE_128B_BASE32B: 4 atoms of
// 32 features x 4 row groups of 4 rows), no transpose anywhere.  3xTF32: the raw block is the hi operand as it is (the tensor core
// ignores the low 13 mantissa bits), split warps write the lo part to a second ring; one elected thread issues, per 8-row group,
//     D += G_lo^T X_hi,  D += G_hi^T X_lo,  D += G_hi^T X_hi        (tcgen05.mma kind::tf32, M = N = 128, K = 8)
// into a 128-column TMEM accumulator.  The tensor core does not round its fp32 accumulation to nearest, so a
// long accumulation chain drifts (measured: 2 x 288,084 rows in ONE chain per CTA missed the 1e-5 bound against
// fp64 by 1.5x); the accumulation therefore runs in SEGMENTS of kSegBlocks blocks that alternate between two
// TMEM accumulators, and four epilogue warps drain each finished segment into the CTA's fp32 partial in global
// memory (L2-resident, round-to-nearest adds, fixed order) while the next segment accumulates.  A fixed-order
// second stage sums the [148][128][128] partials in double: no atomics, bit-reproducible.
//
// Rings: RAW/hi ring of kRaw = 5 blocks (32 KB: 32 rows of G and X), LO ring of kLo = 2 blocks (224 KB in all).
// Warps: 0-3 loaders, 4 MMA issuer, 5-12 split, 13-16 epilogue (TMEM lane quadrant = warp id % 4).
#include "common.cuh"
#include "tc05.cuh"

using namespace eg;
using namespace eg::tc;

namespace {

#ifndef EG_WG_ROWS
#define EG_WG_ROWS 32  // measured at batch 64: 8 rows 1.65 ms, 16 rows 1.18 ms, 32 rows 1.07 ms (per-block handshakes)
#define EG_WG_RAW 5
#define EG_WG_LO 2
#endif
#ifndef EG_WG_SPLIT
#define EG_WG_SPLIT 8
#endif
constexpr int kRows = EG_WG_ROWS;              // rows per block (kRows / 8 MMA k-steps)
constexpr int kRaw = EG_WG_RAW;                // raw / hi ring depth
constexpr int kLo = EG_WG_LO;                  // lo ring depth
constexpr int kSegBlocks = 2048 / kRows;       // blocks (2048 rows, 768 MMAs) per accumulation segment
constexpr int kLoadWarps = 4, kSplitWarps = EG_WG_SPLIT, kEpiWarps = 4;
static_assert(kRows % kSplitWarps == 0 && kRows % kLoadWarps == 0 && kRows % 8 == 0, "block shape");
constexpr int kMmaWarp = kLoadWarps;
constexpr int kSplitWarp0 = kMmaWarp + 1;
constexpr int kEpiWarp0 = kSplitWarp0 + kSplitWarps;
constexpr int kThreads = (kLoadWarps + 1 + kSplitWarps + kEpiWarps) * 32;
constexpr uint32_t kOpBytes = kRows * 512;     // one operand (G or X) of a block: 8 KB
constexpr uint32_t kBlockBytes = 2 * kOpBytes; // G + X
constexpr uint32_t kOffLo = kRaw * kBlockBytes;
constexpr uint32_t kOffBars = kOffLo + kLo * kBlockBytes;
constexpr uint32_t kSmemBytes = kOffBars + 8 * (2 * kRaw + 2 * kLo + 4) + 16 + 1024;
constexpr uint32_t kLbo = (kRows / 4) * 512;   // byte stride between the 32-feature atoms of an operand
constexpr uint32_t kSbo = 512;                 // byte stride between 4-row groups
static_assert(kSmemBytes <= 232448, "shared memory budget");

// byte offset of 16-byte chunk c (0..31: features 4c..4c+3) of row r (0..kRows-1) inside an operand block:
// [atom a = c / 8 (32 features)][row group r / 4][4 rows x 128 B]; inside the 512-byte swizzle atom the 32-byte
// unit (c & 7) >> 1 of row r sits at unit position ((c & 7) >> 1) ^ (r & 3)  (SWIZZLE_128B_BASE32B).
__device__ __forceinline__ uint32_t op_off(int r, int c) {
  return (uint32_t)(((c >> 3) * (kRows / 4) + (r >> 2)) * 512 + (r & 3) * 128 + (((((c & 7) >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 3xTF32 with a TRUNCATING split: the tensor core reads the upper 19 bits of a 32-bit tf32 operand, so the raw
// fp32 block in shared memory already IS the hi operand (hi = x with the low 13 mantissa bits dropped) and only
// lo = x - hi (exact in fp32, then rounded to tf32) has to be written: one shared-memory store pass instead of
// two (the kernel runs at 75 % LSU utilisation, ncu r01g).
__device__ __forceinline__ uint32_t lo1(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ uint4 lo4(const float4& x) { return make_uint4(lo1(x.x), lo1(x.y), lo1(x.z), lo1(x.w)); }

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(long long rows, const float* __restrict__ G, const float* __restrict__ X,
                float* __restrict__ parts, double* __restrict__ colsum_parts) {
  extern __shared__ uint8_t smem_raw[];
  // dynamic shared memory of a CTA without static shared memory / cluster starts at shared-window address 0x400
  // (1 KB system reserve), already 1024-byte aligned: a literal base makes every ring / barrier address an immediate
  // (see gcn_tc.cu); any other layout traps.
  constexpr uint32_t sm = 0x400;
  if (((smem_u32(smem_raw) + 1023u) & ~1023u) != sm) __trap();
  uint8_t* smem = smem_raw + (sm - smem_u32(smem_raw));
  const uint32_t bar_raw_full = sm + kOffBars, bar_raw_empty = bar_raw_full + 8 * kRaw,
                 bar_lo_full = bar_raw_empty + 8 * kRaw, bar_lo_empty = bar_lo_full + 8 * kLo,
                 bar_acc_full = bar_lo_empty + 8 * kLo, bar_acc_empty = bar_acc_full + 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* acc_full = bars + 2 * kRaw + 2 * kLo;  // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;              // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nblocks = (rows + kRows - 1) / kRows;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kRaw; ++s) {
      mbar_init(bars + s, kLoadWarps * 32);  // raw_full: one cp.async arrival per loader thread
      mbar_init(bars + kRaw + s, 1);         // raw_empty: tcgen05.commit
    }
    for (int s = 0; s < kLo; ++s) {
      mbar_init(bars + 2 * kRaw + s, kSplitWarps);  // lo_full
      mbar_init(bars + 2 * kRaw + kLo + s, 1);      // lo_empty: tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, kEpiWarps);
    }
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kLoadWarps) {
    // ===== loaders: 16 rows x 512 B of G and of X per block, one global row per warp instruction ================
    const int c = lane;
    uint32_t it = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
      const uint32_t rs = it % kRaw, rphase = (it / kRaw) & 1u;
      mbar_wait_a(bar_raw_empty + rs * 8, rphase ^ 1u);
      const uint32_t dst = sm + rs * kBlockBytes;
      const long long row0 = blk * kRows;
#pragma unroll
      for (int m = 0; m < kRows / kLoadWarps; ++m) {
        const int r = warp + kLoadWarps * m;
        const bool ok = row0 + r < rows;
        const long long off = ok ? (row0 + r) * 128 + c * 4 : 0;
        cp_async16_zfill(dst + op_off(r, c), G + off, ok);
        cp_async16_zfill(dst + kOpBytes + op_off(r, c), X + off, ok);
      }
      cp_async_mbar_arrive_a(bar_raw_full + rs * 8);
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer ==========================================================================================
    constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 1, 1);  // both operands MN-major
    uint32_t it = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
      const uint32_t rs = it % kRaw;
      const uint32_t ls = it % kLo, lphase = (it / kLo) & 1u;
      const uint32_t seg = it / kSegBlocks, sb = it % kSegBlocks, buf = seg & 1u;
      if (sb == 0) {  // new segment: its accumulator must have been drained (segment seg - 2)
        mbar_wait_a(bar_acc_empty + buf * 8, ((seg >> 1) & 1u) ^ 1u);
        tc_fence_after();
      }
      mbar_wait_a(bar_lo_full + ls * 8, lphase);  // the split warps saw the raw block land and wrote its lo part
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d = tmem_base + buf * 128;
        const uint32_t g_hi = sm + rs * kBlockBytes, x_hi = g_hi + kOpBytes;
        const uint32_t g_lo = sm + kOffLo + ls * kBlockBytes, x_lo = g_lo + kOpBytes;
#pragma unroll
        for (int kg = 0; kg < kRows / 8; ++kg) {
          const uint32_t o = kg * 1024;
          umma_tf32(d, umma_desc_mn128_b32(g_lo + o, kLbo, kSbo), umma_desc_mn128_b32(x_hi + o, kLbo, kSbo), idesc,
                    (sb | (uint32_t)kg) != 0);
          umma_tf32(d, umma_desc_mn128_b32(g_hi + o, kLbo, kSbo), umma_desc_mn128_b32(x_lo + o, kLbo, kSbo), idesc, 1u);
          umma_tf32(d, umma_desc_mn128_b32(g_hi + o, kLbo, kSbo), umma_desc_mn128_b32(x_hi + o, kLbo, kSbo), idesc, 1u);
        }
        umma_commit(bars + kRaw + rs);            // raw_empty
        umma_commit(bars + 2 * kRaw + kLo + ls);  // lo_empty
        if (sb == kSegBlocks - 1 || blk + gridDim.x >= nblocks) umma_commit(acc_full + buf);  // segment complete
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ===== epilogue: finished segments -> the CTA's partial dW (fp32, round-to-nearest adds, fixed order) =========
    // thread <-> output feature o = TMEM lane; P[o][0..127] is this thread's 512 bytes of the partial
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int o = q * 32 + lane;
    float* P = parts + (size_t)blockIdx.x * 128 * 128 + (size_t)o * 128;
    const long long my_blocks = (nblocks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t nseg = (uint32_t)((my_blocks + kSegBlocks - 1) / kSegBlocks);
    for (uint32_t seg = 0; seg < nseg; ++seg) {
      const uint32_t buf = seg & 1u;
      mbar_wait_a(bar_acc_full + buf * 8, (seg >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int sl = 0; sl < 4; ++sl) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + sl * 32, v);
        float4 old[8];
        if (seg) {
#pragma unroll
          for (int c = 0; c < 8; ++c) old[c] = *reinterpret_cast<const float4*>(P + sl * 32 + c * 4);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 r = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                 __uint_as_float(v[4 * c + 3]));
          if (seg) r.x += old[c].x, r.y += old[c].y, r.z += old[c].z, r.w += old[c].w;
          *reinterpret_cast<float4*>(P + sl * 32 + c * 4) = r;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar_acc_empty + buf * 8);
    }
  } else {
    // ===== split warps: lo = x - trunc_tf32(x) of the raw block (which is the hi operand as it is); column sums of G ========================================
    const int t = tid - kSplitWarp0 * 32;
    const int c = t & 31, r0 = (t >> 5) * (kRows / kSplitWarps);
    uint32_t off[kRows / kSplitWarps];
#pragma unroll
    for (int m = 0; m < kRows / kSplitWarps; ++m) off[m] = op_off(r0 + m, c);
    double cs[4] = {0.0, 0.0, 0.0, 0.0};
    uint32_t it = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
      const uint32_t rs = it % kRaw, rphase = (it / kRaw) & 1u;
      const uint32_t ls = it % kLo, lphase = (it / kLo) & 1u;
      mbar_wait_a(bar_raw_full + rs * 8, rphase);
      mbar_wait_a(bar_lo_empty + ls * 8, lphase ^ 1u);
      const uint32_t raw = sm + rs * kBlockBytes, lo_t = sm + kOffLo + ls * kBlockBytes;
      float4 sg = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int m = 0; m < kRows / kSplitWarps; ++m) {
        const float4 g = lds4(raw + off[m]);
        const float4 x = lds4(raw + kOpBytes + off[m]);
        sts4(lo_t + off[m], lo4(g));
        sts4(lo_t + kOpBytes + off[m], lo4(x));
        sg.x += g.x, sg.y += g.y, sg.z += g.z, sg.w += g.w;
      }
      cs[0] += (double)sg.x, cs[1] += (double)sg.y, cs[2] += (double)sg.z, cs[3] += (double)sg.w;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar_lo_full + ls * 8);
    }
    if (colsum_parts) {  // [cta][split warp][128]; thread owns features 4c..4c+3 of its warp's rows
      double* dst = colsum_parts + ((size_t)blockIdx.x * kSplitWarps + (t >> 5)) * 128 + c * 4;
      dst[0] = cs[0], dst[1] = cs[1], dst[2] = cs[2], dst[3] = cs[3];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// fixed-order second stage: dW = sum of the per-CTA partials, dbias = sum of the per-warp column sums (double)
__global__ void wgrad_tc_reduce_kernel(int nparts, const float* __restrict__ parts,
                                       const double* __restrict__ colsum_parts, float* __restrict__ dW,
                                       float* __restrict__ dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * 128) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)parts[(size_t)p * 128 * 128 + i];
    dW[i] = (float)s;
  }
  if (dbias && i < 128) {
    double s = 0.0;
    for (int p = 0; p < nparts * kSplitWarps; ++p) s += colsum_parts[(size_t)p * 128 + i];
    dbias[i] = (float)s;
  }
}

}  // namespace

namespace eg {

int launch_wgrad_tc(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s) {
  static_assert((size_t)kNumSMs * kSplitWarps * 128 * sizeof(double) <= kStatsBytes, "column-sum partials fit the stats area");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  static std::atomic<unsigned long long> attr_mask{0};
  if (first_use_on_current_device(attr_mask))
    EG_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  const long long nblocks = (rows + kRows - 1) / kRows;
  const int sms = num_sms();
  const int grid = (int)(nblocks < sms ? nblocks : sms);
  double* colsum = reinterpret_cast<double*>(ws);
  float* parts = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  ProfileScope prof("wgrad_tc", s);
  wgrad_tc_kernel<<<grid, kThreads, kSmemBytes, s>>>(rows, G, X, parts, dbias ? colsum : nullptr);
  EG_LAUNCH_CHECK();
  wgrad_tc_reduce_kernel<<<(128 * 128 + 255) / 256, 256, 0, s>>>(grid, parts, dbias ? colsum : nullptr, dW, dbias);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // namespace eg
