// Classifier tail: the four independent heads of node_classifiers (src/core/models.py:363-377) evaluated
// together as block-diagonal transforms.  Layer 0 of the four heads is one stacked [128 -> 4x32]
// tensor-core GEMM (eg_linear128); this file holds layer 4 (4 x [32 -> 16]) and layer 8 (4 x [16 -> 1]).
// The BN/ReLU/Dropout between them is eg_bn_act_* with cols = 128 / 64.
#include "common.cuh"

using namespace eg;

namespace {

// ---- layer 4: four block-diagonal [32 -> 16] transforms ---------------------------------------------------------
// Lane mapping of both kernels: a warp owns whole rows; lane = 8 k + c  <->  head k, inputs 4c..4c+3 of that head,
// i.e. the lane's float4 of the row's 512 contiguous bytes.  Every global access is a fully coalesced row (the
// first version gave a thread its own rows: 32 different 128-byte lines per load instruction, LSU-bound at 89 %).
// The lane's 64 weights W2[k][0..15][4c..4c+3] live in registers; there is no shared-memory traffic in the loops.
constexpr int kMidFwdThreads = 256, kMidFwdBlocks = 2;   // per SM
constexpr int kMidBwdThreads = 128, kMidBwdBlocks = 4;
constexpr int kMidGrid = kNumSMs * 4;
static_assert((size_t)kMidGrid * 128 * sizeof(double) <= kStatsBytes, "clf_mid_fwd partials fit the stats area");

struct P2 {  // packed fp32 pair: one FFMA2 / FADD2 issue slot for two values
  unsigned long long u;
};
__device__ __forceinline__ P2 p2(float lo, float hi) {
  return P2{((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo)};
}
__device__ __forceinline__ float p2_lo(P2 a) { return __uint_as_float((uint32_t)a.u); }
__device__ __forceinline__ float p2_hi(P2 a) { return __uint_as_float((uint32_t)(a.u >> 32)); }
__device__ __forceinline__ P2 p2_fma(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.u) : "l"(a.u), "l"(b.u), "l"(c.u));
  return r;
}
__device__ __forceinline__ P2 p2_mul(P2 a, P2 b) {
  P2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
  return r;
}
__device__ __forceinline__ P2 p2_add(P2 a, P2 b) {
  P2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.u) : "l"(a.u), "l"(b.u));
  return r;
}
__device__ __forceinline__ P2 p2_shfl_xor(P2 a, int m) {
  return P2{(unsigned long long)__shfl_xor_sync(0xffffffffu, a.u, m)};
}

// Z2[r][16k+j] = b2[k][j] + sum_i A1[r][32k+i] * W2[k][j][i]
// Each lane forms the partial sums of all 16 outputs of its head over its 4 inputs (32 FFMA2), then the 8 lanes of
// a head transpose-reduce them (8 + 4 + 2 shuffles) so that lane c ends with outputs 2c, 2c+1 = its float2 of the
// row's 256 contiguous output bytes.  The exchange needs no selects: lane c keeps its partials in the order
// m -> output m ^ 2c (a permutation of its register-resident weights), so in every stage every lane keeps the low
// half of its accumulators and sends the high half.  Column statistics: a lane owns two columns for its whole life.
__global__ void __launch_bounds__(kMidFwdThreads, kMidFwdBlocks)
clf_mid_fwd_kernel(long long rows, const float* __restrict__ A1, const float* __restrict__ W2,
                   const float* __restrict__ b2, float* __restrict__ Z2, double* __restrict__ parts) {
  constexpr int kWarps = kMidFwdThreads / 32;
  constexpr int kFwdDepth = 8;
  __shared__ double red[kWarps][128];
  __shared__ __align__(16) float ring[kWarps * kFwdDepth * 128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, k = lane >> 3, c = lane & 7;
  // wp[jp][e] = (W2[k][2(jp^c)][4c+e], W2[k][2(jp^c)+1][4c+e])
  P2 wp[8][4];
#pragma unroll
  for (int jp = 0; jp < 8; ++jp) {
    const int j = 2 * (jp ^ c);
    const float4 w0 = ldg4(W2 + (k * 16 + j) * 32 + c * 4), w1 = ldg4(W2 + (k * 16 + j + 1) * 32 + c * 4);
    wp[jp][0] = p2(w0.x, w1.x);
    wp[jp][1] = p2(w0.y, w1.y);
    wp[jp][2] = p2(w0.z, w1.z);
    wp[jp][3] = p2(w0.w, w1.w);
  }
  const P2 bias = p2(__ldg(b2 + lane * 2), __ldg(b2 + lane * 2 + 1));
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  double ds0 = 0.0, ds1 = 0.0, dq0 = 0.0, dq1 = 0.0;
  int since_flush = 0;
  // rows stream through a private per-warp shared-memory ring (cp.async, kFwdDepth rows = 4 KB per warp in flight)
  const long long wid = (long long)blockIdx.x * kWarps + warp, nw = (long long)gridDim.x * kWarps;
  const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring) + warp * kFwdDepth * 512 + lane * 16;
  auto issue = [&](long long r, int slot) {
    if (r < rows)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_u + slot * 512), "l"(A1 + r * 128 + lane * 4) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int d = 0; d < kFwdDepth - 1; ++d) issue(wid + d * nw, d);
  int slot = 0;
  for (long long r = wid; r < rows; r += nw) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kFwdDepth - 2) : "memory");
    // each lane reads back exactly the 16 bytes it copied itself: no warp barrier needed
    const float4 a = *reinterpret_cast<const float4*>(ring + (warp * kFwdDepth + slot) * 128 + lane * 4);
    issue(r + (long long)(kFwdDepth - 1) * nw, (slot + kFwdDepth - 1) % kFwdDepth);
    slot = (slot + 1) % kFwdDepth;
    const P2 ax = p2(a.x, a.x), ay = p2(a.y, a.y), az = p2(a.z, a.z), aw = p2(a.w, a.w);
    P2 acc[8];
#pragma unroll
    for (int jp = 0; jp < 8; ++jp) {
      acc[jp] = p2_mul(ax, wp[jp][0]);
      acc[jp] = p2_fma(ay, wp[jp][1], acc[jp]);
      acc[jp] = p2_fma(az, wp[jp][2], acc[jp]);
      acc[jp] = p2_fma(aw, wp[jp][3], acc[jp]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = p2_add(acc[i], p2_shfl_xor(acc[i + 4], 4));
#pragma unroll
    for (int i = 0; i < 2; ++i) acc[i] = p2_add(acc[i], p2_shfl_xor(acc[i + 2], 2));
    const P2 z = p2_add(p2_add(acc[0], p2_shfl_xor(acc[1], 1)), bias);
    const float z0 = p2_lo(z), z1 = p2_hi(z);
    *reinterpret_cast<float2*>(Z2 + r * 64 + lane * 2) = make_float2(z0, z1);
    s0 += z0;
    s1 += z1;
    q0 = fmaf(z0, z0, q0);
    q1 = fmaf(z1, z1, q1);
    if (++since_flush == 64) {  // 64 rows per fp32 run, then double
      ds0 += s0; ds1 += s1; dq0 += q0; dq1 += q1;
      s0 = s1 = q0 = q1 = 0.f;
      since_flush = 0;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (parts) {  // parts[block][0..63] sums, [64..127] sums of squares; warps combined in fixed order
    red[warp][lane * 2] = ds0 + (double)s0;
    red[warp][lane * 2 + 1] = ds1 + (double)s1;
    red[warp][64 + lane * 2] = dq0 + (double)q0;
    red[warp][64 + lane * 2 + 1] = dq1 + (double)q1;
    __syncthreads();
    if (tid < 128) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += red[w][tid];
      parts[(size_t)blockIdx.x * 128 + tid] = t;
    }
  }
}

// dA1[r][32k+i] = sum_j dZ2[r][16k+j] W2[k][j][i];  per-block partials of dW2[k][j][i], db2[k][j].
// A warp owns HALF rows (heads 2h, 2h+1; h = warp & 1): lane = 16 kk + c  <->  head 2h + kk, inputs 2c, 2c+1 of that
// head = the lane's float2 of the half row's 256 contiguous bytes.  Per half row a lane needs its float2 of A1 and
// the 16 dZ2 values of its head; the packed FFMA2 run over PAIRS OF j, so the dZ2 pairs are used as they sit in the
// registers of a 128-bit load (no broadcast packing): dA1 = even-j + odd-j partial sums (16 FFMA2 + 2 adds), the
// lane's 16 x 2 block of dW2 (16 FFMA2).  Weights + accumulators take 64 registers, which leaves room for 16 warps
// per SM; every warp streams its half rows through a PRIVATE shared-memory ring with cp.async (kMidDepth slots
// in flight, no block-level barrier in the loop).
constexpr int kMidPart = 2048 + 64;
constexpr int kMidDepth = 8;                 // ring slots per warp
constexpr int kMidSlotBytes = 256 + 128;     // half a row of A1 + half a row of dZ2
constexpr int kMidHalfPart = 1024 + 32;      // one half's share of a partial
__global__ void __launch_bounds__(kMidBwdThreads, kMidBwdBlocks)
clf_mid_bwd_kernel(long long rows, const float* __restrict__ A1, const float* __restrict__ W2,
                   const float* __restrict__ dZ2, float* __restrict__ dA1, float* __restrict__ parts) {
  constexpr int kWarps = kMidBwdThreads / 32;
  constexpr int kRingFloats = kWarps * kMidDepth * kMidSlotBytes / 4, kRedFloats = kWarps * kMidHalfPart;
  __shared__ __align__(16) float smem[kRedFloats > kRingFloats ? kRedFloats : kRingFloats];  // rings, then the block reduction
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, h = warp & 1, kk = lane >> 4, c = lane & 15;
  const int head = 2 * h + kk;
  // wq[m][e] = (W2[head][2m][2c+e], W2[head][2m+1][2c+e])
  P2 wq[8][2];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float2 w0 = __ldg(reinterpret_cast<const float2*>(W2 + (head * 16 + 2 * m) * 32 + 2 * c));
    const float2 w1 = __ldg(reinterpret_cast<const float2*>(W2 + (head * 16 + 2 * m + 1) * 32 + 2 * c));
    wq[m][0] = p2(w0.x, w1.x);
    wq[m][1] = p2(w0.y, w1.y);
  }
  P2 wacc[8][2];  // (dW2[head][2m][2c+e], dW2[head][2m+1][2c+e])
#pragma unroll
  for (int m = 0; m < 8; ++m) wacc[m][0] = wacc[m][1] = p2(0.f, 0.f);
  float bacc = 0.f;  // db2 of column 32h + lane
  const long long wid = (long long)blockIdx.x * (kWarps / 2) + (warp >> 1), nw = (long long)gridDim.x * (kWarps / 2);
  float* ringf = smem + warp * (kMidDepth * kMidSlotBytes / 4);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ringf);
  // slot layout: [256 B of A1][128 B of dZ2]; lanes 0-15 copy A1, lanes 16-23 dZ2, 16 bytes each
  const float* src0 = lane < 16 ? A1 + 64 * h + lane * 4 : dZ2 + 32 * h + (lane - 16) * 4;
  const long long src_stride = lane < 16 ? 128 : 64;
  const uint32_t dst0 = ring + (lane < 16 ? lane * 16 : 256 + (lane - 16) * 16);
  auto issue = [&](long long r, int slot) {
    if (r < rows && lane < 24)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + slot * kMidSlotBytes), "l"(src0 + r * src_stride) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int d = 0; d < kMidDepth - 1; ++d) issue(wid + d * nw, d);
  int slot = 0;
  for (long long r = wid; r < rows; r += nw) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kMidDepth - 2) : "memory");
    __syncwarp();
    const float* row = ringf + slot * (kMidSlotBytes / 4);
    const float2 a = *reinterpret_cast<const float2*>(row + lane * 2);
    const float4 g0 = *reinterpret_cast<const float4*>(row + 64 + kk * 16),
                 g1 = *reinterpret_cast<const float4*>(row + 64 + kk * 16 + 4),
                 g2 = *reinterpret_cast<const float4*>(row + 64 + kk * 16 + 8),
                 g3 = *reinterpret_cast<const float4*>(row + 64 + kk * 16 + 12);
    const float gb = row[64 + lane];
    __syncwarp();  // every lane has read the slot: refill it with the row kMidDepth - 1 ahead
    issue(r + (long long)(kMidDepth - 1) * nw, (slot + kMidDepth - 1) % kMidDepth);
    slot = (slot + 1) % kMidDepth;
    const P2 dz[8] = {p2(g0.x, g0.y), p2(g0.z, g0.w), p2(g1.x, g1.y), p2(g1.z, g1.w),
                      p2(g2.x, g2.y), p2(g2.z, g2.w), p2(g3.x, g3.y), p2(g3.z, g3.w)};
    const P2 a0 = p2(a.x, a.x), a1 = p2(a.y, a.y);
    P2 o0 = p2(0.f, 0.f), o1 = o0, o2 = o0, o3 = o0;  // (even-j, odd-j) partial sums of inputs 2c, 2c+1; two chains each
#pragma unroll
    for (int m = 0; m < 8; m += 2) {
      o0 = p2_fma(dz[m], wq[m][0], o0);
      o1 = p2_fma(dz[m], wq[m][1], o1);
      o2 = p2_fma(dz[m + 1], wq[m + 1][0], o2);
      o3 = p2_fma(dz[m + 1], wq[m + 1][1], o3);
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      wacc[m][0] = p2_fma(dz[m], a0, wacc[m][0]);
      wacc[m][1] = p2_fma(dz[m], a1, wacc[m][1]);
    }
    bacc += gb;
    o0 = p2_add(o0, o2);
    o1 = p2_add(o1, o3);
    *reinterpret_cast<float2*>(dA1 + r * 128 + 64 * h + lane * 2) =
        make_float2(p2_lo(o0) + p2_hi(o0), p2_lo(o1) + p2_hi(o1));
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();  // all rings are drained: reuse the memory for the block reduction
  // red[warp][(kk*16 + j)*32 + i] for the warp's two heads, then [1024 + lane] for db2
  float* red = smem + warp * kMidHalfPart;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    *reinterpret_cast<float2*>(red + (kk * 16 + 2 * m) * 32 + 2 * c) = make_float2(p2_lo(wacc[m][0]), p2_lo(wacc[m][1]));
    *reinterpret_cast<float2*>(red + (kk * 16 + 2 * m + 1) * 32 + 2 * c) = make_float2(p2_hi(wacc[m][0]), p2_hi(wacc[m][1]));
  }
  red[1024 + lane] = bacc;
  __syncthreads();
  // block partial, warps of a half combined in fixed order.  Layout P[(k*16+j)*32 + i], P[2048 + 16k + j]
  float* P = parts + (size_t)blockIdx.x * kMidPart;
  for (int i = tid; i < 2 * kMidHalfPart; i += kMidBwdThreads) {
    const int hh = i / kMidHalfPart, q = i - hh * kMidHalfPart;
    float t = 0.f;
#pragma unroll
    for (int wq2 = 0; wq2 < kWarps / 2; ++wq2) t += smem[(2 * wq2 + hh) * kMidHalfPart + q];
    P[q < 1024 ? hh * 1024 + q : 2048 + hh * 32 + (q - 1024)] = t;
  }
}

__global__ void parts_reduce_kernel(int nparts, int width, const float* __restrict__ parts, int n_a,
                                    float* __restrict__ out_a, float* __restrict__ out_b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)parts[(size_t)p * width + i];
  if (i < n_a) out_a[i] = (float)s; else if (out_b) out_b[i - n_a] = (float)s;
}

// logits[r][k] = b3[k] + sum_j A2[r][16k+j] W3[k][j]      (16 lanes per row, one float4 each)
__global__ void __launch_bounds__(256)
clf_out_fwd_kernel(long long rows, const float* __restrict__ A2, const float* __restrict__ W3,
                   const float* __restrict__ b3, int sigmoid, float* __restrict__ out) {
  const int q = threadIdx.x & 15, half = (threadIdx.x >> 4) & 1;
  const float4 w = ldg4(W3 + q * 4);
  const float4 b = ldg4(b3);
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = warp_id * 2; r0 < rows; r0 += nwarps * 2) {  // warp-uniform trip count
    const long long r = r0 + half;
    const bool valid = r < rows;
    const float4 a = valid ? ldg4(A2 + r * 64 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float p = a.x * w.x;
    p = fmaf(a.y, w.y, p);
    p = fmaf(a.z, w.z, p);
    p = fmaf(a.w, w.w, p);
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    p += __shfl_xor_sync(0xffffffffu, p, 2);
    const int base = half * 16;  // first lane of this half warp
    float l0 = __shfl_sync(0xffffffffu, p, base + 0), l1 = __shfl_sync(0xffffffffu, p, base + 4);
    float l2 = __shfl_sync(0xffffffffu, p, base + 8), l3 = __shfl_sync(0xffffffffu, p, base + 12);
    if (q == 0 && valid) {
      float4 o = make_float4(l0 + b.x, l1 + b.y, l2 + b.z, l3 + b.w);
      if (sigmoid) {
        o.x = 1.f / (1.f + expf(-o.x)); o.y = 1.f / (1.f + expf(-o.y));
        o.z = 1.f / (1.f + expf(-o.z)); o.w = 1.f / (1.f + expf(-o.w));
      }
      st4(out + r * 4, o);
    }
  }
}

// dA2[r][16k+j] = dz[r][k] W3[k][j]; partial sums of dW3[k][j] = sum_r dz A2, db3[k] = sum_r dz.
// Block = 16 column groups x 16 row lanes.
constexpr int kOutPart = 64 + 4;
__global__ void __launch_bounds__(256)
clf_out_bwd_kernel(long long rows, const float* __restrict__ A2, const float* __restrict__ W3,
                   const float* __restrict__ out, const float* __restrict__ dout, int sigmoid,
                   float* __restrict__ dA2, float* __restrict__ parts) {
  __shared__ float red[256 * 5];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int k = tx >> 2;
  const float4 w = ldg4(W3 + tx * 4);
  float4 sw = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
  for (long long r = (long long)blockIdx.x * 16 + ty; r < rows; r += (long long)gridDim.x * 16) {
    float dz = __ldg(dout + r * 4 + k);
    if (sigmoid) {
      float s = __ldg(out + r * 4 + k);
      dz *= s * (1.f - s);
    }
    const float4 a = ldg4(A2 + r * 64 + tx * 4);
    st4(dA2 + r * 64 + tx * 4, make_float4(dz * w.x, dz * w.y, dz * w.z, dz * w.w));
    sw.x = fmaf(dz, a.x, sw.x); sw.y = fmaf(dz, a.y, sw.y);
    sw.z = fmaf(dz, a.z, sw.z); sw.w = fmaf(dz, a.w, sw.w);
    sb += dz;
  }
  float* my = red + threadIdx.x * 5;
  my[0] = sw.x; my[1] = sw.y; my[2] = sw.z; my[3] = sw.w; my[4] = sb;
  __syncthreads();
  if (threadIdx.x < 64) {  // column c = threadIdx.x
    const int gx = threadIdx.x >> 2, kk = threadIdx.x & 3;
    float s = 0.f;
    for (int y = 0; y < 16; ++y) s += red[(y * 16 + gx) * 5 + kk];
    parts[(size_t)blockIdx.x * kOutPart + threadIdx.x] = s;
  } else if (threadIdx.x < 68) {
    const int kk = threadIdx.x - 64;
    float s = 0.f;
    for (int y = 0; y < 16; ++y) s += red[(y * 16 + kk * 4) * 5 + 4];
    parts[(size_t)blockIdx.x * kOutPart + threadIdx.x] = s;
  }
}

}  // namespace

namespace eg {
int launch_stats_finalize(int nparts, int cols, int stride, long long rows, const double* parts, float* mean,
                          float* var, cudaStream_t s);
}

extern "C" {

int eg_clf_mid_fwd(int64_t rows, const float* A1, const float* W2, const float* b2, float* Z2, float* mean,
                   float* var, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && A1 && W2 && b2 && Z2, "eg_clf_mid_fwd: NULL argument");
  const bool stats = mean && var;
  if (stats && (!ws || ws_bytes < kWorkspaceBytes)) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  const long long want = (rows + 31) / 32;
  const int cap = num_sms() * kMidFwdBlocks;
  const int grid = (int)(want < cap ? want : cap);
  double* parts = stats ? reinterpret_cast<double*>(ws) : nullptr;
  ProfileScope prof("clf_mid_fwd", as_stream(stream));
  clf_mid_fwd_kernel<<<grid, kMidFwdThreads, 0, as_stream(stream)>>>(rows, A1, W2, b2, Z2, parts);
  EG_LAUNCH_CHECK();
  if (stats) return launch_stats_finalize(grid, 64, 64, rows, parts, mean, var, as_stream(stream));
  return EG_OK;
}

int eg_clf_mid_bwd(int64_t rows, const float* A1, const float* W2, const float* dZ2, float* dA1, float* dW2,
                   float* db2, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && A1 && W2 && dZ2 && dA1 && dW2 && db2, "eg_clf_mid_bwd: NULL argument");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  static_assert((size_t)kMidGrid * kMidPart * sizeof(float) <= kWgradBytes, "clf_mid_bwd partials fit the workspace");
  const long long want = (rows + 15) / 16;
  const int cap = num_sms() * kMidBwdBlocks;
  const int grid = (int)(want < cap ? want : cap);
  float* parts = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  ProfileScope prof("clf_mid_bwd", as_stream(stream));
  clf_mid_bwd_kernel<<<grid, kMidBwdThreads, 0, as_stream(stream)>>>(rows, A1, W2, dZ2, dA1, parts);
  EG_LAUNCH_CHECK();
  parts_reduce_kernel<<<(kMidPart + 255) / 256, 256, 0, as_stream(stream)>>>(grid, kMidPart, parts, 2048, dW2, db2);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_clf_out_fwd(int64_t rows, const float* A2, const float* W3, const float* b3, int sigmoid, float* out,
                   void* stream) {
  EG_CHECK_ARG(rows >= 1 && A2 && W3 && b3 && out, "eg_clf_out_fwd: NULL argument");
  long long blocks = (rows * 16 + 255) / 256;
  long long cap = (long long)num_sms() * 16;
  int grid = (int)(blocks < cap ? blocks : cap);
  ProfileScope prof("clf_out_fwd", as_stream(stream));
  clf_out_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, A2, W3, b3, sigmoid, out);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_clf_out_bwd(int64_t rows, const float* A2, const float* W3, const float* out, const float* dout,
                   int sigmoid, float* dA2, float* dW3, float* db3, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && A2 && W3 && dout && dA2 && dW3 && db3, "eg_clf_out_bwd: NULL argument");
  EG_CHECK_ARG(!sigmoid || out, "eg_clf_out_bwd: sigmoid head needs the forward output");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  long long blocks = (rows + 15) / 16;
  int grid = (int)(blocks < kMaxParts ? blocks : kMaxParts);
  float* parts = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  ProfileScope prof("clf_out_bwd", as_stream(stream));
  clf_out_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, A2, W3, out, dout, sigmoid, dA2, parts);
  EG_LAUNCH_CHECK();
  parts_reduce_kernel<<<1, 128, 0, as_stream(stream)>>>(grid, kOutPart, parts, 64, dW3, db3);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // extern "C"
