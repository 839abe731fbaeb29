// The four node classifiers (src/core/models.py:363-377,488-490) as ONE chain per direction that never
// materialises an activated tensor: every BatchNorm / ReLU / Dropout is recomputed from the pre-activation where it
// is consumed, and every BatchNorm-backward reduction runs inside the kernel that produces its addends.
//
//   forward   z1 = h W1^T + b1 (+ column stats)          eg::launch_linear_tc (tcgen05, gcn_tc.cu)        2   U
//             z2 = blockdiag(W2) act1(z1) + b2 (+ stats)  clf_mid_act_fwd_kernel                           1.5 U
//             out = blockdiag(W3) act2(z2) + b3           clf_tail_fwd_kernel                              0.5 U
//   backward  BN2 sums, dW3, db3 from (z2, dout)          clf_tail_bwd_kernel (reduction only)             0.5 U
//             dz2 = BN2-backward(mask2 * dout W3)         clf_tail_dz2_kernel (reads z2; writes dz2)       1   U
//             dW2, db2, g1 = mask1 * (W2^T dz2), BN1 sums clf_mid_act_bwd_kernel (reads z1, dz2; writes g1) 2.5 U
//             dz1 = BN1-backward(g1, z1), in place        bn_bwd_apply (bn.cu)                             3   U
//             dh = dz1 W1, dW1 = dz1^T h, db1             eg::launch_linear_tc / eg::launch_wgrad_tc       4   U
// with U = rows * 128 * 4 bytes: 4 U forward + 11 U backward, against 7 U + 15 U for the r01 chain of separate
// linear / BatchNorm-activation / layer kernels (which wrote and re-read a1, a2, da2, dz2, da1).
// (r02f/h: forming dz2 inside clf_mid_act_bwd_kernel -- one column per lane, exchanged through the warp's ring slot,
// software-pipelined one row ahead -- saves that 1 U but made the kernel instruction-bound: 3.2-3.6 ms against
// 1.4 ms for the r01 kernel; the dropout hash and the index arithmetic are integer work at half the FP32 rate.)
// act(z) = relu(drop(gamma (z - mean) rsqrt(var + eps) + beta)) with the arithmetic of bn.cu, so the sign pattern
// and the dropout masks are bit-identical to eg_bn_act_fwd on the same inputs.
#include "common.cuh"

namespace eg {
int launch_linear_tc(long long rows, const float* A, const float* W, int trans_w, const float* bias,
                     const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                     cudaStream_t s);  // gcn_tc.cu
int launch_wgrad_tc(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s);  // wgrad_tc.cu
int launch_stats_finalize(int nparts, int cols, int stride, long long rows, const double* parts, float* mean,
                          float* var, cudaStream_t s);  // bn.cu
int launch_bn_bwd_apply(long long rows, int cols, const float* G, const float* Z, const float* mean, const float* var,
                        const float* gamma, float eps, const float* coef, float* dZ, cudaStream_t s);  // bn.cu
}  // namespace eg

using namespace eg;

namespace {

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

// BatchNorm-apply + dropout + ReLU of one layer, as the kernels below re-evaluate it on load (pointers to the
// per-column vectors; `thr` / `keep_scale` / `seed` as in bn.cu).
struct Act {
  const float* mean;
  const float* var;
  const float* gamma;
  const float* beta;
  float eps;
  uint32_t thr;
  float keep_scale;
  uint64_t seed;
};
struct Col {  // one column's constants
  float mean, sc /* gamma * invstd */, beta, invstd;
};
__device__ __forceinline__ Col load_col(const Act& a, int c) {
  Col o;
  o.invstd = 1.0f / sqrtf(__ldg(a.var + c) + a.eps);  // torch: 1/sqrt(var+eps), rounded once (as bn.cu)
  o.mean = __ldg(a.mean + c);
  o.sc = __ldg(a.gamma + c) * o.invstd;
  o.beta = __ldg(a.beta + c);
  return o;
}
// value after BN + dropout + ReLU, and whether the gradient passes (kept and BN output > 0)
__device__ __forceinline__ float act_fwd(float z, const Col& c, bool keep, float keep_scale, bool& pass) {
  const float bn = fmaf(z - c.mean, c.sc, c.beta);
  pass = keep && bn > 0.f;
  const float y = keep ? bn * keep_scale : 0.f;
  return fmaxf(y, 0.f);
}

constexpr int kMidFwdThreads = 256, kMidFwdBlocks = 2;  // per SM
constexpr int kMidBwdThreads = 128, kMidBwdBlocks = 4;
constexpr int kMidGrid = kNumSMs * 4;
static_assert((size_t)kMidGrid * 128 * sizeof(double) <= kStatsBytes, "clf_mid_act_fwd partials fit the stats area");

// ---- forward, layer 4:  Z2[r][16k+j] = b2[k][j] + sum_i act1(Z1)[r][32k+i] W2[k][j][i]  (+ column statistics) ------
// Lane mapping of classifier.cu's clf_mid_fwd_kernel: a warp owns whole rows, lane = 8k + c <-> head k, inputs
// 4c..4c+3 = the lane's float4 of the row's 512 contiguous bytes (columns 4 lane .. 4 lane + 3 of Z1, whose
// BatchNorm constants the lane keeps in registers); 32 FFMA2 per row and lane, 14-shuffle transpose-reduce.
__global__ void __launch_bounds__(kMidFwdThreads, kMidFwdBlocks)
clf_mid_act_fwd_kernel(long long rows, const float* __restrict__ Z1, const Act act, const float* __restrict__ W2,
                       const float* __restrict__ b2, float* __restrict__ Z2, double* __restrict__ parts) {
  constexpr int kWarps = kMidFwdThreads / 32;
  constexpr int kDepth = 8;
  __shared__ double red[kWarps][128];
  __shared__ __align__(16) float ring[kWarps * kDepth * 128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, k = lane >> 3, c = lane & 7;
  P2 wp[8][4];  // wp[jp][e] = (W2[k][2(jp^c)][4c+e], W2[k][2(jp^c)+1][4c+e])
#pragma unroll
  for (int jp = 0; jp < 8; ++jp) {
    const int j = 2 * (jp ^ c);
    const float4 w0 = ldg4(W2 + (k * 16 + j) * 32 + c * 4), w1 = ldg4(W2 + (k * 16 + j + 1) * 32 + c * 4);
    wp[jp][0] = p2(w0.x, w1.x);
    wp[jp][1] = p2(w0.y, w1.y);
    wp[jp][2] = p2(w0.z, w1.z);
    wp[jp][3] = p2(w0.w, w1.w);
  }
  Col col[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) col[e] = load_col(act, lane * 4 + e);
  const P2 bias = p2(__ldg(b2 + lane * 2), __ldg(b2 + lane * 2 + 1));
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  double ds0 = 0.0, ds1 = 0.0, dq0 = 0.0, dq1 = 0.0;
  int since_flush = 0;
  const long long wid = (long long)blockIdx.x * kWarps + warp, nw = (long long)gridDim.x * kWarps;
  const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring) + warp * kDepth * 512 + lane * 16;
  auto issue = [&](long long r, int slot) {
    if (r < rows)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_u + slot * 512), "l"(Z1 + r * 128 + lane * 4) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int d = 0; d < kDepth - 1; ++d) issue(wid + d * nw, d);
  int slot = 0;
  for (long long r = wid; r < rows; r += nw) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kDepth - 2) : "memory");
    // each lane reads back exactly the 16 bytes it copied itself: no warp barrier needed
    const float4 z = *reinterpret_cast<const float4*>(ring + (warp * kDepth + slot) * 128 + lane * 4);
    issue(r + (long long)(kDepth - 1) * nw, (slot + kDepth - 1) % kDepth);
    slot = (slot + 1) % kDepth;
    const uint32_t keep = act.thr ? drop_keep4(act.seed, (uint64_t)(r * 32 + lane), act.thr) : 0xfu;
    bool pass;
    const float a0 = act_fwd(z.x, col[0], keep & 1u, act.keep_scale, pass);
    const float a1 = act_fwd(z.y, col[1], keep & 2u, act.keep_scale, pass);
    const float a2 = act_fwd(z.z, col[2], keep & 4u, act.keep_scale, pass);
    const float a3 = act_fwd(z.w, col[3], keep & 8u, act.keep_scale, pass);
    const P2 ax = p2(a0, a0), ay = p2(a1, a1), az = p2(a2, a2), aw = p2(a3, a3);
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
    const P2 zz = p2_add(p2_add(acc[0], p2_shfl_xor(acc[1], 1)), bias);
    const float z0 = p2_lo(zz), z1 = p2_hi(zz);
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

// ---- forward, layer 8:  out[r][k] = b3[k] + sum_j act2(Z2)[r][16k+j] W3[k][j]   (16 lanes per row) ---------------
__global__ void __launch_bounds__(256)
clf_tail_fwd_kernel(long long rows, const float* __restrict__ Z2, const Act act, const float* __restrict__ W3,
                    const float* __restrict__ b3, int sigmoid, float* __restrict__ out) {
  const int q = threadIdx.x & 15, half = (threadIdx.x >> 4) & 1;
  const float4 w = ldg4(W3 + q * 4);
  const float4 b = ldg4(b3);
  Col col[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) col[e] = load_col(act, q * 4 + e);
  const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  constexpr int kU = 4;  // row pairs in flight per warp: one 128-bit load per thread and row is not enough to cover the
                         // DRAM latency (one pair per iteration ran at 2.9 TB/s)
  for (long long r0 = warp_id * 2; r0 < rows; r0 += nwarps * 2 * kU) {  // warp-uniform trip count
    float4 z[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * nwarps * 2 + half;
      z[u] = r < rows ? ldg4(Z2 + r * 64 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * nwarps * 2 + half;
      if (r0 + u * nwarps * 2 >= rows) break;  // warp-uniform
      const bool valid = r < rows;
      const uint32_t keep = act.thr ? drop_keep4(act.seed, (uint64_t)(r * 16 + q), act.thr) : 0xfu;
      bool pass;
      float p = act_fwd(z[u].x, col[0], keep & 1u, act.keep_scale, pass) * w.x;
      p = fmaf(act_fwd(z[u].y, col[1], keep & 2u, act.keep_scale, pass), w.y, p);
      p = fmaf(act_fwd(z[u].z, col[2], keep & 4u, act.keep_scale, pass), w.z, p);
      p = fmaf(act_fwd(z[u].w, col[3], keep & 8u, act.keep_scale, pass), w.w, p);
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
}

// ---- backward, layer 8 (reduction only): per-block partials of dW3[k][j] = sum_r dz a2, db3[k] = sum_r dz and of
// the BatchNorm-2 backward sums s1[c] = sum_r dA, s2[c] = sum_r dA xhat, where dz = dout (* s (1 - s) for the sigmoid
// head), a2 = act2(z2), dA = pass ? dz W3 keep_scale : 0.  Nothing of rows x 64 is written.
// Block = 16 column groups x 16 row lanes.  Partial layout: [64 dW3][4 db3][64 s1][64 s2].
constexpr int kTailPart = 64 + 4 + 64 + 64;
__global__ void __launch_bounds__(256, 2)
clf_tail_bwd_kernel(long long rows, const float* __restrict__ Z2, const Act act, const float* __restrict__ W3,
                    const float* __restrict__ out, const float* __restrict__ dout, int sigmoid,
                    float* __restrict__ parts) {
  __shared__ float red[256 * 13];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int k = tx >> 2;
  const float4 w4 = ldg4(W3 + tx * 4);
  const float w[4] = {w4.x, w4.y, w4.z, w4.w};
  Col col[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) col[e] = load_col(act, tx * 4 + e);
  float sw[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f}, sb = 0.f;
  float tw[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f}, tb = 0.f;
  int run = 0;
  constexpr int kU = 4;  // rows in flight per thread (memory-level parallelism: the loop is a dependent load -> math chain)
  const long long rstep = (long long)gridDim.x * 16;
  for (long long r0 = (long long)blockIdx.x * 16 + ty; r0 < rows; r0 += rstep * kU) {
    float4 z4[kU];
    float dzu[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * rstep;
      const bool ok = r < rows;
      z4[u] = ok ? ldg4(Z2 + r * 64 + tx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      dzu[u] = ok ? __ldg(dout + r * 4 + k) : 0.f;
      if (sigmoid && ok) {
        const float s = __ldg(out + r * 4 + k);
        dzu[u] *= s * (1.f - s);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = r0 + u * rstep;
      if (r >= rows) break;
      const float dz = dzu[u];
      const float z[4] = {z4[u].x, z4[u].y, z4[u].z, z4[u].w};
      const uint32_t keep = act.thr ? drop_keep4(act.seed, (uint64_t)(r * 16 + tx), act.thr) : 0xfu;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        bool pass;
        const float a2 = act_fwd(z[e], col[e], (keep >> e) & 1u, act.keep_scale, pass);
        const float dA = pass ? dz * w[e] * act.keep_scale : 0.f;
        sw[e] = fmaf(dz, a2, sw[e]);
        s1[e] += dA;
        s2[e] = fmaf(dA, (z[e] - col[e].mean) * col[e].invstd, s2[e]);
      }
      sb += dz;
    }
    if (++run == 16) {  // bounded fp32 runs (64 rows)
#pragma unroll
      for (int e = 0; e < 4; ++e) { tw[e] += sw[e]; t1[e] += s1[e]; t2[e] += s2[e]; sw[e] = s1[e] = s2[e] = 0.f; }
      tb += sb;
      sb = 0.f;
      run = 0;
    }
  }
  float* my = red + threadIdx.x * 13;
#pragma unroll
  for (int e = 0; e < 4; ++e) { my[e] = tw[e] + sw[e]; my[4 + e] = t1[e] + s1[e]; my[8 + e] = t2[e] + s2[e]; }
  my[12] = tb + sb;
  __syncthreads();
  if (threadIdx.x < 64) {  // column c = threadIdx.x: fixed-order sum over the 16 row lanes
    const int gx = threadIdx.x >> 2, e = threadIdx.x & 3;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int y = 0; y < 16; ++y) {
      const float* src = red + (y * 16 + gx) * 13;
      a += src[e];
      b += src[4 + e];
      c += src[8 + e];
    }
    float* P = parts + (size_t)blockIdx.x * kTailPart;
    P[threadIdx.x] = a;
    P[68 + threadIdx.x] = b;
    P[132 + threadIdx.x] = c;
  } else if (threadIdx.x < 68) {
    const int kk = threadIdx.x - 64;
    float s = 0.f;
    for (int y = 0; y < 16; ++y) s += red[(y * 16 + kk * 4) * 13 + 12];
    parts[(size_t)blockIdx.x * kTailPart + threadIdx.x] = s;
  }
}

// dz2[r][c] = sc2[c] (dA - c1[c] - xhat c2[c]),  dA = pass2 ? dout[r][head] (* s (1 - s)) W3[c] keep_scale : 0: the
// BatchNorm-2 backward applied to the masked layer-8 gradient, written once ([rows,64]) for clf_mid_act_bwd_kernel.
__global__ void __launch_bounds__(256)
clf_tail_dz2_kernel(long long rows, const float* __restrict__ Z2, const Act act, const float* __restrict__ W3,
                    const float* __restrict__ out, const float* __restrict__ dout, int sigmoid,
                    const float* __restrict__ coef, float* __restrict__ dZ2) {
  const int tx = threadIdx.x & 15, k = tx >> 2;  // 256 % 16 == 0 and the stride is a multiple of 16: one column group per thread
  const float4 w4 = ldg4(W3 + tx * 4);
  const float w[4] = {w4.x, w4.y, w4.z, w4.w};
  Col col[4];
  float c1[4], c2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    col[e] = load_col(act, tx * 4 + e);
    c1[e] = __ldg(coef + tx * 4 + e);
    c2[e] = __ldg(coef + 64 + tx * 4 + e);
  }
  const long long total = rows * 16, step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    const bool two = i + step < total;
    const long long i2 = two ? i + step : i;
    const float4 za = ldg4(Z2 + i * 4), zb = ldg4(Z2 + i2 * 4);
    float da = __ldg(dout + (i >> 4) * 4 + k), db = __ldg(dout + (i2 >> 4) * 4 + k);
    if (sigmoid) {
      const float sa = __ldg(out + (i >> 4) * 4 + k), sb = __ldg(out + (i2 >> 4) * 4 + k);
      da *= sa * (1.f - sa);
      db *= sb * (1.f - sb);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long iu = u ? i2 : i;
      const float4 z4 = u ? zb : za;
      const float dz = u ? db : da;
      const float z[4] = {z4.x, z4.y, z4.z, z4.w};
      float o[4];
      const uint32_t keep = act.thr ? drop_keep4(act.seed, (uint64_t)iu, act.thr) : 0xfu;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xc = z[e] - col[e].mean;
        const float bn = fmaf(xc, col[e].sc, col[e].beta);
        const bool pass = ((keep >> e) & 1u) && bn > 0.f;
        const float dA = pass ? dz * w[e] * act.keep_scale : 0.f;
        o[e] = col[e].sc * (dA - c1[e] - xc * col[e].invstd * c2[e]);
      }
      st4(dZ2 + iu * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

// Sums per-block float partials in double (fixed order): out_a[0..n_a), out_b[0..n_b) from the first n_a + n_b
// entries; the next `cols` entries are BatchNorm sums s1 and the `cols` after them s2: dbeta = s1, dgamma = s2,
// coef = (s1 / rows, s2 / rows) for a train-mode BatchNorm, zeros for an eval-mode one (dz = sc * dA).
__global__ void clf_parts_finalize_kernel(int nparts, int width, const float* __restrict__ parts, int n_a, int n_b,
                                          int cols, long long rows, int batch_stats, float* __restrict__ out_a,
                                          float* __restrict__ out_b, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta, float* __restrict__ coef) {
  // block = 32 outputs x 8 splits: split k sums partials k, k + 8, ... in order, the 8 sub-sums are combined in order
  // (one thread per output walked all <= 592 partials: 36 us)
  __shared__ double red[8][32];
  const int i = blockIdx.x * 32 + threadIdx.x, k = threadIdx.y;
  double s = 0.0;
  if (i < width)
    for (int p = k; p < nparts; p += 8) s += (double)parts[(size_t)p * width + i];
  red[k][threadIdx.x] = s;
  __syncthreads();
  if (k != 0 || i >= width) return;
  s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += red[q][threadIdx.x];
  if (i < n_a) {
    out_a[i] = (float)s;
  } else if (i < n_a + n_b) {
    out_b[i - n_a] = (float)s;
  } else if (i < n_a + n_b + cols) {
    const int c = i - n_a - n_b;
    dbeta[c] = (float)s;
    coef[c] = batch_stats ? (float)(s / (double)rows) : 0.f;
  } else {
    const int c = i - n_a - n_b - cols;
    dgamma[c] = (float)s;
    coef[cols + c] = batch_stats ? (float)(s / (double)rows) : 0.f;
  }
}

// ---- backward, layer 4 ---------------------------------------------------------------------------------------------
// Per half row (heads 2h, 2h+1; a warp owns half rows, as classifier.cu's clf_mid_bwd_kernel):
//   a1[i]   = act1(z1)[i], mask pass1[i]                                             (lane = inputs 2c, 2c+1 of its head)
//   dW2    += dz2 (x) a1,  db2 += dz2,  da1 = W2^T dz2,  g1 = pass1 ? da1 keep_scale : 0   (written, [rows,128])
//   BN1 sums s1 += g1, s2 += g1 xhat1
// Slot of the per-warp cp.async ring: [256 B z1 half row][128 B dz2 half row].
constexpr int kMidDepth = 8;
constexpr int kMidSlotBytes = 256 + 128;
constexpr int kMidHalfPart = 1024 + 32 + 64 + 64;  // dW2 of 2 heads, db2 of 32 columns, s1 / s2 of 64 columns
constexpr int kMidPart = 2 * kMidHalfPart;         // [2048 dW2][64 db2][128 s1][128 s2]
static_assert((size_t)kMidGrid * kMidPart * sizeof(float) <= kWgradBytes, "clf_mid_act_bwd partials fit the workspace");
__global__ void __launch_bounds__(kMidBwdThreads, kMidBwdBlocks)
clf_mid_act_bwd_kernel(long long rows, const float* __restrict__ Z1, const float* __restrict__ dZ2,
                       const float* __restrict__ W2, const Act act1, float* __restrict__ G1,
                       float* __restrict__ parts) {
  constexpr int kWarps = kMidBwdThreads / 32;
  constexpr int kRingFloats = kWarps * kMidDepth * kMidSlotBytes / 4, kRedFloats = kWarps * kMidHalfPart;
  __shared__ __align__(16) float smem[kRedFloats > kRingFloats ? kRedFloats : kRingFloats];  // rings, then the block reduction
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, h = warp & 1, kk = lane >> 4, c = lane & 15;
  const int head = 2 * h + kk;
  P2 wq[8][2];  // wq[m][e] = (W2[head][2m][2c+e], W2[head][2m+1][2c+e])
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
  // this lane owns columns 64h + 2 lane, + 1 of Z1 / g1 (= inputs 2c, 2c+1 of its head) and column 32h + lane of dz2 (db2)
  const int c1i = 64 * h + 2 * lane;
  const Col col1a = load_col(act1, c1i), col1b = load_col(act1, c1i + 1);
  float bacc = 0.f, s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;       // current fp32 runs
  float tbacc = 0.f, t1a = 0.f, t1b = 0.f, t2a = 0.f, t2b = 0.f;      // sums of finished runs
  int run = 0;
  const long long wid = (long long)blockIdx.x * (kWarps / 2) + (warp >> 1), nw = (long long)gridDim.x * (kWarps / 2);
  float* ringf = smem + warp * (kMidDepth * kMidSlotBytes / 4);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ringf);
  // lanes 0-15 copy z1, lanes 16-23 dz2; 16 bytes each
  const float* src0 = lane < 16 ? Z1 + 64 * h + lane * 4 : dZ2 + 32 * h + (lane - 16) * 4;
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
    const float2 z1 = *reinterpret_cast<const float2*>(row + lane * 2);
    const float4 g0 = *reinterpret_cast<const float4*>(row + 64 + kk * 16),
                 g1 = *reinterpret_cast<const float4*>(row + 64 + kk * 16 + 4),
                 g2 = *reinterpret_cast<const float4*>(row + 64 + kk * 16 + 8),
                 g3 = *reinterpret_cast<const float4*>(row + 64 + kk * 16 + 12);
    const float dz_own = row[64 + lane];
    __syncwarp();  // every lane has read the slot: refill it with the row kMidDepth - 1 ahead
    issue(r + (long long)(kMidDepth - 1) * nw, (slot + kMidDepth - 1) % kMidDepth);
    slot = (slot + 1) % kMidDepth;
    // a1 of inputs 2c, 2c+1 (columns 64h + 2 lane, + 1 of Z1)
    const uint32_t keep1 = act1.thr ? drop_keep4(act1.seed, (uint64_t)(r * 32 + 16 * h + (lane >> 1)), act1.thr) : 0xfu;
    bool pass1a, pass1b;
    const float a1a = act_fwd(z1.x, col1a, (keep1 >> ((lane & 1) * 2)) & 1u, act1.keep_scale, pass1a);
    const float a1b = act_fwd(z1.y, col1b, (keep1 >> ((lane & 1) * 2 + 1)) & 1u, act1.keep_scale, pass1b);
    const P2 dz[8] = {p2(g0.x, g0.y), p2(g0.z, g0.w), p2(g1.x, g1.y), p2(g1.z, g1.w),
                      p2(g2.x, g2.y), p2(g2.z, g2.w), p2(g3.x, g3.y), p2(g3.z, g3.w)};
    const P2 a0 = p2(a1a, a1a), a1 = p2(a1b, a1b);
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
    o0 = p2_add(o0, o2);
    o1 = p2_add(o1, o3);
    const float ga = pass1a ? (p2_lo(o0) + p2_hi(o0)) * act1.keep_scale : 0.f;
    const float gb = pass1b ? (p2_lo(o1) + p2_hi(o1)) * act1.keep_scale : 0.f;
    *reinterpret_cast<float2*>(G1 + r * 128 + c1i) = make_float2(ga, gb);
    bacc += dz_own;
    s1a += ga;
    s1b += gb;
    s2a = fmaf(ga, (z1.x - col1a.mean) * col1a.invstd, s2a);
    s2b = fmaf(gb, (z1.y - col1b.mean) * col1b.invstd, s2b);
    if (++run == 64) {
      tbacc += bacc; t1a += s1a; t1b += s1b; t2a += s2a; t2b += s2b;
      bacc = s1a = s1b = s2a = s2b = 0.f;
      run = 0;
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();  // all rings are drained: reuse the memory for the block reduction
  // red[warp]: [(kk*16 + j)*32 + i] dW2 of the warp's two heads, [1024 + lane] db2, [1056 + 2 lane + e] s1, [1120 + ..] s2
  float* red = smem + warp * kMidHalfPart;
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    *reinterpret_cast<float2*>(red + (kk * 16 + 2 * m) * 32 + 2 * c) = make_float2(p2_lo(wacc[m][0]), p2_lo(wacc[m][1]));
    *reinterpret_cast<float2*>(red + (kk * 16 + 2 * m + 1) * 32 + 2 * c) = make_float2(p2_hi(wacc[m][0]), p2_hi(wacc[m][1]));
  }
  red[1024 + lane] = tbacc + bacc;
  *reinterpret_cast<float2*>(red + 1056 + 2 * lane) = make_float2(t1a + s1a, t1b + s1b);
  *reinterpret_cast<float2*>(red + 1120 + 2 * lane) = make_float2(t2a + s2a, t2b + s2b);
  __syncthreads();
  // block partial, warps of a half combined in fixed order.  Layout: [2048 dW2][64 db2][128 s1][128 s2]
  float* P = parts + (size_t)blockIdx.x * kMidPart;
  for (int i = tid; i < 2 * kMidHalfPart; i += kMidBwdThreads) {
    const int hh = i / kMidHalfPart, q = i - hh * kMidHalfPart;
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < kWarps / 2; ++w2) t += smem[(2 * w2 + hh) * kMidHalfPart + q];
    int dst;
    if (q < 1024) dst = hh * 1024 + q;
    else if (q < 1056) dst = 2048 + hh * 32 + (q - 1024);
    else if (q < 1120) dst = 2112 + hh * 64 + (q - 1056);
    else dst = 2240 + hh * 64 + (q - 1120);
    P[dst] = t;
  }
}

Act make_act(const float* mean, const float* var, const float* gamma, const float* beta, float eps, float drop_p,
             uint64_t seed) {
  Act a;
  a.mean = mean, a.var = var, a.gamma = gamma, a.beta = beta, a.eps = eps;
  a.thr = drop_threshold(drop_p);
  a.keep_scale = a.thr ? 1.0f / (1.0f - drop_p) : 1.0f;
  a.seed = seed;
  return a;
}

}  // namespace

extern "C" {

int eg_classifier_fwd(int64_t rows, const float* h, const eg_classifier_params* p, float* mean1, float* var1,
                      float* mean2, float* var2, float* z1, float* z2, float* out, void* ws, size_t ws_bytes,
                      void* stream) {
  EG_CHECK_ARG(rows >= 1 && h && p && mean1 && var1 && mean2 && var2 && z1 && z2 && out, "eg_classifier_fwd: NULL argument");
  EG_CHECK_ARG(p->w1 && p->b1 && p->g1 && p->be1 && p->w2 && p->b2 && p->g2 && p->be2 && p->w3 && p->b3,
               "eg_classifier_fwd: NULL parameter");
  EG_CHECK_ARG(p->drop_p >= 0.f && p->drop_p < 1.f, "eg_classifier_fwd: drop_p must be in [0,1)");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  const bool stats = p->batch_stats != 0;
  int rc = launch_linear_tc(rows, h, p->w1, 1, p->b1, nullptr, z1, stats ? mean1 : nullptr, stats ? var1 : nullptr, ws,
                            ws_bytes, s);
  if (rc != EG_OK) return rc;
  const Act act1 = make_act(mean1, var1, p->g1, p->be1, p->eps, p->drop_p, p->seed);
  const Act act2 = make_act(mean2, var2, p->g2, p->be2, p->eps, p->drop_p, p->seed + 1);
  {
    const long long want = (rows + 31) / 32;
    const int cap = num_sms() * kMidFwdBlocks;
    const int grid = (int)(want < cap ? want : cap);
    double* parts = stats ? reinterpret_cast<double*>(ws) : nullptr;
    {
      ProfileScope prof("clf_mid_act_fwd", s);
      clf_mid_act_fwd_kernel<<<grid, kMidFwdThreads, 0, s>>>(rows, z1, act1, p->w2, p->b2, z2, parts);
      EG_LAUNCH_CHECK();
    }
    if (stats) {
      rc = launch_stats_finalize(grid, 64, 64, rows, parts, mean2, var2, s);
      if (rc != EG_OK) return rc;
    }
  }
  {
    long long blocks = (rows * 16 + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    ProfileScope prof("clf_tail_fwd", s);
    clf_tail_fwd_kernel<<<grid, 256, 0, s>>>(rows, z2, act2, p->w3, p->b3, p->sigmoid, out);
    EG_LAUNCH_CHECK();
  }
  return EG_OK;
}

int eg_classifier_bwd(int64_t rows, const float* h, const eg_classifier_params* p, const float* mean1,
                      const float* var1, const float* mean2, const float* var2, const float* z1, const float* z2,
                      const float* out, const float* dout, float* scratch, float* dh,
                      const eg_classifier_grads* g, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && h && p && mean1 && var1 && mean2 && var2 && z1 && z2 && dout && scratch && g,
               "eg_classifier_bwd: NULL argument");
  EG_CHECK_ARG(!p->sigmoid || out, "eg_classifier_bwd: the sigmoid head needs the forward output");
  EG_CHECK_ARG(g->dw1 && g->db1 && g->dg1 && g->dbe1 && g->dw2 && g->db2 && g->dg2 && g->dbe2 && g->dw3 && g->db3,
               "eg_classifier_bwd: NULL gradient output");
  EG_CHECK_ARG(p->drop_p >= 0.f && p->drop_p < 1.f, "eg_classifier_bwd: drop_p must be in [0,1)");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  const Act act1 = make_act(mean1, var1, p->g1, p->be1, p->eps, p->drop_p, p->seed);
  const Act act2 = make_act(mean2, var2, p->g2, p->be2, p->eps, p->drop_p, p->seed + 1);
  float* parts = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  // BatchNorm backward coefficients live in the slack behind the two partial areas (untouched by the kernels below)
  float* coef1 = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes + kWgradBytes);  // [256]
  float* coef2 = coef1 + 256;                                                                        // [128]
  {
    long long blocks = (rows + 15) / 16;
    const int cap = num_sms() * 2;  // __launch_bounds__(256, 2): one resident wave
    const int grid = (int)(blocks < cap ? blocks : cap);
    ProfileScope prof("clf_tail_bwd", s);
    clf_tail_bwd_kernel<<<grid, 256, 0, s>>>(rows, z2, act2, p->w3, out, dout, p->sigmoid, parts);
    EG_LAUNCH_CHECK();
    clf_parts_finalize_kernel<<<(kTailPart + 31) / 32, dim3(32, 8), 0, s>>>(grid, kTailPart, parts, 64, 4, 64, rows,
                                                                      p->batch_stats, g->dw3, g->db3, g->dg2, g->dbe2,
                                                                      coef2);
    EG_LAUNCH_CHECK();
  }
  {
    float* dz2 = scratch + (size_t)rows * 128;  // scratch = [rows,128] g1 / dz1 followed by [rows,64] dz2
    const long long groups = rows * 16;
    long long blocks = (groups + 255) / 256;
    const long long capb = (long long)num_sms() * 16;
    const int grid_e = (int)(blocks < capb ? (blocks < 1 ? 1 : blocks) : capb);
    {
      ProfileScope prof("clf_tail_dz2", s);
      clf_tail_dz2_kernel<<<grid_e, 256, 0, s>>>(rows, z2, act2, p->w3, out, dout, p->sigmoid, coef2, dz2);
      EG_LAUNCH_CHECK();
    }
    const long long want = (rows + 15) / 16;
    const int cap = num_sms() * kMidBwdBlocks;
    const int grid = (int)(want < cap ? want : cap);
    ProfileScope prof("clf_mid_act_bwd", s);
    clf_mid_act_bwd_kernel<<<grid, kMidBwdThreads, 0, s>>>(rows, z1, dz2, p->w2, act1, scratch, parts);
    EG_LAUNCH_CHECK();
    clf_parts_finalize_kernel<<<(kMidPart + 31) / 32, dim3(32, 8), 0, s>>>(grid, kMidPart, parts, 2048, 64, 128, rows,
                                                                     p->batch_stats, g->dw2, g->db2, g->dg1, g->dbe1,
                                                                     coef1);
    EG_LAUNCH_CHECK();
  }
  int rc = launch_bn_bwd_apply(rows, 128, scratch, z1, mean1, var1, p->g1, p->eps, coef1, scratch, s);  // g1 -> dz1 in place
  if (rc != EG_OK) return rc;
  rc = launch_wgrad_tc(rows, scratch, h, g->dw1, g->db1, ws, ws_bytes, s);
  if (rc != EG_OK) return rc;
  if (dh) rc = launch_linear_tc(rows, scratch, p->w1, 0, nullptr, nullptr, dh, nullptr, nullptr, nullptr, 0, s);
  return rc;
}

}  // extern "C"
