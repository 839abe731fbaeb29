// Classifier tail: the four independent heads of node_classifiers (src/core/models.py:363-377) evaluated
// together as block-diagonal transforms.  Layer 0 of the four heads is one stacked [128 -> 4x32]
// tensor-core GEMM (eg_linear128); this file holds layer 4 (4 x [32 -> 16]) and layer 8 (4 x [16 -> 1]).
// The BN/ReLU/Dropout between them is eg_bn_act_* with cols = 128 / 64.
#include "common.cuh"

using namespace eg;

namespace {

constexpr int kMidThreads = 128;  // warp k <-> head k, lane <-> row of the 32-row tile
constexpr int kMidGrid = kNumSMs * 6;  // 6 resident blocks per SM (34 KB of shared memory each)
static_assert((size_t)kMidGrid * 128 * sizeof(double) <= kStatsBytes, "clf_mid_fwd partials fit the stats area");
constexpr int TR = 32;
constexpr int LDA = 132;          // A1 tile stride
constexpr int LDO = 68;           // Z2 tile stride

// sum over the 32 lanes of v[i] for i = 0..31; on return lane l holds the total of v[l] (31 shuffles)
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}
// (x, y) += a * (w.x, w.y): one packed FFMA2
__device__ __forceinline__ void fma2(float& x, float& y, float a, float wx, float wy) {
  unsigned long long c = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(x);
  const unsigned long long a2 = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
  const unsigned long long w2 = ((unsigned long long)__float_as_uint(wy) << 32) | __float_as_uint(wx);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a2), "l"(w2));
  x = __uint_as_float((uint32_t)c);
  y = __uint_as_float((uint32_t)(c >> 32));
}

// Z2[r][16k+j] = b2[k][j] + sum_i A1[r][32k+i] * W2[k][j][i]
// Warp k <-> head k; a thread owns TWO rows (lane, lane + 32 of a 64-row tile) and reads its 128-byte slices of
// A1 straight from global (16 independent 128-bit loads in flight per thread, no shared-memory staging, no block
// barrier in the loop); the transposed weight Wt[k][i][0..15] is read as 4 broadcast LDS.128 per input feature and
// shared by both rows; packed FFMA2.  Column statistics: 31-shuffle transpose-reduce per tile, double per thread.
constexpr int TR2 = 64;
__global__ void __launch_bounds__(kMidThreads)
clf_mid_fwd_kernel(long long rows, const float* __restrict__ A1, const float* __restrict__ W2,
                   const float* __restrict__ b2, float* __restrict__ Z2, double* __restrict__ parts) {
  __shared__ __align__(16) float Wt[4 * 32 * 16];  // Wt[k][i][j] = W2[k][j][i]
  __shared__ float bs[64];
  const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;
  for (int i = tid; i < 2048; i += kMidThreads) {
    const int kk = i >> 9, j = (i >> 5) & 15, ii = i & 31;
    Wt[(kk * 32 + ii) * 16 + j] = __ldg(W2 + i);
  }
  if (tid < 64) bs[tid] = __ldg(b2 + tid);
  __syncthreads();
  double run = 0.0;  // lane l < 16: sum of column 16k + l; lane 16 + l: sum of squares of that column
  const long long ntiles = (rows + TR2 - 1) / TR2;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * TR2 + lane, r1 = r0 + 32;
    const bool ok0 = r0 < rows, ok1 = r1 < rows;
    float4 a0[8], a1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      a0[q] = ok0 ? ldg4(A1 + r0 * 128 + k * 32 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      a1[q] = ok1 ? ldg4(A1 + r1 * 128 + k * 32 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float z0[16], z1[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z0[j] = z1[j] = bs[k * 16 + j];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float av0[4] = {a0[q].x, a0[q].y, a0[q].z, a0[q].w}, av1[4] = {a1[q].x, a1[q].y, a1[q].z, a1[q].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float* wrow = Wt + (k * 32 + q * 4 + e) * 16;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(wrow + j4 * 4);
          fma2(z0[4 * j4], z0[4 * j4 + 1], av0[e], w.x, w.y);
          fma2(z0[4 * j4 + 2], z0[4 * j4 + 3], av0[e], w.z, w.w);
          fma2(z1[4 * j4], z1[4 * j4 + 1], av1[e], w.x, w.y);
          fma2(z1[4 * j4 + 2], z1[4 * j4 + 3], av1[e], w.z, w.w);
        }
      }
    }
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      if (ok0) st4(Z2 + r0 * 64 + k * 16 + j4 * 4, make_float4(z0[4 * j4], z0[4 * j4 + 1], z0[4 * j4 + 2], z0[4 * j4 + 3]));
      if (ok1) st4(Z2 + r1 * 64 + k * 16 + j4 * 4, make_float4(z1[4 * j4], z1[4 * j4 + 1], z1[4 * j4 + 2], z1[4 * j4 + 3]));
    }
    if (parts) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x0 = ok0 ? z0[j] : 0.f, x1 = ok1 ? z1[j] : 0.f;
        v[j] = x0 + x1;
        v[16 + j] = fmaf(x0, x0, x1 * x1);
      }
      run += (double)transpose_reduce32(v, lane);
    }
  }
  if (parts) {  // parts[block][0..63] sums, [64..127] sums of squares
    const int col = k * 16 + (lane & 15);
    parts[(size_t)blockIdx.x * 128 + (lane < 16 ? col : 64 + col)] = run;
  }
}

// dA1[r][32k+i] = sum_j dZ2[r][16k+j] W2[k][j][i];  per-block partials of dW2[k][j][i], db2[k][j].
constexpr int kMidPart = 2048 + 64;
__global__ void __launch_bounds__(kMidThreads, 6)
clf_mid_bwd_kernel(long long rows, const float* __restrict__ A1, const float* __restrict__ W2,
                   const float* __restrict__ dZ2, float* __restrict__ dA1, float* __restrict__ parts) {
  __shared__ __align__(16) float As[TR * LDA];   // A1 tile, then reused for the dA1 tile
  __shared__ __align__(16) float Ws[4 * 16 * 32];
  __shared__ __align__(16) float Gs[TR * LDO];   // dZ2 tile
  const int tid = threadIdx.x, lane = tid & 31, k = tid >> 5;
  for (int i = tid; i < 2048; i += kMidThreads) Ws[i] = __ldg(W2 + i);
  float wacc[16];  // dW2[k][j][lane]
#pragma unroll
  for (int j = 0; j < 16; ++j) wacc[j] = 0.f;
  float bacc = 0.f;  // db2[k][lane] for lane < 16
  const long long ntiles = (rows + TR - 1) / TR;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TR;
    const int nvalid = (int)min((long long)TR, rows - row0);
    __syncthreads();
    for (int i = tid; i < TR * 32; i += kMidThreads) {
      int r = i >> 5, c4 = i & 31;
      float4 v = r < nvalid ? ldg4(A1 + (row0 + r) * 128 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(As + r * LDA + c4 * 4) = v;
    }
    for (int i = tid; i < TR * 16; i += kMidThreads) {
      int r = i >> 4, c4 = i & 15;
      float4 v = r < nvalid ? ldg4(dZ2 + (row0 + r) * 64 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(Gs + r * LDO + c4 * 4) = v;
    }
    __syncthreads();
    // weight / bias gradient: lane <-> input feature i
#pragma unroll 4
    for (int r = 0; r < TR; ++r) {
      const float a = As[r * LDA + k * 32 + lane];
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // packed FFMA2: two output features per issue slot
        const float4 gz = *reinterpret_cast<const float4*>(Gs + r * LDO + k * 16 + q * 4);
        fma2(wacc[4 * q], wacc[4 * q + 1], a, gz.x, gz.y);
        fma2(wacc[4 * q + 2], wacc[4 * q + 3], a, gz.z, gz.w);
      }
      if (lane < 16) bacc += Gs[r * LDO + k * 16 + lane];
    }
    // input gradient: lane <-> row
    float dz[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 gz = *reinterpret_cast<const float4*>(Gs + lane * LDO + k * 16 + q * 4);
      dz[4 * q] = gz.x; dz[4 * q + 1] = gz.y; dz[4 * q + 2] = gz.z; dz[4 * q + 3] = gz.w;
    }
    __syncthreads();  // everyone is done reading the A1 tile -> reuse As for dA1
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + (k * 16 + j) * 32 + i4 * 4);
        fma2(o.x, o.y, dz[j], w.x, w.y);
        fma2(o.z, o.w, dz[j], w.z, w.w);
      }
      *reinterpret_cast<float4*>(As + lane * LDA + k * 32 + i4 * 4) = o;
    }
    __syncthreads();
    for (int i = tid; i < TR * 32; i += kMidThreads) {
      int r = i >> 5, c4 = i & 31;
      if (r < nvalid) st4(dA1 + (row0 + r) * 128 + c4 * 4, *reinterpret_cast<const float4*>(As + r * LDA + c4 * 4));
    }
  }
  float* P = parts + (size_t)blockIdx.x * kMidPart;
#pragma unroll
  for (int j = 0; j < 16; ++j) P[(k * 16 + j) * 32 + lane] = wacc[j];
  if (lane < 16) P[2048 + k * 16 + lane] = bacc;
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
  long long ntiles = (rows + TR2 - 1) / TR2;
  int grid = (int)(ntiles < kMidGrid ? ntiles : kMidGrid);
  double* parts = stats ? reinterpret_cast<double*>(ws) : nullptr;
  ProfileScope prof("clf_mid_fwd", as_stream(stream));
  clf_mid_fwd_kernel<<<grid, kMidThreads, 0, as_stream(stream)>>>(rows, A1, W2, b2, Z2, parts);
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
  long long ntiles = (rows + TR - 1) / TR;
  int grid = (int)(ntiles < kMidGrid ? ntiles : kMidGrid);
  float* parts = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  ProfileScope prof("clf_mid_bwd", as_stream(stream));
  clf_mid_bwd_kernel<<<grid, kMidThreads, 0, as_stream(stream)>>>(rows, A1, W2, dZ2, dA1, parts);
  EG_LAUNCH_CHECK();
  parts_reduce_kernel<<<(kMidPart + 255) / 256, 256, 0, as_stream(stream)>>>(grid, kMidPart, parts, 2048, dW2, db2);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_clf_out_fwd(int64_t rows, const float* A2, const float* W3, const float* b3, int sigmoid, float* out,
                   void* stream) {
  EG_CHECK_ARG(rows >= 1 && A2 && W3 && b3 && out, "eg_clf_out_fwd: NULL argument");
  long long blocks = (rows * 16 + 255) / 256;
  long long cap = (long long)kNumSMs * 16;
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
