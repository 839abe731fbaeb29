// Dense per-node transforms [rows,128] x [128,128] on tensor cores with 3xTF32 error compensation
// (a = a_hi + a_lo, b = b_hi + b_lo; a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate), which
// keeps fp32-class accuracy (plain TF32/bf16 fail the 1e-4 parity bar).
// Replaces nn.Linear inside PyG GCNConv.lin and node_classifiers[k][0] (src/core/models.py:330,364).
//
// v1 of this file uses warp-level mma.sync.m16n8k8.tf32 (legacy tensor path).  The tcgen05/TMEM
// version fused with the aggregation lives in gcn_fused.cu when present.
#include <stdlib.h>

#include "common.cuh"

using namespace eg;

namespace {

constexpr int kThreads = 256;
constexpr int TM = 128;   // rows per tile
constexpr int LDS = 132;  // smem row stride (floats): conflict-free for the (row=g, k=t) fragment pattern
constexpr int LDT = 136;  // stride for the K-major-rows pattern of the weight-gradient kernel

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// async copy of a [TM x 128] row tile (rows beyond `rows` are zero-filled)
__device__ __forceinline__ void load_tile_async(float* dst, int ld, const float* src, long long row0,
                                                long long rows, int nrows) {
  for (int i = threadIdx.x; i < nrows * 32; i += kThreads) {
    int r = i >> 5, c4 = i & 31;
    bool ok = row0 + r < rows;
    cp_async16(dst + r * ld + c4 * 4, ok ? src + (row0 + r) * 128 + c4 * 4 : src, ok);
  }
}

// C = A * op(W) + bias (+ addend); optional per-column sum / sum-of-squares partials (double).
__global__ void __launch_bounds__(kThreads, 1)
linear128_kernel(long long rows, const float* __restrict__ A, const float* __restrict__ W, int trans_w,
                 const float* __restrict__ bias, const float* __restrict__ addend, float* __restrict__ C,
                 double* __restrict__ stat_parts) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                  // [128][LDS]  Ws[n][k]
  float* As0 = Ws + 128 * LDS;       // two A stages
  float* As1 = As0 + TM * LDS;
  float* red = As1 + TM * LDS;       // [8 warps][64] x2 (sum, sq)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const long long ntiles = (rows + TM - 1) / TM;

  // B operand: Ws[n][k] = trans_w ? W[n][k] : W[k][n]
  for (int i = tid; i < 128 * 128; i += kThreads) {
    int r = i >> 7, c = i & 127;
    float v = __ldg(W + i);
    if (trans_w) Ws[r * LDS + c] = v; else Ws[c * LDS + r] = v;
  }
  long long tile = blockIdx.x;
  if (tile < ntiles) load_tile_async(As0, LDS, A, tile * TM, rows, TM);
  cp_async_commit();
  double run_sum = 0.0, run_sq = 0.0;  // thread c < 128 owns column c
  int stage = 0;
  for (; tile < ntiles; tile += gridDim.x, stage ^= 1) {
    float* As = stage ? As1 : As0;
    float* An = stage ? As0 : As1;
    long long next = tile + gridDim.x;
    if (next < ntiles) load_tile_async(An, LDS, A, next * TM, rows, TM);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    float acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll 2
    for (int k0 = 0; k0 < 128; k0 += 8) {
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* ap = As + (wm * 32 + mt * 16 + g) * LDS + k0 + t;
        split_tf32(ap[0], ahi[mt][0], alo[mt][0]);
        split_tf32(ap[8 * LDS], ahi[mt][1], alo[mt][1]);
        split_tf32(ap[4], ahi[mt][2], alo[mt][2]);
        split_tf32(ap[8 * LDS + 4], ahi[mt][3], alo[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float* bp = Ws + (wn * 64 + nt * 8 + g) * LDS + k0 + t;
        uint32_t bhi[2], blo[2];
        split_tf32(bp[0], bhi[0], blo[0]);
        split_tf32(bp[4], bhi[1], blo[1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], alo[mt], bhi);
          mma_tf32(acc[mt][nt], ahi[mt], blo);
          mma_tf32(acc[mt][nt], ahi[mt], bhi);
        }
      }
    }

    // epilogue: bias / addend, store, statistics
    const long long row_base = tile * TM + wm * 32;
    float csum[16], csq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) csum[i] = csq[i] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = wn * 64 + nt * 8 + 2 * t;
      float b0 = bias ? __ldg(bias + col) : 0.f, b1 = bias ? __ldg(bias + col + 1) : 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          long long r = row_base + mt * 16 + h * 8 + g;
          if (r < rows) {
            float v0 = acc[mt][nt][2 * h] + b0, v1 = acc[mt][nt][2 * h + 1] + b1;
            if (addend) {
              float2 ad = __ldg(reinterpret_cast<const float2*>(addend + r * 128 + col));
              v0 += ad.x; v1 += ad.y;
            }
            *reinterpret_cast<float2*>(C + r * 128 + col) = make_float2(v0, v1);
            csum[2 * nt] += v0; csq[2 * nt] += v0 * v0;
            csum[2 * nt + 1] += v1; csq[2 * nt + 1] += v1 * v1;
          }
        }
      }
    }
    if (stat_parts) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          csum[i] += __shfl_xor_sync(0xffffffffu, csum[i], o);
          csq[i] += __shfl_xor_sync(0xffffffffu, csq[i], o);
        }
      }
      if (g == 0) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          red[warp * 64 + nt * 8 + 2 * t] = csum[2 * nt];
          red[warp * 64 + nt * 8 + 2 * t + 1] = csum[2 * nt + 1];
          red[512 + warp * 64 + nt * 8 + 2 * t] = csq[2 * nt];
          red[512 + warp * 64 + nt * 8 + 2 * t + 1] = csq[2 * nt + 1];
        }
      }
    }
    __syncthreads();  // all warps done with As (and red written)
    if (stat_parts && tid < 128) {
      const int cn = tid >> 6, cc = tid & 63;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        run_sum += (double)red[(m * 2 + cn) * 64 + cc];
        run_sq += (double)red[512 + (m * 2 + cn) * 64 + cc];
      }
    }
  }
  cp_async_wait<0>();
  if (stat_parts && tid < 128) {
    stat_parts[(size_t)blockIdx.x * 256 + tid] = run_sum;
    stat_parts[(size_t)blockIdx.x * 256 + 128 + tid] = run_sq;
  }
}

// Weight gradient partials: P[cta][o][i] = sum over the CTA's row tiles of G[n][o] * X[n][i];
// column sums of G into colsum_parts[cta][o].
constexpr int TK = 64;  // rows per stage
__global__ void __launch_bounds__(kThreads, 1)
wgrad128_kernel(long long rows, const float* __restrict__ G, const float* __restrict__ X,
                float* __restrict__ parts, double* __restrict__ colsum_parts) {
  extern __shared__ __align__(16) float smem[];
  float* Gs[2] = {smem, smem + TK * LDT};
  float* Xs[2] = {smem + 2 * TK * LDT, smem + 3 * TK * LDT};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const long long ntiles = (rows + TK - 1) / TK;
  float acc[2][8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  double colsum = 0.0;

  long long tile = blockIdx.x;
  if (tile < ntiles) {
    load_tile_async(Gs[0], LDT, G, tile * TK, rows, TK);
    load_tile_async(Xs[0], LDT, X, tile * TK, rows, TK);
  }
  cp_async_commit();
  int stage = 0;
  for (; tile < ntiles; tile += gridDim.x, stage ^= 1) {
    long long next = tile + gridDim.x;
    if (next < ntiles) {
      load_tile_async(Gs[stage ^ 1], LDT, G, next * TK, rows, TK);
      load_tile_async(Xs[stage ^ 1], LDT, X, next * TK, rows, TK);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* gs = Gs[stage];
    const float* xs = Xs[stage];
#pragma unroll 2
    for (int k0 = 0; k0 < TK; k0 += 8) {
      uint32_t ahi[2][4], alo[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {  // A(row=o, k=n) = G[n][o]
        const float* ap = gs + (k0 + t) * LDT + wm * 32 + mt * 16 + g;
        split_tf32(ap[0], ahi[mt][0], alo[mt][0]);
        split_tf32(ap[8], ahi[mt][1], alo[mt][1]);
        split_tf32(ap[4 * LDT], ahi[mt][2], alo[mt][2]);
        split_tf32(ap[4 * LDT + 8], ahi[mt][3], alo[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {  // B(k=n, col=i) = X[n][i]
        const float* bp = xs + (k0 + t) * LDT + wn * 64 + nt * 8 + g;
        uint32_t bhi[2], blo[2];
        split_tf32(bp[0], bhi[0], blo[0]);
        split_tf32(bp[4 * LDT], bhi[1], blo[1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], alo[mt], bhi);
          mma_tf32(acc[mt][nt], ahi[mt], blo);
          mma_tf32(acc[mt][nt], ahi[mt], bhi);
        }
      }
    }
    if (colsum_parts && tid < 128) {
      float s = 0.f;
#pragma unroll 8
      for (int n = 0; n < TK; ++n) s += gs[n * LDT + tid];
      colsum += (double)s;
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  float* P = parts + (size_t)blockIdx.x * 128 * 128;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int r = wm * 32 + mt * 16 + h * 8 + g, c = wn * 64 + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(P + r * 128 + c) = make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
      }
  if (colsum_parts && tid < 128) colsum_parts[(size_t)blockIdx.x * 128 + tid] = colsum;
}

__global__ void wgrad_reduce_kernel(int nparts, const float* __restrict__ parts,
                                    const double* __restrict__ colsum_parts, float* __restrict__ dW,
                                    float* __restrict__ dbias) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * 128) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)parts[(size_t)p * 128 * 128 + i];
    dW[i] = (float)s;
  }
  if (dbias && colsum_parts && i < 128) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += colsum_parts[(size_t)p * 128 + i];
    dbias[i] = (float)s;
  }
}

}  // namespace

namespace eg {

// EG_LEGACY_MMA=1 routes the dense transforms through the v1 mma.sync kernels (debug / A-B switch only)
bool legacy_mma() {
  static const bool v = [] {
    const char* e = getenv("EG_LEGACY_MMA");
    return e && e[0] == '1';
  }();
  return v;
}
int launch_linear_tc(long long rows, const float* A, const float* W, int trans_w, const float* bias,
                     const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                     cudaStream_t s);

int launch_stats_finalize(int nparts, int cols, int stride, long long rows, const double* parts, float* mean,
                          float* var, cudaStream_t s);  // bn.cu

int launch_linear128(long long rows, const float* A, const float* W, int trans_w, const float* bias,
                     const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                     cudaStream_t s) {
  const bool stats = mean && var;
  if (stats && (!ws || ws_bytes < kWorkspaceBytes)) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  const size_t smem = (size_t)(128 * LDS + 2 * TM * LDS + 1024) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    EG_CUDA(cudaFuncSetAttribute(linear128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  long long ntiles = (rows + TM - 1) / TM;
  int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  if (grid < 1) grid = 1;
  ProfileScope prof("linear128", s);
  double* parts = stats ? reinterpret_cast<double*>(ws) : nullptr;
  linear128_kernel<<<grid, kThreads, smem, s>>>(rows, A, W, trans_w, bias, addend, C, parts);
  EG_LAUNCH_CHECK();
  if (stats) {
    int rc = launch_stats_finalize(grid, 128, 128, rows, parts, mean, var, s);
    if (rc) return rc;
  }
  return EG_OK;
}

int launch_wgrad_tc(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s);  // wgrad_tc.cu

int launch_wgrad128(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s) {
  if (!legacy_mma()) return launch_wgrad_tc(rows, G, X, dW, dbias, ws, ws_bytes, s);
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  const size_t smem = (size_t)(4 * TK * LDT) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    EG_CUDA(cudaFuncSetAttribute(wgrad128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  long long ntiles = (rows + TK - 1) / TK;
  int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  if (grid < 1) grid = 1;
  double* colsum = reinterpret_cast<double*>(ws);
  float* parts = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  ProfileScope prof("wgrad128", s);
  wgrad128_kernel<<<grid, kThreads, smem, s>>>(rows, G, X, parts, dbias ? colsum : nullptr);
  EG_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<(128 * 128 + 255) / 256, 256, 0, s>>>(grid, parts, dbias ? colsum : nullptr, dW, dbias);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // namespace eg

extern "C" {

int eg_linear128(int64_t rows, const float* A, const float* W, int trans_w, const float* bias,
                 const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                 void* stream) {
  EG_CHECK_ARG(rows >= 1 && A && W && C, "eg_linear128: bad arguments");
  if (!legacy_mma())
    return launch_linear_tc(rows, A, W, trans_w, bias, addend, C, mean, var, ws, ws_bytes, as_stream(stream));
  return launch_linear128(rows, A, W, trans_w, bias, addend, C, mean, var, ws, ws_bytes, as_stream(stream));
}

int eg_linear128_wgrad(int64_t rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                       size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && G && X && dW, "eg_linear128_wgrad: bad arguments");
  return launch_wgrad128(rows, G, X, dW, dbias, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"
