// Fused BatchNorm1d-apply + Dropout + ReLU|Identity + residual, forward and backward, and the
// deterministic two-stage column statistics they need.
// Replaces gnn_layers[i].module_1..3 + `h + hidden_embeds[i]` (src/core/models.py:332-335,434-435) and
// the BN/ReLU/Dropout triples inside node_classifiers (src/core/models.py:366-373).
// Everything here is HBM-bound elementwise work: 128-bit loads/stores, one pass per tensor.
#include "common.cuh"

using namespace eg;

namespace {

constexpr int kThreads = 256;

// Per-column constants of the 4 columns a thread owns.  Every kernel below strides by a multiple of the row
// width, so a thread keeps ONE column group for its whole life and the constants live in registers.  (They used to
// be staged in shared memory as a 16-byte struct per column: a thread reading 4 consecutive structs strides the
// warp by 64 B = a 4-way bank conflict on every load, and the kernels ran at 97 % LSU instead of at the HBM
// roof -- ncu, r01h.)
struct ColParams4 {
  float mean[4], sc[4] /* gamma*invstd */, beta[4], invstd[4];
};

__device__ __forceinline__ ColParams4 load_params(int c0, const float* mean, const float* var, const float* gamma,
                                                  const float* beta, float eps) {
  ColParams4 p;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float inv = 1.0f / sqrtf(__ldg(var + c0 + k) + eps);  // torch: 1/sqrt(var+eps), rounded once
    p.mean[k] = __ldg(mean + c0 + k);
    p.invstd[k] = inv;
    p.sc[k] = __ldg(gamma + c0 + k) * inv;
    p.beta[k] = __ldg(beta + c0 + k);
  }
  return p;
}

__global__ void __launch_bounds__(kThreads)
bn_act_fwd_kernel(long long rows, int cols, const float* __restrict__ H, const float* __restrict__ mean,
                  const float* __restrict__ var, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, uint32_t thr, float keep_scale, uint64_t seed,
                  int relu, const float* __restrict__ res, float* __restrict__ Y) {
  const int cg = cols >> 2;
  const ColParams4 p = load_params((threadIdx.x % cg) * 4, mean, var, gamma, beta, eps);  // kThreads % cg == 0
  const long long total = rows * cg;
  const long long step = (long long)gridDim.x * blockDim.x;
  // two groups per iteration: up to 4 independent 128-bit loads in flight per thread
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    const bool two = i + step < total;
    const long long i2 = two ? i + step : i;
    const float4 ha = ldg4(H + i * 4), hb = ldg4(H + i2 * 4);
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
    if (res) {
      ra = ldg4(res + i * 4);
      rb = ldg4(res + i2 * 4);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long iu = u ? i2 : i;
      const float4 h = u ? hb : ha, r = u ? rb : ra;
      float v[4] = {h.x, h.y, h.z, h.w};
      const uint32_t keep = thr ? drop_keep4(seed, (uint64_t)iu, thr) : 0xfu;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float y = fmaf(v[k] - p.mean[k], p.sc[k], p.beta[k]);
        y = ((keep >> k) & 1u) ? y * keep_scale : 0.f;
        if (relu) y = fmaxf(y, 0.f);
        v[k] = y;
      }
      st4(Y + iu * 4, make_float4(v[0] + r.x, v[1] + r.y, v[2] + r.z, v[3] + r.w));
    }
  }
}

// Pass 1 of the backward: per-column sums of dA and dA*xhat, where dA = dY * dropmask * relumask.
// Block = (cols/4) column groups x RY row lanes; each thread keeps its column group for all its rows.
// If dH_eval != NULL (eval-mode BN) dH = sc * dA is written in the same pass.
__global__ void __launch_bounds__(kThreads, 3)
bn_act_bwd_reduce_kernel(long long rows, int cols, const float* __restrict__ dY, const float* __restrict__ H,
                         const float* __restrict__ mean, const float* __restrict__ var,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         uint32_t thr, float keep_scale, uint64_t seed, int relu,
                         float* __restrict__ dH_eval, double* __restrict__ parts) {
  __shared__ double red[2][kThreads * 4];
  const int cg = cols >> 2;
  const int ry = kThreads / cg;            // row lanes per block (cols=128 -> 8)
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const ColParams4 p = load_params(tx * 4, mean, var, gamma, beta, eps);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
  int since_flush = 0;
  if (ty < ry) {
    // two rows per iteration: 4 independent 128-bit loads in flight per thread (the pass is pure streaming)
    const long long rstep = (long long)gridDim.x * ry;
    for (long long r = (long long)blockIdx.x * ry + ty; r < rows; r += 2 * rstep) {
      const long long ia = r * cg + tx;
      const bool two = r + rstep < rows;
      const long long ib = two ? ia + rstep * cg : ia;
      const float4 ha = ldg4(H + ia * 4), ga = ldg4(dY + ia * 4);
      const float4 hb = ldg4(H + ib * 4), gb = ldg4(dY + ib * 4);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const long long i = u ? ib : ia;
        const float4 h = u ? hb : ha, g = u ? gb : ga;
        float hv[4] = {h.x, h.y, h.z, h.w}, gv[4] = {g.x, g.y, g.z, g.w}, o[4];
        uint32_t keep = thr ? drop_keep4(seed, (uint64_t)i, thr) : 0xfu;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float xc = hv[k] - p.mean[k];
          float bn = fmaf(xc, p.sc[k], p.beta[k]);
          bool pass = ((keep >> k) & 1u) && (!relu || bn > 0.f);
          float dA = pass ? gv[k] * keep_scale : 0.f;
          s1[k] += dA;
          s2[k] = fmaf(dA, xc * p.invstd[k], s2[k]);
          o[k] = dA * p.sc[k];
        }
        if (dH_eval) st4(dH_eval + i * 4, make_float4(o[0], o[1], o[2], o[3]));
      }
      if (++since_flush == 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { d1[k] += s1[k]; d2[k] += s2[k]; s1[k] = s2[k] = 0.f; }
        since_flush = 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    red[0][threadIdx.x * 4 + k] = d1[k] + (double)s1[k];
    red[1][threadIdx.x * 4 + k] = d2[k] + (double)s2[k];
  }
  __syncthreads();
  if (threadIdx.x < cols) {  // column c: sum over row lanes in fixed order
    const int c = threadIdx.x, gx = c >> 2, k = c & 3;
    double a = 0.0, b = 0.0;
    for (int y = 0; y < ry; ++y) {
      a += red[0][(y * cg + gx) * 4 + k];
      b += red[1][(y * cg + gx) * 4 + k];
    }
    parts[(size_t)blockIdx.x * 2 * cols + c] = a;
    parts[(size_t)blockIdx.x * 2 * cols + cols + c] = b;
  }
}

// Fixed-order sum of per-CTA partials parts[p][2][stride] (p < nparts) for column c, split over kSplit threads:
// thread k sums p = k, k + kSplit, ... in order, the kSplit sub-sums are combined in order through shared memory.
// Deterministic for a given nparts.  Blocks of kFinCols columns x kSplit splits (one block of 128 x 8 walked up to 74
// dependent L2 round trips per thread: 51 us for 592 partials; 32 splits over 4 blocks: 19).
constexpr int kSplit = 32, kFinCols = 32;
__device__ __forceinline__ void split_sum2(int nparts, int stride, const double* __restrict__ parts, int c, int k,
                                           double (*red)[2][kFinCols], double& a, double& b) {
  const int cl = threadIdx.x;  // column inside the block
  double sa = 0.0, sb = 0.0;
  for (int p = k; p < nparts; p += kSplit) {
    sa += parts[(size_t)p * 2 * stride + c];
    sb += parts[(size_t)p * 2 * stride + stride + c];
  }
  red[k][0][cl] = sa;
  red[k][1][cl] = sb;
  __syncthreads();
  a = b = 0.0;
  if (k == 0) {
#pragma unroll
    for (int q = 0; q < kSplit; ++q) {
      a += red[q][0][cl];
      b += red[q][1][cl];
    }
  }
}

// finalize: dbeta = sum dA, dgamma = sum dA*xhat, coef = (sum dA / rows, sum dA*xhat / rows)
// launch: <<<(cols + kFinCols - 1) / kFinCols, dim3(kFinCols, kSplit)>>>
__global__ void bn_bwd_finalize_kernel(int nparts, int cols, long long rows, const double* __restrict__ parts,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef) {
  __shared__ double red[kSplit][2][kFinCols];
  const int cg = blockIdx.x * kFinCols + threadIdx.x;
  const int c = cg < cols ? cg : cols - 1, k = threadIdx.y;
  double a, b;
  split_sum2(nparts, cols, parts, c, k, red, a, b);
  if (k != 0 || cg >= cols) return;
  if (dbeta) dbeta[c] = (float)a;
  if (dgamma) dgamma[c] = (float)b;
  coef[c] = (float)(a / (double)rows);
  coef[cols + c] = (float)(b / (double)rows);
}

// Pass 2 (train-mode BN): dH = sc * (dA - c1 - xhat * c2)
__global__ void __launch_bounds__(kThreads)
bn_act_bwd_apply_kernel(long long rows, int cols, const float* __restrict__ dY, const float* __restrict__ H,
                        const float* __restrict__ mean, const float* __restrict__ var,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        uint32_t thr, float keep_scale, uint64_t seed, int relu,
                        const float* __restrict__ coef, float* __restrict__ dH) {
  const int cg = cols >> 2;
  const int c0 = (threadIdx.x % cg) * 4;  // kThreads % cg == 0: one column group per thread
  const ColParams4 p = load_params(c0, mean, var, gamma, beta, eps);
  float c1[4], c2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c1[k] = __ldg(coef + c0 + k);
    c2[k] = __ldg(coef + cols + c0 + k);
  }
  const long long total = rows * cg;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    const bool two = i + step < total;
    const long long i2 = two ? i + step : i;
    const float4 ha = ldg4(H + i * 4), ga = ldg4(dY + i * 4);
    const float4 hb = ldg4(H + i2 * 4), gb = ldg4(dY + i2 * 4);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long iu = u ? i2 : i;
      const float4 h = u ? hb : ha, g = u ? gb : ga;
      float hv[4] = {h.x, h.y, h.z, h.w}, gv[4] = {g.x, g.y, g.z, g.w}, o[4];
      const uint32_t keep = thr ? drop_keep4(seed, (uint64_t)iu, thr) : 0xfu;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xc = hv[k] - p.mean[k];
        float bn = fmaf(xc, p.sc[k], p.beta[k]);
        bool pass = ((keep >> k) & 1u) && (!relu || bn > 0.f);
        float dA = pass ? gv[k] * keep_scale : 0.f;
        o[k] = p.sc[k] * (dA - c1[k] - xc * p.invstd[k] * c2[k]);
      }
      st4(dH + iu * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

__global__ void __launch_bounds__(kThreads)
col_stats_kernel(long long rows, int cols, const float* __restrict__ Z, double* __restrict__ parts) {
  __shared__ double red[2][kThreads * 4];
  const int cg = cols >> 2;
  const int ry = kThreads / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  double d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
  int since_flush = 0;
  if (ty < ry) {
    for (long long r = (long long)blockIdx.x * ry + ty; r < rows; r += (long long)gridDim.x * ry) {
      float4 z = ldg4(Z + (r * cg + tx) * 4);
      float v[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { s1[k] += v[k]; s2[k] = fmaf(v[k], v[k], s2[k]); }
      if (++since_flush == 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { d1[k] += s1[k]; d2[k] += s2[k]; s1[k] = s2[k] = 0.f; }
        since_flush = 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    red[0][threadIdx.x * 4 + k] = d1[k] + (double)s1[k];
    red[1][threadIdx.x * 4 + k] = d2[k] + (double)s2[k];
  }
  __syncthreads();
  if (threadIdx.x < cols) {
    const int c = threadIdx.x, gx = c >> 2, k = c & 3;
    double a = 0.0, b = 0.0;
    for (int y = 0; y < ry; ++y) {
      a += red[0][(y * cg + gx) * 4 + k];
      b += red[1][(y * cg + gx) * 4 + k];
    }
    parts[(size_t)blockIdx.x * 2 * cols + c] = a;
    parts[(size_t)blockIdx.x * 2 * cols + cols + c] = b;
  }
}

// launch: <<<(cols + kFinCols - 1) / kFinCols, dim3(kFinCols, kSplit)>>>
__global__ void stats_finalize_kernel(int nparts, int cols, int stride, long long rows,
                                      const double* __restrict__ parts, float* __restrict__ mean,
                                      float* __restrict__ var) {
  __shared__ double red[kSplit][2][kFinCols];
  const int cg = blockIdx.x * kFinCols + threadIdx.x;
  const int c = cg < cols ? cg : cols - 1, k = threadIdx.y;
  double s, q;
  split_sum2(nparts, stride, parts, c, k, red, s, q);
  if (k != 0 || cg >= cols) return;
  double m = s / (double)rows;
  double v = q / (double)rows - m * m;
  mean[c] = (float)m;
  var[c] = (float)(v > 0.0 ? v : 0.0);
}

__global__ void sums_finalize_kernel(int nparts, int cols, const double* __restrict__ parts,
                                     float* __restrict__ sums) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += parts[(size_t)p * 2 * cols + c];
  sums[c] = (float)s;
}

__global__ void dropout_mask_kernel(long long groups, uint32_t thr, float keep_scale, uint64_t seed,
                                    float* __restrict__ mask) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < groups;
       i += (long long)gridDim.x * blockDim.x) {
    uint32_t keep = thr ? drop_keep4(seed, (uint64_t)i, thr) : 0xfu;
    st4(mask + i * 4, make_float4((keep & 1u) ? keep_scale : 0.f, (keep & 2u) ? keep_scale : 0.f,
                                  (keep & 4u) ? keep_scale : 0.f, (keep & 8u) ? keep_scale : 0.f));
  }
}

inline bool cols_ok(int cols) { return cols >= 4 && cols <= 128 && (cols % 4) == 0 && (kThreads % (cols / 4)) == 0; }

inline int elem_grid(long long total_groups) {
  long long b = (total_groups + kThreads - 1) / kThreads;
  long long cap = (long long)num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

inline int reduce_grid(long long rows, int cols, int blocks_per_sm = kMaxParts / kNumSMs) {
  int ry = kThreads / (cols / 4);
  long long b = (rows + ry - 1) / ry;
  long long cap = (long long)blocks_per_sm * num_sms();  // one resident wave (<= kMaxParts partial slots)
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

namespace eg {
int launch_stats_finalize(int nparts, int cols, int stride, long long rows, const double* parts, float* mean,
                          float* var, cudaStream_t s) {
  stats_finalize_kernel<<<(cols + kFinCols - 1) / kFinCols, dim3(kFinCols, kSplit), 0, s>>>(nparts, cols, stride, rows, parts, mean, var);
  EG_LAUNCH_CHECK();
  return EG_OK;
}
// dZ = sc * (G - c1 - xhat * c2): the apply pass of a train-mode BatchNorm backward whose masked input gradient G and
// coefficients coef = (c1[cols], c2[cols]) were produced elsewhere (classifier_fused.cu).  dZ may alias G.
int launch_bn_bwd_apply(long long rows, int cols, const float* G, const float* Z, const float* mean, const float* var,
                        const float* gamma, float eps, const float* coef, float* dZ, cudaStream_t s) {
  if (!cols_ok(cols)) {
    set_error("launch_bn_bwd_apply: unsupported cols %d", cols);
    return EG_ERR_INVALID;
  }
  const long long groups = rows * (cols / 4);
  ProfileScope prof("bn_bwd_apply", s);
  // no dropout (thr = 0), no ReLU: the kernel's mask is all-pass and dA = G
  bn_act_bwd_apply_kernel<<<elem_grid(groups), kThreads, 0, s>>>(rows, cols, G, Z, mean, var, gamma, /*beta=*/gamma, eps,
                                                                0u, 1.0f, 0ull, 0, coef, dZ);
  EG_LAUNCH_CHECK();
  return EG_OK;
}
int launch_col_sums(long long rows, int cols, const float* Z, float* sums, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  double* parts = reinterpret_cast<double*>(ws);
  const int grid = reduce_grid(rows, cols);
  ProfileScope prof("col_stats", s);
  col_stats_kernel<<<grid, kThreads, 0, s>>>(rows, cols, Z, parts);
  EG_LAUNCH_CHECK();
  sums_finalize_kernel<<<1, 128, 0, s>>>(grid, cols, parts, sums);
  EG_LAUNCH_CHECK();
  return EG_OK;
}
}  // namespace eg

extern "C" {

int eg_bn_act_fwd(int64_t rows, int cols, const float* H, const float* mean, const float* var,
                  const float* gamma, const float* beta, float eps, float drop_p, uint64_t seed, int relu,
                  const float* res, float* Y, void* stream) {
  EG_CHECK_ARG(rows >= 1 && H && mean && var && gamma && beta && Y, "eg_bn_act_fwd: NULL argument");
  EG_CHECK_ARG(cols_ok(cols), "eg_bn_act_fwd: cols must be a multiple of 4 dividing 1024, <= 128 (got %d)", cols);
  EG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "eg_bn_act_fwd: drop_p must be in [0,1)");
  const uint32_t thr = drop_threshold(drop_p);
  const float ks = thr ? 1.0f / (1.0f - drop_p) : 1.0f;
  const long long groups = rows * (cols / 4);
  ProfileScope prof("bn_act_fwd", as_stream(stream));
  bn_act_fwd_kernel<<<elem_grid(groups), kThreads, 0, as_stream(stream)>>>(rows, cols, H, mean, var, gamma, beta,
                                                                           eps, thr, ks, seed, relu, res, Y);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_bn_act_bwd(int64_t rows, int cols, const float* dY, const float* H, const float* mean,
                  const float* var, const float* gamma, const float* beta, float eps, float drop_p,
                  uint64_t seed, int relu, int batch_stats, float* dH, float* dgamma, float* dbeta, void* ws,
                  size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && dY && H && mean && var && gamma && beta && dH, "eg_bn_act_bwd: NULL argument");
  EG_CHECK_ARG(cols_ok(cols), "eg_bn_act_bwd: unsupported cols %d", cols);
  EG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "eg_bn_act_bwd: drop_p must be in [0,1)");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  const uint32_t thr = drop_threshold(drop_p);
  const float ks = thr ? 1.0f / (1.0f - drop_p) : 1.0f;
  double* parts = reinterpret_cast<double*>(ws);
  float* coef = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);  // 2*cols floats
  const int grid = reduce_grid(rows, cols, 3);  // __launch_bounds__(kThreads, 3) of the reduce kernel
  ProfileScope prof("bn_act_bwd", s);
  bn_act_bwd_reduce_kernel<<<grid, kThreads, 0, s>>>(rows, cols, dY, H, mean, var, gamma, beta, eps, thr, ks, seed,
                                                     relu, batch_stats ? nullptr : dH, parts);
  EG_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<(cols + kFinCols - 1) / kFinCols, dim3(kFinCols, kSplit), 0, s>>>(grid, cols, rows, parts, dgamma, dbeta, coef);
  EG_LAUNCH_CHECK();
  if (batch_stats) {
    const long long groups = rows * (cols / 4);
    bn_act_bwd_apply_kernel<<<elem_grid(groups), kThreads, 0, s>>>(rows, cols, dY, H, mean, var, gamma, beta, eps,
                                                                  thr, ks, seed, relu, coef, dH);
    EG_LAUNCH_CHECK();
  }
  return EG_OK;
}

int eg_dropout_mask(int64_t rows, int cols, float drop_p, uint64_t seed, float* mask, void* stream) {
  EG_CHECK_ARG(rows >= 1 && mask && cols % 4 == 0, "eg_dropout_mask: bad arguments");
  const uint32_t thr = drop_threshold(drop_p);
  const float ks = thr ? 1.0f / (1.0f - drop_p) : 1.0f;
  const long long groups = rows * (cols / 4);
  dropout_mask_kernel<<<elem_grid(groups), kThreads, 0, as_stream(stream)>>>(groups, thr, ks, seed, mask);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_col_stats(int64_t rows, int cols, const float* Z, float* mean, float* var, void* ws, size_t ws_bytes,
                 void* stream) {
  EG_CHECK_ARG(rows >= 1 && Z && mean && var, "eg_col_stats: NULL argument");
  EG_CHECK_ARG(cols_ok(cols), "eg_col_stats: unsupported cols %d", cols);
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  double* parts = reinterpret_cast<double*>(ws);
  const int grid = reduce_grid(rows, cols);
  col_stats_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(rows, cols, Z, parts);
  EG_LAUNCH_CHECK();
  return launch_stats_finalize(grid, cols, cols, rows, parts, mean, var, as_stream(stream));
}

}  // extern "C"
