// Train-mode BatchNorm2d over NCHW maps with FEW channels and a large spatial extent, with the ReLU that precedes it and
// the bias of the convolution before that folded in: y = BN(relu(x + b)).  These are the conv3x3 -> ReLU -> BatchNorm2d
// blocks of the UNet pyramid in front of the graph path (src/core/models.py:841-876): with 4..64 channels cuDNN's
// spatial BN kernels run one CTA per channel (8..64 of 148 SMs), and composed from PyTorch element-wise ops the layer
// costs ~23 tensor passes.  Here: forward = statistics (1 read) + apply (1 read, 1 write), backward = sums (2 reads) +
// apply (2 reads, 1 write); the ReLU and the bias cost no pass of their own, the bias gradient is a by-product of the
// backward apply.  HBM / L2 bound element-wise work: 128-bit accesses, fixed-order two-stage reductions in double
// (deterministic).
// Work unit = (frame n, channel c, chunk of kChunk float4 of the H*W plane); plane size must be a multiple of 4.
#include "common.cuh"

using namespace eg;

namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 4;                     // float4 per thread and unit
constexpr int kChunk = kThreads * kPerThread;     // float4 per unit (16 KB)
constexpr int kMaxSplits = 1024;                  // partial sums per channel

struct Shape {
  int n, c, hw4, chunks;  // frames, channels, float4 per plane, units per plane
};

// block-wide sum of two doubles in a fixed order; result valid in thread 0
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double red[2][kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[0][warp] = a, red[1][warp] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) a += red[0][w], b += red[1][w];
  }
}

// parts[c][split][2]: per channel, block `split` sums the units split, split + splits, ... of that channel.
// DY == nullptr: (sum v, sum v^2) of v = relu?(x); else (sum dy, sum dy v).
__global__ void __launch_bounds__(kThreads)
bn2d_sums_kernel(Shape sh, int splits, const float* __restrict__ X, const float* __restrict__ DY, int relu_in,
                 const float* __restrict__ pre_bias, double* __restrict__ parts) {
  const int c = blockIdx.y, split = blockIdx.x;
  const float pb = pre_bias ? __ldg(pre_bias + c) : 0.f;
  const int units = sh.n * sh.chunks;
  float s0 = 0.f, s1 = 0.f;
  double d0 = 0.0, d1 = 0.0;
  for (int u = split; u < units; u += splits) {
    const int n = u / sh.chunks, ch = u - n * sh.chunks;
    const size_t plane = ((size_t)n * sh.c + c) * sh.hw4;
    const int i0 = ch * kChunk + threadIdx.x;
    float4 x[kPerThread], g[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
      const int i = i0 + k * kThreads;
      x[k] = i < sh.hw4 ? ldg4(X + (plane + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[k] = (DY && i < sh.hw4) ? ldg4(DY + (plane + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
      if (i0 + k * kThreads >= sh.hw4) continue;  // (past the plane: with a folded bias even a zero would count)
      const float xv[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
      const float gv[4] = {g[k].x, g[k].y, g[k].z, g[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = relu_in ? fmaxf(xv[e] + pb, 0.f) : xv[e] + pb;
        if (DY) {  // BatchNorm's input is v = relu(x): its sums take dy as it is; the ReLU mask applies to dx only
          s0 += gv[e];
          s1 = fmaf(gv[e], v, s1);
        } else {
          s0 += v;
          s1 = fmaf(v, v, s1);
        }
      }
    }
    d0 += (double)s0, d1 += (double)s1;  // one unit (16 values per thread) per fp32 run
    s0 = s1 = 0.f;
  }
  block_sum2(d0, d1);
  if (threadIdx.x == 0) {
    parts[((size_t)c * splits + split) * 2] = d0;
    parts[((size_t)c * splits + split) * 2 + 1] = d1;
  }
}

// One block per channel sums the channel's partials (thread t takes splits t, t + 256, ...; fixed-order combine: the
// result does not depend on timing) -- a single thread walking up to 1024 partials took 21 us per launch.
// forward: mean / biased variance per channel.  backward: dbeta = sum dy, dgamma = invstd (sum dy v - mean sum dy), and
// the apply coefficients coef[c] = (a, b, k) of dx = mask (a dy + b v + k).
__global__ void __launch_bounds__(kThreads)
bn2d_finalize_kernel(int splits, double count, const double* __restrict__ parts, float* __restrict__ mean_out,
                     float* __restrict__ var_out, const float* __restrict__ mean, const float* __restrict__ var,
                     const float* __restrict__ gamma, float eps, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ coef) {
  const int c = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int s = threadIdx.x; s < splits; s += kThreads) {
    a += parts[((size_t)c * splits + s) * 2];
    b += parts[((size_t)c * splits + s) * 2 + 1];
  }
  block_sum2(a, b);
  if (threadIdx.x != 0) return;
  if (mean_out) {
    const double m = a / count;
    mean_out[c] = (float)m;
    var_out[c] = (float)fmax(b / count - m * m, 0.0);
    return;
  }
  const float invstd = 1.0f / sqrtf(var[c] + eps);
  const double m = (double)mean[c];
  const double dg = (b - m * a) * (double)invstd;
  dbeta[c] = (float)a;
  dgamma[c] = (float)dg;
  const double sc = (double)gamma[c] * (double)invstd;
  const double bb = -sc * (double)invstd * dg / count;
  coef[3 * c] = (float)sc;
  coef[3 * c + 1] = (float)bb;
  coef[3 * c + 2] = (float)(-sc * a / count - bb * m);
}

// forward: y = relu?(x) * sc + sh.  backward (DY != nullptr): out = mask (a dy + b relu?(x) + k).
__global__ void __launch_bounds__(kThreads)
bn2d_apply_kernel(Shape sh, const float* __restrict__ X, const float* __restrict__ DY, int relu_in,
                  const float* __restrict__ pre_bias, const float* __restrict__ mean, const float* __restrict__ var,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  const float* __restrict__ coef, float* __restrict__ out, double* __restrict__ unit_sums) {
  const long long units = (long long)sh.n * sh.c * sh.chunks;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const long long pl = u / sh.chunks;  // plane = n * C + c
    const int ch = (int)(u - pl * sh.chunks), c = (int)(pl % sh.c);
    const float pb = pre_bias ? __ldg(pre_bias + c) : 0.f;
    float usum = 0.f;  // backward: this thread's share of the unit's sum of dx (gradient of the folded conv bias)
    float a, b, k;
    if (DY) {
      a = __ldg(coef + 3 * c), b = __ldg(coef + 3 * c + 1), k = __ldg(coef + 3 * c + 2);
    } else {
      const float invstd = 1.0f / sqrtf(__ldg(var + c) + eps);
      b = __ldg(gamma + c) * invstd;
      k = fmaf(-__ldg(mean + c), b, __ldg(beta + c));
      a = 0.f;
    }
    const size_t plane = (size_t)pl * sh.hw4;
    const int i0 = ch * kChunk + threadIdx.x;
    float4 x[kPerThread], g[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int i = i0 + q * kThreads;
      x[q] = i < sh.hw4 ? ldg4(X + (plane + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[q] = (DY && i < sh.hw4) ? ldg4(DY + (plane + i) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int i = i0 + q * kThreads;
      if (i >= sh.hw4) continue;
      const float xv[4] = {x[q].x, x[q].y, x[q].z, x[q].w};
      const float gv[4] = {g[q].x, g[q].y, g[q].z, g[q].w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = relu_in ? fmaxf(xv[e] + pb, 0.f) : xv[e] + pb;
        if (DY) {
          const float t = fmaf(a, gv[e], fmaf(b, v, k));
          o[e] = (relu_in && !(xv[e] + pb > 0.f)) ? 0.f : t;
          usum += o[e];
        } else {
          o[e] = fmaf(v, b, k);
        }
      }
      st4(out + (plane + i) * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
    if (unit_sums) {  // (block-uniform) fixed-order sum of the unit
      double d = (double)usum, zero = 0.0;
      __syncthreads();  // block_sum2's scratch may still be read by thread 0 from the previous unit
      block_sum2(d, zero);
      if (threadIdx.x == 0) unit_sums[u] = d;
    }
  }
}

// dpre_bias[c] = sum over the channel's units (n, chunk) of unit_sums, one block per channel, fixed order
__global__ void __launch_bounds__(kThreads)
bn2d_bias_grad_kernel(Shape sh, const double* __restrict__ unit_sums, float* __restrict__ dpre_bias) {
  const int c = blockIdx.x, per = sh.n * sh.chunks;
  double a = 0.0, zero = 0.0;
  for (int i = threadIdx.x; i < per; i += kThreads) {
    const int n = i / sh.chunks, ch = i - n * sh.chunks;
    a += unit_sums[((size_t)n * sh.c + c) * sh.chunks + ch];
  }
  block_sum2(a, zero);
  if (threadIdx.x == 0) dpre_bias[c] = (float)a;
}

int check_shape(int n, int c, long long hw, size_t ws_bytes, const void* ws, Shape& sh, int& splits) {
  if (n < 1 || c < 1 || c > 64 || hw < 4 || hw % 4 != 0 || hw / 4 > (1LL << 28)) {
    set_error("eg_bn2d: need n >= 1, 1 <= channels <= 64, H*W a positive multiple of 4 (got n=%d, c=%d, hw=%lld)", n, c, hw);
    return EG_ERR_INVALID;
  }
  sh.n = n, sh.c = c, sh.hw4 = (int)(hw / 4);
  sh.chunks = (sh.hw4 + kChunk - 1) / kChunk;
  const long long units = (long long)n * sh.chunks;
  long long want = ((long long)num_sms() * 8 + c - 1) / c;
  splits = (int)std::min<long long>(std::min<long long>(want, units), kMaxSplits);
  // [c][splits][2] doubles, 3 x 64 floats of coefficients, one double per unit (bias gradient)
  if (!ws || ws_bytes < (size_t)c * splits * 2 * sizeof(double) + 3 * 64 * sizeof(float) + (size_t)units * c * sizeof(double)) {
    set_error("eg_bn2d: workspace too small");
    return EG_ERR_WORKSPACE;
  }
  return EG_OK;
}

}  // namespace

extern "C" {

int eg_bn2d_fwd(int n, int channels, int64_t hw, const float* x, const float* pre_bias, int relu_in, const float* gamma,
                const float* beta, float eps, float* y, float* mean, float* var, void* ws, size_t ws_bytes,
                void* stream) {
  EG_CHECK_ARG(x && gamma && beta && y && mean && var, "eg_bn2d_fwd: NULL argument");
  Shape sh;
  int splits;
  if (int rc = check_shape(n, channels, hw, ws_bytes, ws, sh, splits)) return rc;
  cudaStream_t s = as_stream(stream);
  double* parts = reinterpret_cast<double*>(ws);
  {
    ProfileScope prof("bn2d_fwd", s);
    bn2d_sums_kernel<<<dim3(splits, channels), kThreads, 0, s>>>(sh, splits, x, nullptr, relu_in, pre_bias, parts);
    EG_LAUNCH_CHECK();
    bn2d_finalize_kernel<<<channels, kThreads, 0, s>>>(splits, (double)n * (double)hw, parts, mean, var, nullptr, nullptr,
                                                       nullptr, eps, nullptr, nullptr, nullptr);
    EG_LAUNCH_CHECK();
    const long long units = (long long)n * channels * sh.chunks;
    const int grid = (int)std::min<long long>(units, (long long)num_sms() * 16);
    bn2d_apply_kernel<<<grid, kThreads, 0, s>>>(sh, x, nullptr, relu_in, pre_bias, mean, var, gamma, beta, eps, nullptr, y,
                                                nullptr);
    EG_LAUNCH_CHECK();
  }
  return EG_OK;
}

int eg_bn2d_bwd(int n, int channels, int64_t hw, const float* x, const float* pre_bias, int relu_in, const float* dy,
                const float* mean, const float* var, const float* gamma, float eps, float* dx, float* dgamma,
                float* dbeta, float* dpre_bias, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(x && dy && mean && var && gamma && dgamma && dbeta, "eg_bn2d_bwd: NULL argument");
  EG_CHECK_ARG(!dpre_bias || dx, "eg_bn2d_bwd: the gradient of the folded bias is a by-product of dx");
  Shape sh;
  int splits;
  if (int rc = check_shape(n, channels, hw, ws_bytes, ws, sh, splits)) return rc;
  cudaStream_t s = as_stream(stream);
  double* parts = reinterpret_cast<double*>(ws);
  float* coef = reinterpret_cast<float*>(parts + (size_t)channels * splits * 2);
  double* unit_sums = reinterpret_cast<double*>(coef + 3 * 64);
  {
    ProfileScope prof("bn2d_bwd", s);
    bn2d_sums_kernel<<<dim3(splits, channels), kThreads, 0, s>>>(sh, splits, x, dy, relu_in, pre_bias, parts);
    EG_LAUNCH_CHECK();
    bn2d_finalize_kernel<<<channels, kThreads, 0, s>>>(splits, (double)n * (double)hw, parts, nullptr, nullptr, mean, var,
                                                       gamma, eps, dgamma, dbeta, coef);
    EG_LAUNCH_CHECK();
    if (dx) {
      const long long units = (long long)n * channels * sh.chunks;
      const int grid = (int)std::min<long long>(units, (long long)num_sms() * 16);
      bn2d_apply_kernel<<<grid, kThreads, 0, s>>>(sh, x, dy, relu_in, pre_bias, mean, var, gamma, nullptr, eps, coef, dx,
                                                  dpre_bias ? unit_sums : nullptr);
      EG_LAUNCH_CHECK();
      if (dpre_bias) {
        bn2d_bias_grad_kernel<<<channels, kThreads, 0, s>>>(sh, unit_sums, dpre_bias);
        EG_LAUNCH_CHECK();
      }
    }
  }
  return EG_OK;
}

}  // extern "C"
