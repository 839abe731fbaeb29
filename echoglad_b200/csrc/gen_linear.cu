// Generic-width dense transforms over strided views (fp32 SIMT, exact fp32 FMA chains):
//
//     Y = act(A_eff op(W) + bias + addend)           eg_linear_fwd   (also the input gradient: op(W) = W)
//     dW = G_eff^T A,  db = column sums of G_eff      eg_linear_wgrad
//
// where X_eff = X masked by (gate > 0) -- the ReLU mask of a forward output, so the backward of `relu(linear(x))`
// needs no separate masking pass.  A view addresses element (r, c) as
//     base[(r / frame_rows) * frame_stride + (r % frame_rows) * row_stride + c * col_stride]
// which covers row-major node tensors ([rows, F]), one lattice level inside the node tensor (frame_rows = s*s,
// frame_stride = N*F) and NCHW feature maps (row_stride = 1, col_stride = s*s) without a transposition pass.
//
// Users: (1) the 1x1 convolution + ReLU + packing of the SMALL pyramid levels (cin 16..512, 8 % of the nodes;
// src/core/models.py:708-710,728-741 -- the two big levels have their own kernel in embed.cu), forward and backward;
// (2) every dense transform of a model whose widths are not the tensor-core kernels' 128 / 32 (reference constructor
// defaults node_hidden_dim = 64, classifier_hidden_dim = 16, src/core/models.py:290-296).
#include "common.cuh"

using namespace eg;

namespace {

constexpr int kThreads = 256;
constexpr int TM = 64, TN = 64, KC = 16;

struct View {
  float* base;
  unsigned frame_rows;  // rows per frame (>= 1)
  long long frame_stride, row_stride, col_stride;
  __device__ __forceinline__ long long off(unsigned r, int c) const {
    const unsigned f = r / frame_rows;
    return (long long)f * frame_stride + (long long)(r - f * frame_rows) * row_stride + (long long)c * col_stride;
  }
};

struct FwdArgs {
  unsigned rows;
  int K, N;
  View a, gate, addend, y;
  const float* w;
  const float* bias;
  int trans_w, relu, has_gate, has_addend;
};

// tile loader: dst[k][r] (k-major, padded) <- src(r0 + r, k0 + k) for r < TR, k < KC; zero outside [rows) x [K)
template <int TR>
__device__ __forceinline__ void load_tile(float (*dst)[TR + 4], const View& v, const View* gate, unsigned r0, int k0,
                                          unsigned rows, int K) {
  const bool rows_fast = v.row_stride == 1 && v.col_stride != 1;
#pragma unroll
  for (int i = 0; i < TR * KC / kThreads; ++i) {
    const int idx = threadIdx.x + kThreads * i;
    const int r = rows_fast ? idx % TR : idx / KC;
    const int k = rows_fast ? idx / TR : idx % KC;
    float x = 0.f;
    if (r0 + r < rows && k0 + k < K) {
      const long long o = v.off(r0 + r, k0 + k);
      x = v.base[o];
      if (gate && !(gate->base[gate->off(r0 + r, k0 + k)] > 0.f)) x = 0.f;
    }
    dst[k][r] = x;
  }
}

__global__ void __launch_bounds__(kThreads) glin_fwd_kernel(const FwdArgs p) {
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Ws[KC][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const unsigned r0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += KC) {
    load_tile<TM>(As, p.a, p.has_gate ? &p.gate : nullptr, r0, k0, p.rows, p.K);
#pragma unroll
    for (int i = 0; i < TN * KC / kThreads; ++i) {  // op(W)(k, n): trans_w ? w[n][k] : w[k][n]
      const int idx = threadIdx.x + kThreads * i;
      const int n = p.trans_w ? idx / KC : idx % TN;
      const int k = p.trans_w ? idx % KC : idx / TN;
      float x = 0.f;
      if (n0 + n < p.N && k0 + k < p.K)
        x = __ldg(p.w + (p.trans_w ? (long long)(n0 + n) * p.K + k0 + k : (long long)(k0 + k) * p.N + n0 + n));
      Ws[k][n] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned r = r0 + ty * 4 + i;
    if (r >= p.rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= p.N) continue;
      float o = acc[i][j];
      if (p.bias) o += __ldg(p.bias + c);
      if (p.has_addend) o += p.addend.base[p.addend.off(r, c)];
      if (p.relu) o = fmaxf(o, 0.f);
      p.y.base[p.y.off(r, c)] = o;
    }
  }
}

struct WgradArgs {
  unsigned rows, rows_per_split;
  int K, N, splits;
  View g, gate, a;
  int has_gate;
  float* parts;  // [splits][N*K + N]
};

__global__ void __launch_bounds__(kThreads) glin_wgrad_kernel(const WgradArgs p) {
  __shared__ __align__(16) float Gs[KC][TN + 4];  // [row chunk][n]
  __shared__ __align__(16) float As[KC][TM + 4];  // [row chunk][k]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // ty -> n, tx -> k
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TM;
  const unsigned rbeg = blockIdx.z * p.rows_per_split;
  const unsigned rend = min(p.rows, rbeg + p.rows_per_split);
  float acc[4][4] = {};
  float accb[4] = {};
  for (unsigned r0 = rbeg; r0 < rend; r0 += KC) {
    // both tiles are [row][col] with the row chunk as the slow index: dst[r][c]
    const bool g_rows_fast = p.g.row_stride == 1 && p.g.col_stride != 1;
    const bool a_rows_fast = p.a.row_stride == 1 && p.a.col_stride != 1;
#pragma unroll
    for (int i = 0; i < TN * KC / kThreads; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      const int r = g_rows_fast ? idx % KC : idx / TN, c = g_rows_fast ? idx / KC : idx % TN;
      float x = 0.f;
      if (r0 + r < rend && n0 + c < p.N) {
        x = p.g.base[p.g.off(r0 + r, n0 + c)];
        if (p.has_gate && !(p.gate.base[p.gate.off(r0 + r, n0 + c)] > 0.f)) x = 0.f;
      }
      Gs[r][c] = x;
    }
#pragma unroll
    for (int i = 0; i < TM * KC / kThreads; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      const int r = a_rows_fast ? idx % KC : idx / TM, c = a_rows_fast ? idx / KC : idx % TM;
      float x = 0.f;
      if (r0 + r < rend && k0 + c < p.K) x = p.a.base[p.a.off(r0 + r, k0 + c)];
      As[r][c] = x;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < KC; ++r) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[r][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[r][tx * 4]);
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
        accb[i] += gv[i];
      }
    }
    __syncthreads();
  }
  float* part = p.parts + (size_t)blockIdx.z * ((size_t)p.N * p.K + p.N);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < p.K) part[(size_t)n * p.K + k] = acc[i][j];
    }
    if (blockIdx.y == 0 && tx == 0) part[(size_t)p.N * p.K + n] = accb[i];
  }
}

__global__ void glin_wgrad_finalize_kernel(int total, int nk, int splits, const float* parts, float* dw, float* db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double s = 0.0;
  for (int z = 0; z < splits; ++z) s += (double)parts[(size_t)z * total + i];
  if (i < nk)
    dw[i] = (float)s;
  else if (db)
    db[i - nk] = (float)s;
}

int to_view(const char* what, const eg_view* v, long long rows, View& out) {
  EG_CHECK_ARG(v && v->base, "%s: NULL view", what);
  EG_CHECK_ARG(v->frame_rows >= 1, "%s: frame_rows must be >= 1", what);
  out.base = v->base;
  out.frame_rows = (unsigned)(v->frame_rows > rows ? (rows > 0 ? rows : 1) : v->frame_rows);
  out.frame_stride = v->frame_stride;
  out.row_stride = v->row_stride;
  out.col_stride = v->col_stride;
  return EG_OK;
}

}  // namespace

extern "C" {

int eg_linear_fwd(int64_t rows, int k, int n, const eg_view* a, const eg_view* gate, const float* w, int trans_w,
                  const float* bias, const eg_view* addend, int relu, const eg_view* y, void* stream) {
  EG_CHECK_ARG(rows >= 1 && rows < (1LL << 31) && k >= 1 && n >= 1 && k <= 4096 && n <= 4096 && w,
               "eg_linear_fwd: bad arguments (rows %lld, k %d, n %d)", (long long)rows, k, n);
  FwdArgs p{};
  p.rows = (unsigned)rows;
  p.K = k;
  p.N = n;
  if (int rc = to_view("eg_linear_fwd: a", a, rows, p.a)) return rc;
  if (int rc = to_view("eg_linear_fwd: y", y, rows, p.y)) return rc;
  if (gate) {
    if (int rc = to_view("eg_linear_fwd: gate", gate, rows, p.gate)) return rc;
    p.has_gate = 1;
  }
  if (addend) {
    if (int rc = to_view("eg_linear_fwd: addend", addend, rows, p.addend)) return rc;
    p.has_addend = 1;
  }
  p.w = w;
  p.bias = bias;
  p.trans_w = trans_w;
  p.relu = relu;
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("linear_generic", s);
  dim3 grid((unsigned)((rows + TM - 1) / TM), (unsigned)((n + TN - 1) / TN));
  glin_fwd_kernel<<<grid, kThreads, 0, s>>>(p);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_linear_wgrad(int64_t rows, int k, int n, const eg_view* g, const eg_view* gate, const eg_view* a, float* dw,
                    float* db, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && rows < (1LL << 31) && k >= 1 && n >= 1 && k <= 4096 && n <= 4096 && dw,
               "eg_linear_wgrad: bad arguments (rows %lld, k %d, n %d)", (long long)rows, k, n);
  const size_t per = (size_t)n * k + n;
  if (!ws || ws_bytes < kWorkspaceBytes || per * sizeof(float) > kWgradBytes) {
    set_error("eg_linear_wgrad: workspace too small (need %zu bytes; n*k = %zu)", kWorkspaceBytes, (size_t)n * k);
    return EG_ERR_WORKSPACE;
  }
  WgradArgs p{};
  p.rows = (unsigned)rows;
  p.K = k;
  p.N = n;
  if (int rc = to_view("eg_linear_wgrad: g", g, rows, p.g)) return rc;
  if (int rc = to_view("eg_linear_wgrad: a", a, rows, p.a)) return rc;
  if (gate) {
    if (int rc = to_view("eg_linear_wgrad: gate", gate, rows, p.gate)) return rc;
    p.has_gate = 1;
  }
  long long splits = (rows + 2047) / 2048;
  const long long cap = (long long)(kWgradBytes / (per * sizeof(float)));
  const long long want = 2LL * num_sms();
  splits = splits < 1 ? 1 : splits;
  splits = splits > cap ? cap : splits;
  splits = splits > want ? want : splits;
  long long rps = (rows + splits - 1) / splits;
  rps = (rps + KC - 1) / KC * KC;
  splits = (rows + rps - 1) / rps;
  p.rows_per_split = (unsigned)rps;
  p.splits = (int)splits;
  p.parts = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + kStatsBytes);
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("linear_generic_wgrad", s);
  dim3 grid((unsigned)((n + TN - 1) / TN), (unsigned)((k + TM - 1) / TM), (unsigned)splits);
  glin_wgrad_kernel<<<grid, kThreads, 0, s>>>(p);
  EG_LAUNCH_CHECK();
  const int total = (int)per;
  glin_wgrad_finalize_kernel<<<(total + 255) / 256, 256, 0, s>>>(total, n * k, (int)splits, p.parts, dw, db);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // extern "C"
