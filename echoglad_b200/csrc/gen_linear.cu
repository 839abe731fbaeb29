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

// row part of a view's offset for tile rows [r0, r0 + n): one integer division per row and block instead of one per
// loaded element (the address arithmetic, not the FMAs, dominated the first version of these kernels)
__device__ __forceinline__ void row_offsets(long long* tab, const View& v, unsigned r0, int n, unsigned rows) {
  for (int r = threadIdx.x; r < n; r += kThreads) tab[r] = r0 + r < rows ? v.off(r0 + r, 0) : -1;
}

__global__ void __launch_bounds__(kThreads) glin_fwd_kernel(const FwdArgs p) {
  // operand chunks during the K loop, then the [TN][TM] output tile of a rows-fast (NCHW) destination
  __shared__ __align__(16) float smem_f[TN * (TM + 1)];
  static_assert(2 * KC * (TM + 4) <= TN * (TM + 1), "operand chunks fit the output staging area");
  float (*As)[TM + 4] = reinterpret_cast<float (*)[TM + 4]>(smem_f);
  float (*Ws)[TN + 4] = reinterpret_cast<float (*)[TN + 4]>(smem_f + KC * (TM + 4));
  __shared__ long long arow[TM], grow[TM];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const unsigned r0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  row_offsets(arow, p.a, r0, TM, p.rows);
  if (p.has_gate) row_offsets(grow, p.gate, r0, TM, p.rows);
  __syncthreads();
  const bool rows_fast = p.a.row_stride == 1 && p.a.col_stride != 1;
  constexpr int NA = TM * KC / kThreads, NW = TN * KC / kThreads;
  float aq[NA], wq[NW];
  auto fetch = [&](int k0) {  // the next chunk's loads are in flight during the current chunk's FMAs
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      const int r = rows_fast ? idx % TM : idx / KC;
      const int k = rows_fast ? idx / TM : idx % KC;
      float x = 0.f;
      const long long ro = arow[r];
      if (ro >= 0 && k0 + k < p.K) {
        x = p.a.base[ro + (long long)(k0 + k) * p.a.col_stride];
        if (p.has_gate && !(p.gate.base[grow[r] + (long long)(k0 + k) * p.gate.col_stride] > 0.f)) x = 0.f;
      }
      aq[i] = x;
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {  // op(W)(k, n): trans_w ? w[n][k] : w[k][n]
      const int idx = threadIdx.x + kThreads * i;
      const int n = p.trans_w ? idx / KC : idx % TN;
      const int k = p.trans_w ? idx % KC : idx / TN;
      float x = 0.f;
      if (n0 + n < p.N && k0 + k < p.K)
        x = __ldg(p.w + (p.trans_w ? (long long)(n0 + n) * p.K + k0 + k : (long long)(k0 + k) * p.N + n0 + n));
      wq[i] = x;
    }
  };
  float acc[4][4] = {};
  fetch(0);
  for (int k0 = 0; k0 < p.K; k0 += KC) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      As[rows_fast ? idx / TM : idx % KC][rows_fast ? idx % TM : idx / KC] = aq[i];
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      Ws[p.trans_w ? idx % KC : idx / TN][p.trans_w ? idx / KC : idx % TN] = wq[i];
    }
    __syncthreads();
    if (k0 + KC < p.K) fetch(k0 + KC);
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
  const bool y_rows_fast = p.y.row_stride == 1 && p.y.col_stride != 1 && !p.has_addend;
  if (y_rows_fast) {
    // NCHW destination: a thread's 4 x 4 block would scatter 4-byte stores over 16 channel planes; the tile goes
    // through shared memory and leaves with the row index fastest (64 consecutive floats per channel plane)
    float (*Ys)[TM + 1] = reinterpret_cast<float (*)[TM + 1]>(smem_f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = n0 + tx * 4 + j;
        float o = acc[i][j];
        if (p.bias && c < p.N) o += __ldg(p.bias + c);
        if (p.relu) o = fmaxf(o, 0.f);
        Ys[tx * 4 + j][ty * 4 + i] = o;
      }
    __syncthreads();
    __shared__ long long yrow[TM];
    row_offsets(yrow, p.y, r0, TM, p.rows);
    __syncthreads();
    for (int idx = threadIdx.x; idx < TN * TM; idx += kThreads) {
      const int r = idx % TM, c = idx / TM;
      if (yrow[r] >= 0 && n0 + c < p.N) p.y.base[yrow[r] + (long long)(n0 + c) * p.y.col_stride] = Ys[c][r];
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned r = r0 + ty * 4 + i;
    if (r >= p.rows) continue;
    const long long yo = p.y.off(r, 0), ao = p.has_addend ? p.addend.off(r, 0) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= p.N) continue;
      float o = acc[i][j];
      if (p.bias) o += __ldg(p.bias + c);
      if (p.has_addend) o += p.addend.base[ao + (long long)c * p.addend.col_stride];
      if (p.relu) o = fmaxf(o, 0.f);
      p.y.base[yo + (long long)c * p.y.col_stride] = o;
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

// Rows are the reduction index: a block owns a [TN x TM] block of dW and one split of the rows, and walks its rows in
// chunks of WR.  The next chunk's global loads are issued into registers before the current chunk's FMAs (the first
// version loaded, synchronised and computed 16 rows at a time: one exposed DRAM round trip per 16 rows, 2.5 ms per
// step for 0.5 GB of traffic on the small pyramid levels).
constexpr int WR = 32;
__global__ void __launch_bounds__(kThreads) glin_wgrad_kernel(const WgradArgs p) {
  __shared__ __align__(16) float Gs[WR][TN + 4];  // [row chunk][n]
  __shared__ __align__(16) float As[WR][TM + 4];  // [row chunk][k]
  __shared__ long long grow[2][WR], gtrow[2][WR], arow[2][WR];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // ty -> n, tx -> k
  const int n0 = blockIdx.x * TN, k0 = blockIdx.y * TM;
  const unsigned rbeg = blockIdx.z * p.rows_per_split;
  const unsigned rend = min(p.rows, rbeg + p.rows_per_split);
  float acc[4][4] = {};
  float accb[4] = {};
  const bool g_rows_fast = p.g.row_stride == 1 && p.g.col_stride != 1;
  const bool a_rows_fast = p.a.row_stride == 1 && p.a.col_stride != 1;
  constexpr int NG = TN * WR / kThreads, NA = TM * WR / kThreads;
  float gq[NG], aq[NA];
  auto row_tab = [&](unsigned r0, int buf) {
    if (threadIdx.x < WR) {
      const unsigned r = r0 + threadIdx.x;
      const bool in = r < rend;
      grow[buf][threadIdx.x] = in ? p.g.off(r, 0) : -1;
      arow[buf][threadIdx.x] = in ? p.a.off(r, 0) : -1;
      gtrow[buf][threadIdx.x] = in && p.has_gate ? p.gate.off(r, 0) : -1;
    }
  };
  auto fetch = [&](int buf) {  // both tiles are [row][col] with the row chunk as the slow index
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      const int r = g_rows_fast ? idx % WR : idx / TN, c = g_rows_fast ? idx / WR : idx % TN;
      float x = 0.f;
      if (grow[buf][r] >= 0 && n0 + c < p.N) {
        x = p.g.base[grow[buf][r] + (long long)(n0 + c) * p.g.col_stride];
        if (p.has_gate && !(p.gate.base[gtrow[buf][r] + (long long)(n0 + c) * p.gate.col_stride] > 0.f)) x = 0.f;
      }
      gq[i] = x;
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      const int r = a_rows_fast ? idx % WR : idx / TM, c = a_rows_fast ? idx / WR : idx % TM;
      float x = 0.f;
      if (arow[buf][r] >= 0 && k0 + c < p.K) x = p.a.base[arow[buf][r] + (long long)(k0 + c) * p.a.col_stride];
      aq[i] = x;
    }
  };
  row_tab(rbeg, 0);
  __syncthreads();
  fetch(0);
  int buf = 0;
  for (unsigned r0 = rbeg; r0 < rend; r0 += WR, buf ^= 1) {
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      Gs[g_rows_fast ? idx % WR : idx / TN][g_rows_fast ? idx / WR : idx % TN] = gq[i];
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = threadIdx.x + kThreads * i;
      As[a_rows_fast ? idx % WR : idx / TM][a_rows_fast ? idx / WR : idx % TM] = aq[i];
    }
    row_tab(r0 + WR, buf ^ 1);
    __syncthreads();
    if (r0 + WR < rend) fetch(buf ^ 1);  // in flight during the FMAs below
#pragma unroll
    for (int r = 0; r < WR; ++r) {
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
  // enough blocks to fill the machine several times over (a block is 256 threads and 19 KB of shared memory), bounded
  // by the partial-sum area of the workspace
  const long long tiles = (long long)((n + TN - 1) / TN) * ((k + TM - 1) / TM);
  const long long cap = (long long)(kWgradBytes / (per * sizeof(float)));
  long long splits = (6LL * num_sms() + tiles - 1) / tiles;
  splits = std::min<long long>(splits, ((long long)rows + WR - 1) / WR);
  splits = std::max<long long>(1, std::min<long long>(splits, cap));
  long long rps = (rows + splits - 1) / splits;
  rps = (rps + WR - 1) / WR * WR;
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
