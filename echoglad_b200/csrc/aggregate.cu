// out = A_hat * in over the batched block-diagonal graph — atomic-free segmented reduction over the
// per-frame sorted CSR (replaces PyG GCNConv.propagate + torch_scatter.scatter_add; called from
// src/core/models.py:431 through GCNConv).
//
// Layout: in/out are node-major float[batch*N, F].  One warp owns one output row; lane l owns the
// 16-byte column group(s) l, l+32, ...  so every neighbour row is fetched with fully coalesced
// 128-bit loads (512 B per warp request at F=128).  Sum order per row: ascending source index, GCN
// self loop last — the order the reference's CPU scatter_add uses, and deterministic run to run.
#include "common.cuh"

struct eg_graph;
namespace eg {
struct Topo;
const eg_graph_info& graph_info(const eg_graph* g);
const int32_t* graph_rowptr(const eg_graph* g);
const int32_t* graph_col(const eg_graph* g);
const float* graph_w(const eg_graph* g);
}  // namespace eg

using namespace eg;

namespace {

constexpr int kWarpsPerBlock = 8;

template <int VEC>  // VEC float4 per lane: F = 128*VEC
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
agg_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               const float* __restrict__ w, int N, long long total_rows, const float* __restrict__ in,
               float* __restrict__ out) {
  constexpr int F = 128 * VEC;
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= total_rows) return;
  const int frame = (int)(row / N);
  const int v = (int)(row - (long long)frame * N);
  const float* base = in + (long long)frame * N * F;
  const int beg = __ldg(rowptr + v), end = __ldg(rowptr + v + 1);
  float4 acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int chunk = beg; chunk < end; chunk += 32) {
    const int n = min(32, end - chunk);
    int my_col = 0;
    float my_w = 0.f;
    if (lane < n) {
      my_col = __ldg(col + chunk + lane);
      my_w = __ldg(w + chunk + lane);
    }
    int e = 0;
    for (; e + 4 <= n; e += 4) {  // 4 neighbour rows in flight
      int c0 = __shfl_sync(0xffffffffu, my_col, e), c1 = __shfl_sync(0xffffffffu, my_col, e + 1);
      int c2 = __shfl_sync(0xffffffffu, my_col, e + 2), c3 = __shfl_sync(0xffffffffu, my_col, e + 3);
      float w0 = __shfl_sync(0xffffffffu, my_w, e), w1 = __shfl_sync(0xffffffffu, my_w, e + 1);
      float w2 = __shfl_sync(0xffffffffu, my_w, e + 2), w3 = __shfl_sync(0xffffffffu, my_w, e + 3);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const int o = (lane + 32 * i) * 4;
        float4 x0 = ldg4(base + (long long)c0 * F + o), x1 = ldg4(base + (long long)c1 * F + o);
        float4 x2 = ldg4(base + (long long)c2 * F + o), x3 = ldg4(base + (long long)c3 * F + o);
        acc[i].x = fmaf(w0, x0.x, acc[i].x); acc[i].y = fmaf(w0, x0.y, acc[i].y);
        acc[i].z = fmaf(w0, x0.z, acc[i].z); acc[i].w = fmaf(w0, x0.w, acc[i].w);
        acc[i].x = fmaf(w1, x1.x, acc[i].x); acc[i].y = fmaf(w1, x1.y, acc[i].y);
        acc[i].z = fmaf(w1, x1.z, acc[i].z); acc[i].w = fmaf(w1, x1.w, acc[i].w);
        acc[i].x = fmaf(w2, x2.x, acc[i].x); acc[i].y = fmaf(w2, x2.y, acc[i].y);
        acc[i].z = fmaf(w2, x2.z, acc[i].z); acc[i].w = fmaf(w2, x2.w, acc[i].w);
        acc[i].x = fmaf(w3, x3.x, acc[i].x); acc[i].y = fmaf(w3, x3.y, acc[i].y);
        acc[i].z = fmaf(w3, x3.z, acc[i].z); acc[i].w = fmaf(w3, x3.w, acc[i].w);
      }
    }
    for (; e < n; ++e) {
      int c0 = __shfl_sync(0xffffffffu, my_col, e);
      float w0 = __shfl_sync(0xffffffffu, my_w, e);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float4 x0 = ldg4(base + (long long)c0 * F + (lane + 32 * i) * 4);
        acc[i].x = fmaf(w0, x0.x, acc[i].x); acc[i].y = fmaf(w0, x0.y, acc[i].y);
        acc[i].z = fmaf(w0, x0.z, acc[i].z); acc[i].w = fmaf(w0, x0.w, acc[i].w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) st4_stream(out + row * F + (lane + 32 * i) * 4, acc[i]);
}

// F = 64: half a warp per row
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
agg_csr64_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                 const float* __restrict__ w, int N, long long total_rows, const float* __restrict__ in,
                 float* __restrict__ out) {
  constexpr int F = 64;
  const int sub = threadIdx.x & 15;
  long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (row >= total_rows) return;
  const int frame = (int)(row / N);
  const int v = (int)(row - (long long)frame * N);
  const float* base = in + (long long)frame * N * F;
  const int beg = __ldg(rowptr + v), end = __ldg(rowptr + v + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = beg; e < end; ++e) {
    const int c = __ldg(col + e);
    const float ww = __ldg(w + e);
    float4 x = ldg4(base + (long long)c * F + sub * 4);
    acc.x = fmaf(ww, x.x, acc.x); acc.y = fmaf(ww, x.y, acc.y);
    acc.z = fmaf(ww, x.z, acc.z); acc.w = fmaf(ww, x.w, acc.w);
  }
  st4_stream(out + row * F + sub * 4, acc);
}

}  // namespace

namespace eg {
int launch_aggregate(const eg_graph* g, int batch, int feat, const float* in, float* out, cudaStream_t s) {
  const eg_graph_info& info = graph_info(g);
  const long long rows = (long long)batch * info.num_nodes;
  const int threads = kWarpsPerBlock * 32;
  ProfileScope prof("aggregate", s);
  if (feat == 128) {
    unsigned blocks = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    agg_csr_kernel<1><<<blocks, threads, 0, s>>>(graph_rowptr(g), graph_col(g), graph_w(g), info.num_nodes, rows, in, out);
  } else if (feat == 256) {
    unsigned blocks = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    agg_csr_kernel<2><<<blocks, threads, 0, s>>>(graph_rowptr(g), graph_col(g), graph_w(g), info.num_nodes, rows, in, out);
  } else if (feat == 64) {
    unsigned blocks = (unsigned)((rows * 16 + threads - 1) / threads);
    agg_csr64_kernel<<<blocks, threads, 0, s>>>(graph_rowptr(g), graph_col(g), graph_w(g), info.num_nodes, rows, in, out);
  } else {
    set_error("eg_gcn_aggregate: feat must be 64, 128 or 256 (got %d)", feat);
    return EG_ERR_INVALID;
  }
  EG_LAUNCH_CHECK();
  return EG_OK;
}
}  // namespace eg

extern "C" int eg_gcn_aggregate(const eg_graph* g, int batch, int feat, const float* in, float* out,
                                void* stream) {
  EG_CHECK_ARG(g && in && out && batch >= 1, "eg_gcn_aggregate: bad arguments");
  EG_CHECK_ARG(in != out, "eg_gcn_aggregate: in-place aggregation is not supported");
  return launch_aggregate(g, batch, feat, in, out, as_stream(stream));
}
