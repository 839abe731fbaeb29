// Node-feature packing: 8 NCHW pyramid maps -> node-major X[batch*N, F] (and the reverse for the
// gradient).  Replaces the per-frame, per-level permute/reshape/torch.cat loop of create_node_pixels
// (src/core/models.py:722-756), which re-copies the growing tensor B*8 times.
// Tiled transpose through shared memory: reads are coalesced along the spatial axis of the NCHW map,
// writes along the feature axis of the node rows.
#include "common.cuh"

struct eg_graph;
namespace eg {
const eg_graph_info& graph_info(const eg_graph* g);
}
using namespace eg;

namespace {

constexpr int TP = 32;  // positions per tile

// to_nodes != 0: map[b][c][p] -> X[(b*N + off + p)][c];  else the reverse (gradient scatter)
template <int F>
__global__ void __launch_bounds__(256)
pack_level_kernel(int N, int off, int P, float* __restrict__ map, float* __restrict__ X, int to_nodes) {
  __shared__ float tile[TP][F + 1];
  const int b = blockIdx.y, p0 = blockIdx.x * TP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* mb = map + (long long)b * F * P;
  float* xb = X + ((long long)b * N + off) * F;
  if (to_nodes) {
    for (int c = warp; c < F; c += 8) {
      int p = p0 + lane;
      tile[lane][c] = p < P ? mb[(long long)c * P + p] : 0.f;
    }
    __syncthreads();
    for (int pp = warp; pp < TP; pp += 8) {
      int p = p0 + pp;
      if (p < P)
        for (int c = lane; c < F; c += 32) xb[(long long)p * F + c] = tile[pp][c];
    }
  } else {
    for (int pp = warp; pp < TP; pp += 8) {
      int p = p0 + pp;
      if (p < P)
        for (int c = lane; c < F; c += 32) tile[pp][c] = xb[(long long)p * F + c];
    }
    __syncthreads();
    for (int c = warp; c < F; c += 8) {
      int p = p0 + lane;
      if (p < P) mb[(long long)c * P + p] = tile[lane][c];
    }
  }
}

// rows [row0, row0+cnt) of every frame <-> compact [batch, cnt, F]
__global__ void rows_copy_kernel(int batch, int N, int row0, int cnt, int F, float* __restrict__ compact,
                                 float* __restrict__ X, int to_nodes) {
  long long total = (long long)batch * cnt * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % F);
    long long r = i / F;
    int j = (int)(r % cnt), b = (int)(r / cnt);
    long long xi = ((long long)b * N + row0 + j) * F + c;
    if (to_nodes) X[xi] = compact[i]; else compact[i] = X[xi];
  }
}

int run(const eg_graph* g, int batch, float* const* maps, float* head, float* tail, float* X, int to_nodes,
        cudaStream_t s) {
  const eg_graph_info& info = graph_info(g);
  constexpr int F = EG_F;
  ProfileScope prof(to_nodes ? "pack_nodes" : "pack_nodes_grad", s);
  for (int l = 0; l < info.num_levels; ++l) {
    if (!maps[l]) continue;  // level without a map (gradient not requested)
    const int P = info.level_size[l] * info.level_size[l];
    dim3 grid((P + TP - 1) / TP, batch);
    pack_level_kernel<F><<<grid, 256, 0, s>>>(info.num_nodes, info.level_offset[l], P, maps[l], X, to_nodes);
    EG_LAUNCH_CHECK();
  }
  if (info.first_pixel_node > 0) {
    EG_CHECK_ARG(head, "pack: graph has connection nodes but head rows are NULL");
    rows_copy_kernel<<<64, 256, 0, s>>>(batch, info.num_nodes, 0, info.first_pixel_node, F, head, X, to_nodes);
    EG_LAUNCH_CHECK();
  }
  if (info.num_coord_nodes > 0 && tail) {  // tail == NULL: the caller fills them (eg_coord_sample_fwd) / has consumed them
    rows_copy_kernel<<<64, 256, 0, s>>>(batch, info.num_nodes, info.num_nodes - info.num_coord_nodes,
                                        info.num_coord_nodes, F, tail, X, to_nodes);
    EG_LAUNCH_CHECK();
  }
  return EG_OK;
}

}  // namespace

extern "C" {

int eg_pack_nodes(const eg_graph* g, int batch, const float* const* maps, const float* head, const float* tail,
                  float* X, void* stream) {
  EG_CHECK_ARG(g && maps && X && batch >= 1, "eg_pack_nodes: bad arguments");
  return run(g, batch, const_cast<float* const*>(reinterpret_cast<const float* const*>(maps)),
             const_cast<float*>(head), const_cast<float*>(tail), X, 1, as_stream(stream));
}

int eg_pack_nodes_grad(const eg_graph* g, int batch, const float* dX, float* const* d_maps, float* d_head,
                       float* d_tail, void* stream) {
  EG_CHECK_ARG(g && d_maps && dX && batch >= 1, "eg_pack_nodes_grad: bad arguments");
  return run(g, batch, d_maps, d_head, d_tail, const_cast<float*>(dX), 0, as_stream(stream));
}

}  // extern "C"
