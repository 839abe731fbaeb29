// Fused pyramid-level embedding: 1x1 convolution (cin -> 128) + bias + ReLU + NCHW -> node-row packing in
// one pass, forward and backward.  Replaces, per lattice level,
//     new_features[l] = F.relu(self.linears[l](features[l]))          src/core/models.py:708-710
//     x = cat([x, reshape(new_features[l][i].permute(1,2,0), (-1,128))])  src/core/models.py:728-741
// i.e. SURVEY.md §8(f) row 1: the step immediately before the message-passing path.  The reference writes the
// [B,128,s,s] map (conv), reads+writes it (ReLU), reads it and writes the node rows (permute/cat): ~5 U of
// traffic for the two big levels; here the narrow input map (cin = 4 or 8 channels) is read once and the node
// rows are written once (1 U), and the backward reads dX once.
//
// Built for the levels that hold 92 % of the nodes of default.yml (main grid: cin 4, 128x128 level: cin 8);
// the tiny deep levels (cin 16..512, <= 4096 nodes per frame) keep the PyTorch conv + eg_pack_nodes path.
//
// Tile = 64 consecutive positions of one frame (P % 64 == 0) x 128 features; thread (warp w, lane l) owns
// positions 8w..8w+7 and features 4l..4l+3, so a warp writes / reads whole 512-byte node rows.
#include "common.cuh"

struct eg_graph;
namespace eg {
const eg_graph_info& graph_info(const eg_graph* g);
}
using namespace eg;

namespace {

constexpr int kThreads = 256;
constexpr int TP = 64;  // positions per tile

template <int CIN>
struct EmbedSmem {
  float Ws[CIN][128];  // Ws[c][o] = W[o][c]
  float in_s[CIN][TP];
};

template <int CIN>
__device__ __forceinline__ void stage_weight(EmbedSmem<CIN>& sm, const float* __restrict__ W) {
  for (int i = threadIdx.x; i < CIN * 128; i += kThreads) sm.Ws[i % CIN][i / CIN] = __ldg(W + i);
}

// in_s[c][pp] = in[b][c][p0 + pp] (coalesced along pp)
template <int CIN>
__device__ __forceinline__ void stage_input(EmbedSmem<CIN>& sm, const float* __restrict__ in_frame, int P, int p0) {
  for (int e = threadIdx.x; e < CIN * TP; e += kThreads) {
    const int c = e / TP, pp = e % TP;
    sm.in_s[c][pp] = __ldg(in_frame + (long long)c * P + p0 + pp);
  }
}

// pre-activation of this thread's 8 positions x 4 features: bias + sum_c W[o][c] in[c][pos], c ascending.
// One fixed fmaf chain, shared by forward and backward so the ReLU mask is bit-identical.
template <int CIN>
__device__ __forceinline__ void preact(const EmbedSmem<CIN>& sm, int w, int lane, const float4& b4, float4 (&acc)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = b4;
#pragma unroll
  for (int c = 0; c < CIN; ++c) {
    const float4 w4 = *reinterpret_cast<const float4*>(&sm.Ws[c][lane * 4]);
    const float4 xa = *reinterpret_cast<const float4*>(&sm.in_s[c][w * 8]);
    const float4 xb = *reinterpret_cast<const float4*>(&sm.in_s[c][w * 8 + 4]);
    const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      acc[q].x = fmaf(w4.x, xs[q], acc[q].x);
      acc[q].y = fmaf(w4.y, xs[q], acc[q].y);
      acc[q].z = fmaf(w4.z, xs[q], acc[q].z);
      acc[q].w = fmaf(w4.w, xs[q], acc[q].w);
    }
  }
}

template <int CIN>
__global__ void __launch_bounds__(kThreads)
level_embed_fwd_kernel(int N, int off, int P, int tiles_per_frame, int num_tiles, const float* __restrict__ in,
                       const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ X) {
  __shared__ EmbedSmem<CIN> sm;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  stage_weight<CIN>(sm, W);
  const float4 b4 = ldg4(bias + lane * 4);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_frame, p0 = (tile - b * tiles_per_frame) * TP;
    __syncthreads();  // previous tile's in_s fully consumed (and Ws staged)
    stage_input<CIN>(sm, in + (long long)b * CIN * P, P, p0);
    __syncthreads();
    float4 acc[8];
    preact<CIN>(sm, w, lane, b4, acc);
    float* xr = X + ((long long)b * N + off + p0 + w * 8) * 128 + lane * 4;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      st4(xr + q * 128, make_float4(fmaxf(acc[q].x, 0.f), fmaxf(acc[q].y, 0.f), fmaxf(acc[q].z, 0.f), fmaxf(acc[q].w, 0.f)));
  }
}

// sum over the 32 lanes of v[i] for i = 0..31; on return lane l holds the total of v[l] in v[0] (31 shuffles)
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

// Backward.  g = dX * (preact > 0);  d_in[b][c][p] = sum_o W[o][c] g[p][o];  per-block partials of
// dW[o][c] = sum g[p][o] in[c][p] and dbias[o] = sum g[p][o]   (parts: [grid][128][CIN + 1], bias last).
template <int CIN>
__global__ void __launch_bounds__(kThreads, 2)
level_embed_bwd_kernel(int N, int off, int P, int tiles_per_frame, int num_tiles, const float* __restrict__ in,
                       const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ dX,
                       float* __restrict__ d_in, float* __restrict__ parts) {
  __shared__ EmbedSmem<CIN> sm;
  __shared__ float red[8][128][CIN + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  stage_weight<CIN>(sm, W);
  const float4 b4 = ldg4(bias + lane * 4);
  float dw[CIN][4], db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) dw[c][k] = 0.f;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_frame, p0 = (tile - b * tiles_per_frame) * TP;
    __syncthreads();
    stage_input<CIN>(sm, in + (long long)b * CIN * P, P, p0);
    const float* gr = dX + ((long long)b * N + off + p0 + w * 8) * 128 + lane * 4;
    float4 g[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) g[q] = ldg4(gr + q * 128);
    __syncthreads();
    {
      float4 acc[8];
      preact<CIN>(sm, w, lane, b4, acc);
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // ReLU backward: grad * (out > 0)
        g[q].x = acc[q].x > 0.f ? g[q].x : 0.f;
        g[q].y = acc[q].y > 0.f ? g[q].y : 0.f;
        g[q].z = acc[q].z > 0.f ? g[q].z : 0.f;
        g[q].w = acc[q].w > 0.f ? g[q].w : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) db[0] += g[q].x, db[1] += g[q].y, db[2] += g[q].z, db[3] += g[q].w;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float4 xa = *reinterpret_cast<const float4*>(&sm.in_s[c][w * 8]);
      const float4 xb = *reinterpret_cast<const float4*>(&sm.in_s[c][w * 8 + 4]);
      const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        dw[c][0] = fmaf(g[q].x, xs[q], dw[c][0]);
        dw[c][1] = fmaf(g[q].y, xs[q], dw[c][1]);
        dw[c][2] = fmaf(g[q].z, xs[q], dw[c][2]);
        dw[c][3] = fmaf(g[q].w, xs[q], dw[c][3]);
      }
    }
    if (d_in) {
#pragma unroll
      for (int c0 = 0; c0 < CIN; c0 += 4) {  // 4 channels x 8 positions = 32 values per transpose-reduce
        float v[32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 w4 = *reinterpret_cast<const float4*>(&sm.Ws[c0 + cc][lane * 4]);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            v[cc * 8 + q] = fmaf(w4.x, g[q].x, fmaf(w4.y, g[q].y, fmaf(w4.z, g[q].z, w4.w * g[q].w)));
        }
        const float tot = transpose_reduce32(v, lane);  // lane = cc * 8 + q
        d_in[((long long)b * CIN + c0 + (lane >> 3)) * P + p0 + w * 8 + (lane & 7)] = tot;
      }
    }
  }
  // block partial: sum over the 8 warps in fixed order
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int c = 0; c < CIN; ++c) red[w][lane * 4 + k][c] = dw[c][k];
    red[w][lane * 4 + k][CIN] = db[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * (CIN + 1); i += kThreads) {
    const int o = i / (CIN + 1), c = i % (CIN + 1);
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += red[ww][o][c];
    parts[(size_t)blockIdx.x * 128 * (CIN + 1) + i] = s;
  }
}

// fixed-order second stage in double: dW[o][c], dbias[o]
__global__ void level_embed_reduce_kernel(int nparts, int cin, const float* __restrict__ parts,
                                          float* __restrict__ dW, float* __restrict__ dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * (cin + 1)) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)parts[(size_t)p * 128 * (cin + 1) + i];
  const int o = i / (cin + 1), c = i % (cin + 1);
  if (c < cin) dW[o * cin + c] = (float)s;
  else dbias[o] = (float)s;
}

constexpr int kMaxGrid = kNumSMs * 8;
// the backward kernel keeps per-block partials that a second stage sums serially per output: two resident blocks per SM
// (its launch bound) are one full wave, and 296 partials instead of 1184 cut the second stage from 93 us to ~25 us
constexpr int kMaxGridBwd = kNumSMs * 2;

struct LevelGeom {
  int N, off, P, tiles_per_frame, num_tiles, grid;
};

int level_geom(const eg_graph* g, int batch, int level, int cin, LevelGeom& lg, const char* who) {
  const eg_graph_info& info = graph_info(g);
  if (level < 0 || level >= info.num_levels) {
    set_error("%s: level %d out of range (graph has %d lattice levels)", who, level, info.num_levels);
    return EG_ERR_INVALID;
  }
  if (cin != 4 && cin != 8) {
    set_error("%s: built for cin in {4, 8} (got %d); wider levels use the conv + eg_pack_nodes path", who, cin);
    return EG_ERR_INVALID;
  }
  lg.N = info.num_nodes;
  lg.off = info.level_offset[level];
  lg.P = info.level_size[level] * info.level_size[level];
  if (lg.P % TP) {
    set_error("%s: level %d has %d positions, not a multiple of %d", who, level, lg.P, TP);
    return EG_ERR_INVALID;
  }
  lg.tiles_per_frame = lg.P / TP;
  const long long nt = (long long)batch * lg.tiles_per_frame;
  if (nt >= (1LL << 31)) {
    set_error("%s: batch * positions too large", who);
    return EG_ERR_INVALID;
  }
  lg.num_tiles = (int)nt;
  lg.grid = lg.num_tiles < kMaxGrid ? lg.num_tiles : kMaxGrid;
  return EG_OK;
}

}  // namespace

extern "C" {

int eg_level_embed_supported(const eg_graph* g, int level, int cin) {
  if (!g) return 0;
  const eg_graph_info& info = graph_info(g);
  if (level < 0 || level >= info.num_levels || (cin != 4 && cin != 8)) return 0;
  return (info.level_size[level] * info.level_size[level]) % TP == 0;
}

int eg_level_embed_fwd(const eg_graph* g, int batch, int level, int cin, const float* in, const float* W,
                       const float* bias, float* X, void* stream) {
  EG_CHECK_ARG(g && in && W && bias && X && batch >= 1, "eg_level_embed_fwd: bad arguments");
  LevelGeom lg;
  int rc = level_geom(g, batch, level, cin, lg, "eg_level_embed_fwd");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("level_embed_fwd", s);
  if (cin == 4)
    level_embed_fwd_kernel<4><<<lg.grid, kThreads, 0, s>>>(lg.N, lg.off, lg.P, lg.tiles_per_frame, lg.num_tiles, in, W, bias, X);
  else
    level_embed_fwd_kernel<8><<<lg.grid, kThreads, 0, s>>>(lg.N, lg.off, lg.P, lg.tiles_per_frame, lg.num_tiles, in, W, bias, X);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_level_embed_bwd(const eg_graph* g, int batch, int level, int cin, const float* in, const float* W,
                       const float* bias, const float* dX, float* d_in, float* dW, float* dbias, void* ws,
                       size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(g && in && W && bias && dX && dW && dbias && batch >= 1, "eg_level_embed_bwd: bad arguments");
  LevelGeom lg;
  int rc = level_geom(g, batch, level, cin, lg, "eg_level_embed_bwd");
  if (rc) return rc;
  static_assert((size_t)kMaxGrid * 128 * 9 * sizeof(float) <= kWorkspaceBytes, "partials fit the workspace");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  float* parts = reinterpret_cast<float*>(ws);
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("level_embed_bwd", s);
  const int grid = lg.num_tiles < kMaxGridBwd ? lg.num_tiles : kMaxGridBwd;
  if (cin == 4)
    level_embed_bwd_kernel<4><<<grid, kThreads, 0, s>>>(lg.N, lg.off, lg.P, lg.tiles_per_frame, lg.num_tiles, in, W, bias, dX, d_in, parts);
  else
    level_embed_bwd_kernel<8><<<grid, kThreads, 0, s>>>(lg.N, lg.off, lg.P, lg.tiles_per_frame, lg.num_tiles, in, W, bias, dX, d_in, parts);
  EG_LAUNCH_CHECK();
  level_embed_reduce_kernel<<<(128 * (cin + 1) + 63) / 64, 64, 0, s>>>(grid, cin, parts, dW, dbias);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // extern "C"
