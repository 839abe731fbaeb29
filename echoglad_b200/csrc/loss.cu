// Multi-level losses, forward + gradient in fused passes, no host round trips.
//  * eg_bce_multilevel        replaces WeightedBCEWithLogitsLoss.compute (src/core/criterion.py:13-27,30-34)
//    which builds its weight tensor on the host through numpy (2 D2H + 1 H2D per step).
//  * eg_expected_landmark_mse replaces ExpectedLandmarkMSE.compute (src/core/criterion.py:93-151), ~15 small
//    kernels per level in the reference.
//  * eg_node_labels           replaces create_node_labels (src/core/datasets.py:523-549).
#include "common.cuh"

using namespace eg;

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];  // fixed order
  return t;
}

// ---- BCE ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void bce_elem(float x, float y, float v, float ones_weight, float& loss, float& grad) {
  const float w = (ones_weight > 1.f && y == 1.f) ? ones_weight : 1.f;
  const float e = expf(-fabsf(x));
  const float l = fmaxf(x, 0.f) - x * y + log1pf(e);
  const float sig = x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
  loss = v * w * l;
  grad = v * w * (sig - y);
}

__global__ void __launch_bounds__(kThreads)
bce_kernel(long long n4, long long n, const float* __restrict__ x, const float* __restrict__ y,
           const float* __restrict__ valid, float ones_weight, float* __restrict__ grad,
           double* __restrict__ parts) {
  __shared__ double sh[8];
  double num = 0.0, den = 0.0;
  float fnum = 0.f, fden = 0.f;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float xv[4], yv[4], vv[4], gv[4];
    if (i * 4 + 3 < n) {
      float4 a = ldg4(x + i * 4), b = ldg4(y + i * 4), c = ldg4(valid + i * 4);
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
      yv[0] = b.x; yv[1] = b.y; yv[2] = b.z; yv[3] = b.w;
      vv[0] = c.x; vv[1] = c.y; vv[2] = c.z; vv[3] = c.w;
    } else {
      for (int k = 0; k < 4; ++k) {
        bool ok = i * 4 + k < n;
        xv[k] = ok ? x[i * 4 + k] : 0.f;
        yv[k] = ok ? y[i * 4 + k] : 0.f;
        vv[k] = ok ? valid[i * 4 + k] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float l;
      bce_elem(xv[k], yv[k], vv[k], ones_weight, l, gv[k]);
      fnum += l;
      fden += vv[k];
    }
    if (grad) {
      if (i * 4 + 3 < n) st4(grad + i * 4, make_float4(gv[0], gv[1], gv[2], gv[3]));
      else for (int k = 0; k < 4; ++k) if (i * 4 + k < n) grad[i * 4 + k] = gv[k];
    }
    if (++cnt == 16) { num += fnum; den += fden; fnum = fden = 0.f; cnt = 0; }
  }
  num += fnum; den += fden;
  num = block_sum(num, sh);
  den = block_sum(den, sh);
  if (threadIdx.x == 0) {
    parts[(size_t)blockIdx.x * 2] = num;
    parts[(size_t)blockIdx.x * 2 + 1] = den;
  }
}

// one block of kThreads: thread t sums partials t, t + kThreads, ... (ascending), block_sum combines in fixed order
// (a single thread walking all <= 592 partials took 42 us)
__global__ void __launch_bounds__(kThreads)
bce_finalize_kernel(int nparts, const double* __restrict__ parts, float loss_weight,
                    float* __restrict__ loss_out, float* __restrict__ scale_out) {
  __shared__ double sh[8];
  double num = 0.0, den = 0.0;
  for (int p = threadIdx.x; p < nparts; p += blockDim.x) { num += parts[2 * p]; den += parts[2 * p + 1]; }
  num = block_sum(num, sh);
  den = block_sum(den, sh);
  if (threadIdx.x == 0) {
    *loss_out = (float)((double)loss_weight * num / den);
    *scale_out = (float)((double)loss_weight / den);
  }
}

__global__ void __launch_bounds__(kThreads)
scale_kernel(long long n, float* __restrict__ g, const float* __restrict__ scale) {
  const float s = __ldg(scale);
  const long long n4 = n >> 2;  // 128-bit passes (the tensor is 16-byte aligned: a torch allocation), scalar tail
  float4* g4 = reinterpret_cast<float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = g4[i];
    v.x *= s, v.y *= s, v.z *= s, v.w *= s;
    g4[i] = v;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) g[i] *= s;
}

// ---- expected-landmark MSE (channels == 4: one float4 per node) ------------------------------------------
struct SegStat {  // per (frame, level)
  float e_h[4], e_w[4], gt_h[4], gt_w[4], v[4], xmax[4], sumexp[4], c_h[4], c_w[4];
};

__device__ __forceinline__ float block_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, sh[w]);
  return t;
}
__device__ __forceinline__ int block_min(int v, int* sh) {
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  int t = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = min(t, sh[w]);
  return t;
}

struct Levels {
  int n;
  int size[EG_MAX_LEVELS];
  int start[EG_MAX_LEVELS];
  int total;
};

// Softmax moments of one (frame, level) segment, all 4 channels.  grid = (batch * levels, kSegSplits): a segment of more
// than kSplitMin nodes is cut into kSegSplits slices, one block each (one block per segment left 84 of 148 SMs idle
// while 64 blocks walked the 50176-node main level three times: 0.26 ms); every block keeps its own running maximum
// ("online softmax") and elmse_combine_kernel merges the slices in fixed order.
constexpr int kSegSplits = 8, kSplitMin = 4096;
struct SegPart {  // one slice
  float xmax[4], ymax[4], sumexp[4], sum_h[4], sum_w[4], vsum[4];
  int min_h[4], min_w[4];
};

__global__ void __launch_bounds__(kThreads)
elmse_stats_kernel(Levels lv, const float* __restrict__ x, const float* __restrict__ y,
                   const float* __restrict__ valid, SegPart* __restrict__ part) {
  __shared__ double shd[8];
  __shared__ float shf[8];
  __shared__ int shi[8];
  const int b = blockIdx.x / lv.n, l = blockIdx.x % lv.n;
  const int g = lv.size[l], n = g * g;
  const int splits = n > kSplitMin ? kSegSplits : 1;
  if ((int)blockIdx.y >= splits) return;
  const int per = (n + splits - 1) / splits, i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  const long long base = ((long long)b * lv.total + lv.start[l]) * 4;
  float xm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float ym[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  float vs[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    float4 a = ldg4(x + base + (long long)i * 4), c = ldg4(y + base + (long long)i * 4);
    float4 d = ldg4(valid + base + (long long)i * 4);
    xm[0] = fmaxf(xm[0], a.x); xm[1] = fmaxf(xm[1], a.y); xm[2] = fmaxf(xm[2], a.z); xm[3] = fmaxf(xm[3], a.w);
    ym[0] = fmaxf(ym[0], c.x); ym[1] = fmaxf(ym[1], c.y); ym[2] = fmaxf(ym[2], c.z); ym[3] = fmaxf(ym[3], c.w);
    vs[0] += d.x; vs[1] += d.y; vs[2] += d.z; vs[3] += d.w;
  }
  double vsum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    xm[k] = block_max(xm[k], shf);
    ym[k] = block_max(ym[k], shf);
    vsum[k] = block_sum((double)vs[k], shd);
  }
  float se[4] = {0.f, 0.f, 0.f, 0.f}, sh_[4] = {0.f, 0.f, 0.f, 0.f}, sw_[4] = {0.f, 0.f, 0.f, 0.f};
  int mh[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX}, mw[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {  // (the slice is L2-resident from the first pass)
    float4 a = ldg4(x + base + (long long)i * 4), c = ldg4(y + base + (long long)i * 4);
    const int h = i / g, w = i - h * g;
    float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float e = expf(av[k] - xm[k]);
      se[k] += e;
      sh_[k] = fmaf(e, (float)h, sh_[k]);
      sw_[k] = fmaf(e, (float)w, sw_[k]);
      if (cv[k] == ym[k]) { mh[k] = min(mh[k], h); mw[k] = min(mw[k], w); }
    }
  }
  SegPart& o = part[(size_t)blockIdx.x * kSegSplits + blockIdx.y];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double s = block_sum((double)se[k], shd);
    double eh = block_sum((double)sh_[k], shd);
    double ew = block_sum((double)sw_[k], shd);
    int gh = block_min(mh[k], shi), gw = block_min(mw[k], shi);
    if (threadIdx.x == 0) {
      o.xmax[k] = xm[k], o.ymax[k] = ym[k];
      o.sumexp[k] = (float)s, o.sum_h[k] = (float)eh, o.sum_w[k] = (float)ew;
      o.vsum[k] = (float)vsum[k];
      o.min_h[k] = gh, o.min_w[k] = gw;
    }
  }
}

// thread = (segment, channel): merges the slices of a segment in ascending order (double), rescaling every slice's sums
// to the segment maximum
__global__ void elmse_combine_kernel(Levels lv, int nseg, const SegPart* __restrict__ part, SegStat* __restrict__ seg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nseg * 4) return;
  const int sgi = idx >> 2, k = idx & 3, l = sgi % lv.n;
  const int n = lv.size[l] * lv.size[l], splits = n > kSplitMin ? kSegSplits : 1;
  const SegPart* p = part + (size_t)sgi * kSegSplits;
  float xm = -INFINITY, ym = -INFINITY;
  for (int j = 0; j < splits; ++j) xm = fmaxf(xm, p[j].xmax[k]), ym = fmaxf(ym, p[j].ymax[k]);
  double s = 0.0, eh = 0.0, ew = 0.0, vs = 0.0;
  int gh = INT_MAX, gw = INT_MAX;
  for (int j = 0; j < splits; ++j) {
    const double r = (double)expf(p[j].xmax[k] - xm);  // 1 for the slice that holds the maximum
    s += r * (double)p[j].sumexp[k];
    eh += r * (double)p[j].sum_h[k];
    ew += r * (double)p[j].sum_w[k];
    vs += (double)p[j].vsum[k];
    if (p[j].ymax[k] == ym) gh = min(gh, p[j].min_h[k]), gw = min(gw, p[j].min_w[k]);
  }
  SegStat& o = seg[sgi];
  o.e_h[k] = (float)(eh / s);
  o.e_w[k] = (float)(ew / s);
  o.gt_h[k] = (float)gh;
  o.gt_w[k] = (float)gw;
  o.v[k] = (float)(vs / (double)n);
  o.xmax[k] = xm;
  o.sumexp[k] = (float)s;
}

// one block: loss and the per-segment gradient coefficients.  A WARP per (level, channel), its lanes over the frames
// (64 threads walking the frames one by one took 95 us); sums in fixed order (lane-strided partials, xor tree).
__global__ void __launch_bounds__(kThreads)
elmse_finalize_kernel(Levels lv, int batch, float loss_weight, SegStat* __restrict__ seg,
                      float* __restrict__ loss_out) {
  __shared__ double sh[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  double total = 0.0;
  for (int idx = warp; idx < lv.n * 4; idx += nwarps) {
    const int l = idx / 4, k = idx % 4;
    const double g = (double)lv.size[l];
    double nv = 0.0;
    for (int b = lane; b < batch; b += 32) nv += (double)seg[b * lv.n + l].v[k];
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    if (nv == 0.0) nv = 1.0;
    double acc = 0.0;
    for (int b = lane; b < batch; b += 32) {
      SegStat& s = seg[b * lv.n + l];
      const double dh = (double)s.e_h[k] / g - (double)s.gt_h[k] / g;
      const double dw = (double)s.e_w[k] / g - (double)s.gt_w[k] / g;
      acc += (dh * dh + dw * dw) * (double)s.v[k];
      s.c_h[k] = (float)((double)loss_weight * 2.0 * dh / g * (double)s.v[k] / nv);
      s.c_w[k] = (float)((double)loss_weight * 2.0 * dw / g * (double)s.v[k] / nv);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) total += acc / nv;
  }
  total = block_sum(total, sh);
  if (threadIdx.x == 0) *loss_out = (float)((double)loss_weight * total);
}

__global__ void __launch_bounds__(kThreads)
elmse_grad_kernel(Levels lv, const float* __restrict__ x, const SegStat* __restrict__ seg,
                  float* __restrict__ grad) {
  const int b = blockIdx.x / lv.n, l = blockIdx.x % lv.n;
  const int g = lv.size[l], n = g * g;
  const long long base = ((long long)b * lv.total + lv.start[l]) * 4;
  const SegStat s = seg[blockIdx.x];
  float inv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) inv[k] = 1.0f / s.sumexp[k];
  for (int i = threadIdx.x + blockIdx.y * blockDim.x; i < n; i += blockDim.x * gridDim.y) {
    float4 a = ldg4(x + base + (long long)i * 4);
    const int h = i / g, w = i - h * g;
    float av[4] = {a.x, a.y, a.z, a.w}, o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float p = expf(av[k] - s.xmax[k]) * inv[k];
      o[k] = p * (((float)h - s.e_h[k]) * s.c_h[k] + ((float)w - s.e_w[k]) * s.c_w[k]);
    }
    st4(grad + base + (long long)i * 4, make_float4(o[0], o[1], o[2], o[3]));
  }
}

__global__ void labels_kernel(int batch, int channels, int frame, Levels lv, const int32_t* __restrict__ coords,
                              float* __restrict__ y, int32_t* __restrict__ oob_count) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * channels * lv.n) return;
  const int l = idx % lv.n, c = (idx / lv.n) % channels, b = idx / (lv.n * channels);
  const int g = lv.size[l];
  int h = coords[(b * channels + c) * 2], w = coords[(b * channels + c) * 2 + 1];
  if (h >= frame || w >= frame || h < -frame || w < -frame) {
    // the reference raises IndexError here (datasets.py:536-537); a kernel cannot raise, and an all-zero heat map
    // would silently train on gt = (0, 0): poison the level's first node (every loss of the frame turns NaN) and
    // count the coordinate for the caller
    y[((long long)b * lv.total + lv.start[l]) * channels + c] = __int_as_float(0x7fc00000);
    if (oob_count && l == 0) atomicAdd(oob_count, 1);
    return;
  }
  // np.digitize(h, linspace(0, frame, g+1)) - 1 == floor(h*g/frame) for 0 <= h < frame; a negative
  // coordinate lands in bin -1, which numpy indexing wraps to the last row/column (datasets.py:532-537)
  int bh = (l == lv.n - 1) ? h : (h < 0 ? -1 : (int)(((long long)h * g) / frame));
  int bw = (l == lv.n - 1) ? w : (w < 0 ? -1 : (int)(((long long)w * g) / frame));
  if (bh < 0) bh += g;
  if (bw < 0) bw += g;
  y[((long long)b * lv.total + lv.start[l] + (long long)bh * g + bw) * channels + c] = 1.0f;
}

int make_levels(int num_levels, const int32_t* level_size, Levels& lv) {
  if (num_levels < 1 || num_levels > EG_MAX_LEVELS || !level_size) return -1;
  lv.n = num_levels;
  int off = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (level_size[l] < 1) return -1;
    lv.size[l] = level_size[l];
    lv.start[l] = off;
    off += level_size[l] * level_size[l];
  }
  lv.total = off;
  return 0;
}

}  // namespace


namespace {
// Expected landmark coordinates of the main level (post-path metric, src/core/evaluators.py:310-348): one block per
// frame, all four channels at once (one float4 per node).  Pass 1: channel maxima of the logits and of the label
// heat map; pass 2: softmax normaliser and first moments in h and w, smallest row / column holding the label
// maximum (torch.max returns the first maximum), sum of `valid`.  Block-wide reductions in a fixed order.
constexpr int kEcThreads = 512;
struct Ec4 { float v[4]; };
__device__ __forceinline__ float ec_red(float x, float* red, bool is_max, bool is_min) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float y = __shfl_xor_sync(0xffffffffu, x, o);
    x = is_max ? fmaxf(x, y) : (is_min ? fminf(x, y) : x + y);
  }
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < kEcThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : (is_min ? fminf(r, red[w]) : r + red[w]);
  return r;
}
__global__ void __launch_bounds__(kEcThreads)
expected_coords_kernel(int n0, int frame, const float* __restrict__ logits, const float* __restrict__ y,
                       const float* __restrict__ valid, float* __restrict__ pred_hw, int* __restrict__ gt_hw,
                       float* __restrict__ valid_mean) {
  __shared__ float red[kEcThreads / 32];
  const int b = blockIdx.x, P = frame * frame;
  const float4* L = reinterpret_cast<const float4*>(logits) + (size_t)b * n0 + (n0 - P);
  const float4* Y = reinterpret_cast<const float4*>(y) + (size_t)b * n0 + (n0 - P);
  const float4* V = valid ? reinterpret_cast<const float4*>(valid) + (size_t)b * n0 + (n0 - P) : nullptr;
  float ml[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, my[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int i = threadIdx.x; i < P; i += kEcThreads) {
    const float4 l = __ldg(L + i), t = __ldg(Y + i);
    ml[0] = fmaxf(ml[0], l.x); ml[1] = fmaxf(ml[1], l.y); ml[2] = fmaxf(ml[2], l.z); ml[3] = fmaxf(ml[3], l.w);
    my[0] = fmaxf(my[0], t.x); my[1] = fmaxf(my[1], t.y); my[2] = fmaxf(my[2], t.z); my[3] = fmaxf(my[3], t.w);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ml[c] = ec_red(ml[c], red, true, false);
    my[c] = ec_red(my[c], red, true, false);
  }
  float se[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0}, sw[4] = {0, 0, 0, 0}, sv[4] = {0, 0, 0, 0};
  float gh[4], gw[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) gh[c] = gw[c] = (float)frame;
  for (int i = threadIdx.x; i < P; i += kEcThreads) {
    const float4 l4 = __ldg(L + i), t4 = __ldg(Y + i);
    const float4 v4 = V ? __ldg(V + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float l[4] = {l4.x, l4.y, l4.z, l4.w}, t[4] = {t4.x, t4.y, t4.z, t4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
    const float h = (float)(i / frame), w = (float)(i % frame);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float e = expf(l[c] - ml[c]);
      se[c] += e;
      sh[c] = fmaf(e, h, sh[c]);
      sw[c] = fmaf(e, w, sw[c]);
      sv[c] += v[c];
      if (t[c] == my[c]) {
        gh[c] = fminf(gh[c], h);
        gw[c] = fminf(gw[c], w);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float e = ec_red(se[c], red, false, false), a = ec_red(sh[c], red, false, false),
                bw = ec_red(sw[c], red, false, false), vv = ec_red(sv[c], red, false, false),
                h0 = ec_red(gh[c], red, false, true), w0 = ec_red(gw[c], red, false, true);
    if (threadIdx.x == 0) {
      pred_hw[(b * 4 + c) * 2] = a / e;
      pred_hw[(b * 4 + c) * 2 + 1] = bw / e;
      gt_hw[(b * 4 + c) * 2] = (int)h0;
      gt_hw[(b * 4 + c) * 2 + 1] = (int)w0;
      valid_mean[b * 4 + c] = vv / (float)P;
    }
  }
}
}  // namespace

extern "C" {

int eg_bce_multilevel(int64_t n, const float* logits, const float* y, const float* valid, float ones_weight,
                      float loss_weight, float* loss_out, float* dlogits, void* ws, size_t ws_bytes,
                      void* stream) {
  EG_CHECK_ARG(n >= 1 && logits && y && valid && loss_out, "eg_bce_multilevel: NULL argument");
  if (!ws || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  const long long n4 = (n + 3) / 4;
  long long blocks = (n4 + kThreads - 1) / kThreads;
  int grid = (int)(blocks < kMaxParts ? blocks : kMaxParts);
  double* parts = reinterpret_cast<double*>(ws);
  float* scale = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kStatsBytes);
  ProfileScope prof("bce", s);
  bce_kernel<<<grid, kThreads, 0, s>>>(n4, n, logits, y, valid, ones_weight, dlogits, parts);
  EG_LAUNCH_CHECK();
  bce_finalize_kernel<<<1, kThreads, 0, s>>>(grid, parts, loss_weight, loss_out, scale);
  EG_LAUNCH_CHECK();
  if (dlogits) {
    scale_kernel<<<grid, kThreads, 0, s>>>(n, dlogits, scale);
    EG_LAUNCH_CHECK();
  }
  return EG_OK;
}

int eg_expected_landmark_mse(int batch, int channels, int num_levels, const int32_t* level_size,
                             const float* logits, const float* y, const float* valid, float loss_weight,
                             float* loss_out, float* dlogits, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(batch >= 1 && logits && y && valid && loss_out, "eg_expected_landmark_mse: bad argument");
  EG_CHECK_ARG(channels == 4, "eg_expected_landmark_mse: num_output_channels must be 4 (engine.py:92), got %d",
               channels);
  Levels lv;
  EG_CHECK_ARG(make_levels(num_levels, level_size, lv) == 0, "eg_expected_landmark_mse: bad level list");
  const size_t nseg = (size_t)batch * num_levels;
  const size_t seg_bytes = (sizeof(SegStat) * nseg + 255) / 256 * 256;
  const size_t need = seg_bytes + sizeof(SegPart) * nseg * kSegSplits;
  if (!ws || ws_bytes < need || ws_bytes < kWorkspaceBytes) {
    set_error("workspace too small: need %zu bytes", need > kWorkspaceBytes ? need : kWorkspaceBytes);
    return EG_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  SegStat* seg = reinterpret_cast<SegStat*>(ws);
  SegPart* part = reinterpret_cast<SegPart*>(reinterpret_cast<char*>(ws) + seg_bytes);
  ProfileScope prof("elmse", s);
  elmse_stats_kernel<<<dim3((unsigned)nseg, kSegSplits), kThreads, 0, s>>>(lv, logits, y, valid, part);
  EG_LAUNCH_CHECK();
  elmse_combine_kernel<<<(unsigned)((nseg * 4 + 127) / 128), 128, 0, s>>>(lv, (int)nseg, part, seg);
  EG_LAUNCH_CHECK();
  elmse_finalize_kernel<<<1, kThreads, 0, s>>>(lv, batch, loss_weight, seg, loss_out);
  EG_LAUNCH_CHECK();
  if (dlogits) {
    dim3 grid(batch * num_levels, 8);
    elmse_grad_kernel<<<grid, kThreads, 0, s>>>(lv, logits, seg, dlogits);
    EG_LAUNCH_CHECK();
  }
  return EG_OK;
}

int eg_node_labels(int batch, int channels, int frame_size, int num_levels, const int32_t* level_size,
                   const int32_t* coords, float* y, int32_t* oob_count, void* stream) {
  EG_CHECK_ARG(batch >= 1 && channels >= 1 && coords && y, "eg_node_labels: bad argument");
  Levels lv;
  EG_CHECK_ARG(make_levels(num_levels, level_size, lv) == 0, "eg_node_labels: bad level list");
  EG_CHECK_ARG(lv.size[num_levels - 1] == frame_size, "eg_node_labels: last level must be the pixel grid");
  cudaStream_t s = as_stream(stream);
  EG_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)batch * lv.total * channels, s));
  const int total = batch * channels * num_levels;
  labels_kernel<<<(total + 127) / 128, 128, 0, s>>>(batch, channels, frame_size, lv, coords, y, oob_count);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_expected_coords(int batch, int channels, int nodes_per_frame, int frame_size, const float* logits,
                       const float* y, const float* valid, float* pred_hw, int32_t* gt_hw, float* valid_mean,
                       void* stream) {
  EG_CHECK_ARG(batch >= 1 && logits && y && pred_hw && gt_hw && valid_mean, "eg_expected_coords: NULL argument");
  EG_CHECK_ARG(channels == 4, "eg_expected_coords: built for 4 landmark channels (got %d)", channels);
  EG_CHECK_ARG(frame_size >= 1 && (long long)frame_size * frame_size <= nodes_per_frame,
               "eg_expected_coords: frame_size^2 exceeds nodes_per_frame");
  ProfileScope prof("expected_coords", as_stream(stream));
  expected_coords_kernel<<<batch, kEcThreads, 0, as_stream(stream)>>>(nodes_per_frame, frame_size, logits, y, valid,
                                                                      pred_hw, gt_hw, valid_mean);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // extern "C"
