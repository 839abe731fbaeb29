// Coordinate-graph branch (`use_coordinate_graph`): the per-layer coordinate update of
// src/core/models.py:438-473 and `bilinear_interpolation` (:539-553) as three launches forward (gather the 4 rows
// per frame, MLP, re-sample) and two backward, and the `MAE` coordinate criterion (src/core/criterion.py:52-64).
//
// After GNN layer i the reference (a) builds relative-position features of the 4 landmark coordinates of a frame,
// (b) feeds [coordinate-node embedding (128) | relative positions (8)] through node_coordinate_mlp[i]
// (Linear 136->32, BN, ReLU, Dropout, Linear 32->16, BN, ReLU, Dropout, Linear 16->2), (c) adds the result to the
// coordinates and clamps them to [0, S-1], (d) re-samples the coordinate nodes' embeddings from the main-level rows
// of the layer output with a tent-weight bilinear map and writes them back in place.  It does so with ~30 eager
// ops, a dense [4,S,S] weight map per frame and three device->host syncs per layer (np.where on node_type).
//
// Here: R = 4*batch rows in total, so the MLP with its train-mode BatchNorm statistics (which couple all rows) is
// one CTA (no grid barrier), 16 warps, a warp per row; the row gathers / scatters around it, whose cost is the
// latency of cold rows of a multi-GB tensor, are multi-CTA launches with a warp per row / frame and all of a
// landmark's row reads in flight at once (r02k: one CTA doing everything took 85 us forward / 310 us backward at
// batch 64).  The bilinear map is the 4 taps where the tent `relu(1 - |c - g|)` is non-zero (same weights and
// sub-gradients as the dense formula).  The backward turns the gradient of the updated node tensor IN PLACE into
// the gradient of the layer output: the 4 tap rows of each landmark receive `weight * d(new row)`, the coordinate
// rows receive the MLP's input gradient.  Everything is fixed-order (no atomics): bit-reproducible.
#include "common.cuh"

using namespace eg;

namespace {

constexpr int kThreads = 512, kWarps = kThreads / 32;
constexpr int kIn = EG_F + 8, kH1 = 32, kH2 = 16;
constexpr unsigned kFull = 0xffffffffu;

struct Args {
  float* Y;             // [batch*N, 128] node tensor (forward: updated in place); backward: post-update values
  int batch, N, c0, m0, S;
  const float* coords_in;  // [R,2]
  eg_coord_mlp_params p;
  float *mean1, *var1, *mean2, *var2;
  float *feat_in, *z1, *z2, *pre, *coords_out;
};

struct BwdArgs {
  Args a;
  float* dY;                 // [batch*N,128] in/out
  const float* dcoords_out;  // [R,2] or NULL
  float* scratch;            // [R*64]
  eg_coord_mlp_grads g;
  float* dcoords_in;         // [R,2] or NULL
};

__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// the two grid lines with a non-zero tent weight for coordinate c, their weights and d weight / d c
__device__ __forceinline__ void tent(float c, int S, int (&line)[2], float (&w)[2], float (&d)[2]) {
  const float top = (float)(S - 1);
  const float base = fminf(fmaxf(floorf(c), 0.f), top);
  const bool dup = base + 1.f > top;  // second line clamped onto the first: it does not exist
  const float nxt = dup ? base : base + 1.f;
  const float e0 = c - base, e1 = c - nxt;
  const float t0 = 1.f - fabsf(e0), t1 = 1.f - fabsf(e1);
  w[0] = fmaxf(t0, 0.f);
  w[1] = dup ? 0.f : fmaxf(t1, 0.f);
  d[0] = t0 > 0.f ? -sgnf(e0) : 0.f;
  d[1] = (dup || !(t1 > 0.f)) ? 0.f : -sgnf(e1);
  line[0] = (int)base;
  line[1] = (int)nxt;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ float keep_scale(uint64_t seed, long long e, uint32_t thr, float ks) {
  if (!thr) return 1.f;
  return ((drop_keep4(seed, (uint64_t)(e >> 2), thr) >> (e & 3)) & 1u) ? ks : 0.f;
}

// column statistics of Z[R, C] (C = 32 or 16) by the whole CTA: fixed-order double sums
template <int C>
__device__ void cta_col_sums(int R, const float* Z, const float* Z2, double* red, double* out_s, double* out_q,
                             bool product) {
  // out_s[c] = sum_r Z[r][c]; out_q[c] = sum_r Z[r][c] * (product ? Z2[r][c] : Z[r][c])
  constexpr int P = kThreads / C;
  const int t = threadIdx.x, c = t % C, part = t / C;
  double s = 0.0, q = 0.0;
  for (int r = part; r < R; r += P) {
    const float v = Z[(long long)r * C + c];
    s += (double)v;
    q += (double)v * (double)(product ? Z2[(long long)r * C + c] : v);
  }
  red[part * C + c] = s;
  red[kThreads + part * C + c] = q;
  __syncthreads();
  if (t < C) {
    double ss = 0.0, qq = 0.0;
    for (int k = 0; k < P; ++k) {
      ss += red[k * C + t];
      qq += red[kThreads + k * C + t];
    }
    out_s[t] = ss;
    out_q[t] = qq;
  }
  __syncthreads();
}

struct Smem {
  float w1t[kIn][kH1];   // [k][j]
  float w1[kH1][kIn];    // [j][k]  (backward)
  float w2t[kH1][kH2];   // [k][j]
  float w2[kH2][kH1];    // [j][k]  (backward)
  float w3[2][kH2];
  float in[kWarps][kIn];
  float sc1[kH1], sh1[kH1], mu1[kH1], inv1[kH1];
  float sc2[kH2], sh2[kH2], mu2[kH2], inv2[kH2];
  float c1a[kH1], c1b[kH1], c2a[kH2], c2b[kH2];  // backward BatchNorm coefficients
  double red[2 * kThreads];
  double cs[kH1], cq[kH1];
};

__device__ void load_weights(Smem& sm, const eg_coord_mlp_params& p) {
  for (int i = threadIdx.x; i < kH1 * kIn; i += kThreads) {
    const int j = i / kIn, k = i - j * kIn;
    const float v = __ldg(p.w1 + i);
    sm.w1[j][k] = v;
    sm.w1t[k][j] = v;
  }
  for (int i = threadIdx.x; i < kH2 * kH1; i += kThreads) {
    const int j = i / kH1, k = i - j * kH1;
    const float v = __ldg(p.w2 + i);
    sm.w2[j][k] = v;
    sm.w2t[k][j] = v;
  }
  if (threadIdx.x < 2 * kH2) sm.w3[threadIdx.x / kH2][threadIdx.x % kH2] = __ldg(p.w3 + threadIdx.x);
}

// BatchNorm constants of a layer from (mean, var): scale = gamma * invstd, shift = beta - mean * scale
__device__ __forceinline__ void bn_consts(float mean, float var, float gamma, float beta, float eps, float& sc,
                                          float& sh, float& inv) {
  inv = 1.0f / sqrtf(var + eps);
  sc = gamma * inv;
  sh = fmaf(-mean, sc, beta);
}

// Batch statistics (batch_stats != 0: computed and written to mean/var) or given ones -> per-column constants.
template <int C>
__device__ void layer_stats(Smem& sm, int R, const float* Z, float* mean, float* var, const float* gamma,
                            const float* beta, float eps, int batch_stats, float* sc, float* sh, float* mu,
                            float* inv) {
  if (batch_stats) cta_col_sums<C>(R, Z, nullptr, sm.red, sm.cs, sm.cq, false);
  const int t = threadIdx.x;
  if (t < C) {
    float m, v;
    if (batch_stats) {
      const double dm = sm.cs[t] / R;
      double dv = sm.cq[t] / R - dm * dm;
      dv = dv < 0.0 ? 0.0 : dv;
      m = (float)dm;
      v = (float)dv;
      mean[t] = m;
      var[t] = v;
    } else {
      m = mean[t];
      v = var[t];
    }
    bn_consts(m, v, __ldg(gamma + t), __ldg(beta + t), eps, sc[t], sh[t], inv[t]);
    mu[t] = m;
  }
  __syncthreads();
}

// rows of landmark r's bilinear sample: new[f] = sum_{a,b} wh[a] ww[b] Y[main(line_h[a], line_w[b])][f]
__device__ __forceinline__ float4 sample_row(const float* Ymain, int S, float ch, float cw, int lane) {
  int lh[2], lw[2];
  float wh[2], ww[2], dh[2], dw[2];
  tent(ch, S, lh, wh, dh);
  tent(cw, S, lw, ww, dw);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const float w = wh[a] * ww[b];
      const float4 v = *reinterpret_cast<const float4*>(Ymain + ((long long)lh[a] * S + lw[b]) * EG_F + lane * 4);
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z);
      acc.w = fmaf(w, v.w, acc.w);
    }
  return acc;
}

__global__ void __launch_bounds__(kThreads, 1) coord_sample_fwd_kernel(float* Y, int batch, int N, int c0, int m0,
                                                                         int S, const float* coords) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = 4 * batch;
  for (int r = blockIdx.x * kWarps + warp; r < R; r += gridDim.x * kWarps) {
    const int b = r >> 2, k = r & 3;
    const float* Ymain = Y + ((long long)b * N + m0) * EG_F;
    const float4 v = sample_row(Ymain, S, coords[2 * r], coords[2 * r + 1], lane);
    *reinterpret_cast<float4*>(Y + ((long long)b * N + c0 + k) * EG_F + lane * 4) = v;
  }
}

// Backward of the bilinear sample for the 4 landmarks of frame b (one warp, landmarks and taps in order, so rows
// shared by several taps are updated without a race): dY[tap row] += w * dnew, returns d loss / d coords.
// zero_rows: the coordinate rows of dY are cleared afterwards (their forward values were overwritten).
__device__ __forceinline__ void sample_bwd_frame(float* dY, const float* Y, int b, int N, int c0, int m0, int S,
                                                 const float* coords, int lane, float (&dch)[4], float (&dcw)[4],
                                                 bool zero_rows) {
  const float* Ymain = Y + ((long long)b * N + m0) * EG_F;
  float* dmain = dY + ((long long)b * N + m0) * EG_F;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int r = 4 * b + k;
    float* drow = dY + ((long long)b * N + c0 + k) * EG_F + lane * 4;
    const float4 dn = *reinterpret_cast<const float4*>(drow);
    int lh[2], lw[2];
    float wh[2], ww[2], dh[2], dw[2];
    tent(coords[2 * r], S, lh, wh, dh);
    tent(coords[2 * r + 1], S, lw, ww, dw);
    float gh = 0.f, gw = 0.f;
    long long off[4];
    float4 v[4], gacc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // the landmark's 4 + 4 row reads are independent: all in flight at once
      off[i] = ((long long)lh[i >> 1] * S + lw[i & 1]) * EG_F + lane * 4;
      v[i] = *reinterpret_cast<const float4*>(Ymain + off[i]);
      gacc[i] = *reinterpret_cast<const float4*>(dmain + off[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int a = i >> 1, bb = i & 1;
      const float dot = warp_sum(dn.x * v[i].x + dn.y * v[i].y + dn.z * v[i].z + dn.w * v[i].w);
      gh = fmaf(dot, dh[a] * ww[bb], gh);
      gw = fmaf(dot, wh[a] * dw[bb], gw);
      const float w = wh[a] * ww[bb];
      if (w != 0.f) {  // warp-uniform; taps with a non-zero weight are 4 different rows
        float4 g = gacc[i];
        g.x = fmaf(w, dn.x, g.x);
        g.y = fmaf(w, dn.y, g.y);
        g.z = fmaf(w, dn.z, g.z);
        g.w = fmaf(w, dn.w, g.w);
        *reinterpret_cast<float4*>(dmain + off[i]) = g;
      }
    }
    dch[k] = gh;
    dcw[k] = gw;
    if (zero_rows) *reinterpret_cast<float4*>(drow) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ZERO: the coordinate rows of dY are cleared (initial sample); otherwise they are left for the MLP backward
template <bool ZERO>
__global__ void __launch_bounds__(kThreads, 1) coord_sample_bwd_kernel(float* dY, const float* Y, int batch, int N,
                                                                         int c0, int m0, int S, const float* coords,
                                                                         float* dcoords) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * kWarps + warp; b < batch; b += gridDim.x * kWarps) {
    float dch[4], dcw[4];
    sample_bwd_frame(dY, Y, b, N, c0, m0, S, coords, lane, dch, dcw, ZERO);
    if (dcoords && lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dcoords[2 * (4 * b + k)] = dch[k];
        dcoords[2 * (4 * b + k) + 1] = dcw[k];
      }
    }
  }
}

// activation of layer-1 / layer-2 column `c` of row r recomputed from the saved pre-activation
__device__ __forceinline__ float act(float z, float sc, float sh, float ks) { return fmaxf(fmaf(z, sc, sh), 0.f) * ks; }

__global__ void __launch_bounds__(kThreads, 1) coord_update_fwd_kernel(const Args a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = 4 * a.batch;
  const eg_coord_mlp_params& p = a.p;
  const uint32_t thr = drop_threshold(p.drop_p);
  const float ks = thr ? 1.0f / (1.0f - p.drop_p) : 1.0f;
  load_weights(sm, p);
  __syncthreads();
  // ---- layer 1: z1 = [embedding | relative positions] W1^T + b1 -------------------------------------------------
  float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp < R) nxt = *reinterpret_cast<const float4*>(a.feat_in + (long long)warp * EG_F + lane * 4);
  for (int r = warp; r < R; r += kWarps) {
    const int b = r >> 2;
    *reinterpret_cast<float4*>(&sm.in[warp][lane * 4]) = nxt;
    if (r + kWarps < R)  // next row of this warp: in flight during the 136-step dot products
      nxt = *reinterpret_cast<const float4*>(a.feat_in + (long long)(r + kWarps) * EG_F + lane * 4);
    if (lane < 8) {  // -(c_k - c_j) for j = 0..3, (h, w)   (src/core/models.py:441-444)
      const int j = lane >> 1, d = lane & 1;
      sm.in[warp][EG_F + lane] = a.coords_in[2 * (4 * b + j) + d] - a.coords_in[2 * r + d];
    }
    __syncwarp();
    float acc = __ldg(p.b1 + lane);
#pragma unroll 8
    for (int kk = 0; kk < kIn; ++kk) acc = fmaf(sm.in[warp][kk], sm.w1t[kk][lane], acc);
    a.z1[(long long)r * kH1 + lane] = acc;
    __syncwarp();
  }
  __syncthreads();
  layer_stats<kH1>(sm, R, a.z1, a.mean1, a.var1, p.g1, p.be1, p.eps, p.batch_stats, sm.sc1, sm.sh1, sm.mu1, sm.inv1);
  // ---- layer 2 ------------------------------------------------------------------------------------------------
  for (int r = warp; r < R; r += kWarps) {
    const long long e = (long long)r * kH1 + lane;
    const float a1 = act(a.z1[e], sm.sc1[lane], sm.sh1[lane], keep_scale(p.seed, e, thr, ks));
    float acc = __ldg(p.b2 + (lane & 15));
#pragma unroll
    for (int kk = 0; kk < kH1; ++kk) acc = fmaf(__shfl_sync(kFull, a1, kk), sm.w2t[kk][lane & 15], acc);
    if (lane < kH2) a.z2[(long long)r * kH2 + lane] = acc;
  }
  __syncthreads();
  layer_stats<kH2>(sm, R, a.z2, a.mean2, a.var2, p.g2, p.be2, p.eps, p.batch_stats, sm.sc2, sm.sh2, sm.mu2, sm.inv2);
  // ---- layer 3, clamp (the re-sampling at coords_out is the next launch: coord_sample_fwd_kernel) ------------------
  for (int r = warp; r < R; r += kWarps) {
    float a2 = 0.f;
    if (lane < kH2) {
      const long long e = (long long)r * kH2 + lane;
      a2 = act(a.z2[e], sm.sc2[lane], sm.sh2[lane], keep_scale(p.seed + 1, e, thr, ks));
    }
    const float d0 = warp_sum(a2 * sm.w3[0][lane & 15] * (lane < kH2 ? 1.f : 0.f)) + __ldg(p.b3);
    const float d1 = warp_sum(a2 * sm.w3[1][lane & 15] * (lane < kH2 ? 1.f : 0.f)) + __ldg(p.b3 + 1);
    const float ph = a.coords_in[2 * r] + d0, pw = a.coords_in[2 * r + 1] + d1;
    const float top = (float)(a.S - 1);
    const float ch = fminf(fmaxf(ph, 0.f), top), cw = fminf(fmaxf(pw, 0.f), top);
    if (lane == 0) {
      a.pre[2 * r] = ph;
      a.pre[2 * r + 1] = pw;
      a.coords_out[2 * r] = ch;
      a.coords_out[2 * r + 1] = cw;
    }
  }
}

// feat_in[r] = Y[coordinate row r]: the MLP's inputs, kept for the backward (the rows are overwritten afterwards)
__global__ void __launch_bounds__(kThreads) coord_gather_kernel(const float* Y, int batch, int N, int c0,
                                                                float* feat_in) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = 4 * batch;
  for (int r = blockIdx.x * kWarps + warp; r < R; r += gridDim.x * kWarps) {
    const int b = r >> 2, k = r & 3;
    *reinterpret_cast<float4*>(feat_in + (long long)r * EG_F + lane * 4) =
        *reinterpret_cast<const float4*>(Y + ((long long)b * N + c0 + k) * EG_F + lane * 4);
  }
}

// scratch layout (floats): ddelta [R,2] | g2 -> dz2 [R,16] | g1 -> dz1 [R,32] | drel [R,8] | dsample [R,2]
__global__ void __launch_bounds__(kThreads, 1) coord_update_bwd_kernel(const BwdArgs q) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const Args& a = q.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
  const int R = 4 * a.batch;
  const eg_coord_mlp_params& p = a.p;
  const uint32_t thr = drop_threshold(p.drop_p);
  const float ks = thr ? 1.0f / (1.0f - p.drop_p) : 1.0f;
  float* ddelta = q.scratch;
  float* dz2 = ddelta + (long long)R * 2;
  float* dz1 = dz2 + (long long)R * kH2;
  float* drel = dz1 + (long long)R * kH1;
  const float* dsample = drel + (long long)R * 8;
  load_weights(sm, p);
  if (t < kH1) {
    bn_consts(a.mean1[t], a.var1[t], __ldg(p.g1 + t), __ldg(p.be1 + t), p.eps, sm.sc1[t], sm.sh1[t], sm.inv1[t]);
    sm.mu1[t] = a.mean1[t];
  }
  if (t < kH2) {
    bn_consts(a.mean2[t], a.var2[t], __ldg(p.g2 + t), __ldg(p.be2 + t), p.eps, sm.sc2[t], sm.sh2[t], sm.inv2[t]);
    sm.mu2[t] = a.mean2[t];
  }
  __syncthreads();
  // ---- clamp backward -> d delta (the re-sampling backward ran as the previous launch: its d coords are in dsample) --
  for (int i = t; i < R * 2; i += kThreads) {
    float g = dsample[i];
    if (q.dcoords_out) g += q.dcoords_out[i];
    const float pr = a.pre[i];
    ddelta[i] = (pr >= 0.f && pr <= (float)(a.S - 1)) ? g : 0.f;  // torch.clamp passes the bounds
  }
  __syncthreads();
  // ---- layer 3 backward: g2 = d a2 through Dropout / ReLU ------------------------------------------------------------
  for (int i = t; i < R * kH2; i += kThreads) {
    const int r = i / kH2, j = i - r * kH2;
    const float kp = keep_scale(p.seed + 1, i, thr, ks);
    const float pre = fmaf(a.z2[i], sm.sc2[j], sm.sh2[j]);
    const float da2 = ddelta[2 * r] * sm.w3[0][j] + ddelta[2 * r + 1] * sm.w3[1][j];
    dz2[i] = pre > 0.f ? da2 * kp : 0.f;
  }
  // dW3[d][j] = sum_r ddelta[r][d] a2[r][j], db3[d] = sum_r ddelta[r][d]
  if (t < 2 * kH2) {
    const int d = t / kH2, j = t % kH2;
    double s = 0.0;
    for (int r = 0; r < R; ++r) {
      const long long e = (long long)r * kH2 + j;
      s += (double)ddelta[2 * r + d] * (double)act(a.z2[e], sm.sc2[j], sm.sh2[j], keep_scale(p.seed + 1, e, thr, ks));
    }
    q.g.dw3[t] = (float)s;
  } else if (t < 2 * kH2 + 2) {
    const int d = t - 2 * kH2;
    double s = 0.0;
    for (int r = 0; r < R; ++r) s += (double)ddelta[2 * r + d];
    q.g.db3[d] = (float)s;
  }
  __syncthreads();
  // ---- BatchNorm 2 backward ---------------------------------------------------------------------------------------
  cta_col_sums<kH2>(R, dz2, a.z2, sm.red, sm.cs, sm.cq, true);  // cs = sum g, cq = sum g * z2
  if (t < kH2) {
    const double sg = sm.cs[t], sgz = sm.cq[t];
    const double dgamma = (sgz - (double)sm.mu2[t] * sg) * (double)sm.inv2[t];  // sum g * zhat
    q.g.dbe2[t] = (float)sg;
    q.g.dg2[t] = (float)dgamma;
    sm.c2a[t] = p.batch_stats ? (float)(sg / R) : 0.f;
    sm.c2b[t] = p.batch_stats ? (float)(dgamma / R) : 0.f;
  }
  __syncthreads();
  for (int i = t; i < R * kH2; i += kThreads) {
    const int j = i % kH2;
    const float zh = (a.z2[i] - sm.mu2[j]) * sm.inv2[j];
    dz2[i] = sm.sc2[j] * (dz2[i] - sm.c2a[j] - zh * sm.c2b[j]);
  }
  __syncthreads();
  // ---- layer 2 backward: dW2, db2, g1 = d a1 through Dropout / ReLU ---------------------------------------------
  {
    const int j = t / kH1, k = t % kH1;  // 512 threads = 16 x 32 elements of dW2
    double s = 0.0;
    for (int r = 0; r < R; ++r) {
      const long long e = (long long)r * kH1 + k;
      s += (double)dz2[(long long)r * kH2 + j] *
           (double)act(a.z1[e], sm.sc1[k], sm.sh1[k], keep_scale(p.seed, e, thr, ks));
    }
    q.g.dw2[t] = (float)s;
    if (t < kH2) {
      double sb = 0.0;
      for (int r = 0; r < R; ++r) sb += (double)dz2[(long long)r * kH2 + t];
      q.g.db2[t] = (float)sb;
    }
  }
  for (int i = t; i < R * kH1; i += kThreads) {
    const int r = i / kH1, k = i - r * kH1;
    float da1 = 0.f;
#pragma unroll
    for (int j = 0; j < kH2; ++j) da1 = fmaf(dz2[(long long)r * kH2 + j], sm.w2[j][k], da1);
    const float pre = fmaf(a.z1[i], sm.sc1[k], sm.sh1[k]);
    dz1[i] = pre > 0.f ? da1 * keep_scale(p.seed, i, thr, ks) : 0.f;
  }
  __syncthreads();
  // ---- BatchNorm 1 backward ---------------------------------------------------------------------------------------
  cta_col_sums<kH1>(R, dz1, a.z1, sm.red, sm.cs, sm.cq, true);
  if (t < kH1) {
    const double sg = sm.cs[t], sgz = sm.cq[t];
    const double dgamma = (sgz - (double)sm.mu1[t] * sg) * (double)sm.inv1[t];
    q.g.dbe1[t] = (float)sg;
    q.g.dg1[t] = (float)dgamma;
    sm.c1a[t] = p.batch_stats ? (float)(sg / R) : 0.f;
    sm.c1b[t] = p.batch_stats ? (float)(dgamma / R) : 0.f;
  }
  __syncthreads();
  for (int i = t; i < R * kH1; i += kThreads) {
    const int j = i % kH1;
    const float zh = (a.z1[i] - sm.mu1[j]) * sm.inv1[j];
    dz1[i] = sm.sc1[j] * (dz1[i] - sm.c1a[j] - zh * sm.c1b[j]);
  }
  __syncthreads();
  // ---- layer 1 backward: dW1 [32,136], db1, input gradient -> coordinate rows of dY, relative positions ----------
  for (int i = t; i < kH1 * kIn; i += kThreads) {
    const int j = i / kIn, k = i - j * kIn;
    double s = 0.0;
    if (k < EG_F) {
      for (int r = 0; r < R; ++r) s += (double)dz1[(long long)r * kH1 + j] * (double)a.feat_in[(long long)r * EG_F + k];
    } else {
      const int jj = (k - EG_F) >> 1, d = (k - EG_F) & 1;
      for (int r = 0; r < R; ++r) {
        const int b = r >> 2;
        const float rel = a.coords_in[2 * (4 * b + jj) + d] - a.coords_in[2 * r + d];
        s += (double)dz1[(long long)r * kH1 + j] * (double)rel;
      }
    }
    q.g.dw1[i] = (float)s;
  }
  if (t < kH1) {
    double sb = 0.0;
    for (int r = 0; r < R; ++r) sb += (double)dz1[(long long)r * kH1 + t];
    q.g.db1[t] = (float)sb;
  }
  for (int r = warp; r < R; r += kWarps) {
    const int b = r >> 2, k = r & 3;
    const float g = dz1[(long long)r * kH1 + lane];
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j = 0; j < kH1; ++j) {
      const float gj = __shfl_sync(kFull, g, j);
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const int kk = lane + 32 * m;
        if (kk < kIn) acc[m] = fmaf(gj, sm.w1[j][kk], acc[m]);
      }
    }
    float* drow = q.dY + ((long long)b * a.N + a.c0 + k) * EG_F;
#pragma unroll
    for (int m = 0; m < 4; ++m) drow[lane + 32 * m] = acc[m];  // the rows' forward values were overwritten: assign
    if (lane < 8) drel[r * 8 + lane] = acc[4];
  }
  __syncthreads();
  // ---- d coords_in: identity path of the clamp + relative-position features (rel[(b,k)][j] = c_j - c_k) -----------
  if (q.dcoords_in) {
    for (int i = t; i < R * 2; i += kThreads) {
      const int r = i >> 1, d = i & 1, b = r >> 2, k = r & 3;
      float g = ddelta[i];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) g += drel[(4 * b + kk) * 8 + 2 * k + d];  // as c_j of landmark kk
#pragma unroll
      for (int j = 0; j < 4; ++j) g -= drel[r * 8 + 2 * j + d];                  // as c_k of its own row
      q.dcoords_in[i] = g;
    }
  }
}

// loss = weight * mean |pred - y| over n values, grad = weight * sign(pred - y) / n   (nn.L1Loss)
__global__ void __launch_bounds__(256, 1) mae_kernel(int n, const float* pred, const float* y, float weight,
                                                     float* loss, float* grad) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = pred[i] - y[i];
    s += (double)fabsf(d);
    if (grad) grad[i] = weight * sgnf(d) / (float)n;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 256; ++i) tot += red[i];
    *loss = (float)((double)weight * tot / n);
  }
}

int check_geom(const char* what, const float* Y, int batch, int N, int c0, int m0, int S) {
  EG_CHECK_ARG(Y && batch >= 1 && S >= 1, "%s: bad arguments", what);
  EG_CHECK_ARG(c0 >= 0 && c0 + 4 <= N && m0 >= 0 && (long long)m0 + (long long)S * S <= N &&
                   (c0 >= m0 + S * S || c0 + 4 <= m0),
               "%s: coordinate rows [%d,%d) / main rows [%d,%d) do not fit a frame of %d nodes", what, c0, c0 + 4, m0,
               m0 + S * S, N);
  return EG_OK;
}

int check_params(const char* what, const eg_coord_mlp_params* p) {
  EG_CHECK_ARG(p && p->w1 && p->b1 && p->g1 && p->be1 && p->w2 && p->b2 && p->g2 && p->be2 && p->w3 && p->b3,
               "%s: NULL parameter", what);
  EG_CHECK_ARG(p->drop_p >= 0.f && p->drop_p < 1.f, "%s: drop_p %f outside [0,1)", what, p->drop_p);
  return EG_OK;
}

template <typename K>
int opt_in_smem(K kernel, std::atomic<unsigned long long>& mask) {
  if (first_use_on_current_device(mask))
    EG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  return EG_OK;
}

}  // namespace

extern "C" {

int eg_coord_sample_fwd(float* Y, int batch, int nodes_per_frame, int coord_row0, int main_row0, int frame_size,
                        const float* coords, void* stream) {
  if (int rc = check_geom("eg_coord_sample_fwd", Y, batch, nodes_per_frame, coord_row0, main_row0, frame_size)) return rc;
  EG_CHECK_ARG(coords, "eg_coord_sample_fwd: coords is NULL");
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("coord_sample_fwd", s);
  const int grid = (4 * batch + kWarps - 1) / kWarps;
  coord_sample_fwd_kernel<<<grid < 148 ? grid : 148, kThreads, 0, s>>>(Y, batch, nodes_per_frame, coord_row0,
                                                                      main_row0, frame_size, coords);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_coord_sample_bwd(float* dY, const float* Y, int batch, int nodes_per_frame, int coord_row0, int main_row0,
                        int frame_size, const float* coords, float* dcoords, void* stream) {
  if (int rc = check_geom("eg_coord_sample_bwd", Y, batch, nodes_per_frame, coord_row0, main_row0, frame_size)) return rc;
  EG_CHECK_ARG(dY && coords, "eg_coord_sample_bwd: NULL argument");
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("coord_sample_bwd", s);
  const int grid = (batch + kWarps - 1) / kWarps;
  coord_sample_bwd_kernel<true><<<grid < 148 ? grid : 148, kThreads, 0, s>>>(dY, Y, batch, nodes_per_frame, coord_row0,
                                                                            main_row0, frame_size, coords, dcoords);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_coord_update_fwd(float* Y, int batch, int nodes_per_frame, int coord_row0, int main_row0, int frame_size,
                        const float* coords_in, const eg_coord_mlp_params* p, float* mean1, float* var1, float* mean2,
                        float* var2, float* feat_in, float* z1, float* z2, float* pre, float* coords_out,
                        void* stream) {
  if (int rc = check_geom("eg_coord_update_fwd", Y, batch, nodes_per_frame, coord_row0, main_row0, frame_size)) return rc;
  if (int rc = check_params("eg_coord_update_fwd", p)) return rc;
  EG_CHECK_ARG(coords_in && mean1 && var1 && mean2 && var2 && feat_in && z1 && z2 && pre && coords_out,
               "eg_coord_update_fwd: NULL argument");
  static std::atomic<unsigned long long> mask{0};
  if (int rc = opt_in_smem(coord_update_fwd_kernel, mask)) return rc;
  Args a{Y, batch, nodes_per_frame, coord_row0, main_row0, frame_size, coords_in, *p, mean1, var1, mean2, var2,
         feat_in, z1, z2, pre, coords_out};
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("coord_update_fwd", s);
  const int grid = (4 * batch + kWarps - 1) / kWarps;
  coord_gather_kernel<<<grid < 148 ? grid : 148, kThreads, 0, s>>>(Y, batch, nodes_per_frame, coord_row0, feat_in);
  EG_LAUNCH_CHECK();
  coord_update_fwd_kernel<<<1, kThreads, sizeof(Smem), s>>>(a);
  EG_LAUNCH_CHECK();
  coord_sample_fwd_kernel<<<grid < 148 ? grid : 148, kThreads, 0, s>>>(Y, batch, nodes_per_frame, coord_row0,
                                                                      main_row0, frame_size, coords_out);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_coord_update_bwd(float* dY, const float* dcoords_out, const float* Y, int batch, int nodes_per_frame,
                        int coord_row0, int main_row0, int frame_size, const float* coords_in,
                        const eg_coord_mlp_params* p, const float* mean1, const float* var1, const float* mean2,
                        const float* var2, const float* feat_in, const float* z1, const float* z2, const float* pre,
                        const float* coords_out, float* scratch, const eg_coord_mlp_grads* g, float* dcoords_in,
                        void* stream) {
  if (int rc = check_geom("eg_coord_update_bwd", Y, batch, nodes_per_frame, coord_row0, main_row0, frame_size)) return rc;
  if (int rc = check_params("eg_coord_update_bwd", p)) return rc;
  EG_CHECK_ARG(dY && coords_in && mean1 && var1 && mean2 && var2 && feat_in && z1 && z2 && pre && coords_out && scratch,
               "eg_coord_update_bwd: NULL argument");
  EG_CHECK_ARG(g && g->dw1 && g->db1 && g->dg1 && g->dbe1 && g->dw2 && g->db2 && g->dg2 && g->dbe2 && g->dw3 && g->db3,
               "eg_coord_update_bwd: NULL gradient output");
  static std::atomic<unsigned long long> mask{0};
  if (int rc = opt_in_smem(coord_update_bwd_kernel, mask)) return rc;
  BwdArgs q{};
  q.a = Args{const_cast<float*>(Y), batch, nodes_per_frame, coord_row0, main_row0, frame_size, coords_in, *p,
             const_cast<float*>(mean1), const_cast<float*>(var1), const_cast<float*>(mean2), const_cast<float*>(var2),
             const_cast<float*>(feat_in), const_cast<float*>(z1), const_cast<float*>(z2), const_cast<float*>(pre),
             const_cast<float*>(coords_out)};
  q.dY = dY;
  q.dcoords_out = dcoords_out;
  q.scratch = scratch;
  q.g = *g;
  q.dcoords_in = dcoords_in;
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("coord_update_bwd", s);
  const int grid = (batch + kWarps - 1) / kWarps;
  coord_sample_bwd_kernel<false><<<grid < 148 ? grid : 148, kThreads, 0, s>>>(
      dY, Y, batch, nodes_per_frame, coord_row0, main_row0, frame_size, coords_out, scratch + (size_t)4 * batch * 58);
  EG_LAUNCH_CHECK();
  coord_update_bwd_kernel<<<1, kThreads, sizeof(Smem), s>>>(q);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

int eg_mae(int64_t n, const float* pred, const float* y, float loss_weight, float* loss, float* grad, void* stream) {
  EG_CHECK_ARG(n >= 1 && n < (1LL << 31) && pred && y && loss, "eg_mae: bad arguments");
  cudaStream_t s = as_stream(stream);
  ProfileScope prof("mae", s);
  mae_kernel<<<1, 256, 0, s>>>((int)n, pred, y, loss_weight, loss, grad);
  EG_LAUNCH_CHECK();
  return EG_OK;
}

}  // extern "C"
