// One GCNConv, forward and backward: C-ABI entry points over the fused tensor-core kernel.
// Replaces gnn_layers[i].module_0 (PyG GCNConv) — src/core/models.py:329-331,431.
//
// Re-association used throughout: A_hat (X W^T) == (A_hat X) W^T, so the aggregation runs on the layer
// input and feeds the transform directly (fp32 error of the re-association measured at 4e-7 rms,
// SURVEY.md §7.3).  A_hat is symmetric, hence in the backward
//     G = A_hat dH,   dX = G W (+ residual gradient),   dW = G^T X,   dbias = colsum(dH).
#include "common.cuh"

struct eg_graph;
namespace eg {
const eg_graph_info& graph_info(const eg_graph* g);
int launch_aggregate(const eg_graph* g, int batch, int feat, const float* in, float* out, cudaStream_t s);
int launch_wgrad_tc(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s);
int launch_gcn_tc(const eg_graph* g, int batch, const float* X, const float* W, int trans_w, const float* bias,
                  const float* addend, float* Out, float* AggOut, float* mean, float* var, void* ws, size_t ws_bytes,
                  cudaStream_t s);
int launch_col_sums(long long rows, int cols, const float* Z, float* sums, void* ws, size_t ws_bytes,
                    cudaStream_t s);
int launch_gcn_tc_eval(const eg_graph* g, int batch, const float* X, const float* W, const float* scale,
                       const float* shift, int relu, const float* addend, float* Out, cudaStream_t s);
}  // namespace eg
using namespace eg;

namespace {
// conv bias + eval-mode BatchNorm as one affine map per feature: y = acc * sc + sh with sc = gamma / sqrt(var + eps)
// (the rounding of bn.cu: invstd = 1 / sqrtf(var + eps), sc = gamma * invstd) and sh = (bias - mean) * sc + beta.
__global__ void fold_bn_kernel(const float* __restrict__ bias, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, float* __restrict__ sc, float* __restrict__ sh) {
  const int f = threadIdx.x;
  const float invstd = 1.0f / sqrtf(var[f] + eps);
  const float c = gamma[f] * invstd;
  sc[f] = c;
  sh[f] = fmaf((bias ? bias[f] : 0.f) - mean[f], c, beta[f]);
}
}  // namespace

extern "C" {

int eg_gcn_conv_fwd(const eg_graph* g, int batch, const float* X, const float* W, const float* bias, float* H,
                    float* mean, float* var, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(g && X && W && H && batch >= 1, "eg_gcn_conv_fwd: bad arguments");
  EG_CHECK_ARG(X != H, "eg_gcn_conv_fwd: X and H must not alias");
  // one fused tcgen05 kernel: gather -> 3xTF32 MMA -> bias + statistics epilogue
  return launch_gcn_tc(g, batch, X, W, 1, bias, nullptr, H, nullptr, mean, var, ws, ws_bytes, as_stream(stream));
}

int eg_gcn_layer_eval_fwd(const eg_graph* g, int batch, const float* X, const float* W, const float* bias,
                          const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                          float eps, int relu, int residual, float* Y, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(g && X && W && gamma && beta && running_mean && running_var && Y && batch >= 1,
               "eg_gcn_layer_eval_fwd: bad arguments");
  EG_CHECK_ARG(X != Y, "eg_gcn_layer_eval_fwd: X and Y must not alias");
  if (!ws || ws_bytes < 2 * EG_F * sizeof(float)) {
    set_error("eg_gcn_layer_eval_fwd: workspace too small: need %zu bytes", 2 * EG_F * sizeof(float));
    return EG_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  float* sc = reinterpret_cast<float*>(ws);
  fold_bn_kernel<<<1, EG_F, 0, s>>>(bias, gamma, beta, running_mean, running_var, eps, sc, sc + EG_F);
  EG_LAUNCH_CHECK();
  return launch_gcn_tc_eval(g, batch, X, W, sc, sc + EG_F, relu, residual ? X : nullptr, Y, s);
}

int eg_gcn_conv_bwd(const eg_graph* g, int batch, const float* X, const float* W, const float* dH,
                    const float* dX_add, float* dX, float* dW, float* dbias, float* scratch, void* ws,
                    size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(g && W && dH && scratch && batch >= 1, "eg_gcn_conv_bwd: bad arguments");
  EG_CHECK_ARG(scratch != dH && scratch != dX, "eg_gcn_conv_bwd: scratch must not alias dH / dX");
  cudaStream_t s = as_stream(stream);
  const long long rows = (long long)batch * graph_info(g).num_nodes;
  int rc;
  if (dX)  // G = A_hat dH (side output) and dX = G W + dX_add in one pass over dH
    rc = launch_gcn_tc(g, batch, dH, W, 0, nullptr, dX_add, dX, scratch, nullptr, nullptr, ws, ws_bytes, s);
  else     // only the weight gradient is wanted: bare aggregation
    rc = launch_aggregate(g, batch, EG_F, dH, scratch, s);
  if (rc) return rc;
  if (dW) {
    EG_CHECK_ARG(X, "eg_gcn_conv_bwd: dW requested but X is NULL");
    rc = launch_wgrad_tc(rows, scratch, X, dW, nullptr, ws, ws_bytes, s);
    if (rc) return rc;
  }
  if (dbias) {
    rc = launch_col_sums(rows, EG_F, dH, dbias, ws, ws_bytes, s);
    if (rc) return rc;
  }
  return EG_OK;
}

}  // extern "C"
