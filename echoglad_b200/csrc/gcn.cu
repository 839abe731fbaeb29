// One GCNConv, forward and backward, composed from the aggregation and the tensor-core transform.
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
int launch_linear128(long long rows, const float* A, const float* W, int trans_w, const float* bias,
                     const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                     cudaStream_t s);
int launch_wgrad128(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s);
int launch_gcn_tc(const eg_graph* g, int batch, const float* X, const float* W, int trans_w, const float* bias,
                  const float* addend, float* Out, float* AggOut, float* mean, float* var, void* ws, size_t ws_bytes,
                  cudaStream_t s);
bool legacy_mma();
int launch_col_sums(long long rows, int cols, const float* Z, float* sums, void* ws, size_t ws_bytes,
                    cudaStream_t s);
}  // namespace eg
using namespace eg;

extern "C" {

int eg_gcn_conv_fwd(const eg_graph* g, int batch, const float* X, const float* W, const float* bias, float* H,
                    float* mean, float* var, void* ws, size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(g && X && W && H && batch >= 1, "eg_gcn_conv_fwd: bad arguments");
  EG_CHECK_ARG(X != H, "eg_gcn_conv_fwd: X and H must not alias");
  cudaStream_t s = as_stream(stream);
  const long long rows = (long long)batch * graph_info(g).num_nodes;
  if (!legacy_mma())  // one fused tcgen05 kernel: gather -> 3xTF32 MMA -> bias + statistics epilogue
    return launch_gcn_tc(g, batch, X, W, 1, bias, nullptr, H, nullptr, mean, var, ws, ws_bytes, s);
  int rc = launch_aggregate(g, batch, EG_F, X, H, s);  // H <- A_hat X
  if (rc) return rc;
  // H <- H W^T + b, in place: every CTA reads its row tile into shared memory before writing it back
  return launch_linear128(rows, H, W, 1, bias, nullptr, H, mean, var, ws, ws_bytes, s);
}

int eg_gcn_conv_bwd(const eg_graph* g, int batch, const float* X, const float* W, const float* dH,
                    const float* dX_add, float* dX, float* dW, float* dbias, float* scratch, void* ws,
                    size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(g && W && dH && scratch && batch >= 1, "eg_gcn_conv_bwd: bad arguments");
  EG_CHECK_ARG(scratch != dH && scratch != dX, "eg_gcn_conv_bwd: scratch must not alias dH / dX");
  cudaStream_t s = as_stream(stream);
  const long long rows = (long long)batch * graph_info(g).num_nodes;
  int rc;
  const bool fused = dX && !legacy_mma();
  if (fused)  // G = A_hat dH (side output) and dX = G W + dX_add in one pass over dH
    rc = launch_gcn_tc(g, batch, dH, W, 0, nullptr, dX_add, dX, scratch, nullptr, nullptr, ws, ws_bytes, s);
  else
    rc = launch_aggregate(g, batch, EG_F, dH, scratch, s);  // G = A_hat dH
  if (rc) return rc;
  if (dW) {
    EG_CHECK_ARG(X, "eg_gcn_conv_bwd: dW requested but X is NULL");
    rc = launch_wgrad128(rows, scratch, X, dW, nullptr, ws, ws_bytes, s);
    if (rc) return rc;
  }
  if (dbias) {
    rc = launch_col_sums(rows, EG_F, dH, dbias, ws, ws_bytes, s);
    if (rc) return rc;
  }
  if (dX && !fused) {
    rc = launch_linear128(rows, scratch, W, 0, nullptr, dX_add, dX, nullptr, nullptr, ws, ws_bytes, s);
    if (rc) return rc;
  }
  return EG_OK;
}

}  // extern "C"
