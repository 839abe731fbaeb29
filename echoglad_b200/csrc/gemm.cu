// Dense per-node transforms [rows,128] x [128,128]: C-ABI entry points.  Both run on the 5th-generation tensor
// cores with 3xTF32 error compensation (a = a_hi + a_lo, b = b_hi + b_lo; a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,
// fp32 accumulate), which keeps fp32-class accuracy (plain TF32 / bf16 fail the 1e-4 parity bar):
//   eg_linear128        -> gcn_tc.cu, linear mode of the fused kernel (weight resident in TMEM)
//   eg_linear128_wgrad  -> wgrad_tc.cu (MN-major operands, segmented accumulation)
// Replaces nn.Linear inside PyG GCNConv.lin and node_classifiers[k][0] (src/core/models.py:330,364) and its autograd.
#include "common.cuh"

namespace eg {
int launch_linear_tc(long long rows, const float* A, const float* W, int trans_w, const float* bias,
                     const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                     cudaStream_t s);  // gcn_tc.cu
int launch_wgrad_tc(long long rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t s);  // wgrad_tc.cu
}  // namespace eg
using namespace eg;

extern "C" {

int eg_linear128(int64_t rows, const float* A, const float* W, int trans_w, const float* bias,
                 const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                 void* stream) {
  EG_CHECK_ARG(rows >= 1 && A && W && C, "eg_linear128: bad arguments");
  return launch_linear_tc(rows, A, W, trans_w, bias, addend, C, mean, var, ws, ws_bytes, as_stream(stream));
}

int eg_linear128_wgrad(int64_t rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                       size_t ws_bytes, void* stream) {
  EG_CHECK_ARG(rows >= 1 && G && X && dW, "eg_linear128_wgrad: bad arguments");
  return launch_wgrad_tc(rows, G, X, dW, dbias, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"
