/*
 * echoglad_b200 — C ABI of the B200-native (sm_100a) EchoGLAD GNN hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every device pointer is a
 * caller-owned, contiguous allocation on the device the graph handle was created on; the library
 * never frees or retains a caller pointer past the call.  All work is enqueued on the caller's
 * stream (a `cudaStream_t` passed as `void*`; NULL = legacy default stream); nothing synchronises
 * the device except eg_graph_create / eg_graph_destroy.  Every entry point returns 0 (EG_OK) or a
 * negative EG_ERR_* code; the message is available from eg_last_error() (thread-local).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the upstream
 * DSL-Lab/echoglad tree).  F = feature width; the tensor-core paths require F == 128
 * (`node_embedding_dim == node_hidden_dim == 128`, configs/default.yml:14-15).
 */
#ifndef ECHOGLAD_B200_H
#define ECHOGLAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EG_OK 0
#define EG_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define EG_ERR_CUDA (-2)        /* CUDA runtime error (message has the cudaError string) */
#define EG_ERR_UNSUPPORTED (-3) /* spec the reference itself builds malformed (e.g. 2^naux < frame/2) */
#define EG_ERR_WORKSPACE (-4)   /* caller workspace too small */

#define EG_MAX_LEVELS 16
#define EG_F 128 /* feature width of the tensor-core paths */

typedef struct eg_graph eg_graph; /* opaque, owns the device CSR of ONE frame's static graph */

/* Arguments of the reference's create_graphs (src/core/datasets.py:375-467) + data.* flags
 * (configs/default.yml:67-75). */
typedef struct eg_graph_spec {
  int32_t frame_size;            /* data.transform.image_size */
  int32_t num_aux_graphs;        /* data.num_aux_graphs (ignored when use_main_graph_only) */
  int32_t use_main_graph_only;   /* data.use_main_graph_only */
  int32_t use_coordinate_graph;  /* data.use_coordinate_graph: 4 isolated K4 nodes appended */
  int32_t use_connection_nodes;  /* data.use_connection_nodes: num_aux_graphs+1 hub nodes prepended */
  int32_t main_diagonal;         /* data.main_graph_type == 'grid-diagonal' */
  int32_t aux_diagonal;          /* data.aux_graph_type == 'grid-diagonal' */
} eg_graph_spec;

typedef struct eg_graph_info {
  int32_t num_nodes;        /* N per frame, all node types */
  int32_t num_edges;        /* directed edges per frame, no self loops (== edge_index.shape[1]) */
  int32_t num_pixel_nodes;  /* N0: nodes with node_type == 0 (contiguous range) */
  int32_t first_pixel_node; /* == number of connection nodes */
  int32_t num_coord_nodes;  /* 0 or 4, at the end of the frame */
  int32_t num_levels;       /* lattice levels: aux 1..n then main */
  int32_t level_size[EG_MAX_LEVELS];   /* side length of each lattice */
  int32_t level_offset[EG_MAX_LEVELS]; /* first node of each lattice inside the frame */
  int32_t max_degree;       /* largest in-degree incl. the GCN self loop */
  int32_t crop_offset;      /* centre-crop offset c of the finest aux level (datasets.py:502) */
} eg_graph_info;

const char* eg_version(void);
const char* eg_last_error(void);
/* Bytes of scratch the compute entry points below need (independent of the batch size). */
size_t eg_workspace_bytes(void);

/* ---- static hierarchical graph: replaces create_graphs + from_networkx + Batch collate ------------
 * (src/core/datasets.py:375-521 and :258; PyG Batch.from_data_list offsets).  Built once, on device,
 * as a symmetric CSR sorted by (target, source) with the GCN self loop last in every row and the
 * symmetric normalisation deg^-1/2[u] * deg^-1/2[v] stored per entry (PyG gcn_norm, recomputed every
 * layer every step by the reference, src/core/models.py:330). */
int eg_graph_create(const eg_graph_spec* spec, int device, eg_graph** out);
void eg_graph_destroy(eg_graph* g);
int eg_graph_get_info(const eg_graph* g, eg_graph_info* info);
/* Host-side closed form of the same graph (no GPU needed): fills info only. */
int eg_graph_spec_info(const eg_graph_spec* spec, eg_graph_info* info);
/* Device pointers of the per-frame CSR (rowptr int32[N+1], col int32[nnz], w float[nnz], dis float[N]);
 * nnz = num_edges + num_nodes. */
int eg_graph_csr(const eg_graph* g, const int32_t** rowptr, const int32_t** col, const float** w,
                 const float** dis);
/* Output-tile table used by the tensor-core kernels: DEVICE int32[tiles_per_frame][128] node ids of one
 * frame (-1 = padding), every node exactly once; 8x16 lattice patches where the level allows it. */
int eg_graph_tiles(const eg_graph* g, const int32_t** tile_nodes, int32_t* tiles_per_frame);
/* Host-only consistency check of the tile table and of the per-tile gather plan of the fused kernel against
 * the closed-form neighbour lists and gcn_norm weights (no GPU needed).  stats (optional): int64[8] = tiles,
 * plan rows, CSR rows, far edges, staged edges, largest staged-source count, and the number of tiles in each
 * class of the fused kernel (lattice, general).  Returns the number of
 * violations (0 = consistent) or a negative EG_ERR_* code. */
int eg_graph_plan_check(const eg_graph_spec* spec, int64_t* stats);
/* The fused GCN kernel has two staging plans: the PATCH plan (regular 4-neighbour lattices: TMA box copies of the haloed
 * 8x16 patch + parents, 2x2 node blocks) and the GATHER plan (any graph: per-row slot plan, cp.async row copies).
 * mode 0 = automatic (patch wherever the graph allows it; the default, or EG_GCN_PLAN=gather in the environment),
 * 1 = always gather.  Returns the previous mode.  Same results either way (within the summation-order tolerance):
 * a test / A-B hook, not a tuning knob. */
int eg_gcn_plan_select(int mode);
/* Host-only consistency check of the PATCH plan (the TMA path of the fused kernel on regular 4-neighbour lattices:
 * haloed 8x16 patch + parents staged by tensor-map box copies, one 2x2 node block per half-warp; children either read
 * directly or -- for the patches of the last aux level -- as ONE pooled row per node written by the main patches of the
 * same processing unit): the sources every patch node reads through its box positions / child window / pool rows and
 * their weights against the closed-form neighbour lists and gcn_norm weights, and the unit sequence (a permutation of
 * the tiles; a unit of several tiles = pool writers followed by their reader).  No GPU needed.  stats (optional):
 * int64[6] = plan usable for this spec (0/1), plain patch tiles, patch tiles with directly loaded children, CSR tiles,
 * patch tiles reading their unit's pool, units per frame.  Returns the number of violations or a negative EG_ERR_* code. */
int eg_graph_patch_check(const eg_graph_spec* spec, int64_t* stats);
/* edge_index exactly as the reference's loader produces it for a batch of `batch` frames:
 * int64[2, batch*num_edges], grouped by source in the networkx insertion order, frame b offset by
 * b*num_nodes.  `out` is a DEVICE pointer. */
int eg_graph_export_edge_index(const eg_graph* g, int batch, int64_t* out, void* stream);
/* Same on the host, no GPU needed (`out` is a HOST pointer): lets a data loader emit the tensor the
 * reference builds with networkx in ~3 s per sample. */
int eg_graph_host_edge_index(const eg_graph_spec* spec, int batch, int64_t* out);
/* node_type float64[batch*N] (0 pixel, 1 coordinate, 2 connection; datasets.py:386-460), HOST pointer. */
int eg_graph_host_node_type(const eg_graph_spec* spec, int batch, double* out);
/* Validates a caller-supplied batched edge_index (DEVICE int64[2, num_cols]) against the static graph;
 * writes the number of mismatching entries to *mismatch (DEVICE int32, set to -1 on a shape mismatch). */
int eg_graph_check_edge_index(const eg_graph* g, int batch, const int64_t* edge_index, int64_t num_cols,
                              int32_t* mismatch, void* stream);

/* ---- node-feature packing: replaces the per-frame permute/reshape/cat loop --------------------------
 * (src/core/models.py:722-756).  maps[l] = DEVICE float[batch, F, s_l, s_l] (NCHW) for lattice level l;
 * `maps` itself is a HOST array of num_levels pointers.  head = float[batch, first_pixel_node, F]
 * (connection-node rows) or NULL; tail = float[batch, num_coord_nodes, F] or NULL (coordinate rows left to
 * eg_coord_sample_fwd / already consumed by eg_coord_sample_bwd).  A NULL maps[l] skips level l (its rows are
 * written by eg_level_embed_fwd).
 * X = float[batch*N, F] node-major.  The _grad form scatters dX back (d_maps etc. are outputs). */
int eg_pack_nodes(const eg_graph* g, int batch, const float* const* maps, const float* head,
                  const float* tail, float* X, void* stream);
int eg_pack_nodes_grad(const eg_graph* g, int batch, const float* dX, float* const* d_maps, float* d_head,
                       float* d_tail, void* stream);

/* ---- fused level embedding: 1x1 conv (cin -> F) + bias + ReLU + packing of ONE lattice level -----------
 * replaces `new_features[l] = F.relu(self.linears[l](features[l]))` (src/core/models.py:708-710) and that
 * level's share of the packing loop (:728-741).  in = DEVICE float[batch, cin, s_l, s_l] (the UNet decoder
 * map), W = float[F, cin] (Conv2d 1x1 weight), bias = float[F]; writes rows [off_l, off_l + s_l^2) of every
 * frame of X.  Built for cin in {4, 8} and s_l^2 % 64 == 0 (eg_level_embed_supported); levels packed this
 * way are passed as NULL to eg_pack_nodes / eg_pack_nodes_grad.
 * _bwd: d_in (nullable) = float[batch, cin, s_l, s_l], dW = float[F, cin], dbias = float[F]; the ReLU mask
 * is recomputed from `in` (bit-identical to the forward), nothing is saved between the passes. */
int eg_level_embed_supported(const eg_graph* g, int level, int cin);
int eg_level_embed_fwd(const eg_graph* g, int batch, int level, int cin, const float* in, const float* W,
                       const float* bias, float* X, void* stream);
int eg_level_embed_bwd(const eg_graph* g, int batch, int level, int cin, const float* in, const float* W,
                       const float* bias, const float* dX, float* d_in, float* dW, float* dbias, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- message passing: out = A_hat * in, A_hat = D^-1/2 (A+I) D^-1/2 -------------------------------
 * replaces PyG GCNConv.propagate / torch_scatter.scatter_add (called from src/core/models.py:431).
 * in/out: float[batch*N, F], F in {64,128,256}.  Atomic-free segmented reduction over the sorted CSR;
 * A_hat is symmetric, so the same call is the backward. */
int eg_gcn_aggregate(const eg_graph* g, int batch, int feat, const float* in, float* out, void* stream);

/* ---- one GCNConv: H = A_hat * X * W^T + bias, plus the batch statistics BatchNorm1d needs ----------
 * replaces gnn_layers[i].module_0 (+ the statistics pass of module_1), src/core/models.py:329-332.
 * X,H: float[batch*N,128]; W: float[128(out),128(in)]; bias: float[128] or NULL.
 * If mean/var are non-NULL they receive the per-column mean and BIASED variance of H over all rows
 * (deterministic two-stage reduction).  ws: eg_workspace_bytes() bytes of device scratch. */
int eg_gcn_conv_fwd(const eg_graph* g, int batch, const float* X, const float* W, const float* bias,
                    float* H, float* mean, float* var, void* ws, size_t ws_bytes, void* stream);
/* Backward of the above given dH: dX = A_hat*(dH*W) (+ dX_add if non-NULL, the residual branch),
 * dW = (A_hat*dH)^T X, dbias = column sums of dH.  scratch: float[batch*N,128] for A_hat*dH. */
int eg_gcn_conv_bwd(const eg_graph* g, int batch, const float* X, const float* W, const float* dH,
                    const float* dX_add, float* dX, float* dW, float* dbias, float* scratch, void* ws,
                    size_t ws_bytes, void* stream);
/* One whole GNN layer in EVAL mode (model.eval(): BatchNorm1d with its running statistics, Dropout off) as ONE launch of
 * the fused kernel: Y = act(gamma * (A_hat X W^T + bias - running_mean) / sqrt(running_var + eps) + beta) (+ X if
 * residual) -- replaces gnn_layers[i].module_0..3 and `h + hidden_embeds[i]` of src/core/models.py:329-335,431-435 when
 * no gradient is needed (validation / inference under torch.no_grad(), src/engine.py:343-350,383-405).  The pre-activation H is never written:
 * 2 (+1 with the residual) node tensors cross HBM per layer instead of 5 for eg_gcn_conv_fwd + eg_bn_act_fwd.  Measured
 * at default.yml / batch 64: 1.22 ms without the residual, 2.65 ms with it (the two launches: 2.32 ms) -- the module takes
 * this route for residual-free layers only.  bias may be NULL; relu: 1 = ReLU,
 * 0 = identity (the last layer).  X, Y: float[batch*N,128], not aliased.  ws: at least 1024 bytes of device scratch. */
int eg_gcn_layer_eval_fwd(const eg_graph* g, int batch, const float* X, const float* W, const float* bias,
                          const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                          float eps, int relu, int residual, float* Y, void* ws, size_t ws_bytes, void* stream);

/* ---- train-mode BatchNorm2d over NCHW maps with few channels, the preceding ReLU and conv bias folded in -----------
 * y = BN(relu(x + pre_bias)): the full-resolution levels of the UNet pyramid in front of the graph path
 * (conv3x3 -> ReLU -> BatchNorm2d, src/core/models.py:841-876; 4..64 channels, where cuDNN's spatial BN runs one CTA per
 * channel).  x, y, dy, dx: float[n, channels, H*W] contiguous (NCHW), hw = H*W a multiple of 4, channels <= 64.
 * pre_bias: float[channels] or NULL -- the bias of the convolution that produced x, when that convolution was run
 * WITHOUT it (saves the bias-add pass; dpre_bias = sum of dx per channel is its gradient, a by-product of the backward).
 * relu_in != 0: BatchNorm's input is relu(x + pre_bias) (and dx carries the ReLU mask).  mean / var: float[channels]
 * outputs of the forward (batch mean, BIASED variance), inputs of the backward.  dx / dpre_bias may be NULL.
 * Deterministic two-stage reductions.  ws: eg_workspace_bytes() bytes of device scratch. */
int eg_bn2d_fwd(int n, int channels, int64_t hw, const float* x, const float* pre_bias, int relu_in, const float* gamma,
                const float* beta, float eps, float* y, float* mean, float* var, void* ws, size_t ws_bytes,
                void* stream);
int eg_bn2d_bwd(int n, int channels, int64_t hw, const float* x, const float* pre_bias, int relu_in, const float* dy,
                const float* mean, const float* var, const float* gamma, float eps, float* dx, float* dgamma,
                float* dbeta, float* dpre_bias, void* ws, size_t ws_bytes, void* stream);

/* ---- fused BatchNorm1d-apply + Dropout + ReLU|Identity + residual ----------------------------------
 * replaces gnn_layers[i].module_1..3 and `h + hidden_embeds[i]` (src/core/models.py:332-335,434-435)
 * and the BN/ReLU/Dropout triples of the classifiers (:366-373).
 * Y = act(drop(gamma*(H-mean)*rsqrt(var+eps)+beta)) + res.   rows x cols, cols in {16..128, %4==0}.
 * Dropout keeps an element when hash(seed, row*cols+col) >= p (counter-based, recomputed in backward;
 * eg_dropout_mask exports it for parity tests).  res may be NULL. */
int eg_bn_act_fwd(int64_t rows, int cols, const float* H, const float* mean, const float* var,
                  const float* gamma, const float* beta, float eps, float drop_p, uint64_t seed, int relu,
                  const float* res, float* Y, void* stream);
/* Backward.  batch_stats != 0: train-mode BN (gradient flows through mean/var); 0: eval mode.
 * Outputs dH (may alias dY), dgamma, dbeta (float[cols]).  The residual gradient is dY itself. */
int eg_bn_act_bwd(int64_t rows, int cols, const float* dY, const float* H, const float* mean,
                  const float* var, const float* gamma, const float* beta, float eps, float drop_p,
                  uint64_t seed, int relu, int batch_stats, float* dH, float* dgamma, float* dbeta, void* ws,
                  size_t ws_bytes, void* stream);
/* mask float[rows*cols] in {0, 1/(1-p)} — the factor eg_bn_act_fwd applied. */
int eg_dropout_mask(int64_t rows, int cols, float drop_p, uint64_t seed, float* mask, void* stream);
/* Per-column mean and biased variance of Z float[rows, cols]. */
int eg_col_stats(int64_t rows, int cols, const float* Z, float* mean, float* var, void* ws, size_t ws_bytes,
                 void* stream);

/* ---- dense per-node transforms (tensor cores, split precision: tf32 main term + bf16 correction terms, fp32 accumulate => fp32-class accuracy, within 2e-6 of fp64) -----
 * replaces nn.Linear inside GCNConv.lin and node_classifiers[k][0] stacked over k
 * (src/core/models.py:330,364).  C[rows,128] = A[rows,128] * op(W) + bias, op(W) = W^T when
 * trans_w != 0 (nn.Linear forward), W otherwise (its input gradient).  mean/var optional as above. */
int eg_linear128(int64_t rows, const float* A, const float* W, int trans_w, const float* bias,
                 const float* addend, float* C, float* mean, float* var, void* ws, size_t ws_bytes,
                 void* stream);
/* dW[128,128] = G^T X (nn.Linear weight gradient, G = dOut[rows,128], X = input[rows,128]);
 * dbias = column sums of G (may be NULL). */
int eg_linear128_wgrad(int64_t rows, const float* G, const float* X, float* dW, float* dbias, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- classifier tail: 4 block-diagonal heads 32 -> 16 -> 1 -----------------------------------------
 * replaces node_classifiers[k][4] and [8] for k = 0..3 (src/core/models.py:369-375).
 * A1: float[rows,128] (activated layer-1 output, head k in columns 32k..32k+31);
 * W2: float[4,16,32], b2: float[4,16]; Z2: float[rows,64] (+ optional column stats).
 * A2: float[rows,64]; W3: float[4,16]; b3: float[4]; logits: float[rows,4]. */
int eg_clf_mid_fwd(int64_t rows, const float* A1, const float* W2, const float* b2, float* Z2, float* mean,
                   float* var, void* ws, size_t ws_bytes, void* stream);
int eg_clf_mid_bwd(int64_t rows, const float* A1, const float* W2, const float* dZ2, float* dA1, float* dW2,
                   float* db2, void* ws, size_t ws_bytes, void* stream);
int eg_clf_out_fwd(int64_t rows, const float* A2, const float* W3, const float* b3, int sigmoid,
                   float* out, void* stream);
int eg_clf_out_bwd(int64_t rows, const float* A2, const float* W3, const float* out, const float* dout,
                   int sigmoid, float* dA2, float* dW3, float* db3, void* ws, size_t ws_bytes, void* stream);

/* ---- losses ----------------------------------------------------------------------------------------
 * eg_bce_multilevel: replaces WeightedBCEWithLogitsLoss.compute (src/core/criterion.py:13-27,30-34):
 * loss = loss_weight * sum(valid * w(y) * bce_with_logits(x, y)) / sum(valid), w = ones_weight if y==1
 * (and ones_weight > 1).  n = number of elements (B*N0*4).  loss_out: DEVICE float[1];
 * dlogits (optional): d loss / d logits. */
int eg_bce_multilevel(int64_t n, const float* logits, const float* y, const float* valid, float ones_weight,
                      float loss_weight, float* loss_out, float* dlogits, void* ws, size_t ws_bytes,
                      void* stream);
/* eg_expected_landmark_mse: replaces ExpectedLandmarkMSE.compute (src/core/criterion.py:93-151).
 * logits/y/valid: float[batch, n0, channels] where n0 = sum(level_size^2); levels given explicitly. */
int eg_expected_landmark_mse(int batch, int channels, int num_levels, const int32_t* level_size,
                             const float* logits, const float* y, const float* valid, float loss_weight,
                             float* loss_out, float* dlogits, void* ws, size_t ws_bytes, void* stream);
/* Multi-level one-hot labels on device from landmark pixel coordinates — replaces create_node_labels
 * (src/core/datasets.py:523-549).  coords: DEVICE int32[batch, channels, 2] (h, w) in [0, frame);
 * y: float[batch, n0, channels].  A coordinate >= frame (the reference raises IndexError, :536-537) or < -frame
 * cannot be labelled: the (frame, channel)'s labels are poisoned with NaN so that every loss computed from them is
 * NaN instead of silently wrong, and *oob_count (DEVICE int32, optional, caller-zeroed) is incremented. */
int eg_node_labels(int batch, int channels, int frame_size, int num_levels, const int32_t* level_size,
                   const int32_t* coords, float* y, int32_t* oob_count, void* stream);

/* ---- post-path metric: expected landmark coordinates (src/core/evaluators.py:310-348) ----------------
 * On the main level (the last frame_size^2 of the nodes_per_frame pixel nodes of every frame): softmax over the
 * nodes of each of the 4 channels and its expected (h, w) -> pred_hw float[batch,4,2]; position of the label heat
 * map maximum (first maximum along each axis, as torch.max) -> gt_hw int32[batch,4,2]; mean of `valid` (NULL =
 * all valid) -> valid_mean float[batch,4].  logits / y / valid: DEVICE float[batch*nodes_per_frame, 4]. */
int eg_expected_coords(int batch, int channels, int nodes_per_frame, int frame_size, const float* logits,
                       const float* y, const float* valid, float* pred_hw, int32_t* gt_hw, float* valid_mean,
                       void* stream);

/* ---- the four node classifiers as one chain (replaces node_classifiers[0..3] applied to h and concatenated,
 * src/core/models.py:363-377,488-490, and their autograd) --------------------------------------------------------
 * Stacked / block-diagonal parameters: w1 float[128,128] = 4 x Linear(128,32).weight stacked, b1 float[128];
 * g1 / be1 float[128] = 4 x BatchNorm1d(32) weight / bias; w2 float[4,16,32], b2 float[4,16]; g2 / be2 float[64];
 * w3 float[4,16], b3 float[4].  batch_stats != 0: train-mode BatchNorm (batch statistics are COMPUTED into
 * mean1/var1/mean2/var2); 0: those four vectors are INPUTS (running statistics).  Dropout (p = drop_p, counter-based
 * with `seed` for layer 1 and `seed + 1` for layer 2, as eg_bn_act_fwd) is applied when drop_p > 0.
 * Forward keeps only the two pre-activations z1 float[rows,128] and z2 float[rows,64] for the backward; no activated
 * tensor is materialised.  out: float[rows,4] (logits, or probabilities when sigmoid != 0). */
typedef struct eg_classifier_params {
  const float *w1, *b1, *g1, *be1;
  const float *w2, *b2, *g2, *be2;
  const float *w3, *b3;
  float eps;
  float drop_p;
  uint64_t seed;
  int32_t batch_stats;
  int32_t sigmoid;
} eg_classifier_params;
typedef struct eg_classifier_grads {  /* all required; same shapes as the parameters */
  float *dw1, *db1, *dg1, *dbe1;
  float *dw2, *db2, *dg2, *dbe2;
  float *dw3, *db3;
} eg_classifier_grads;
int eg_classifier_fwd(int64_t rows, const float* h, const eg_classifier_params* p, float* mean1, float* var1,
                      float* mean2, float* var2, float* z1, float* z2, float* out, void* ws, size_t ws_bytes,
                      void* stream);
/* Backward from dout float[rows,4].  scratch: float[rows * 192] (overwritten: [rows,128] masked layer-1 gradient, then
 * dz1 in place; followed by [rows,64] dz2);
 * dh (optional): float[rows,128] gradient with respect to h; `out` is read for the sigmoid head only. */
int eg_classifier_bwd(int64_t rows, const float* h, const eg_classifier_params* p, const float* mean1,
                      const float* var1, const float* mean2, const float* var2, const float* z1, const float* z2,
                      const float* out, const float* dout, float* scratch, float* dh,
                      const eg_classifier_grads* g, void* ws, size_t ws_bytes, void* stream);

/* ---- coordinate-graph branch (`use_coordinate_graph`; replaces the per-layer coordinate update of
 * HierarchicalPatchModel.forward, src/core/models.py:438-473, `bilinear_interpolation`, :539-553, the initial
 * coordinate-node features of create_node_pixels, :526-527 / :743-744, and their autograd) ---------------------------
 * Geometry of the node tensor Y float[batch * nodes_per_frame, 128]: the 4 coordinate nodes of frame b are rows
 * b*nodes_per_frame + coord_row0 + {0..3}; its row-major frame_size x frame_size main lattice starts at row
 * b*nodes_per_frame + main_row0.  coords: float[4*batch, 2] = (h, w) per landmark, the layout of the reference's
 * `node_coords`.  The bilinear map gathers the <= 4 taps whose tent weight relu(1 - |c - g|) is non-zero (the
 * reference multiplies a dense [4,S,S] weight map into the frame: same weights, same sub-gradients).
 *
 * eg_coord_sample_fwd: Y[coordinate rows] = bilinear sample of the main rows at `coords` (in place).
 * eg_coord_sample_bwd: turns dY (gradient of the tensor AFTER the sample) in place into the gradient of the tensor
 *   before it: tap rows += weight * dY[coordinate row], coordinate rows = 0; dcoords (optional) float[4*batch,2]. */
int eg_coord_sample_fwd(float* Y, int batch, int nodes_per_frame, int coord_row0, int main_row0, int frame_size,
                        const float* coords, void* stream);
int eg_coord_sample_bwd(float* dY, const float* Y, int batch, int nodes_per_frame, int coord_row0, int main_row0,
                        int frame_size, const float* coords, float* dcoords, void* stream);

/* node_coordinate_mlp[i] (src/core/models.py:337-350): Linear(136,32) BN ReLU Dropout Linear(32,16) BN ReLU Dropout
 * Linear(16,2).  w1 float[32,136], w2 float[16,32], w3 float[2,16]; batch_stats / drop_p / seed as in
 * the classifier parameter block: dropout stream `seed` for layer 1, `seed + 1` for layer 2, element index
 * row*width + col. */
typedef struct eg_coord_mlp_params {
  const float *w1, *b1, *g1, *be1;
  const float *w2, *b2, *g2, *be2;
  const float *w3, *b3;
  float eps;
  float drop_p;
  uint64_t seed;
  int32_t batch_stats;
  int32_t reserved;
} eg_coord_mlp_params;
typedef struct eg_coord_mlp_grads { /* all required; same shapes as the parameters */
  float *dw1, *db1, *dg1, *dbe1;
  float *dw2, *db2, *dg2, *dbe2;
  float *dw3, *db3;
} eg_coord_mlp_grads;
/* One coordinate update after a GNN layer (row gather, one-CTA MLP, re-sampling): relative-position features + coordinate-node embeddings ->
 * MLP -> coords_out = clamp(coords_in + delta, 0, frame_size-1) -> the coordinate rows of Y are re-sampled at
 * coords_out and overwritten IN PLACE.  mean1/var1 float[32], mean2/var2 float[16]: outputs when batch_stats != 0
 * (train-mode BatchNorm), inputs (running statistics) otherwise.  Saved for the backward (all outputs): feat_in
 * float[4*batch,128] (the coordinate rows before the overwrite), z1 float[4*batch,32], z2 float[4*batch,16]
 * (pre-activations), pre float[4*batch,2] (coordinates before the clamp), coords_out float[4*batch,2]. */
int eg_coord_update_fwd(float* Y, int batch, int nodes_per_frame, int coord_row0, int main_row0, int frame_size,
                        const float* coords_in, const eg_coord_mlp_params* p, float* mean1, float* var1, float* mean2,
                        float* var2, float* feat_in, float* z1, float* z2, float* pre, float* coords_out,
                        void* stream);
/* Backward: dY float[batch*nodes_per_frame,128] = gradient of the UPDATED tensor, turned in place into the gradient
 * of the layer output (tap rows += weight * d(new row); coordinate rows = MLP input gradient); dcoords_out (optional)
 * = gradient of coords_out; Y = the updated tensor (its main rows are read); scratch float[4*batch*64];
 * dcoords_in (optional) float[4*batch,2]. */
int eg_coord_update_bwd(float* dY, const float* dcoords_out, const float* Y, int batch, int nodes_per_frame,
                        int coord_row0, int main_row0, int frame_size, const float* coords_in,
                        const eg_coord_mlp_params* p, const float* mean1, const float* var1, const float* mean2,
                        const float* var2, const float* feat_in, const float* z1, const float* z2, const float* pre,
                        const float* coords_out, float* scratch, const eg_coord_mlp_grads* g, float* dcoords_in,
                        void* stream);
/* CRITERIA['mae'] = the 'coordinate' loss (src/core/criterion.py:52-64, src/builders/criterion_builder.py:40-41):
 * loss[0] = loss_weight * mean |pred - y| over n values; grad (optional) float[n] = d loss / d pred. */
int eg_mae(int64_t n, const float* pred, const float* y, float loss_weight, float* loss, float* grad, void* stream);

/* ---- generic-width dense transforms over strided views (fp32 FMA; any width) -----------------------------------
 * Element (r, c) of a view = base[(r / frame_rows) * frame_stride + (r % frame_rows) * row_stride + c * col_stride]
 * (strides in elements): a row-major [rows, F] tensor is {p, rows, 0, F, 1}; lattice level l inside the node tensor
 * is {X + off_l * F, s_l^2, N * F, F, 1}; an NCHW map [B, C, s, s] seen as [B * s^2, C] is {p, s^2, C * s^2, 1, s^2}.
 * Replaces (a) `F.relu(self.linears[l](features[l]))` + that level's packing for the SMALL pyramid levels
 * (src/core/models.py:708-710,728-741; the two big levels go through eg_level_embed_*), (b) nn.Linear / GCNConv.lin
 * of models whose widths are not 128 / 32 (constructor defaults 64 / 16, src/core/models.py:290-296).
 * eg_linear_fwd: y = act(a_eff op(w) + bias + addend); w = float[n, k] with trans_w != 0 (nn.Linear forward),
 *   float[k, n] with trans_w == 0 (the input gradient of a Linear whose weight is [k, n]); a_eff = a where
 *   gate > 0, else 0 (gate optional: the ReLU mask of a forward output); bias, addend optional; relu != 0: ReLU.
 * eg_linear_wgrad: dw float[n, k] = g_eff^T a, db (optional) float[n] = column sums of g_eff; ws as everywhere. */
typedef struct eg_view {
  float* base;
  int64_t frame_rows;
  int64_t frame_stride, row_stride, col_stride;
} eg_view;
int eg_linear_fwd(int64_t rows, int k, int n, const eg_view* a, const eg_view* gate, const float* w, int trans_w,
                  const float* bias, const eg_view* addend, int relu, const eg_view* y, void* stream);
int eg_linear_wgrad(int64_t rows, int k, int n, const eg_view* g, const eg_view* gate, const eg_view* a, float* dw,
                    float* db, void* ws, size_t ws_bytes, void* stream);

/* ---- measurement hooks (no reference counterpart) ---------------------------------------------------
 * eg_profile_enable(1) clears and starts recording CUDA-event spans around every launch helper on the
 * caller's stream; eg_profile_read sums the device time and span count recorded under `name`
 * (synchronises on the recorded events); eg_profile_names lists the recorded names, comma separated.
 * eg_launch_count: number of kernels this library has launched in this process. */
int eg_profile_enable(int on);
int eg_profile_read(const char* name, double* total_ms, int64_t* launches);
int eg_profile_names(char* buf, size_t n);
int64_t eg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ECHOGLAD_B200_H */
