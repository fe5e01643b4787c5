/*
 * peagnn.h - C ABI of libpeagnn_sm100.so, the B200 (sm_100a) replacement for the
 * arithmetic underneath PEAGNN's metapath message-passing hot path.
 *
 * The reference (ecml-peagnn/graph_recsys_benchmark) is pure Python: it reaches its
 * arithmetic through torch-geometric 1.5.0 / torch-scatter 2.0.5 (requirements.txt:46-47).
 * Each entry point below names the reference call it stands in for (file:line are relative
 * to the reference checkout).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - the caller owns every buffer (inputs, outputs, workspaces); the library never
 *    allocates or frees device memory and never synchronises the stream;
 *  - every function launches on `stream` (a cudaStream_t passed as void*) and returns 0,
 *    or a negative code with a message retrievable via peagnn_last_error() (thread local);
 *  - floating tensors are fp32 row-major with an explicit leading dimension (in elements);
 *    feature widths must be multiples of 4 and rows 16-byte aligned (128-bit loads);
 *  - index tensors handed over by the reference API are int64 (torch.long); CSR structures
 *    built by this library are int32.
 */
#ifndef PEAGNN_H_
#define PEAGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEAGNN_OK 0
#define PEAGNN_ERR_ARG (-1)
#define PEAGNN_ERR_CUDA (-2)
#define PEAGNN_ERR_WORKSPACE (-3)

typedef void* peagnn_stream_t; /* cudaStream_t */

int peagnn_version(void);
const char* peagnn_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long peagnn_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Graph preparation (integer, bit-exact).
 * Replaces what PyG's MessagePassing.propagate does implicitly with the COO `edge_index`
 * handed over by utils/general_utils.py:280-395 (row 0 = source, row 1 = target) and the
 * self-loop handling of GCNConv/GATConv (add_remaining_self_loops / remove_self_loops).
 * ---------------------------------------------------------------------------------------- */

/* Bytes of scratch peagnn_csr_build needs for E edges. */
size_t peagnn_csr_workspace_bytes(int64_t num_edges, int32_t num_nodes);

/* Group the edges by `key` (stable: ties keep COO order).  rowptr[num_nodes+1]; col[k] = val
 * of the k-th grouped edge; eid[k] = its position in the COO list.  With drop_self_loops != 0,
 * edges with key == val are left out (rowptr[num_nodes] = number kept; col/eid entries beyond
 * that are unspecified).  Grouping by target gives the forward (gather) structure, grouping by
 * source the transposed one used by the backward pass. */
int peagnn_csr_build(const int64_t* key, const int64_t* val, int64_t num_edges, int32_t num_nodes,
                     int drop_self_loops, int32_t* rowptr, int32_t* col, int32_t* eid,
                     void* workspace, size_t workspace_bytes, peagnn_stream_t stream);

/* out[i] = (rowptr[i+1] - rowptr[i] + add) ^ power  (power = -0.5: GCN deg^-1/2 with add = 1,
 * PyG-1.5.0 gcn_norm on the SOURCE index; power = -1: SAGE 1/max(count,1) with add = 0 and
 * clamp_min_one != 0). */
int peagnn_degree_scale(const int32_t* rowptr, int32_t num_nodes, float add, float power,
                        int clamp_min_one, float* out, peagnn_stream_t stream);

/* A CSR-by-destination view (possibly a row shard of it) plus its heavy-row work list.
 * Rows whose degree exceeds heavy_threshold are cut into chunks of at most chunk_edges edges;
 * each chunk is reduced by one CTA into `partial`, then combined in chunk order (deterministic).
 */
typedef struct {
  const int32_t* rowptr;          /* [nrows + 1] offsets into col (absolute)                  */
  const int32_t* col;             /* neighbour (gathered) node ids, GLOBAL                    */
  int32_t nrows;                  /* rows of this view                                        */
  int32_t row_offset;             /* global node id of local row 0                            */
  int32_t heavy_threshold;        /* degree above which a row is on the heavy list            */
  int32_t n_heavy;
  const int32_t* heavy_rows;      /* [n_heavy] LOCAL row ids, ascending                       */
  const int32_t* heavy_chunk_ptr; /* [n_heavy + 1] first chunk of each heavy row              */
  int32_t n_chunks;
  const int32_t* chunk_row;       /* [n_chunks] LOCAL row id of the chunk                     */
  const int32_t* chunk_begin;     /* [n_chunks] edge range of the chunk (absolute offsets)    */
  const int32_t* chunk_end;
  float* partial;                 /* workspace, >= peagnn_partial_floats(...) floats          */
  int64_t nnz;                    /* edges of this view (scheduling hint: sparse views pack
                                     several rows per warp); 0 = unknown                      */
  int32_t explicit_self_loops;    /* != 0: self loops are stored as ordinary edges of the view
                                     (row shards); kernels then add no implicit self loop      */
  int32_t sparse_filter;          /* hint for active_rows: != 0 when only a few percent of the rows are marked (a
                                     training batch) - a warp then scans one 32-row bitmap word instead of one row,
                                     so the launch is ~nrows/256 CTAs that mostly have work, not nrows/8 that exit */
  /* optional per-call filters (NULL = none), bitmaps with bit (id & 31) of word (id >> 5):
     active_rows - only LOCAL rows whose bit is set are computed and written, the rest are left untouched;
     active_cols - edges whose gathered node id (col) has a clear bit are skipped (their table rows are
                   known to be zero, e.g. the gradient rows of nodes outside the batch).
     peagnn_spmm_filtered() fills them in and peagnn_spmm() ignores them; the GAT entry points honour what the
     view carries: active_rows in peagnn_gat_rowmax / _aggregate / _backward_dst (rows outside are not computed;
     their outputs must be pre-zeroed by the caller), active_cols in peagnn_gat_backward_src. */
  const uint32_t* active_rows;
  const uint32_t* active_cols;
} peagnn_csr_t;

/* Floats of `partial` workspace an aggregation of width F (per head) needs on this view. */
size_t peagnn_partial_floats(int32_t n_chunks, int32_t feat, int32_t heads);

/* Per-step sub-structure of a CSR (demand-driven steps): keeps only the edges whose gathered node (col) is marked in
 * active_cols, in their original order.  rowptr_out[nrows + 1]; col_out / perm_out have room for nnz entries (only the
 * first rowptr_out[nrows] are written); perm_out[k] = perm[e] of the kept edge e (perm may be NULL: e itself; perm_out
 * may be NULL).  The view must carry nnz and row_offset = 0.  One flag pass + one CUB scan + one compaction pass over
 * the index array; every metapath that ends with the same relation then walks the small structure instead of all edges. */
size_t peagnn_csr_filter_workspace_bytes(int64_t num_edges);
int peagnn_csr_filter(const peagnn_csr_t* g, const uint32_t* active_cols, const int32_t* perm, int32_t* rowptr_out,
                      int32_t* col_out, int32_t* perm_out, void* workspace, size_t workspace_bytes,
                      peagnn_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * K1/K2: weighted CSR aggregation (GCNConv / SAGEConv message passing, and their transposes).
 *   out[i,:] = rs[i] * ( sum_{e in row i} cs[col_e] * X[col_e,:]  +  self_loop * cs[i] * X[i,:] )
 *              (+ bias) (relu) ; out += previous contents if accumulate.
 * rs / cs may be NULL (= 1).  GCN (models/peagcn.py:16-21 -> GCNConv): rs = cs = deg^-1/2,
 * self_loop = 1.  SAGE mean (models/peasage.py:16-21 -> SAGEConv): rs = 1/max(deg_in,1), cs = NULL,
 * self_loop = 0; its transpose swaps rs and cs.  Replaces index_select + mul + scatter_add of
 * torch_scatter.
 * ---------------------------------------------------------------------------------------- */
int peagnn_spmm(const peagnn_csr_t* g, const float* X, int64_t ldx, int32_t feat,
                float* out, int64_t ldo, const float* rs, const float* cs, int self_loop,
                const float* bias, int relu, int accumulate, peagnn_stream_t stream);

/* Demand-driven form of peagnn_spmm (SURVEY.md section 7 "dead rows": GraphRecsysModel.loss reads
 * only the representation rows of the batch's users and items, models/base.py:209-210).
 * active_rows != NULL: compute / write only the rows whose bit is set (forward of a last step);
 * active_cols != NULL: skip the edges that gather a node whose bit is clear (its transpose: the upstream
 * gradient is zero outside the batch rows).  Same arithmetic and the same in-row edge order as
 * peagnn_spmm on what remains; no atomics.  Bitmaps: uint32 words, bit (id & 31) of word (id >> 5),
 * indexed by local row id (active_rows) / gathered node id (active_cols). */
int peagnn_spmm_filtered(const peagnn_csr_t* g, const float* X, int64_t ldx, int32_t feat,
                         float* out, int64_t ldo, const float* rs, const float* cs, int self_loop,
                         const float* bias, int relu, int accumulate, const uint32_t* active_rows,
                         const uint32_t* active_cols, peagnn_stream_t stream);
/* north_star (2): GCN aggregation of a 64-wide table with the channel's first projection fused into the epilogue
 * (models/base.py:137-139 with models/peagcn.py:16-21: relu(GCNConv(x)) = relu((A_hat x) W + b)):
 *   out[i,:] = rs[i] * ( sum_e cs[col_e] X[col_e,:] + self_loop * cs[i] X[i,:] )            (the aggregate, still written)
 *   H_q[i,:] = act( out[i,:] @ W_q + b_q ),  q < n_proj <= 2,  W_q [64, 64] row-major (GCNConv.weight), b_q may be NULL
 * while the aggregated row is in registers; fp32 FMA.  active_rows as in peagnn_spmm_filtered (NULL = every row). */
int peagnn_spmm_proj(const peagnn_csr_t* g, const float* X, int64_t ldx, float* out, int64_t ldo, const float* rs,
                     const float* cs, int self_loop, const uint32_t* active_rows, int32_t n_proj,
                     const float* W0, const float* b0, float* H0, const float* W1, const float* b1, float* H1,
                     int64_t ldh, int relu, peagnn_stream_t stream);

/* Opt-in bf16 storage of a GATHERED table (north_star: "64-dim bf16/fp32 node rows"): peagnn_to_bf16 rounds an fp32
 * table to bf16 (nearest even; uint16 bit patterns, ldo in elements, multiple of 8), peagnn_spmm_bf16 is peagnn_spmm /
 * peagnn_spmm_filtered reading that table - 16-byte loads of 8 elements, four edges per warp instruction, half the L2
 * bytes; scales, accumulation, bias and the output are fp32.  feat = 64 only (the first-step tables).  The result
 * differs from the fp32 path by the rounding of the gathered values only (2^-9 relative per element; stated
 * tolerances in tests/test_gpu_bf16.py). */
int peagnn_to_bf16(const float* X, int64_t ldx, int64_t n, int32_t feat, uint16_t* out, int64_t ldo,
                   peagnn_stream_t stream);
int peagnn_spmm_bf16(const peagnn_csr_t* g, const uint16_t* Xb, int64_t ldx, int32_t feat, float* out, int64_t ldo,
                     const float* rs, const float* cs, int self_loop, const float* bias, int relu, int accumulate,
                     const uint32_t* active_rows, const uint32_t* active_cols, peagnn_stream_t stream);
/* bitmap[(id / mod) >> 5] |= 1 << ((id / mod) & 31) for every id with id % mod == rem (mod <= 1: every id,
 * bit index = id).  bitmap is pre-zeroed by the caller; ids int64 [n].  mod / rem select and renumber the
 * rows a rank owns under cyclic row sharding. */
int peagnn_mark_rows(const int64_t* ids, int64_t n, int32_t mod, int32_t rem, uint32_t* bitmap,
                     peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3: GAT edge-softmax aggregation (models/peagat.py:16-21 -> GATConv).
 * H is [N, heads*feat]; a_i, a_j are [N, heads] (= <H_n, att_i>, <H_n, att_j>).
 * The edge multiset of target i is its CSR row (self-loop edges already dropped) plus one
 * self loop.  e = leaky_relu(a_i[i] + a_j[j], slope); alpha = exp(e - max) / (sum + 1e-16).
 * ---------------------------------------------------------------------------------------- */
/* rowmax[i,h] = max over the edge multiset of e. */
int peagnn_gat_rowmax(const peagnn_csr_t* g, const float* a_i, const float* a_j, int32_t heads,
                      float slope, float* rowmax, peagnn_stream_t stream);
/* out[i, h*feat:(h+1)*feat] = sum alpha * H[j, h-slice] (+ bias)(relu); denom[i,h] = sum + 1e-16. */
int peagnn_gat_aggregate(const peagnn_csr_t* g, const float* H, int64_t ldh, int32_t feat,
                         int32_t heads, const float* a_i, const float* a_j, float slope,
                         const float* rowmax, float* denom, float* out, int64_t ldo,
                         const float* bias, int relu, peagnn_stream_t stream);
/* Backward, destination side.  dout is the gradient of the aggregate BEFORE bias/relu;
 * `agg` - `agg_bias` (agg_bias may be NULL) is that aggregate; where a relu clamped the output
 * dout is 0, so the forward output can be passed as `agg` with the conv bias as `agg_bias`.  Writes the per-edge
 * pairs ads_e[e, h] = (alpha, ds) (CSR order, [nnz, heads, 2] floats, 8-byte aligned: one coalesced line per 32 edges
 * here, one 8-byte read per edge on the source side), the self-loop terms alpha_self / ds_self [N, heads] and
 * d a_i [N, heads], where ds = d loss / d (a_i[i] + a_j[j]).  Row filter (g->active_rows) only. */
int peagnn_gat_backward_dst(const peagnn_csr_t* g, const float* H, int64_t ldh, int32_t feat,
                            int32_t heads, const float* a_i, const float* a_j, float slope,
                            const float* rowmax, const float* denom, const float* agg,
                            int64_t lda, const float* agg_bias, const float* dout, int64_t ldd,
                            float* ads_e, float* alpha_self, float* ds_self, float* d_ai,
                            peagnn_stream_t stream);
/* Backward, source side, over the TRANSPOSED structure gt (rows = sources, col = targets);
 * perm[k] = position in the destination-ordered per-edge array ads_e of gt's k-th edge.
 * dH[j, h-slice] = sum alpha * dout[i, h-slice] (+ self); d a_j[j,h] = sum ds (+ self). */
int peagnn_gat_backward_src(const peagnn_csr_t* gt, const int32_t* perm, const float* ads_e,
                            const float* alpha_self, const float* ds_self,
                            const float* dout, int64_t ldd, int32_t feat, int32_t heads,
                            float* dH, int64_t ldh, float* d_aj, peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4: node-wise projections (the matmul / nn.Linear inside each conv; fc1/fc2 are in K6).
 *   Y[n,:] = act( X[n,:] @ W (+ bias) ) (+ Y if accumulate);  W is [K, M] row-major, or
 *   [M, K] row-major (nn.Linear layout) when w_is_out_in != 0.   K, M multiples of 4, <= 128.
 *   accumulate adds the previous contents of Y before the activation.
 * If mask != NULL the input is gated on load: X[n,k] * (mask[n,k] > 0); if out_mask != NULL the
 * OUTPUT is gated on store: Y[n,m] * (out_mask[n,m] > 0) - both are relu-backward fusions.
 * ---------------------------------------------------------------------------------------- */
int peagnn_linear(const float* X, int64_t ldx, const float* mask, int64_t ldm, int64_t num_rows,
                  int32_t K, int32_t M, const float* W, int w_is_out_in, const float* bias,
                  int relu, int accumulate, float* Y, int64_t ldy, const float* out_mask,
                  int64_t ldom, peagnn_stream_t stream);

/* One projection problem of a grouped launch (same meaning as the arguments of peagnn_linear; no input gate). */
typedef struct {
  const float* X;
  int64_t ldx;
  int64_t n;                 /* rows; 0 = nothing to do */
  const float* W;
  const float* bias;         /* may be NULL */
  float* Y;
  int64_t ldy;
  const float* out_mask;     /* may be NULL */
  int64_t ldom;
} peagnn_linear_problem_t;
#define PEAGNN_MAX_GROUP 32
/* `count` (<= PEAGNN_MAX_GROUP) independent projections of ONE shape (K, M, weight layout, relu, accumulate) as a single
 * launch: the per-metapath projections of a step (models/base.py:137-139 runs them one conv at a time) are many small
 * problems - 13 x 12 k rows - whose launches each pay a prologue (weight split into shared memory, TMEM allocation) and
 * a tail; grouped, every CTA still works on one problem, but all of them are resident together.  Outputs must not
 * overlap between problems.  Shapes without a grouped kernel run as `count` single launches. */
int peagnn_linear_grouped(const peagnn_linear_problem_t* problems, int32_t count, int32_t K, int32_t M,
                          int w_is_out_in, int relu, int accumulate, peagnn_stream_t stream);

/* Floats of workspace for peagnn_linear_wgrad. */
size_t peagnn_wgrad_workspace_floats(int64_t num_rows, int32_t K, int32_t M);
/* dW = X^T @ (dY * (mask > 0)) stored as [K, M] (or [M, K] if w_is_out_in), db[m] = column sums
 * of the gated dY (db may be NULL).  Deterministic two-stage reduction through `workspace`.
 * K may be 0 (bias gradient only). */
int peagnn_linear_wgrad(const float* X, int64_t ldx, const float* dY, int64_t ldd,
                        const float* mask, int64_t ldm, int64_t num_rows, int32_t K, int32_t M,
                        int w_is_out_in, float* dW, float* db, float* workspace,
                        size_t workspace_floats, peagnn_stream_t stream);

/* One weight-gradient problem of a grouped launch (arguments of peagnn_linear_wgrad; no gate). */
typedef struct {
  const float* X;
  int64_t ldx;
  const float* dY;
  int64_t ldd;
  int64_t n;                 /* rows; 0: dW / db are set to zero */
  float* dW;                 /* may be NULL */
  float* db;                 /* may be NULL */
} peagnn_wgrad_problem_t;
/* `count` (<= PEAGNN_MAX_GROUP per launch; longer lists are cut) weight gradients of ONE shape as two launches
 * (partial sums per CTA, then one fold per problem in CTA order: deterministic).  Workspace >=
 * peagnn_wgrad_grouped_workspace_floats(count, K, M) floats. */
size_t peagnn_wgrad_grouped_workspace_floats(int32_t count, int32_t K, int32_t M);
int peagnn_linear_wgrad_grouped(const peagnn_wgrad_problem_t* problems, int32_t count, int32_t K, int32_t M,
                                int w_is_out_in, float* workspace, size_t workspace_floats,
                                peagnn_stream_t stream);

/* y = dy * (act > 0), elementwise over a [num_rows, feat] block (relu backward). */
int peagnn_relu_backward(const float* dy, int64_t ldd, const float* act, int64_t lda,
                         int64_t num_rows, int32_t feat, float* out, int64_t ldo,
                         peagnn_stream_t stream);

/* GAT attention logits: a_i[n,h] = <H[n,h,:], att_i[h,:]>, a_j likewise (GATConv message()). */
int peagnn_gat_scores(const float* H, int64_t ldh, int64_t num_rows, int32_t feat, int32_t heads,
                      const float* att_i, const float* att_j, float* a_i, float* a_j,
                      peagnn_stream_t stream);
/* Backward of the above: dH[n,h,:] (+)= d_ai[n,h]*att_i[h,:] + d_aj[n,h]*att_j[h,:];
 * d_att_i[h,:] = sum_n d_ai[n,h] * H[n,h,:] (two-stage, workspace >= peagnn_wgrad_workspace_floats
 * (num_rows, heads*feat, 4)). */
int peagnn_gat_scores_backward(const float* H, int64_t ldh, int64_t num_rows, int32_t feat,
                               int32_t heads, const float* att_i, const float* att_j,
                               const float* d_ai, const float* d_aj, float* dH, int64_t ldd,
                               int accumulate, float* d_att_i, float* d_att_j, float* workspace,
                               size_t workspace_floats, peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5: fusion across metapaths (models/base.py:191-206).
 * Z is [N, P, D] (row stride ldz, path stride D).  mode 0 = 'att': w = softmax_p(<Z[n,p],att[p]>),
 * out[n] = sum_p w_p Z[n,p];  mode 1 = 'mean'.  skip_path >= 0 treats that channel as zeros
 * (metapath ablation, base.py:194-195).  P <= 32.
 * ---------------------------------------------------------------------------------------- */
int peagnn_fuse_forward(const float* Z, int64_t ldz, int64_t num_rows, int32_t P, int32_t D,
                        const float* att, int mode, int skip_path, float* out, int64_t ldo,
                        peagnn_stream_t stream);
size_t peagnn_fuse_workspace_floats(int64_t num_rows, int32_t P, int32_t D);
int peagnn_fuse_backward(const float* Z, int64_t ldz, int64_t num_rows, int32_t P, int32_t D,
                         const float* att, int mode, const float* dout, int64_t ldo, float* dZ,
                         int64_t lddz, float* d_att, float* workspace, size_t workspace_floats,
                         peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6: pair scoring + BPR loss (models/base.py:43-48 and 208-214), entity-aware regulariser
 * (models/base.py:50-76).  batch is int64 [B, batch_cols] row-major, cols as in
 * datasets/movielens.py:1179: [u, pos_i, neg_i, e+_i, e-_i, mask_i, e+_u, e-_u, mask_u].
 * ---------------------------------------------------------------------------------------- */
/* scores[b] = fc2(relu(fc1([repr[u_b] || repr[i_b]])));  fc1_w [D, 2D], fc2_w [1, D]. */
int peagnn_predict(const float* repr, int64_t ldr, int32_t D, const int64_t* unids,
                   const int64_t* inids, int64_t B, const float* fc1_w, const float* fc1_b,
                   const float* fc2_w, const float* fc2_b, float* scores, peagnn_stream_t stream);

size_t peagnn_bpr_workspace_floats(int64_t B, int32_t D);
/* loss[0] = -sum_b log sigmoid(score(u,pos) - score(u,neg)), each term evaluated as softplus(-z)
 * (equal to the reference's sigmoid().log() wherever that is finite, and finite beyond).
 * If need_grad: the row gradients are ADDED into d_repr ([N, D], pre-zeroed by the caller) in a
 * fixed order - stable sort of (node id, slot) + one writer per node, no float atomics - and the fc
 * gradients are written (deterministic two-stage).  Node ids must fit int32.  All gradients are
 * for d loss = 1; two calls on the same inputs return bit-identical results. */
int peagnn_bpr_loss(const float* repr, int64_t ldr, int32_t D, const int64_t* batch,
                    int32_t batch_cols, int64_t B, const float* fc1_w, const float* fc1_b,
                    const float* fc2_w, const float* fc2_b, float* loss, int need_grad,
                    float* d_repr, int64_t lddr, float* d_fc1_w, float* d_fc1_b, float* d_fc2_w,
                    float* d_fc2_b, float* workspace, size_t workspace_floats,
                    peagnn_stream_t stream);

size_t peagnn_entity_workspace_floats(int64_t B, int32_t emb, int need_grad);
/* loss[0] += coff * ( -sum log sigmoid(mask_i (|x_pos-x_e+|^2 - |x_pos-x_e-|^2))
 *                     -sum log sigmoid(mask_u (|x_u  -x_e+|^2 - |x_u  -x_e-|^2)) );
 * if need_grad, coff * gradient is ADDED into dx ([N, emb], pre-zeroed by the caller) in a fixed
 * order (same sort-based scatter as peagnn_bpr_loss: bit-reproducible, no atomics). */
int peagnn_entity_reg(const float* x, int64_t ldx, int32_t emb, const int64_t* batch, int64_t B,
                      float coff, float* loss, int need_grad, float* dx, int64_t lddx,
                      float* workspace, size_t workspace_floats, peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K7: evaluation scorer + ranker (solvers.py:33-104, utils/rec_utils.py:7-30).
 * cand is int64 [U, C]: the first n_pos columns are the held-out positives, the rest the
 * sampled negatives (solvers.py:21-31).  Per user: scores, descending stable order
 * (positives first among ties), hit vector, HR@5..20, the reference's "NDCG"@5..20
 * (hits-in-top-K / log2(first-hit-position + 2)), AUC (strict >), BPR eval loss.
 * per_user is fp64 [U, 36]: HR[16] | NDCG[16] | AUC | loss | first-hit rank | reserved
 * (the reference accumulates these per-user rows in fp64 numpy arrays).  C <= 1024, D <= 32.
 * ---------------------------------------------------------------------------------------- */
int peagnn_eval_rank(const float* repr, int64_t ldr, int32_t D, const int64_t* users,
                     const int64_t* cand, int64_t U, int32_t C, int32_t n_pos,
                     const float* fc1_w, const float* fc1_b, const float* fc2_w,
                     const float* fc2_b, double* per_user, float* scores_out /* [U,C] or NULL */,
                     peagnn_stream_t stream);

/* out[c] = mean over rows of A[r, c] in a fixed order (deterministic, fp64 like the reference's
 * np.mean over per-user rows, solvers.py:104); cols <= 64; workspace >= 148 * cols doubles. */
int peagnn_column_mean(const double* A, int64_t lda, int64_t num_rows, int32_t cols, double* out,
                       double* workspace, size_t workspace_doubles, peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8: BPR training rows assembled on the device (replaces the host loops of
 * datasets/movielens.py:920-940 (cf_negative_sampling, BPR branch) and :1153-1177 (entity-aware columns
 * of __getitem__) when train_args['device_sampling'] is set).
 * Row r of the epoch's unshuffled [E * num_neg, cols] table - interaction r / num_neg of user2item
 * ([2, E] int64, row 0 = user nid, row 1 = item nid) plus its sampled columns - is a pure function of
 * (seed, epoch, r): Philox4x32-10 keyed by seed, counter (r, epoch, draw).  out is int64 [B, cols],
 * cols = 3: [u, pos, neg];  cols = 9: + [e+_i, e-_i, mask_i, e+_u, e-_u, mask_u].
 * strategy 0 ("random"): neg uniform over [item_lo, item_lo + num_items);
 * strategy 1 ("unseen"): neg uniform over the items not in the user's train set; seen_ptr [U + 1] /
 *   seen_items are a CSR over users (u - user_lo) of their train item nids, ascending and unique.
 * ifeat_* / ufeat_*: CSR over items / users of their feature node ids; type_starts [num_types + 1]:
 *   ascending first node id of every node type, then num_nodes.  Unused tables may be NULL.
 * ---------------------------------------------------------------------------------------- */
int peagnn_bpr_rows(const int64_t* row_ids, int64_t B, const int64_t* u2i, int64_t E, int32_t num_neg,
                    uint64_t seed, uint64_t epoch, int32_t strategy, int64_t user_lo, int64_t item_lo,
                    int64_t num_items, const int64_t* seen_ptr, const int64_t* seen_items, int32_t cols,
                    const int64_t* ifeat_ptr, const int64_t* ifeat_nids, const int64_t* ufeat_ptr,
                    const int64_t* ufeat_nids, const int64_t* type_starts, int32_t num_types,
                    int64_t* out, peagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Measurement probe (bench.py / tools/l2_gather_probe.py; no product arithmetic): sums the F-float
 * rows table[idx[k]] for a coalesced int32 id stream with the aggregation kernels' access pattern
 * (128-bit loads, F/4 lanes per row) and nothing else - the measured random-row-gather ceiling the
 * aggregation is reported against.  out needs peagnn_probe_out_floats() floats.  feat in {16,32,64,128}.
 * ---------------------------------------------------------------------------------------- */
size_t peagnn_probe_out_floats(void);
int peagnn_probe_gather(const float* table, int64_t ld, int32_t feat, const int32_t* idx, int64_t n_idx,
                        float* out, peagnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PEAGNN_H_ */
