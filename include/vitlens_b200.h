/* vitlens_b200 -- C ABI of the B200 (sm_100a) kernels behind the ViT-Lens hot path.
 *
 * The reference (TencentARC/ViT-Lens) is 100% Python/PyTorch: its hot path has no FFI of its
 * own; every FLOP runs in PyTorch library kernels.  This header is therefore
 * the seam a maintainer would bind from Python (ctypes); each entry point
 * names the reference call(s) it replaces (paths relative to vitlens/src/open_clip/).
 *
 * Conventions: plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 * a name ends in _host.  `stream` is a cudaStream_t passed as void*.  bf16 = __nv_bfloat16 bits.
 * Every function returns 0 on success, a negative VL_E* code on argument errors, or a positive
 * cudaError_t; vl_last_error() returns a static description of the last failure on this thread.
 * Matrices are row-major with an explicit leading dimension (elements).
 */
#ifndef VITLENS_B200_H
#define VITLENS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL_ABI_VERSION 2
#define VL_EINVAL (-1)   /* bad argument (shape / alignment / null pointer) */
#define VL_ENOTSUP (-2)  /* shape outside what the kernel supports */
#define VL_EDRIVER (-3)  /* CUDA driver entry point (tensor-map encode) unavailable */

int vl_abi_version(void);
const char* vl_last_error(void);
/* debug/bring-up knobs (descriptor overrides etc.); key/value are kernel-specific, 0 resets. */
int vl_debug_set(int key, int value);
/* bring-up: device buffer (>= 64 KiB) that instrumented kernels fill with clock64 timelines; NULL disables. */
int vl_debug_buffer(void* dev_ptr);

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA, bf16 x bf16 -> fp32 accumulate):
 *     D[M,N] = epilogue(alpha * sum_k A(m,k) * B(n,k))
 * A is [M,K] row-major (a_mn=0, "K-major") or given transposed as [K,M] row-major (a_mn=1);
 * B is [N,K] row-major (b_mn=0; an nn.Linear weight [out,in]) or [K,N] row-major (b_mn=1).
 * Replaces: every nn.Linear / F.linear / `@` on the path -- in_proj & out_proj of
 * nn.MultiheadAttention (transformer.py:215,252), mlp.c_fc / c_proj (transformer.py:226-234),
 * Lens to_q/to_kv/to_out and FeedForward (perceiver.py:85-154), conv1-as-GEMM (transformer.py:464-470),
 * `pooled @ self.proj` (transformer.py:786-787), and their autograd backward (dgrad / wgrad).
 */
enum {
  VL_EPI_LINEAR = 0,   /* d = alpha*acc + bias                                                   */
  VL_EPI_GELU = 1,     /* u = alpha*acc + bias; aux_out(bf16) = u if given; d = act(u)           */
  VL_EPI_RESIDUAL = 2, /* d = alpha*acc + bias + aux_in                                          */
  VL_EPI_GELU_BWD = 3, /* d = alpha*acc * act'(aux_in)     (dgrad of c_proj fused with GELU bwd) */
  VL_EPI_GEGLU = 4,    /* Lens FeedForward (perceiver.py:85-102): B = the [2F, K] weight with rows PERMUTED so that rows
                          [n*256, n*256+128) are value rows n*128.. and rows [n*256+128, (n+1)*256) the gate rows
                          F + n*128.. (bias permuted alike); u = alpha*acc + bias; aux_out[M, 2F] = u in the ORIGINAL
                          column order (value | gate, needed by the backward); d[M, F] = value * gelu(gate).
                          M >= 512, N = 2F with N % 256 == 0 (VL_ENOTSUP otherwise: use vl_geglu_fwd)          */
  /* Contrastive-loss epilogues (replace `logit_scale * x @ y.T` + F.cross_entropy, loss.py:116-163,
   * 346-385) -- the [rows x cols] logits never reach HBM:                                         */
  VL_EPI_ROWLSE = 5,   /* z = alpha*acc. Per row and per 128/64-column part p: out_vec0[row*nparts+p] = max z,
                          out_vec1[row*nparts+p] = sum exp(z - max); out_vec2[row] = z[row, row+iparam].
                          nparts = ceil(N/BN)*2 is returned through vl_gemm_rowlse_parts(). d unused.     */
  VL_EPI_CLIPGRAD = 6, /* z = alpha*acc; g = exp(z - row_vec[i]) + (col_vec ? exp(z - col_vec[j]) : 0)
                          - (col_vec ? 2 : 1) * [j == i + iparam];  d(bf16) = fparam * g;
                          *scalar_out = sum fparam * g * acc   (d loss / d alpha; written, deterministic);
                          with loss_flags bit 0 the sum runs over the row-softmax term only:
                          fparam * (exp(z - row_vec[i]) - [j == i + iparam]) * acc  (the d/d alpha of THIS rank's rows
                          when the column term carries other ranks' losses: local_loss + gather_with_grad)  */
};

typedef struct {
  const void* a; /* bf16 */
  const void* b; /* bf16 */
  void* d;       /* bf16, or fp32 when d_f32 */
  int32_t M, N, K;
  int64_t lda, ldb, ldd; /* elements */
  int32_t a_mn, b_mn;
  int32_t d_f32;
  int32_t accumulate; /* fp32 output only: D += result (atomic; required when split_k > 1) */
  int32_t split_k;    /* >= 1 */
  int32_t epilogue;   /* VL_EPI_* */
  int32_t act_quick;  /* activation of VL_EPI_GELU / VL_EPI_GELU_BWD: 0 = erf GELU (nn.GELU), 1 = QuickGELU
                         (transformer.py:37-40), 2 = ReLU (grouped PointNet, dvae.py:200-208) */
  float alpha;
  const float* bias;  /* fp32 [N] or NULL */
  const void* aux_in; /* bf16 [M,N] (ld = ldaux) or NULL */
  void* aux_out;      /* bf16 [M,N] (ld = ldaux) or NULL */
  int64_t ldaux;
  /* loss epilogues only */
  const float* row_vec; /* fp32 [M] */
  const float* col_vec; /* fp32 [N] or NULL */
  float* out_vec0;
  float* out_vec1;
  float* out_vec2;
  float* scalar_out;
  int32_t iparam;
  float fparam;
  /* optional device scalars (no host sync): effective alpha = alpha * (*alpha_dev), fparam = fparam * (*fparam_dev) */
  const float* alpha_dev;
  const float* fparam_dev;
  /* LINEAR / RESIDUAL extras (grouped PointNet, dvae.py:196-212): aux_in row = output row / aux_row_div (0 or 1 = same row:
   * the per-group "global feature" term is broadcast over the group's points); relu != 0 clamps the result at 0. */
  int32_t aux_row_div;
  int32_t relu;
  /* Optional fp32 [M]: rowsum_out[m] = sum_k A[m, k] (written; per-tile partials added in a fixed order), computed on the tensor
   * cores from the A tiles already in shared memory (an extra N = 16 MMA against a tile of ones).  With A = dY^T this is the bias gradient of the Linear whose
   * weight gradient the GEMM produces (replaces a separate vl_colsum_bf16 pass over dY).  Requires the CTA-pair kernel:
   * M >= 512, N > 128, LINEAR epilogue, fp32 output; rejected (VL_ENOTSUP) otherwise. */
  float* rowsum_out;
  int32_t loss_flags; /* VL_EPI_CLIPGRAD only, see above */
  /* B sharded by rows over the ranks of the box (the all-gathered features of the contrastive loss, loss.py:55-76): when
   * b_peers != NULL, B's global row r lives at b_peers[r / b_peer_rows] + (r % b_peer_rows) * ldb -- device pointers into the
   * peers' arenas (vl_comm_peer_ptr) -- and `b` is ignored.  The kernel loads every tile from its owner over NVLink; if
   * peer_flags != NULL (device int32 [b_npeers] in THIS rank's arena) the first touch of peer q waits for
   * peer_flags[q] >= peer_flag_value.  b_peer_rows % 256 == 0 (B K-major) or % 64 == 0 (b_mn) when b_npeers > 1. */
  const void* const* b_peers; /* HOST array of b_npeers device pointers */
  int32_t b_npeers, b_peer_rows;
  const int32_t* peer_flags;
  int32_t peer_flag_value;
  /* loss epilogues: optional byte mask [M, N] (row stride ldmask); where it is 0 the logit is replaced by the constant 0 and
   * carries no gradient -- `logits * mask` of ClipLossSimMask / ClipLossLabelMask / TriClipLossLabelMask (loss.py:485-903) */
  const uint8_t* mask;
  int64_t ldmask;
} VlGemmArgs;

int vl_gemm_bf16(const VlGemmArgs* args, void* stream);
/* number of column parts VL_EPI_ROWLSE writes per row for an N-column problem */
int vl_gemm_rowlse_parts(int32_t N);
/* lse[i] = log sum_p out_vec1[i,p]*exp(out_vec0[i,p] - max_p) + max_p;  *loss_sum = sum_i (lse[i] - diag[i]) (written) */
int vl_lse_combine(const float* part_max, const float* part_sum, const float* diag, int32_t M, int32_t nparts,
                   float* lse, float* loss_sum, void* stream);

/* Fused backward of one direction of the gathered InfoNCE loss (reference: gather_features + ClipLoss / TriClipLoss backward,
 * open_clip/loss.py:55-76, 116-138, 158-163; `logits * mask` of the mask variants loss.py:485-903):
 *   dx[i, :] = s * sum_j g[i, j] * y[j, :],   g = gscale * (exp(z - row_lse_i) + [col_lse] exp(z - col_lse_j) - k * onehot(j == i + label_off)),
 *   z = s * x y^T (masked entries: z = 0, g = 0);  *ds_out = sum g * (x y^T)  (with ds_row_only: the row term of g alone).
 * One launch: logits and g stay in TMEM; y is either one bf16 matrix [N, E] or y_npeers row blocks of y_peer_rows rows living in
 * the ranks' peer arenas (read in place over NVLink -- the all-gather fused into the kernel).  E % 64 == 0. */
typedef struct {
  const void* x;            /* bf16 [M, E], row stride ldx */
  const void* y;            /* bf16 [N, E], row stride ldy (or null with y_peers) */
  const void* const* y_peers;
  int32_t y_npeers, y_peer_rows;
  int32_t M, N, E;
  int64_t ldx, ldy;
  float* dx;                /* fp32 [M, E], row stride lddx */
  int64_t lddx;
  const float* row_lse;     /* [M] */
  const float* col_lse;     /* [N] or null */
  int32_t label_off;
  const float* alpha_dev;   /* device scalar: logit scale s */
  float gscale;
  const float* gscale_dev;  /* optional device scalar multiplied into gscale (upstream gradient) */
  float* ds_out;            /* device scalar (written) or null */
  int32_t ds_row_only;
  const uint8_t* mask;      /* [M, N] bytes, row stride ldmask, or null */
  int64_t ldmask;
} VlClipBwdArgs;
int vl_clip_backward(const VlClipBwdArgs* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Symmetric peer memory over NVLink / NVSwitch for the data-parallel step (reference gather_features
 * loss.py:20-78 and the DDP gradient all-reduce pc_tri_main.py:378-380).  One arena per rank (cudaMalloc here, CUDA IPC), mapped
 * into every rank of the box.  The first 2 KiB of each arena are int32 flags[64][8]: flags[idx][src] is written by rank src and
 * polled by the owner; values are monotonically increasing tickets ("ready" = flag >= ticket).
 *   vl_comm_init      allocate + zero this rank's arena, return its 64-byte IPC handle (exchange them out of band, e.g. with
 *                     torch.distributed.all_gather_object) and its address
 *   vl_comm_connect   all_handles = world x 64 bytes in rank order: map the peers' arenas
 *   vl_comm_peer_ptr  base address of rank `peer`'s arena in this process
 *   vl_comm_signal    after all prior work on `stream`: flags[flag_idx][my rank] = value in every arena (system-scope release)
 *   vl_comm_wait      `stream` waits until flags[flag_idx][p] >= value for every p (traps after ~20 s instead of hanging)
 *   vl_comm_peer_reduce_f32 / _gather_f32   out[i] = sum_p arena_p[offset + 4 i]  (rank order, deterministic) / out[p*n + i] = ...
 *   vl_allgather_features   publish the packed bf16 feature block at `offset` of the own arena (no copy: vl_gemm_bf16 with
 *                     b_peers reads the blocks in place)
 *   vl_allreduce_grads      copy-engine push of [offset, offset + bytes) of the own arena to the same range of every peer's
 *                     arena + ticket; vl_adamw_multi_src sums the world copies in rank order
 */
int vl_comm_init(int32_t rank, int32_t world, int64_t arena_bytes, void* handle_out, void** arena_out);
int vl_comm_connect(const void* all_handles);
int vl_comm_peer_ptr(int32_t peer, void** ptr_out);
int vl_comm_destroy(void);
int vl_comm_signal(int32_t flag_idx, int32_t value, void* stream);
int vl_comm_wait(int32_t flag_idx, int32_t value, void* stream);
int vl_comm_peer_reduce_f32(int64_t offset, int64_t n, float* out, void* stream);
int vl_comm_peer_gather_f32(int64_t offset, int64_t n, float* out, void* stream);
int vl_allgather_features(int64_t offset, int64_t bytes, int32_t flag_idx, int32_t ticket, void* stream);
int vl_allreduce_grads(int64_t offset, int64_t bytes, int32_t flag_idx, int32_t ticket, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head attention core, head_dim = 64 (ViT-L/14, ViT-B/32, CLIP text, Lens self/cross).
 * Element (b, t, h, d) of q/k/v/o lives at base + (b*n + t)*ld + h*64 + d (bf16), so the packed
 * in_proj output [T, 3D] is consumed in place (k = qkv + D, v = qkv + 2D, ld = 3D).
 * softmax(scale * q k^T [+ causal mask]) v, fp32 softmax statistics; lse[b,h,t] (natural log) saved.
 * Replaces: F.scaled_dot_product_attention inside nn.MultiheadAttention (transformer.py:241-252,
 * additive causal attn_mask transformer.py:870-876) and the einsum-softmax-einsum of the Lens
 * (perceiver.py:127-145), plus their autograd backward.
 */
int vl_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t H,
                     int32_t nq, int32_t nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale,
                     int32_t causal, void* stream);
int vl_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                     void* dq, void* dk, void* dv, int32_t B, int32_t H, int32_t nq, int32_t nk, int64_t ldq,
                     int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv,
                     float scale, int32_t causal, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row kernels (HBM-bound, one pass).  x/y/dy/dx are bf16 unless noted; parameters and their
 * gradients are fp32.  Reductions over rows (parameter gradients, column sums, moments, loss sums, gradient norms) are
 * DETERMINISTIC: every CTA stores one row of partial sums into stream-ordered scratch memory and a second launch adds the
 * rows in CTA order, so results are bit-identical from run to run and the outputs are WRITTEN (=), never accumulated; they need
 * no zero fill.  The one exception is vl_embed_tokens_bwd (a scatter by token id: fp32 atomics, outputs accumulated).
 */
/* LayerNorm eps affine over the last dim (transformer.py:17-34, perceiver.py:71-72); optional row
 * gather: y[i] = LN(x[row_index[i]]) (cls pooling transformer.py:653-657,783; EOT pooling model.py:537-540). */
int vl_layernorm_fwd(const void* x, int64_t ldx, const int64_t* row_index, const float* w, const float* b, void* y,
                     int64_t ldy, float* mean, float* rstd, int32_t T, int32_t D, float eps, void* stream);
/* dx[row_index[i] or i] = LN'(dy[i]) (+ dres at the same row); dw = sum dy*xhat; db = sum dy (both or neither);
 * dres_sum[d] = sum_i dres[i, d] (optional: the bias gradient of the Linear that produced the residual branch). */
int vl_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const int64_t* row_index, const float* w,
                     const float* mean, const float* rstd, const void* dres, int64_t lddres, void* dx, int64_t lddx,
                     float* dw, float* db, float* dres_sum, int32_t T, int32_t D, void* stream);
/* db[n] = sum_t dy[t,n]: bias gradients of every nn.Linear on the path. */
int vl_colsum_bf16(const void* dy, int64_t ld, float* db, int32_t T, int32_t N, void* stream);
/* Patch gather feeding conv1-as-GEMM (nn.Conv2d bias=False: transformer.py:464-470, AST_tokenizer.py:22-28,
 * DepthTokenizer.py:22-28): out[(b*OH+oh)*OW+ow, (c*kh+i)*kw+j] = in[b*sb + c*sc + (oh*stride_h+i)*sh + (ow*stride_w+j)*sw],
 * zero-padded to Kpad columns; `in` is fp32 or bf16 with arbitrary element strides (the audio
 * unsqueeze/transpose of AST_tokenizer.py:46-47 is just a stride choice). */
int vl_patchify(const void* in, int32_t in_is_bf16, void* out, int32_t B, int32_t C, int32_t OH, int32_t OW, int32_t kh,
                int32_t kw, int32_t stride_h, int32_t stride_w, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                int32_t Kpad, void* stream);
/* out[b, c+l] = tok[b,l] + pos[c+l]; out[b,0] = cls + pos[0] when has_cls (transformer.py:743,756-768). */
int vl_assemble_tokens(const void* tok, const float* cls, const float* pos, void* out, int32_t B, int32_t L, int32_t D,
                       int32_t has_cls, void* stream);
int vl_assemble_tokens_bwd(const void* dx, void* dtok, float* dpos, float* dcls, int32_t B, int32_t L, int32_t D,
                           int32_t has_cls, void* stream);
/* out[r] = table[ids[r]] + pos[r % ctx] (model.py:530-532) and its backward (dtable / dpos ACCUMULATED with atomics: zero them). */
int vl_embed_tokens(const int64_t* ids, const float* table, const float* pos, void* out, int64_t rows, int32_t ctx,
                    int32_t D, void* stream);
int vl_embed_tokens_bwd(const int64_t* ids, const void* dx, float* dtable, float* dpos, int64_t rows, int32_t ctx,
                        int32_t D, void* stream);
/* F.normalize(dim=-1, eps=1e-12) on fp32 [B,E] rows (model.py:522,526,540) and backward. */
int vl_l2norm_fwd(const float* x, float* y, float* inv_norm, int32_t B, int32_t E, float eps, void* stream);
int vl_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int32_t B, int32_t E, void* stream);
/* GEGLU of the Lens FeedForward (perceiver.py:85-88): h[M,2F] = [val|gate] -> out[M,F] = val*gelu(gate). */
int vl_geglu_fwd(const void* h, void* out, int64_t M, int32_t F, void* stream);
int vl_geglu_bwd(const void* h, const void* dout, void* dh, int64_t M, int32_t F, void* stream);
int vl_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);
int vl_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);
/* Decoupled-weight-decay Adam step on one fp32 tensor (optim.AdamW, training/point_cloud/pc_tri_main.py:394-419). */
int vl_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int32_t step, float grad_scale, void* stream);
/* The same update for every parameter in ONE launch.  ptrs: [n_tensors][5] device pointers (p, g or 0, m, v, bf16 copy of p
 * or 0); sizes / wds: per tensor element count / weight decay; lrs: per tensor learning rate (optimizer param_groups; the
 * scheduler's assign_learning_rate, training/scheduler.py) or NULL = `lr` for all; chunk_tab: [n_chunks][2] int32 (tensor id,
 * chunk index), chunks of 16384 elements.  A tensor whose gradient pointer is 0 is skipped (torch.optim semantics for
 * grad None).  The bf16 copy (the tensor-core operand cache) is refreshed in the same pass.  n_src > 1: every gradient is the
 * sum, in source order, of n_src copies src_stride ELEMENTS apart (the ranks' slots of vl_allreduce_grads; the DDP reduction
 * of pc_tri_main.py:378-380 folded into the update) -- n_src = 1, src_stride = 0 for plain gradients. */
int vl_adamw_multi(const int64_t* ptrs, const int64_t* sizes, const float* wds, const float* lrs, const int32_t* chunk_tab, int32_t n_chunks,
                   float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale, int32_t n_src, int64_t src_stride, void* stream);
/* Gradient-norm clipping (torch.nn.utils.clip_grad_norm_, reference training/train.py:212-240) without an extra pass over the
 * gradients: vl_multi_sqnorm writes sum(g^2) over every tensor of the same tables to *sumsq (deterministic); vl_adamw_multi_clip is
 * vl_adamw_multi with every gradient scaled by min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6)). */
int vl_multi_sqnorm(const int64_t* ptrs, const int64_t* sizes, const int32_t* chunk_tab, int32_t n_chunks, float* sumsq, int32_t n_src,
                    int64_t src_stride, void* stream);
int vl_adamw_multi_clip(const int64_t* ptrs, const int64_t* sizes, const float* wds, const float* lrs, const int32_t* chunk_tab,
                        int32_t n_chunks, float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale, const float* sumsq,
                        float max_norm, int32_t n_src, int64_t src_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Point-cloud tokenizer (reference modal_3d/models/pointbert): farthest point sampling with explicit start indices
 * (misc.py:48-68; the reference draws them with torch.randint at :60), kNN grouping minus centre (dvae.py:107-176),
 * K=3 linear + folded BatchNorm + ReLU / GELU (dvae.py:200-203, point_encoder.py:325-327), per-group max (dvae.py:205,210).
 */
int vl_fps(const float* xyz, const int64_t* start, int32_t B, int32_t N, int32_t npoint, int64_t* idx_out, float* centers,
           void* stream);
int vl_knn_group(const float* xyz, const float* centers, int32_t B, int32_t N, int32_t G, int32_t k, float* nb_out,
                 int64_t* idx_out, void* stream);
int vl_linear3(const float* x, const float* w, const float* scale, const float* shift, void* out, void* pre_out, int64_t R,
               int32_t C, int32_t act, void* stream);
int vl_group_max(const void* x, void* out, int32_t* arg, int64_t groups, int32_t G, int32_t C, void* stream);
/* backward pieces of the tokenizer: max scatter, per-group sums, (sum a, sum a*b) column sums, 3-input weight gradient */
int vl_group_max_bwd(const void* dout, const int32_t* arg, void* dx, int64_t groups, int32_t G, int32_t C, void* stream);
int vl_group_sum(const void* x, void* out, int64_t groups, int32_t G, int32_t C, void* stream);
int vl_colsum2_bf16(const void* a, const void* b, float* s1, float* s2, int64_t T, int32_t N, void* stream);
int vl_wgrad3(const void* dy, const float* x, float* dw, int64_t R, int32_t C, void* stream);
/* BatchNorm1d with batch statistics inside the point tokenizer (training mode; reference dvae.py:185-193 under model.train(),
 * SyncBN per pc_tri_main.py:372-373).  vl_moments3: out12 = [sum x (3) | sum x x^T (3x3)] over the R rows of x[R,3] -- the
 * statistics of first_conv.0's outputs follow in closed form.  vl_col_affine_bf16: out = act(p0[c]*a + p1[c]*b + p2[c]) per
 * column (b, p1 optional; act 0 none / 1 relu): the normalise+ReLU pass and the BatchNorm backward correction. */
int vl_moments3(const float* x, float* out12, int64_t R, void* stream);
int vl_col_affine_bf16(const void* a, const void* b, const float* p0, const float* p1, const float* p2, void* out, int64_t R, int32_t C,
                       int32_t act, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Zero-shot evaluation on the device (reference training/zero_shot.py:36-60,155-257,572-789, open_clip/zero_shot_classifier.py:
 * 27-88, open_clip/metrics/{accuracy,map,recall}.py).  Similarities come from vl_gemm_bf16 (fp32 output); these finish the job.
 */
/* out[g] = normalize(mean_t normalize(x[g*T + t])) over the T templates of class g (fp32 rows of E <= 1024); transpose_out
 * writes the [E, G] classifier layout of build_zero_shot_classifier (ldo = row stride of out). */
int vl_template_mean(const float* x, float* out, int32_t G, int32_t T, int32_t E, int64_t ldo, int32_t transpose_out, void* stream);
/* Row-wise top-k (k <= 16) of fp32 scores[rows, cols]: idx_out[rows, k] int32 (descending score, ties -> smaller column, -1 when
 * cols < k), val_out optional.  torch.topk in accuracy()/acc()/Recall.retrieval_eval. */
int vl_topk_rows(const float* scores, int64_t ld, int32_t rows, int32_t cols, int32_t k, int32_t* idx_out, float* val_out, void* stream);
/* Per-class average precision over N samples (sklearn.metrics.average_precision_score(average=None), metrics/map.py:50), ties
 * share a threshold; apply_sigmoid mirrors map.py:36.  ap_out[C] (0 for a class without positives), npos_out[C] optional.
 * N <= 25600 (one class column lives in shared memory). */
int vl_average_precision(const float* scores, int64_t lds, const float* targets, int64_t ldt, int32_t N, int32_t C, int32_t apply_sigmoid,
                         float* ap_out, int32_t* npos_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input pipelines on the device.
 */
/* Kaldi-compatible log-mel filterbank + pad/crop + normalisation (reference modal_audio/processors/at_processor.py:845-872, i.e.
 * torchaudio.compliance.kaldi.fbank(htk_compat, hanning, 25/10 ms, dither 0, power, log) -> ZeroPad/crop to target_len ->
 * (x - mean) / std).  wav: fp32 [n_clips] clips of n_samples, clip_stride apart; window fp32 [frame_len]; mel fp32 [n_mel, 257]
 * (get_mel_banks' matrix padded with a zero column); out fp32 [n_clips, target_len, n_mel].  frame_len <= 512 (FFT size 512). */
int vl_fbank(const float* wav, int64_t clip_stride, int32_t n_clips, int32_t n_samples, int32_t frame_len, int32_t frame_shift, const float* window,
             const float* mel, int32_t n_mel, float preemph, int32_t target_len, float mean, float std, float* out, void* stream);
/* pc_norm (modal_3d/processors/pc_processor.py:32-38): per cloud, xyz -= centroid, xyz /= max distance; in/out fp32 [B, N, C >= 3]. */
int vl_pc_norm(const float* in, float* out, int32_t B, int32_t N, int32_t C, void* stream);
/* DepthNorm + Normalize of the depth (disparity) channel (modal_depth/processors/transforms_rgbd.py:393-413, vt_processor.py:
 * 311-322): out = (clamp(d, min_depth[, max_depth]) / max_depth - mean) / std. */
int vl_depth_norm(const float* in, float* out, int64_t n, float min_depth, float max_depth, int32_t clamp_max, float mean, float std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITLENS_B200_H */
