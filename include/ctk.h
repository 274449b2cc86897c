/* ctk.h - C ABI of libctk.so: hand-written sm_100a kernels for the CT-CLIP training hot path.
 *
 * Drop-in boundary: the reference is pure PyTorch, so the "FFI" a maintainer
 * binds is ctypes from Python. Every entry point takes plain device
 * pointers, sizes and a cudaStream_t (as void*); nothing here knows about torch.
 *
 * Conventions
 *   - return 0 (CTK_OK) or a negative ctk_status; ctk_last_error() gives a thread-local message.
 *   - all buffers (outputs and workspaces) are allocated by the caller; the library never
 *     allocates or frees device memory and keeps no mutable global state.
 *   - stream ordered, no internal synchronisation, re-entrant.
 *   - no CPU path and no other architecture: a device that is not compute capability 10.x
 *     yields CTK_ERR_ARCH.
 *   - "tokens" are rows of a row-major [rows, dim] matrix; fp32 residual stream, bf16 GEMM
 *     operands, fp32 accumulation everywhere.
 *
 * Reference files cited below are relative to the upstream repository root
 * (transformer_maskgit/transformer_maskgit/{attention,ctvit}.py, CT_CLIP/ct_clip/{ct_clip,distributed}.py).
 */
#ifndef CTK_H_
#define CTK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CTK_OK = 0,
    CTK_ERR_SHAPE = -1,
    CTK_ERR_ALIGN = -2,
    CTK_ERR_ARCH = -3,
    CTK_ERR_CUDA = -4
} ctk_status;

const char* ctk_last_error(void);
int ctk_version(void);
/* CTK_OK when the current CUDA device is sm_100-class, CTK_ERR_ARCH otherwise. */
int ctk_device_ok(void);
/* number of kernels this library has launched in the process so far (diagnostic counter only). */
unsigned long long ctk_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense contraction core (tcgen05 + TMEM + TMA).  D = A[M,K] * B[N,K]^T, bf16 in, fp32 accumulate.
 * Replaces every nn.Linear on the path: attention.py:52-57 (FeedForward), :123-129 (to_q,
 * to_kv, to_out), ctvit.py:173 (patch projection), and the matmuls autograd derives from them.
 * *_mn_major = 0: operand is [rows, K] with K contiguous; 1: operand is [K, rows] (rows
 * contiguous) - the layout of activations in a weight-gradient product.  Instantiated: (0, 0) every epilogue;
 * (1, 1) BF16 / F32 / RESID_F32 / GEGLU / GEGLU_BWD / QKV / ATOMIC_F32 / ARGMAX; (a = 0, b = 1) - the input-gradient
 * product dX = dY W with the weight W [K = out, N = in] as stored - BF16 / RESID_F32 / GEGLU_BWD / GELU_BWD.
 * ------------------------------------------------------------------------------------------ */
typedef enum {
    CTK_EPI_BF16 = 0,      /* C bf16 = alpha * (acc + bias)                                        */
    CTK_EPI_F32 = 1,       /* C fp32 = acc + bias                                                  */
    CTK_EPI_RESID_F32 = 2, /* C fp32 = acc + bias + resid   (resid may alias C)  attention.py:445,450 */
    CTK_EPI_GEGLU = 3,     /* C = U bf16 [M,N] pre-activations, aux0 = H bf16 [M,N/2] = gelu(gate)*value;
                              weights interleaved per 128 hidden units by ctk_pack_ff_w1 attention.py:45-48 */
    CTK_EPI_GEGLU_BWD = 4, /* acc = dH; aux0 = U; C = dU (same interleaved layout)                 */
    CTK_EPI_QKV = 5,       /* q / kv projections into the packed bf16 [M, 3*heads*32] buffer C at column
                              offset i1: the first i0 columns of this GEMM are per-head (32 wide) l2-normalised
                              (eps 1e-12) and scaled by alpha*vec0[d]; the rest pass through (v).
                              aux0 = fp32 [M, ld_aux0] reciprocal norms at head index (col+i1)/32.
                              attention.py:145-160 (q from LayerNorm(x), k/v from raw x) */
    CTK_EPI_ATOMIC_F32 = 6,/* C fp32 += alpha * acc via atomics (split-K); row_map permutes output rows */
    CTK_EPI_ARGMAX = 7,    /* C = uint64 [M]: atomicMax of (orderable(acc)<<32 | ~col)  (VQ code search)   */
    CTK_EPI_GELU = 8,      /* text tower (HF BertIntermediate, ct_clip.py:1271): C = U bf16 [M,N] = acc + bias
                              (pre-activations, kept for the backward pass), aux0 = bf16 [M,N] = gelu(U), erf form */
    CTK_EPI_GELU_BWD = 9,  /* acc = dG; aux0 = U bf16 [M,N]; C = dU bf16 = dG * gelu'(U)                      */
    CTK_EPI_LSE_PART = 10, /* contrastive logits x = exp(*vec1) * acc (ct_clip.py:1347), never stored: C = fp32
                              [ceil(N/128), M, 3] online softmax statistics (max, sum e^(x-max), sum x e^(x-max)) of
                              every row over each block of 128 columns; aux0 (fp32 [M], may be NULL) receives the
                              logit with row + i0 == col (the positives, ct_clip.py:1355-1358)                 */
    CTK_EPI_CLIP_GRAD = 11,/* dloss/dlogit of the symmetric InfoNCE, scaled for the latent
                              gradients: g = alpha * exp(*vec1) * (e^(x - vec0[row]) + e^(x - bias[col])
                              - 2 [row + i0 == col + i1]); C = bf16 hi part [M,N], aux0 = bf16 lo part (g - hi) */
    CTK_EPI_ARGMAX_PART = 12 /* exact VQ code search, pass 1 (ctvit.py:403): for every row and every block of 128
                              columns, C = uint64 [ceil(N/128), M] = (orderable(best acc) << 32 | ~col) and aux0 = fp32
                              [ceil(N/128), M] = the block's SECOND best value; ctk_vq_select re-scores in fp32 whatever
                              lies within the bf16 rounding bound of the row maximum, so the chosen code is the fp32 one */
} ctk_epilogue;

typedef struct {
    void* C;
    long long ldc;
    const float* bias;     /* [N] or NULL */
    const float* resid;
    long long ldr;
    void* aux0;
    long long ld_aux0;
    const float* vec0;     /* [32] */
    const float* vec1;     /* [32]; LSE_PART / CLIP_GRAD: pointer to the log of the logit scale (device scalar) */
    const int* row_map;    /* [M] or NULL */
    float alpha;
    int i0;                /* QKV: number of leading columns to normalise */
    int i1;                /* QKV: output column offset */
} ctk_gemm_epilogue_t;

int ctk_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                  int b_mn_major, int M, int N, int K, int epilogue,
                  const ctk_gemm_epilogue_t* epi, int split_k /* 0 = auto, ATOMIC only */,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight preparation (fp32 master parameters -> bf16 tensor-core operands), once per step.
 * ------------------------------------------------------------------------------------------ */
/* dst bf16 [rows, ld_dst] = src fp32 [rows, cols] * (col_scale ? col_scale[col] : 1); pad columns zeroed. */
int ctk_cast_bf16(const float* src, void* dst, long long rows, long long cols, long long ld_dst,
                  const float* col_scale, void* stream);
/* dst bf16 [cols, ld_dst] = src fp32 [rows, cols]^T, pad columns (>= rows) zeroed. */
int ctk_transpose_cast_bf16(const float* src, void* dst, long long rows, long long cols,
                            long long ld_dst, void* stream);
/* same, but ONLY the columns [0, rows) of every dst row are written: dst may be a column slice of a wider matrix
 * (e.g. [Wq^T | Wk^T | Wv^T] packed side by side).  The shipped modules no longer need transposed weight copies - their
 * input-gradient products read the [out, in] operand MN-major (b_mn_major = 1 below) - the entry stays for callers that do. */
int ctk_transpose_cast_bf16_slice(const float* src, void* dst, long long rows, long long cols,
                                  long long ld_dst, void* stream);
/* FeedForward W1 (2*inner, dim) fp32 -> interleaved bf16 (2*inner_pad, dim): for every block of
 * 128 hidden units, 128 value rows then 128 gate rows (attention.py:47: first half value, second
 * half gate). Also writes row_map[2*inner_pad] (interleaved row -> source row, -1 for padding)
 * and, if dst_t != NULL, the transposed bf16 copy (dim, 2*inner_pad) for the input gradient. */
int ctk_pack_ff_w1(const float* w1, void* dst, void* dst_t, int* row_map, int inner, int inner_pad,
                   int dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * CTViT.to_patch_emb, first half (ctvit.py:170-172): tubelet gather
 * 'b c (t pt)(h p1)(w p2) -> b t h w (c pt p1 p2)' + LayerNorm statistics over the patch.
 * video fp32 [B,1,D,H,W]; xhat bf16 [B*T*Hp*Wp, ld] = (x - mean) * rstd  (affine folded into the
 * projection weights by the caller); mean, rstd fp32 per patch (eps 1e-5).
 * ------------------------------------------------------------------------------------------ */
int ctk_patch_norm_fwd(const float* video, void* xhat, long long ld, float* mean, float* rstd,
                       int B, int D, int H, int W, int pt, int p1, int p2, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (<= 1024, multiple of 64) of an fp32 [rows, dim] matrix.
 * attention.py:34-41 (gamma + zero beta buffer), :51 nn.LayerNorm, ctvit.py:174.
 * out_bf16 / out_f32 may be NULL. xraw_bf16 (may be NULL) receives the un-normalised input cast to
 * bf16 in the input row order: the k/v projections read the raw stream (attention.py:145-149). If perm_inner > 0 the output row index is transposed:
 * row = (g*perm_outer + o)*perm_inner + i  ->  (g*perm_inner + i)*perm_outer + o, which is the
 * '(b t)(h w) d -> (b h w) t d' rearrangement of ctvit.py:301 (and its inverse :305).
 * ------------------------------------------------------------------------------------------ */
int ctk_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out_bf16,
                      float* out_f32, void* xraw_bf16, float* mean, float* rstd, long long rows,
                      int dim, float eps, int perm_outer, int perm_inner, void* stream);
/* dx (fp32) = LN backward of dy; dy is bf16 (dy_bf16) or fp32 (dy_f32), indexed through the same
 * row permutation as the forward output. dy_bcast_rows > 0: dy has rows/dy_bcast_rows rows and
 * row r reads dy[r / dy_bcast_rows] * dy_scale (gradient of a mean-pool broadcast back).
 * dx_accum != 0: dx += result. dx_bf16 (may be NULL) receives a bf16 copy of the final dx (the next
 * GEMM's A operand). dgamma/dbeta (fp32 [dim], pre-zeroed) accumulate atomically; dbeta may be NULL. */
int ctk_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                      const float* mean, const float* rstd, float* dx, void* dx_bf16, int dx_accum,
                      float* dgamma, float* dbeta, long long rows, int dim, int perm_outer,
                      int perm_inner, long long dy_bcast_rows, float dy_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * PEG (attention.py:62-90) + residual (attention.py:443): depthwise 3x3x3 conv over the token
 * grid [B, n0, n1, n2, dim] *as laid out in memory* with causal (2,0) padding on axis 0 and
 * (1,1) on axes 1,2.  y = conv(x) + bias + x.  w fp32 [dim,27], b fp32 [dim].
 * ------------------------------------------------------------------------------------------ */
int ctk_peg_fwd(const float* x, const float* w, const float* b, float* y, int B, int n0, int n1,
                int n2, int dim, void* stream);
/* dx = conv^T(dy) + dy (dx_bf16: optional bf16 copy); dw [dim,27], db [dim] accumulate atomically
 * (pre-zeroed by caller). */
int ctk_peg_bwd(const float* dy, const float* x, const float* w, float* dx, void* dx_bf16, float* dw,
                float* db, int B, int n0, int n1, int n2, int dim, void* stream);

/* ------------------------------------------------------------------------------------------
 * ContinuousPositionBias (attention.py:335-382) on the (2*gh-1)*(2*gw-1) distinct offsets
 * instead of (gh*gw)^2 pairs.  table fp32 [heads, 2gh-1, 2gw-1];  bias[h,i,j] =
 * table[h, yi-yj+gh-1, xi-xj+gw-1].  h0/h1 fp32 [(2gh-1)(2gw-1), dim] keep the MLP activations.
 * ------------------------------------------------------------------------------------------ */
int ctk_cpb_fwd(const float* w0, const float* b0, const float* w1, const float* b1,
                const float* w2, const float* b2, float* h0, float* h1, float* table, int gh,
                int gw, int dim, int heads, void* stream);
/* all gradient outputs are overwritten (not accumulated); ws fp32 [2 * n_off * dim] scratch. */
int ctk_cpb_bwd(const float* dtable, const float* w0, const float* w1, const float* w2,
                const float* h0, const float* h1, float* dw0, float* db0, float* dw1, float* db1,
                float* dw2, float* db2, float* ws, int gh, int gw, int dim, int heads,
                void* stream);

/* ------------------------------------------------------------------------------------------
 * Cosine attention core (attention.py:162-184) on the packed qkv buffer written by
 * CTK_EPI_QKV: bf16 [nseq*L, 3*heads*32], q pre-scaled.  softmax(q.k^T + bias) v, no mask,
 * dim_head 32.  out bf16 [nseq*L, heads*32]; lse fp32 [nseq, heads, L] (natural log).
 * table == NULL -> no bias (temporal stack).  L == gh*gw when a table is given.
 * Dispatch (same results, different kernels): 24x24-token slices with a table -> tcgen05/TMEM kernels
 * (attention_tc.cu, attention_tc_bwd.cu); 24-token sequences without a table, 2/4/8 heads -> TMA ring +
 * warp MMA (attention_seq24.cu); anything else -> mma.sync / SIMT kernels (attention.cu).
 * Environment switches read once per process: CTK_ATTN_LEGACY=1 forces the last group, CTK_DBIAS_TC=0
 * selects the mma.sync bias-table-gradient kernel instead of the tcgen05 one.
 * ------------------------------------------------------------------------------------------ */
int ctk_attn_fwd(const void* qkv, const float* table, void* out, float* lse, int nseq, int L,
                 int heads, int gh, int gw, void* stream);
/* dqkv bf16 [nseq*L, 3*heads*32] gets d(q_scaled), d(k_scaled), dv;  dtable fp32 (pre-zeroed,
 * NULL when no bias) accumulates the bias-table gradient;  delta fp32 [nseq,heads,L] scratch. */
int ctk_attn_bwd(const void* qkv, const float* table, const void* out, const void* dout,
                 const float* lse, float* delta, void* dqkv, float* dtable, int nseq, int L,
                 int heads, int gh, int gw, void* stream);
/* Backward of the CTK_EPI_QKV epilogue: turns d(q_scaled), d(k_scaled) into gradients of the raw
 * projections (in place in dqkv) and accumulates dq_scale, dk_scale (fp32 [32], pre-zeroed). */
int ctk_qknorm_bwd(void* dqkv, const void* qkv, const float* rnorm, const float* q_scale,
                   const float* k_scale, float alpha, float* dq_scale, float* dk_scale,
                   long long rows, int heads, void* stream);

/* ------------------------------------------------------------------------------------------
 * VectorQuantize(use_cosine_sim=True) (ctvit.py:188,403; vector-quantize-pytorch==1.1.2).
 * ------------------------------------------------------------------------------------------ */
/* rows of x fp32 [rows, dim] -> l2-normalised bf16 (and fp32 if xn_f32 != NULL). */
int ctk_l2norm_rows(const float* x, void* xn_bf16, float* xn_f32, long long rows, int dim,
                    void* stream);
/* best = uint64 [rows] filled by CTK_EPI_ARGMAX (pre-zeroed). ind int64 [rows];
 * quant fp32 [rows, dim] = embed[ind]. */
/* Exact code selection (pass 2 of the VQ search).  part_key / part_second: output of CTK_EPI_ARGMAX_PART over
 * nblk = ceil(C/128) column blocks (similarities of bf16-rounded unit vectors).  Every code whose bf16 similarity lies
 * within `margin` of the row maximum (margin >= 2^-7 bounds the rounding of two unit vectors) is re-scored as the fp32
 * dot product of x/max(|x|,1e-12) with en (= l2-normalised codebook, fp32 [C, dim]); the arg-max of those (lowest index
 * on ties) goes to ind[rows] (int64) and quant[rows, dim] = embed[ind] (raw fp32 codebook rows). dim % 32 == 0, <= 1024. */
int ctk_vq_select(const void* part_key, const float* part_second, int nblk, const float* x, const float* en,
                  const float* embed, long long* ind, float* quant, long long rows, int dim,
                  int codebook_size, float margin, void* stream);
int ctk_vq_gather(const void* best, const float* embed, long long* ind, float* quant,
                  long long rows, int dim, int codebook_size, void* stream);
/* training-mode EMA update of cluster_size [C] and embed [C, dim] (decay 0.8);
 * ws fp32 [C*dim + C] scratch (zeroed inside). */
int ctk_vq_ema_update(const float* xn_f32, const long long* ind, float* cluster_size, float* embed,
                      float* ws, long long rows, int dim, int codebook_size, float decay,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Contrastive head (ct_clip.py:1280-1388, distributed.py:9-20).
 * ------------------------------------------------------------------------------------------ */
/* pooled fp32 [B, dim] = mean over n tokens of x fp32 [B, n, dim]  (ct_clip.py:1297, applied
 * before the bias-free projection, which commutes with the mean).  Also the pooling of CTCLIP.forward_old
 * (ct_clip.py:1549,1566: mean over the frame axis, n = t = 24, dim = h*w*C = 294 912 - the flatten is the layout);
 * dim % 4 == 0, x 16-byte aligned. */
int ctk_mean_pool_fwd(const float* x, float* pooled, int B, long long n, int dim, void* stream);
/* latent fp32 [B, dl] = l2norm(x[B, din] . W[dl, din]^T)  (ct_clip.py:1290,1313-1316, eps 1e-12);
 * x rows are x_stride floats apart (CLS rows of the text encoder output). rnorm fp32 [B].
 * din > 1024 - to_visual_latent of the original checkpoints, Linear(294912, 512) on forward_old's pooling
 * (ct_clip.py:1614) - additionally needs din % 4 == 0, x_stride % 4 == 0 and 16-byte aligned x, W; every output is
 * summed in a fixed order (bitwise repeatable). */
int ctk_latent_fwd(const float* x, long long x_stride, const float* W, float* latent, float* rnorm,
                   int B, int din, int dl, void* stream);
/* dW fp32 [dl, din] and dx fp32 [B, din] (row pitch dx_stride) are overwritten; either may be NULL. */
int ctk_latent_bwd(const float* dlatent, const float* latent, const float* rnorm, const float* x,
                   long long x_stride, const float* W, float* dW, float* dx, long long dx_stride,
                   int B, int din, int dl, void* stream);
/* Symmetric InfoNCE on gathered latents T, I fp32 [N, d] (rank-major), logit scale exp(log_temp):
 *   loss = ClipLoss(S) / b_local ; also dT, dI for rows [row0, row0+b_local) only (AllGather
 *   backward = local slice, distributed.py:18-20) and d(log_temp) summed over the full matrix.
 * out fp32 [2] = {loss, dlog_temp}; d_local fp32 [2, b_local, d] = {dT_local, dI_local}.
 * ws: ctk_clip_loss_ws_bytes(N, b_local). The full N x N logits never touch HBM; only the rank's
 * two [b_local, N] gradient stripes are staged in ws. */
size_t ctk_clip_loss_ws_bytes(int N, int b_local);
int ctk_clip_loss_fwd_bwd(const float* T, const float* I, const float* log_temp, float* out,
                          float* d_local, void* ws, size_t ws_bytes, int N, int d, int b_local,
                          int row0, void* stream);
/* Zero-shot scoring (ct_clip.py:842-855): out[p] = exp(log_temp) * <text_lat[p], image_lat>. */
int ctk_pair_logits(const float* text_lat, const float* image_lat, const float* log_temp,
                    float* out, int P, int d, void* stream);

/* misc */
int ctk_fill_f32(float* p, float v, long long n, void* stream);
/* y fp32 [rows] (+)= ... helpers used by the patch-embed backward:
 * dW[n,k] = gamma[k] * P[n,k] + beta[k] * db[n];  dgamma[k] = sum_n W[n,k] P[n,k];
 * dbeta[k] = sum_n W[n,k] db[n]   (LayerNorm(4000) affine folded into the projection). */
int ctk_patch_affine_bwd(const float* P, const float* W, const float* gamma, const float* beta,
                         const float* db, float* dW, float* dgamma, float* dbeta, int n, int k,
                         void* stream);
/* db fp32 [cols] += column sums of dy (bf16 or fp32) [rows, cols]. */
int ctk_colsum(const void* dy_bf16, const float* dy_f32, float* out, long long rows, int cols,
               void* stream);

/* ------------------------------------------------------------------------------------------
 * Text tower: the HF BertModel the reference passes as text_encoder, ct_clip.py:1271.
 * ------------------------------------------------------------------------------------------ */
/* BertEmbeddings before its LayerNorm: out fp32 [M, H] = word[ids[m]] + typ[token_type ? token_type[m] : 0] + pos[m % L].
 * ids / token_type int64 [M] (token_type may be NULL), M = B * L. */
int ctk_bert_embed_fwd(const long long* ids, const long long* token_type, const float* word, const float* pos,
                       const float* typ, float* out, long long M, int L, int H, void* stream);
/* backward: dword[ids[m]] += de[m] (zero-initialised by the caller; rows equal to pad_idx are skipped, pad_idx < 0: none),
 * dpos[l] = sum_b de[b*L + l] for l < L (written, not accumulated), dtyp[token_type[m]] += de[m] when token_type != NULL
 * (otherwise the caller takes ctk_colsum of de into dtyp[0]). */
int ctk_bert_embed_bwd(const float* de, const long long* ids, const long long* token_type, float* dword, float* dpos,
                       float* dtyp, long long M, int L, int H, long long pad_idx, void* stream);
/* Softmax attention core on packed bf16 projections, head dim 64 (HF BertSelfAttention of the text tower) or 32
 * (`FlashAttention` of CTViT3D, transformer_maskgit/attention.py:189-284): softmax(q k^T * scale + key mask) ->
 * dropout(p_drop) -> . v on qkv [B*L, 3*heads*dh] (q | k | v, heads contiguous).  key_mask uint8 [B, L] (non-zero = attend)
 * or NULL.  null_k / null_v bf16 [heads, n_null, dh] (n_null <= 64, may be 0 / NULL): learned null key/value pairs every
 * query of every sequence also attends to (attention.py:240-248), processed as one extra key block.
 * out bf16 [B*L, heads*dh]; lse fp32 [B, heads, L] (natural log of the row sums of exp(scaled logits)).
 * Dropout: element (b*heads + h, i, j) is kept iff hash(*seed_ptr + seed_off * c, ...) >= p * 2^32 (csrc/mha_dropout.cuh);
 * the seed is read from DEVICE memory so that CUDA-graph replays see a fresh value; the backward regenerates the mask.
 * mma.sync tensor-core kernels, flash-style (the L x L probabilities never leave registers), deterministic. */
int ctk_mha_fwd(const void* qkv, const unsigned char* key_mask, const void* null_k, const void* null_v, int n_null,
                void* out, float* lse, int B, int L, int heads, int dh, float scale, float p_drop,
                const unsigned long long* seed_ptr, unsigned long long seed_off, void* stream);
/* delta fp32 [B, heads, L] is workspace; dqkv bf16 [B*L, 3*heads*dh] receives dq | dk | dv; dnull_k / dnull_v fp32
 * [B, heads, n_null, dh] receive the null pairs' gradients per sequence (the caller sums over B). */
int ctk_mha_bwd(const void* qkv, const unsigned char* key_mask, const void* null_k, const void* null_v, int n_null,
                const void* out, const void* dout, const float* lse, float* delta, void* dqkv, float* dnull_k,
                float* dnull_v, int B, int L, int heads, int dh, float scale, float p_drop,
                const unsigned long long* seed_ptr, unsigned long long seed_off, void* stream);

/* ------------------------------------------------------------------------------------------
 * Loader-side volume preparation: scripts/data.py:49-111 npz_to_tensor on the device.
 * src = the stored array arr_0, (D, H, W) row-major, float32 (src_is_f16 = 0) or float16 (1), in device or
 * pinned host memory; dst fp32 (Dt, Ht, Wt) = (240, 480, 480) in the reference:
 *   (clip(x, -1, 1) + 1) / 2 in the stored dtype, centre crop, centre pad with -1.  Bit-exact with the reference.
 * ------------------------------------------------------------------------------------------ */
int ctk_volume_prep(const void* src, int src_is_f16, int D, int H, int W, float* dst, int Dt, int Ht, int Wt,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer tail: torch.nn.utils.clip_grad_norm_(params, max_norm)
 * (CTCLIPTrainer.py:711-712) + torch.optim.Adam / AdamW (optimizer.py:14-24) over a device table of
 * tensors.  rows: int64 [ntensors][6] = {param ptr, grad ptr, exp_avg ptr, exp_avg_sq ptr, numel,
 * index of the tensor's first chunk}; a chunk is ctk_opt_chunk_elems() elements, nchunks their total.
 * ctk_multi_sqnorm overwrites *out_sq with the sum of squares of all gradients; ctk_multi_adam scales
 * the gradients by min(1, max_norm / (sqrt(*sq) + 1e-6)) (max_norm <= 0: no clipping) and applies
 * step number `step` (>= 1) of Adam (adamw = 0: weight decay added to the gradient) or AdamW.
 * ------------------------------------------------------------------------------------------ */
int ctk_opt_chunk_elems(void);
/* dst (device) <- src (pinned, device-accessible host memory) by kernel loads, not by a DMA copy: the
 * table upload must not queue behind the input batch's H2D transfer.  bytes % 16 == 0. */
int ctk_copy_from_pinned(void* dst, const void* src_pinned, long long bytes, void* stream);
int ctk_multi_sqnorm(const void* rows, int ntensors, long long nchunks, float* out_sq, void* stream);
int ctk_multi_adam(const void* rows, int ntensors, long long nchunks, const float* sq, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int adamw, long long step, float max_norm,
                   int write_clipped_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTK_H_ */
