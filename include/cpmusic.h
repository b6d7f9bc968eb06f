/* cpmusic.h — C-ABI of libcpmusic.so: sm_100a kernels for the compound-word (CP)
 * causal-linear-attention agent (the hot path of
 * daniel05155/Reinforcement-Learning-in-Music-Generation).
 *
 * This is the drop-in boundary.  The reference reaches this
 * arithmetic through pytorch-fast-transformers 0.4.0's pybind11 extension
 * (`causal_dot_product(Q,K,V,product)` / `causal_dot_product_backward(...)`, fp32
 * (N,H,L,E) contiguous, outputs pre-allocated by Python) and through ATen ops
 * issued from the model files cited next to each entry point.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in `_host`;
 *   - the callee never allocates, never synchronises, never throws; the caller
 *     owns every buffer including workspaces;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - returns 0 on success or a negative CPM_ERR_* code; `cpm_last_error_string()`
 *     gives a thread-local human-readable message for the last failure;
 *   - `dtype` selects the storage type of activations (CPM_F32 | CPM_BF16);
 *     accumulation, normalisers, statistics and recurrent state are always fp32;
 *   - token indices are int64 (the reference passes `.long()` tensors).
 */
#ifndef CPMUSIC_H_
#define CPMUSIC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPM_VERSION 100 /* 0.1.0 */

#define CPM_F32 0
#define CPM_BF16 1

#define CPM_OK 0
#define CPM_ERR_BAD_SHAPE (-1)    /* unsupported / inconsistent dimension */
#define CPM_ERR_BAD_ALIGN (-2)    /* pointer or stride not 16-byte aligned */
#define CPM_ERR_BAD_DTYPE (-3)
#define CPM_ERR_NULL (-4)         /* required pointer is NULL */
#define CPM_ERR_WORKSPACE (-5)    /* workspace too small */
#define CPM_ERR_CUDA (-6)         /* launch failed; see cpm_last_error_string() */
#define CPM_ERR_UNSUPPORTED (-7)  /* e.g. tcgen05 path requested for fp32 */

#define CPM_MAX_ATTR 8 /* CP attributes per compound word (reference: 6; upstream CP: 7) */

int cpm_version(void);
const char *cpm_last_error_string(void);
const char *cpm_error_name(int code);
/* Name of the kernel implementation the last cpm_linattn_fwd/bwd call dispatched to
 * ("tcgen05-cp" | "tcgen05" | "simt"); lets tests assert which path ran. */
const char *cpm_linattn_last_impl(void);

/* ------------------------------------------------------------------------------------------
 * A1/A2 — causal linear attention, parallel (teacher-forced) form.
 * Replaces ft `CausalLinearAttention.forward` + `causal_product_cuda` fwd/bwd, i.e. the
 * encoder built at dqn_policy/model.py:128-137, agent_pretrain.py:244-253,
 * ppo_policy/model.py:129-138,313-321:
 *     Qf = elu(q)+1, Kf = elu(k)+1
 *     den[n,l,h] = Qf[n,l,h,:] . sum_{j<=l} Kf[n,j,h,:] + eps
 *     out[n,l,h,:] = ( sum_{j<=l} (Qf[n,l,h,:].Kf[n,j,h,:]) v[n,j,h,:] ) / den[n,l,h]
 * Layout: q,k,v are (N,L,H,E) with E contiguous, head h at element offset h*E and token
 * stride `ld_qkv` elements (so the three can be column slices of one fused QKV GEMM output);
 * out / gout are (N,L,H,M) with token stride `ld_o`; gq,gk,gv have token stride `ld_g`.
 * den is (N,L,H) fp32, written by fwd and read by bwd.  E = M = 64 (the reference), or 128 (see below).
 * impl: 0 = auto (chunk-parallel tcgen05 when dtype==BF16, else simt), 1 = simt,
 *       3 = tcgen05, chunk-parallel (one CTA per 128-token chunk; streaming state pre-pass, or per-chunk states + scan
 *           when there are fewer than 96 (batch, head) chains).
 * Workspace: cpm_linattn_workspace_bytes(N,L,H) bytes (segment / chunk states), scratch only.
 * saved (optional, impl 3): cpm_linattn_saved_bytes(N,L,H) bytes that fwd fills with the per-chunk
 * prefix states (bf16 tiles + fp32 key sums) and bwd reads back instead of rebuilding them — the
 * one tensor besides `den` the caller keeps for backward.  NULL = not kept (bwd recomputes).
 * ------------------------------------------------------------------------------------------ */
int64_t cpm_linattn_workspace_bytes(int N, int L, int H);
int64_t cpm_linattn_saved_bytes(int N, int L, int H);
/* The same two sizes for head width E (= M) of 64 or 128.  128-wide heads (cfg5 read as 8 heads x 128) run on the
 * chunk-parallel tensor-core kernels only - bf16, impl 0 or 3 - as templates of the 64-wide kernels over the
 * two feature halves of q / k and the two value halves of v / out: one score tile per chunk (K = 128), a 2 x 2 grid of 64 x 64
 * state tiles per chunk (saved: 4 bf16 tiles + 128 fp32 key sums per chunk and head).  Anything else returns 0 here and
 * CPM_ERR_UNSUPPORTED from cpm_linattn_fwd / bwd (the host side then runs two 64-wide passes, ops._linattn_fused_e128). */
int64_t cpm_linattn_workspace_bytes_wide(int N, int L, int H, int E);
int64_t cpm_linattn_saved_bytes_wide(int N, int L, int H, int E);
/* Development aid: device buffer that receives clock64() stamps of the chunk-parallel kernels at their phase boundaries -
 * 64 int64 per CTA of the forward per-chunk kernel, 128 int64 per CTA of the backward one (whichever runs while the buffer
 * is installed; size it for 128 x 3 x SMs); NULL switches it off (the default). */
int cpm_debug_linattn_timing(void *device_buffer);
int cpm_linattn_fwd(const void *q, const void *k, const void *v, void *out, float *den,
                    int N, int L, int H, int E, int M, int64_t ld_qkv, int64_t ld_o,
                    int dtype, float eps, int impl, void *workspace, int64_t workspace_bytes,
                    void *saved, int64_t saved_bytes, void *stream);
int cpm_linattn_bwd(const void *q, const void *k, const void *v, const void *out, const float *den,
                    const void *gout, void *gq, void *gk, void *gv,
                    int N, int L, int H, int E, int M, int64_t ld_qkv, int64_t ld_o, int64_t ld_g,
                    int dtype, float eps, int impl, void *workspace, int64_t workspace_bytes,
                    const void *saved, int64_t saved_bytes, void *stream);

/* B1 — recurrent one-token step.  Replaces ft `RecurrentLinearAttention.forward`
 * (builder at dqn_policy/model.py:141-150; call at dqn_policy/model.py:237,
 * testing-no-type-cp.py:154,167):
 *     Z += Kf ; S += Kf (x) v ; out = Qf^T S / (Qf.Z + eps)
 * q,k,v: (N,H,E) rows with row stride ld_qkv; out (N,H,M) row stride ld_o;
 * S (N,H,E,M) fp32 and Z (N,H,E) fp32 are updated IN PLACE (as ft does under no_grad).
 * E = M = 64 (the reference's heads) runs the streaming 256-bit kernel; other widths (E <= 256, M in {32, 64, 128}; e.g.
 * the 8 x 128 variant of BASELINE cfg5) a generic 128-bit one. */
int cpm_linattn_step(const void *q, const void *k, const void *v, float *S, float *Z, void *out,
                     int N, int H, int E, int M, int64_t ld_qkv, int64_t ld_o,
                     int dtype, float eps, void *stream);

/* ------------------------------------------------------------------------------------------
 * C1 — CP embedding gather + sqrt(emb) scale + concat.  Replaces `Embeddings.forward` x6 and
 * `torch.cat` (agent_pretrain.py:185-192,320-335; dqn_policy/model.py:206-221).
 *   idx (T,n_attr) int64; tables_host[a] -> fp32 table (n_tokens_host[a], emb_sizes_host[a]);
 *   out (T, sum(emb)) dtype.  Out-of-range indices write zeros and set *err_flag (if non-NULL).
 * bwd accumulates (+=) into fp32 gtables_host[a] (caller zeroes). emb sizes must be /8.
 * ------------------------------------------------------------------------------------------ */
int cpm_embed_fwd(const int64_t *idx, const float *const *tables_host, const int *n_tokens_host,
                  const int *emb_sizes_host, int n_attr, int64_t T, void *out, int dtype,
                  int *err_flag, void *stream);
int cpm_embed_bwd(const int64_t *idx, const void *gout, float *const *gtables_host,
                  const int *n_tokens_host, const int *emb_sizes_host, int n_attr, int64_t T,
                  int dtype, void *stream);

/* Positional encoding add + dropout: y[r,:] = drop(x[r,:] + pe[pos(r),:]).
 * Replaces `PositionalEncoding.forward` (agent_pretrain.py:208-210).  pe fp32 (max_len,d).
 * pos(r) = pos_offset + (r % L) when pos_dev==NULL, else pos_dev[0] + (r % L) (device-side
 * step counter for graph-captured rollouts; the reference's recurrent mode always adds
 * position 0 — SURVEY D8 — which is pos_offset=0, L=1). */
int cpm_add_pe(const void *x, const float *pe, void *y, int64_t rows, int L, int d,
               int pos_offset, const int32_t *pos_dev, int max_len, float p_drop, uint64_t seed,
               uint64_t rng_offset, int dtype, void *stream);

/* Inverted dropout with a counter-based (Philox4x32-10) mask: y = x*mask/(1-p); the mask
 * depends only on (seed, rng_offset, element index), so applying it to a gradient
 * reproduces the forward mask without storing it. */
int cpm_dropout(const void *x, void *y, int64_t n, float p_drop, uint64_t seed,
                uint64_t rng_offset, int dtype, void *stream);

/* C2 — fused residual + dropout + LayerNorm (affine, eps 1e-5 in the reference).  Replaces
 * `x = norm(x + dropout(y))` of ft's post-norm TransformerEncoderLayer and the final norm
 * (K6):
 *     s = x + drop(res)   (res may be NULL: plain LayerNorm)
 *     y = (s-mean)*rstd*gamma + beta
 * Saves s (dtype, may be NULL for inference), mean, rstd (fp32, may be NULL).  d % 8 == 0,
 * d <= 8192.  bwd: given gy, s, mean, rstd -> gs (grad wrt x), gres (grad wrt res through the
 * dropout mask; may alias NULL when p_drop==0: then gres==gs), and fp32 partial sums for
 * dgamma/dbeta(/dres_bias) in `partials` (cpm_ln_partials_rows() x 3 x d), reduced into dgamma/dbeta (+=).
 */
/* res_bias (optional, fp32 (d)): bias of the Linear that produced `res`, added before the dropout — lets that GEMM run
 * bias-less and makes its bias gradient a by-product of the backward kernel: dres_bias (optional) += column sums of the
 * residual-branch gradient.  partials: cpm_ln_partials_rows() x 3 x d floats. */
int cpm_ln_residual_fwd(const void *x, const void *res, const float *res_bias, const float *gamma, const float *beta,
                        void *y, void *s_out, float *mean, float *rstd, int64_t rows, int d,
                        float eps, float p_drop, uint64_t seed, uint64_t rng_offset, int dtype,
                        void *stream);
int cpm_ln_partials_rows(void);
int cpm_ln_residual_bwd(const void *gy, const void *s, const float *mean, const float *rstd,
                        const float *gamma, void *gs, void *gres, float *dgamma, float *dbeta, float *dres_bias,
                        float *partials, int64_t rows, int d, float p_drop, uint64_t seed,
                        uint64_t rng_offset, int dtype, void *stream);

/* Device-side dropout RNG base (optional; NULL = off, the default).  When set, every dropout-carrying kernel
 * (cpm_add_pe, cpm_dropout, cpm_ln_residual_fwd/bwd, cpm_gelu_fwd/bwd) adds *device_counter to its rng_offset
 * argument, so a CUDA graph that captured those launches (offsets frozen) still draws fresh masks on every replay
 * as long as the graph advances the counter.  Process-wide; set it before capture and keep the memory alive. */
/* Column sums out[c] = sum_r x[r][c] of a (rows, width) bf16 / fp32 matrix with row stride ld (elements): the bias gradient
 * of a Linear layer from its output gradient (what autograd's `grad_output.sum(0)` computes for nn.Linear,
 * agent_pretrain.py:239 / every ft projection).  partials: caller-owned fp32 scratch of cpm_colsum_partials_rows(width) * width
 * floats.  width must be a multiple of 16 (bf16) / 8 (fp32), ld of 8.  out (width) fp32 is OVERWRITTEN.  Deterministic. */
/* Refresh of the bf16 compute packings of a model's fp32 master parameters (encoder.PackCache: the row-concatenated weights every
 * Linear of agent_pretrain.py:239-253 / dqn_policy/model.py:156-161 runs on) in ONE launch.  items_device: device array of n_items
 * records of cpm_pack_item_bytes() bytes each, in this order (natural alignment, 8-byte pointers first):
 *   const float *w (rows x cols master, contiguous), const float *b (rows, or NULL),
 *   bf16 *wc (row-major destination, already offset to the master's first row), bf16 *wt (transposed destination offset to the
 *   master's first COLUMN, row stride ld_t; or NULL), bf16 *bc (or NULL), float *b32 (or NULL), int rows, cols, ld_t, tile0
 * with tile0 = number of 32 x 32 tiles of all earlier items; n_tiles = their total.  Round-to-nearest-even, as torch's copy_. */
int cpm_pack_item_bytes(void);
int cpm_pack_weights(const void *items_device, int n_items, int n_tiles, void *stream);
int cpm_colsum_partials_rows(int width);
int cpm_colsum(const void *x, int64_t rows, int width, int64_t ld, float *out, float *partials, int dtype, void *stream);

/* Row dot out[r] = h[r,:] . u + *c (fp32 out; c may be NULL): the critic's value read-out collapsed to one d-vector,
 * Critic_Transformer.value_produce (ppo_policy/model.py:345-394: six Linear(512, n_a) heads, six Linear(n_a, 1) value heads,
 * average over attributes) is linear in h.  Backward: dh[r,:] = g[r] u (may be NULL) and du[:] = sum_r g[r] h[r,:]
 * (OVERWRITTEN; partials = caller-owned fp32 scratch of cpm_rowdot_partials_rows() * d floats; deterministic). d % 8 == 0, <= 1024. */
int cpm_rowdot_partials_rows(void);
int cpm_rowdot_fwd(const void *h, const float *u, const float *c, float *out, int64_t rows, int d, int dtype, void *stream);
int cpm_rowdot_bwd(const void *h, const float *g, const float *u, void *dh, float *du, float *partials, int64_t rows, int d,
                   int dtype, void *stream);

int cpm_set_rng_base(const uint64_t *device_counter);

/* ---- Dense Linear layers on tcgen05 (csrc/tc_gemm.cu) ---------------------------------------------------------------
 * Replaces the cuBLAS GEMMs behind every nn.Linear of the reference agent: in_linear (agent_pretrain.py:239,337), ft's
 * query/key/value/out projections and linear1/linear2 (agent_pretrain.py:244-253 -> fast_transformers AttentionLayer /
 * TransformerEncoderLayer) and the output heads (agent_pretrain.py:360-375), forward AND backward.
 * bf16 operands, fp32 accumulation in tensor memory, 2-CTA UMMA (256 x 256 tiles), TMA-fed.  Row-major matrices with a
 * row stride in elements (multiple of 8) and 16-byte aligned bases; ragged M / N / K are handled by TMA (zero fill on
 * loads, clipping on stores).
 *
 * cpm_gemm_nt:  D[M x N] = epilogue(A[M x K] . B[N x K]^T).   Forward: A = activations, B = weight (out x in).  Data
 *   gradient: A = dY, B = the transposed weight copy (in x out).  Epilogues:
 *     CPM_GEMM_EPI_BIAS   D = acc + bias                                   (bias fp32 [N] or NULL)
 *     CPM_GEMM_EPI_GELU   D = h = bf16(acc + bias);  D2 = dropout(gelu(h))  exact-erf GELU (ft activation='gelu'); D must be
 *                         dense (ldd == N); the dropout mask is the one cpm_gelu_fwd draws for (seed, rng_offset)
 *     CPM_GEMM_EPI_DGELU  D = acc * gelu'(aux) * dropout mask               aux = the stored pre-activation h (M x N, dense);
 *                         same mask as the forward for the same (seed, rng_offset): what cpm_gelu_bwd computes
 * cpm_gemm_tn:  dW[N x K] += dY[T x N]^T . X[T x K] (fp32): the weight gradient of a Linear layer, both operands read
 *   MN-major from the row-major activations (no transposes), 256 x 512 output tiles, the token range split over clusters.
 *   Output rows are routed to n_dst destination matrices of rows_per_dst rows each (row stride ldw): the separate q / k / v
 *   master gradients behind ONE fused projection GEMM; n_dst = 1, rows_per_dst = N for a plain layer.  dW_host: host array
 *   of n_dst device pointers.  ACCUMULATES with fp32 atomics (zero the buffers for a fresh gradient; this is also the
 *   accumulation across micro-batches): summation order over token ranges is not fixed, so results are reproducible to fp32
 *   rounding, not bitwise.  (Bias gradients: cpm_colsum.) */
#define CPM_GEMM_TN_MAX_DST 4
#define CPM_GEMM_EPI_BIAS 0
#define CPM_GEMM_EPI_GELU 1
#define CPM_GEMM_EPI_DGELU 2
/* Schedule override for cpm_gemm_nt (A/B measurements and tests): 0 auto | 1 stream (A and B tiles both streamed, one
 * 256 x 256 tile at a time) | 3 wide 256 x 512 tile (N <= 512).  Unsupported requests fall back to auto. */
int cpm_gemm_set_mode(int mode);
int cpm_gemm_nt(const void *A, int64_t lda, const void *B, int64_t ldb, void *D, int64_t ldd, void *D2, int64_t ldd2,
                int M, int N, int K, const float *bias, int epilogue, const void *aux, int64_t ld_aux,
                float p_drop, uint64_t seed, uint64_t rng_offset, void *stream);
int cpm_gemm_tn(const void *dY, int64_t ldy, const void *X, int64_t ldx, float *const *dW_host, int n_dst, int rows_per_dst,
                int64_t ldw, int T, int N, int K, void *stream);

/* The same product for the recurrent token step (M = sequences in flight): 64 x 32 tiles, two CTAs per SM, programmatic
 * dependent launch (the weight tiles are fetched before griddepcontrol.wait).  Epilogues CPM_GEMM_EPI_BIAS, CPM_GEMM_EPI_GELU
 * (single output: gelu of the bf16-rounded pre-activation; no dropout - generation runs in eval mode).
 * Reference loop being served: testing-no-type-cp.py:157-167 (forward_hidden(..., is_training=False) + forward_output). */
int cpm_gemm_nt_small(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd, int M, int N, int K,
                      const float *bias, int epilogue, void *stream);
/* cpm_gemm_nt_small[_ln] split a long K (>= 16 k-blocks of 64, an even number: linear2's K = 2048) in two over a thread-block
 * cluster per output tile; the leader CTA adds the other slice's fp32 tile from its shared memory and runs the epilogue
 * (deterministic).  0 switches the split off (A/B runs); default on. */
int cpm_gemm_small_set_split(int on);
/* Development aid: device log {uint64 count; uint64 stamps[capacity][8]} (zero-filled by the caller) to which block (0,0) of
 * every cpm_gemm_nt_small[_ln] launch appends %globaltimer stamps (kernel entry, set-up done, griddepcontrol.wait returned,
 * first activation block landed, accumulator ready, epilogue stored) and N, K.  The pointer is read at LAUNCH time, so it is
 * baked into captured graphs.  NULL switches it off (the default). */
int cpm_debug_small_timing(void *device_log, int capacity);
/* The token-step Linear with the LayerNorm around it folded in - no LayerNorm kernels in the rollout chain.  ft's recurrent
 * encoder layer is  x = norm1(x + out_proj(attn));  y = linear2(gelu(linear1(x)));  x = norm2(x + y).
 * Exactly one of two forms per call:
 *  FOLD (fold_c1 != NULL): A holds RAW pre-norm rows y, W holds gamma o W (bf16), bias holds c2 = W beta + b, fold_c1 = row sums
 *    of the bf16 gamma o W:   D = epilogue( rstd (A W^T - mean c1) + c2 ),  mean / rstd of each row of A computed in the kernel
 *    from the activation tile it holds (K <= 512; one-pass fp32 sum and sum of squares, eps = ln_eps) and, if stats_out != NULL,
 *    written there as (mean, rstd) float pairs (M x 2).  Epilogues CPM_GEMM_EPI_BIAS / CPM_GEMM_EPI_GELU as above.
 *  RESIDUAL (R != NULL):   D = bf16(A W^T + bias) + res,  res = R (bf16, row stride ldr) when r_stats == NULL, else
 *    res = bf16( (R - mean) rstd r_gamma + r_beta ) with (mean, rstd) = r_stats[row]: the LayerNorm output rebuilt from the
 *    pre-norm rows its consumer normalised.  D is the next pre-norm row.
 * Reference: the same loop as cpm_gemm_nt_small (testing-no-type-cp.py:157-167). */
int cpm_gemm_nt_small_ln(const void *A, int64_t lda, const void *W, int64_t ldw, void *D, int64_t ldd, int M, int N, int K,
                         const float *bias, int epilogue, const float *fold_c1, float *stats_out, float ln_eps, const void *R, int64_t ldr,
                         const float *r_stats, const float *r_gamma, const float *r_beta, void *stream);
/* 1: launch the token-step kernels (cpm_gemm_nt_small, cpm_linattn_step, cpm_ln_residual_fwd, cpm_embed_fwd, cpm_add_pe,
 * cpm_heads_sample, cpm_rollout_advance) with the programmatic-stream-serialization attribute, so each overlaps its set-up
 * with its predecessor's tail; they all block in griddepcontrol.wait before touching chain data.  0 (default): plain launches. */
int cpm_set_chain_pdl(int on);

/* bias + exact-erf GELU + dropout (ft activation='gelu' => F.gelu; K6):
 *     y = drop(gelu(x + bias))   (bias fp32 (d) or NULL)
 * bwd: gx = gy * mask/(1-p) * gelu'(x + bias). */
int cpm_gelu_fwd(const void *x, const float *bias, void *y, int64_t rows, int d, float p_drop,
                 uint64_t seed, uint64_t rng_offset, int dtype, void *stream);
/* dbias (optional, fp32 (d), +=): fused bias gradient = column sums of gx; needs d | 4096 and
 * partials of cpm_gelu_bwd_partials_rows(d) x d floats. */
int cpm_gelu_bwd_partials_rows(int d);
int cpm_gelu_bwd(const void *x, const float *bias, const void *gy, void *gx, float *dbias, float *partials,
                 int64_t rows, int d, float p_drop, uint64_t seed, uint64_t rng_offset, int dtype, void *stream);

/* ------------------------------------------------------------------------------------------
 * C3 — per-attribute decode over concatenated head logits.  Replaces
 * `forward_output_sampling` + host numpy `sampling/softmax_with_temperature/nucleus/
 * weighted_sampling` (dqn_policy/model.py:19-55,259-298) and the softmax/argmax/log read-out of
 * `choose_action` (ppo_train.py:259-267; IRL_dqn_train.py:244-250).
 *   logits (rows, ld_logits) ; attribute a occupies columns [seg_host[a], seg_host[a+1]).
 *   mode 0: greedy argmax (first maximal index).  mode 1: temperature softmax, optional nucleus
 *   (top_p_host[a] in (0,1); <=0 or >=1 disables), inverse-CDF draw with the Philox uniform of
 *   (seed, seq_id = seq_base + row, step = *step_dev or step, attribute a).
 *   tokens (rows,n_attr) int64; logp (rows,n_attr) fp32 = log softmax_{T=1}(logits)[token]
 *   (may be NULL); entropy (rows,n_attr) fp32 of the T=1 softmax (may be NULL).
 * Segment width <= 1024.
 * ------------------------------------------------------------------------------------------ */
int cpm_heads_sample(const void *logits, int64_t rows, int64_t ld_logits, const int *seg_host,
                     int n_attr, const float *temperature_host, const float *top_p_host, int mode,
                     uint64_t seed, int64_t seq_base, int step, const int32_t *step_dev,
                     int64_t *tokens, float *logp, float *entropy, int dtype, void *stream);

/* log-prob / entropy of GIVEN tokens under the T=1 softmax of each segment (PPO update).
 * Also the backward: dlogits += glogp*(onehot - p) + gent*(-p*(log p + H)) when dlogits!=NULL is
 * done by cpm_heads_logp_bwd. */
int cpm_heads_logp(const void *logits, int64_t rows, int64_t ld_logits, const int *seg_host,
                   int n_attr, const int64_t *tokens, float *logp, float *entropy, int dtype,
                   void *stream);
int cpm_heads_logp_bwd(const void *logits, int64_t rows, int64_t ld_logits, const int *seg_host,
                       int n_attr, const int64_t *tokens, const float *glogp, const float *gentropy,
                       void *dlogits, int dtype, void *stream);

/* C4 — masked mean cross-entropy per attribute.  Replaces `compute_loss` x6 inside
 * `train_step` (agent_pretrain.py:279-311): loss_a = sum_t mask_t*CE(logits_t[seg a], tgt_ta)
 * / sum_t mask_t.
 *   fwd: loss_num[a] += numerator (fp32, caller zeroes), mask_sum[0] += sum(mask), lse (T,n_attr).
 *   bwd: dlogits[t,seg a] = gscale[a] * mask_t / denom[0] * (softmax - onehot)   (device scalars
 *   gscale (n_attr) and denom (1): the caller may all-reduce denom across ranks first). */
int cpm_masked_ce_fwd(const void *logits, int64_t T, int64_t ld_logits, const int *seg_host,
                      int n_attr, const int64_t *targets, const float *mask, float *loss_num,
                      float *mask_sum, float *lse, int dtype, void *stream);
int cpm_masked_ce_bwd(const void *logits, int64_t T, int64_t ld_logits, const int *seg_host,
                      int n_attr, const int64_t *targets, const float *mask, const float *lse,
                      const float *gscale, const float *denom, void *dlogits, int dtype,
                      void *stream);

/* ------------------------------------------------------------------------------------------
 * D1 — returns / advantage scans (fp32, (B,T) row-major; one warp per trajectory).
 *   mode CPM_RET_COMPAT : reference `calculate_returns` (ppo_train.py:348-353):
 *        R_i = r_i + gamma*R_{i-1} accumulated forward, stored reversed: ret[t] = R_{T-1-t}.
 *   mode CPM_RET_TOGO   : ret[t] = r_t + gamma*(1-done_t)*ret[t+1].
 *   mode CPM_RET_GAE    : GAE(lambda): adv[t] = delta_t + gamma*lam*(1-done_t)*adv[t+1],
 *        delta_t = r_t + gamma*(1-done_t)*V_{t+1} - V_t (V_T = last_value); ret = adv + V.
 * values/dones/last_value/adv may be NULL where the mode does not use them.
 * ------------------------------------------------------------------------------------------ */
#define CPM_RET_COMPAT 0
#define CPM_RET_TOGO 1
#define CPM_RET_GAE 2
int cpm_returns_scan(const float *rewards, const float *values, const float *dones,
                     const float *last_value, float *ret, float *adv, int B, int T, float gamma,
                     float lam, int mode, void *stream);
/* moments of (x - sub) (sub may be NULL): out3 += {count, sum, sum of squares} in fp64. */
int cpm_moments(const float *x, const float *sub, int64_t n, double *out3, void *stream);
/* out = ((x - sub) - mean)/std with mean/std from moments3 (device; may have been all-reduced);
 * unbiased!=0 uses the n-1 denominator like torch.std (ppo_train.py:356,362). */
int cpm_zscore(const float *x, const float *sub, float *out, int64_t n, const double *moments3,
               int unbiased, float eps, void *stream);

/* D4 — reward / discriminator head of the PPO reward model (ppo_policy/model.py:474-493: six proj_a, six
 * eval_a = Linear(n_a, 1), mean over the sequence, sigmoid, average of the six).  The Longformer
 * body stays outside (HF); this takes its last hidden state h (N,L,d).  eval_a(proj_a(h)).mean(L) is linear in h, so
 * the host collapses each attribute to u_a = W_a^T w_a (d floats) and c_a = w_a.b_a + bias_a:
 *     scores[n,a] = sigmoid( mean_l h[n,l,:] . u_a + c_a ),   reward[n] = mean_a scores[n,a].
 * One launch, h read once.  u (n_attr, d) fp32, c (n_attr) fp32; scores may be NULL. */
int cpm_reward_head(const void *h, const float *u, const float *c, float *reward, float *scores,
                    int N, int L, int d, int n_attr, int dtype, void *stream);

/* D2 — PPO losses, forward + backward in one launch (grad of the scalar loss w.r.t. inputs,
 * scaled by grad_scale).
 *   mode CPM_PPO_COMPAT (ppo_train.py:388-396): new_logp (C) broadcast over T rows,
 *        old_logp (T,C), adv (T):  loss = -mean_{t,c} min(0.2*adv_t, clamp(exp(new_c-old_tc),
 *        1-clip,1+clip)*adv_t);  out[0]=loss;  dnew (C).
 *   mode CPM_PPO_STANDARD: elementwise n = T*C, new/old/adv/entropy (n), value/ret (nv):
 *        loss = -mean(min(r*A, clamp(r)*A)) + vf_coef*mse(value,ret) - ent_coef*mean(entropy);
 *        out[0..3] = {loss, policy, value, entropy}; dnew (n), dvalue (nv), dentropy (n).
 * `out` must be zeroed by the caller. */
#define CPM_PPO_COMPAT 0
#define CPM_PPO_STANDARD 1
int cpm_ppo_loss_fwd_bwd(const float *new_logp, const float *old_logp, const float *adv,
                         const float *entropy, const float *value, const float *ret,
                         float *out, float *dnew, float *dentropy, float *dvalue,
                         int64_t T, int64_t C, int64_t nv, float clip, float vf_coef,
                         float ent_coef, float grad_scale, int mode, void *stream);

/* D3 — DQN TD target + MSE, forward + backward (IRL_dqn_train.py:285-330).
 *   q_logits, next_logits: (B,L,ld) with attribute segments seg_host; action (B,A,n_attr) int64;
 *   reward, done (B) fp32.  target[b,k,a] = r_b + gamma*(1-done_b)*top_k-th( max_vocab
 *   next[b,:,seg a] ) (descending over the L positions, A largest).
 *   mode CPM_TD_COMPAT: Q(s,a)[b,k] = q_logits[0, b, seg a + action[b,k,a]] (the reference's
 *        gather quirk; requires B <= L).  mode CPM_TD_STANDARD: Q[b,k] = q_logits[b, L-1-k, ...]
 *        and the target is max_vocab next[b, L-1-k, seg a] (no top-k).
 *   out[0] += mean over attributes of MSE; dq (B,L,ld) is zero-filled then receives
 *   grad_scale * dloss/dq_logits (same dtype as logits).  targets_out (B,A,n_attr) fp32 optional. */
#define CPM_TD_COMPAT 0
#define CPM_TD_STANDARD 1
int cpm_dqn_td_fwd_bwd(const void *q_logits, const void *next_logits, const int64_t *action,
                       const float *reward, const float *done, float *out, void *dq,
                       float *targets_out, int B, int L, int64_t ld, const int *seg_host,
                       int n_attr, int A, float gamma, float grad_scale, int mode, int dtype,
                       void *stream);

/* Rollout bookkeeping for graph-captured generation (device-side step counter):
 *   history_tok[step*n_tok + i] = tokens[i] ; history_f[step*n_f + i] = vals[i] (optional pair);
 *   then *step_dev += 1.  Nothing is written once *step_dev >= max_steps.  Single CTA. */
int cpm_rollout_advance(const int64_t *tokens, int64_t *history_tok, int64_t n_tok, const float *vals,
                        float *history_f, int64_t n_f, int32_t *step_dev, int32_t max_steps,
                        void *stream);

/* ------------------------------------------------------------------------------------------
 * The whole recurrent token step as ONE persistent cooperative kernel (csrc/rollout_step.cu).
 * Replaces the reference's host-driven generation loop (testing-no-type-cp.py:157-167: forward_hidden(is_training=False)
 * -> forward_output -> six numpy samplers, one token of one song per trip) for `batch` songs at once:
 *   embedding gather + in_linear + positional encoding -> n_layers x [QKV projection, recurrent linear-attention state
 *   update (S += phi(k) v^T, z += phi(k), out = phi(q) S / (phi(q).z + eps)), out-projection + residual, LayerNorm + linear1
 *   + GELU, linear2 + residual] -> LayerNorm, final LayerNorm, 6/7 output heads -> temperature / nucleus sampling with the
 *   Philox stream of (seed, seq_base + row, step, attribute) -> history + step counter,
 * `n_steps` tokens per launch.  One CTA per SM; the phases are separated by a device-wide barrier; every CTA streams the
 * weight tiles of ITS output tiles through a TMA ring that runs ahead of the barriers (weights do not depend on the chain),
 * the Linear layers run as tcgen05 tiles with the weight rows on the UMMA M axis and 16 / 32 songs on the N axis, LayerNorm
 * is applied while the activation tile is staged.  bf16 activations and weights, fp32 accumulation / statistics / state.
 * All pointers are device pointers that must stay valid (and at the same address) until cpm_rollout_destroy; weights are
 * read on every run, so in-place updates of the packed copies are picked up.
 * The only allocation is the small host-side handle; the device-side plan lives in the caller's `plan_dev` buffer.
 * cpm_rollout_create validates the configuration (CPM_ERR_UNSUPPORTED for shapes outside d_head = 64, d_model <= 512,
 * d_model % 64 == 0, d_ff % 64 == 0, segments <= 256 wide) and builds the tensor maps and the device-side phase table. */
#define CPM_ROLLOUT_MAX_LAYERS 32
typedef struct CpmRolloutLayer {
    const void *w_qkv; const float *b_qkv;          /* (3 d_model, d_model) bf16 ; (3 d_model) fp32 : [q; k; v] rows */
    const void *w_out; const float *b_out;          /* (d_model, d_model) */
    const float *ln1_g, *ln1_b;                     /* norm1 */
    const void *w_ff1; const float *b_ff1;          /* (d_ff, d_model) */
    const void *w_ff2; const float *b_ff2;          /* (d_model, d_ff) */
    const float *ln2_g, *ln2_b;                     /* norm2 */
    float *S, *Z;                                   /* recurrent state (batch, H, 64, 64), (batch, H, 64) fp32, updated in place */
} CpmRolloutLayer;
typedef struct CpmRolloutConfig {
    int batch, d_model, n_heads, d_ff, n_layers, n_attr;
    int n_tokens[CPM_MAX_ATTR], emb[CPM_MAX_ATTR];  /* vocabulary and embedding width per attribute */
    const float *tables[CPM_MAX_ATTR];              /* embedding tables (n_tokens[a], emb[a]) fp32 */
    const void *w_in; const float *b_in;            /* in_linear (d_model, sum emb) bf16 ; fp32 bias */
    const float *pe; int pe_len;                    /* positional encoding (pe_len, d_model) fp32 */
    int true_positions;                             /* 1: position = step counter; 0: position 0 every step (reference quirk) */
    CpmRolloutLayer layer[CPM_ROLLOUT_MAX_LAYERS];
    const float *lnf_g, *lnf_b;                     /* the encoder's final LayerNorm */
    const void *w_heads; const float *b_heads;      /* concatenated heads (logits_ld rows, d_model) bf16 (rows beyond seg[n_attr] zero) */
    int seg[CPM_MAX_ATTR + 1]; int logits_ld;       /* column offsets of the attributes in the concatenated logits; row stride */
    float temperature[CPM_MAX_ATTR], top_p[CPM_MAX_ATTR];
    int greedy;                                     /* 1: argmax; 0: sample */
    float ln_eps, attn_eps;
    uint64_t seed; int64_t seq_base;
    int64_t *cur;                                   /* (batch, n_attr) current token per song: input of the step, overwritten by the sample */
    float *logp;                                    /* (batch, n_attr) log-prob (T = 1 policy) of the sampled sub-tokens */
    int64_t *hist_tok; float *hist_logp;            /* (max_steps, batch, n_attr) histories, written at row *step_dev */
    int32_t *step_dev; int32_t max_steps;           /* device step counter, advanced by one per token */
    /* scratch owned by the caller, bf16: x (batch, d_model) x3, qkv (batch, 3 d_model), attn (batch, d_model),
     * g (batch, d_ff), logits (batch, logits_ld) */
    void *x0, *x1, *y, *qkv, *attn, *g, *logits;
    unsigned long long *barrier;                    /* 1 counter, zero-initialised once by the caller */
    int *err_flag;                                  /* set to 1 on an out-of-range token id (optional) */
} CpmRolloutConfig;
int64_t cpm_rollout_plan_bytes(void);              /* size of the device-side plan buffer the caller provides (256-byte aligned) */
int cpm_rollout_create(const CpmRolloutConfig *cfg_host, void *plan_dev, void **handle_out);
int cpm_rollout_run(void *handle, int n_steps, void *stream);
int cpm_rollout_phases(void *handle);               /* device-wide phases per token (diagnostics) */
/* Development aid: device buffer of (SMs x 4 steps x phases x 8) uint64 that receives %globaltimer stamps of every CTA
 * ("stage work done", "barrier passed", "activation tile staged", "accumulator ready") for the first 4 steps of each run;
 * NULL switches it off (the default). */
int cpm_debug_rollout_timing(void *device_buffer);
int cpm_rollout_destroy(void *handle);

#ifdef __cplusplus
}
#endif
#endif /* CPMUSIC_H_ */
