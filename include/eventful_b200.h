/*
 * eventful_b200.h -- C ABI of libeventful_b200.so, the sm_100a implementation of the
 * gated sparse-token update path of Eventful Transformers.
 *
 * This is the drop-in boundary below the reference's Python module API
 * (SURVEY.md 8(b)).  The reference has no FFI of its own (it is pure PyTorch);
 * each entry point below names the reference call site(s) whose ATen dispatches
 * it replaces.  All file:line citations are relative to the reference checkout.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors kept
 *     alive by the Python layer); nothing is allocated or freed in the library;
 *   - tensors are dense, row-major, 16-byte aligned; `dtype` is the element type
 *     of every activation / state / weight pointer of the call;
 *   - indices are int64 (torch.int64), shape (B, k);
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (they
 *     are CUDA-graph capturable) and are thread-safe for distinct streams;
 *   - return value 0 = ok, non-zero = error; et_last_error() (thread-local) says why.
 *     There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef EVENTFUL_B200_H
#define EVENTFUL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { ET_F32 = 0, ET_BF16 = 1, ET_F16 = 2 } et_dtype;
typedef enum { ET_SELECT_TOPK = 0, ET_SELECT_THRESHOLD = 1 } et_select_mode;
typedef enum { ET_ACT_NONE = 0, ET_ACT_GELU = 1 } et_activation;          /* nn.GELU() exact erf, blocks.py:114 */
typedef enum { ET_ATTN_DENSE = 0, ET_ATTN_FIRST = 1, ET_ATTN_DELTA = 2 } et_attn_mode;
typedef enum { ET_OK = 0, ET_ERR_ARG = 1, ET_ERR_CUDA = 2, ET_ERR_UNSUPPORTED = 3 } et_status;

/* ---- library / device queries ------------------------------------------------ */
int         et_version(void);
/* Kernels enqueued by this library so far in this process (bench.py reports the per-run delta as gpu_launches). */
long long   et_launch_count(void);
const char* et_last_error(void);
/* Fails with ET_ERR_UNSUPPORTED unless the device is compute capability 10.x. */
int         et_device_info(int device, int* cc_major, int* cc_minor, int* sm_count);

/*
 * Fused gate selection: [x = xa + xb] -> [c = LayerNorm(x)] -> e = c - p ->
 * per-token L2 norm (fp32 accumulate, rounded to dtype) -> radix top-k or
 * threshold selection, ONE launch.
 * Replaces: CountedAdd (counting.py:16-19), nn.LayerNorm (blocks.py:442-444,458-460),
 *   `c - self.p` (modules.py:149,196), TokenNormTopK.forward (policies.py:58-63),
 *   TokenNormThreshold.forward (policies.py:20-32), TokenNormTopFraction (policies.py:88-95).
 * Selection rule = torch CUDA radix select: strictly greater than the k-th value in
 * ascending index order, then equal to it in ascending index order.
 *   xa       (R, N, D)            gate input (or first addend)
 *   xb       (R, N, D) or NULL    second addend (residual); sum rounded to dtype
 *   xsum_out (R, N, D) or NULL    receives xa + xb
 *   c_out    (R, N, D) or NULL    receives the gate input c = LayerNorm(x) of every token: the gather source of
 *                                 et_linear_gather (no separate gather launch between the gate and its linear layer)
 *   ln_w/b   (D) or NULL          LayerNorm affine, eps = ln_eps
 *   p        (R, N, D) or NULL    gate reference state (NULL: the norm of c itself)
 *   norm_out (R, N) float         per-token norms (value already rounded to dtype)
 *   idx_out  (R, k) / (R, N)      selected indices; threshold mode writes count_out[r] of them
 *   ticket   (R) int32            zero-initialised once; self-resetting
 */
int et_gate_select(const void* xa, const void* xb, void* xsum_out, void* c_out, const void* ln_w, const void* ln_b,
                   float ln_eps, const void* p, int64_t R, int64_t N, int64_t D, int dtype, int mode,
                   int64_t k, float threshold, float* norm_out, int64_t* idx_out, int32_t* count_out,
                   int32_t* ticket, void* stream);

/*
 * Row gather + reference-state advance: c~ = c[idx], e~ = c[idx] - p[idx], p[idx] = c~.
 * With LayerNorm parameters the gathered rows are normalised on the fly (ln_after = 0:
 * LN precedes the gate, blocks.py:459-461; ln_after = 1: gate_before_ln, blocks.py:456-458,
 * the state keeps the un-normalised rows).
 * Replaces: c.gather / e.gather / p.scatter_ in TokenGate / TokenDeltaGate (modules.py:150-151,198-200).
 *   count    (R) int32 or NULL   device-side number of valid indices per row (threshold policy)
 *   full_replace != 0            SimpleSTGTGate: afterwards p <- c for every token (modules.py:44)
 */
int et_gate_gather(const void* x, const void* ln_w, const void* ln_b, float ln_eps, int ln_after, void* p,
                   const int64_t* idx, const int32_t* count, int64_t R, int64_t N, int64_t D, int64_t k,
                   int dtype, void* c_tilde, void* e_tilde, int full_replace, void* stream);

/* Column-structure gate (structure="col", modules.py:161-163) on (R, N, M): gathers columns idx. */
int et_gate_gather_cols(const void* c, void* p, const int64_t* idx, int64_t R, int64_t rows_per_index,
                        int64_t N, int64_t M, int64_t k, int dtype, void* c_tilde, void* e_tilde, void* stream);

/*
 * TokenBuffer.forward_incremental (modules.py:86-97): buf[r, idx[r, j], :] = x[r, j, :]
 * (structure 0 = "row") or buf[r, :, idx[r, j]] = x[r, :, j] (structure 1 = "col").
 * `rows_per_index` lets one (B, k) index drive (B, H, ...) tensors (utils.py:198-211).
 */
int et_buffer_scatter(void* buf, const void* x, const int64_t* idx, const int32_t* count, int64_t R,
                      int64_t rows_per_index, int64_t N, int64_t D, int64_t k, int dtype, int structure,
                      void* stream);

/* CountedAdd (counting.py:9-22): out = a + b elementwise (out may alias a). */
int et_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
/* e = c - p (modules.py:149) for the generic (user-defined policy) gate path. */
int et_sub(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);

/* Test / tuning / measurement hooks.  key 1: force the GEMM tile width BLOCK_N (0 = automatic); key 2: 0 routes
 * attention through the mma.sync kernels even where the tcgen05 kernels apply (1 = default); key 3: device pointer to
 * 8 x uint64 receiving %globaltimer phase stamps of et_gate_select (0 = off); key 5: GEMM pipeline depth (1 deep,
 * 2 shallow = two CTAs per SM, 0 = automatic); key 6: 1 brackets the global-attention apply kernel with CUDA events;
 * key 4 / key 7: device pointer to 8 x 16 / 3 x 16 uint64 cycle buckets per warp role of the attention / GEMM kernels
 * (written only by the profiling build, `make prof`); key 8: GEMM rows per CTA tile (1 = 128, 2 = 256, 0 = automatic);
 * key 9: persistent GEMM kernel (1 = always, 2 = never, 0 = automatic); key 10: CTAs per SM of the gate kernels (default 2);
 * key 11: 1 selects the first-generation tcgen05 window-attention kernel (default 2 = second generation). */
int et_debug_set(int key, long long value);
/* Milliseconds of the last apply-kernel launch bracketed under key 6 (synchronises on its end event). */
float et_debug_elapsed_ms(void);

/*
 * Gathered-row linear with scatter epilogue on tcgen05 tensor cores (TMA operand
 * staging, TMEM accumulators):  y = act(A @ W^T + bias), then
 *     out[(m / k) * n_out_rows + idx[m], :] = y[m, :]      (idx != NULL: TokenBuffer scatter)
 *     out[m, :] = y[m, :]                                    (idx == NULL)
 * Replaces: CountedLinear.forward (counting.py:157-162) at blocks.py:122,433,462 and
 *   blocks.py:242-246 (mlp_1 + GELU + mlp_2), fused with TokenBuffer.scatter_ (modules.py:96).
 *   A (M, K), W (n_feat, K) torch Linear layout (counting.py:142-143), bias (n_feat).
 *   count (M / k) int32 or NULL: device-side valid rows per batch entry.
 * ET_BF16 / ET_F16: tcgen05 tensor cores, fp32 accumulate.  ET_F32: fp32 SGEMM on the CUDA cores (fp32 models,
 * BASELINE configs[0] and the reference's timed CUDA config, configs/time/vitdet_vid/_cuda.yml).
 */
int et_linear(const void* A, int64_t M, int64_t K, const void* W, const void* bias, int64_t n_feat, int act,
              void* out, int64_t ld_out, const int64_t* idx, const int32_t* count, int64_t k,
              int64_t n_out_rows, int dtype, void* stream);

/*
 * The same linear layer with the gate's gather AND its state advance fused in (16-bit dtypes):
 *     A[m, :] = a_src[(m / k) * a_rows + a_idx[m], :]       A-operand rows fetched by TMA tile::gather4 straight into the
 *                                                            128B-swizzled tcgen05 operand tiles (no c~ tensor in HBM)
 *     state[(m / k) * a_rows + a_idx[m], :] = A[m, :]        (state != NULL) the gate reference advances in the same kernel
 * Replaces: `c.gather(index)` + `self.p.scatter_(index, c~)` of TokenGate.forward_incremental (modules.py:150-151) followed
 *   by CountedLinear.forward (counting.py:157-162) and TokenBuffer.scatter_ (modules.py:96) -- one launch per gate site.
 *   a_src (M / k * a_rows, K): the gate input of every token (c_out of et_gate_select, or the un-normalised input);
 *   a_idx (M) int64: the gate index; state: same shape as a_src or NULL; the other arguments as et_linear
 *   (idx may be NULL: mlp_1 gathers but does not scatter).  count applies to both the gather and the scatter.
 */
int et_linear_gather(const void* a_src, int64_t a_rows, const int64_t* a_idx, void* state, int64_t M, int64_t K,
                     const void* W, const void* bias, int64_t n_feat, int act, void* out, int64_t ld_out,
                     const int64_t* idx, const int32_t* count, int64_t k, int64_t n_out_rows, int dtype, void* stream);

/*
 * Dense windowed self-attention of EventfulTokenwiseBlock / Block, fused
 * (window partition with qkv-bias pad tokens, head split, q/sqrt(dh) . k^T, decomposed
 * rel-pos bias from the unscaled q, softmax, a . v, head merge, window recombine + crop).
 * Replaces: Block._forward_attention and helpers (blocks.py:205-240,248-301,329-376),
 *   RelativePositionEmbedding.forward (eventful_transformer/utils.py:139-171).
 *   qkv (B, gh*gw + extra, 3*H*dh); pad_token (3*H*dh) = qkv.bias (blocks.py:275-287);
 *   rel_y (wh, wh, dh), rel_x (ww, ww, dh) relative tables or NULL; out (B, N, H*dh).
 *   window (wh, ww) = (0, 0) means one global window over all tokens (incl. class token).
 *   ET_F32 runs in fp32 arithmetic on the CUDA cores and always needs the workspace (softmax statistics).
 */
int et_window_attention(const void* qkv, const void* pad_token, const void* rel_y, const void* rel_x, void* out,
                        void* workspace, int64_t B, int64_t N, int64_t gh, int64_t gw, int64_t wh, int64_t ww,
                        int64_t heads, int64_t dh, int dtype, void* stream);

/* Bytes of caller-provided scratch (`workspace`) the two attention entry points need (upper bound over all code
 * paths): rel-pos bias tables (B, H, N, gh + gw), for the global DELTA mode the v-gate deltas 2 x (B, k, H*dh), and
 * the fp32 softmax statistics of the general-precision path.  Pass wh = ww = 0 for global attention. */
int64_t et_attn_workspace_bytes(int64_t B, int64_t N, int64_t gh, int64_t gw, int64_t wh, int64_t ww,
                                int64_t heads, int64_t dh, int64_t k, int has_relpos);

/*
 * Global self-attention of the Eventful blocks over the QKV TokenBuffer.
 *   mode DENSE : out = softmax(q k^T / sqrt(dh) + relpos) v                (Block / EventfulMatmul1Block)
 *   mode FIRST : as DENSE, and initialises the gate / accumulator state (frame 0)
 *   mode DELTA : A-gate + v-gate (forced by idx) + MatmulDeltaAccumulator update:
 *        a_n = softmax(...)[:, idx];  dA = a_n - a_state[:, idx];  a_state[:, idx] = a_n
 *        v_n = v[idx];  dV = v_n - v_state[idx];  v_state[idx] = v_n
 *        acc += a_n . dV + dA . (v_n - dV);   out = acc
 * Replaces: MatmulBuffer (modules.py:204-252; the product is recomputed, it always equals
 *   (q/scale) k^T of the current buffer), softmax (blocks.py:522), TokenDeltaGate x2
 *   (blocks.py:567-568, modules.py:187-201), MatmulDeltaAccumulator (modules.py:285-295),
 *   RelativePositionEmbedding.forward(inplace=False) (blocks.py:521), Block._cast_matmul_2 /
 *   _uncast_matmul_2 (blocks.py:183-189,393-396) and the pooled-key variants (blocks.py:509-511).
 *   qkv (B, N, 3*H*dh);
 *   kv_pooled NULL, or (B, Nk, 2*H*dh) = [k | v] averaged over pool_h x pool_w cells of the token grid
 *           (et_pool_kv; Nk = (gh / pool_h) * (gw / pool_w)); idx then holds pooled-cell ids (et_pool_index);
 *   rel_y (gh, kh, dh) / rel_x (gw, kw, dh) or NULL (class-token models); kh = gh / pool_h, kw = gw / pool_w
 *           (tables averaged along the key axis, utils.py:185-188; kh = gh, kw = gw without pooling);
 *   count NULL, or (B) int32 on the device: number of valid entries of idx per batch entry (threshold policy,
 *           pooled unique indices); k is then the row stride of idx and the upper bound;
 *   state_dtype: element type of a_state / v_state / acc (matmul_2_cast); out has the model dtype;
 *   a_state (B, H, Nk, NP) stored COLUMN-major per head: a_state[b][h][col][row], row stride
 *           NP = N rounded up to a multiple of 8 (16-byte segments of a selected column);
 *   v_state (B, Nk, H*dh); acc (B, N, H*dh); out (B, N, H*dh); idx (B, k).
 *   row_stats (B, H, N, 2) float workspace (row max, row sum).
 * 16-bit models whose state has the model dtype and un-pooled keys run on the tcgen05 / mma.sync tensor-core
 * kernels; fp32 models, state_dtype != dtype and pooled keys run in fp32 arithmetic on the CUDA cores.
 */
int et_global_attention(const void* qkv, const void* kv_pooled, int64_t pool_h, int64_t pool_w, const void* rel_y,
                        const void* rel_x, int mode, const int64_t* idx, const int32_t* count, int64_t k, void* a_state,
                        void* v_state, void* acc, void* out, float* row_stats, void* workspace, int64_t B, int64_t N,
                        int64_t gh, int64_t gw, int64_t heads, int64_t dh, int dtype, int state_dtype, void* stream);

/*
 * Adaptive token sampling, scoring pass (Block._adaptive_token_sampling, blocks.py:150-157): one statistics pass over the
 * QKV buffer that also emits raw_scores[b, h, t] = softmax(q k^T / sqrt(dh))[b, h, t, 0] * |v[b, h, t]| -- the attention each
 * token pays to the class token times the norm of its value vector -- with every factor rounded to `score_dtype` where the
 * reference holds it in that dtype (the model dtype, or the matmul_2_cast dtype in EventfulBlock, blocks.py:561-562).
 * The normalisation, the sum over axis -3, the top-k and the index stabilisation (blocks.py:157-176,378-391; a host-side
 * loop in the reference too) run on the (B, H, N) scores in the host mirror.
 *   qkv (B, N, 3*H*dh); row_stats (B, H, N, 2) float (row max, row sum); raw_scores (B, H, N) float.  No rel-pos (ATS
 *   models carry a class token).
 */
int et_ats_scores(const void* qkv, int64_t B, int64_t N, int64_t heads, int64_t dh, int dtype, int score_dtype,
                  float* row_stats, float* raw_scores, void* stream);

/*
 * et_global_attention for a subset of the query rows (adaptive token sampling: `a.gather(dim=-2, ats_indices)`,
 * blocks.py:178-181, ahead of the A-gate / v-gate / accumulator, blocks.py:562-569): the Nq queries are the tokens
 * q_index[b][0 .. Nq) of the QKV buffer (the stabilised ATS index), keys and values are all N tokens.
 * a_state (B, H, N, NP) column-major with NP = Nq rounded up to 8; acc / out (B, Nq, H*dh); v_state (B, N, H*dh);
 * idx (B, k) selects keys as in et_global_attention.  Runs on the general-precision kernels (any dtype / state dtype).
 */
int et_global_attention_rows(const void* qkv, const int64_t* q_index, int64_t Nq, int mode, const int64_t* idx,
                             const int32_t* count, int64_t k, void* a_state, void* v_state, void* acc, void* out,
                             float* row_stats, void* workspace, int64_t B, int64_t N, int64_t heads, int64_t dh, int dtype,
                             int state_dtype, void* stream);

/*
 * K/V token pooling (Block._pool_tokens, blocks.py:303-326): out (B, Nk, 2*D) = avg_pool2d of the k and v parts of
 * qkv (B, gh*gw, 3*D) over pool_h x pool_w cells (fp32 average, rounded to dtype).
 */
int et_pool_kv(const void* qkv, void* out, int64_t B, int64_t gh, int64_t gw, int64_t D, int64_t pool_h, int64_t pool_w,
               int dtype, void* stream);

/*
 * EventfulMatmul1Block._pool_index (blocks.py:525-540): token indices idx (B, k) [first count_in[b] valid, or all
 * k when count_in is NULL] -> pooled-cell ids, ascending and unique per batch entry: out_idx (B, k), out_count (B).
 * (For B > 1 the reference's unique(dim=-1) de-duplicates columns jointly and can leave duplicates inside a row,
 * which double-counts deltas downstream; this entry point de-duplicates per row.)
 */
int et_pool_index(const int64_t* idx, const int32_t* count_in, int64_t B, int64_t k, int64_t gh, int64_t gw, int64_t pool_h,
                  int64_t pool_w, int64_t* out_idx, int32_t* out_count, void* stream);

/* Generic strided batched matmul C = alpha * (A @ B) [+ C] (fp32 accumulate, result rounded to dtype before the
 * optional in-place add) for the stand-alone MatmulBuffer / MatmulDeltaAccumulator / CountedMatmul modules
 * (modules.py:204-299, counting.py:165-175).  Two batch levels (batch_outer x batch_inner, e.g. batch x heads) so
 * that permuted views of a QKV buffer need no copy; strides_* are HOST arrays of 4 element strides each:
 * {outer batch, inner batch, row, column} of A (M x K), B (K x N) and C (M x N). */
int et_bmm(const void* A, const void* Bm, void* C, int64_t batch_outer, int64_t batch_inner, int64_t M, int64_t N,
           int64_t K, const int64_t* strides_a, const int64_t* strides_b, const int64_t* strides_c, int accumulate,
           float alpha, int dtype, void* stream);

/*
 * Patch / tubelet extraction for the embedding GEMM: LinearEmbedding (models/vitdet.py:17-52, Conv2d with
 * kernel = stride = patch) and TubeletEmbedding (models/vivit.py:153-192, Conv3d with kernel = stride = tubelet) are GEMMs
 * over non-overlapping patches.  x (B, T, C, H, W) [T = pt = 1 for images] -> out (B, T/pt, (H/ph)(W/pw), C*pt*ph*pw) with
 * the feature order of the flattened conv weight (dim, C, pt, ph, pw); the projection then runs on et_linear.
 */
int et_patchify(const void* x, void* out, int64_t B, int64_t T, int64_t C, int64_t H, int64_t W, int64_t pt, int64_t ph,
                int64_t pw, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVENTFUL_B200_H */
