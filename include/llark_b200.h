/* llark_b200 C-ABI: the inner boundary of the B200-native LLark hot path.
 *
 * The reference (spotify-research/llark) has no C/FFI layer: its boundary is the Python API of
 * m2t/models/llamav2.py, m2t/data_modules.py and jukebox/main.py, and every FLOP is a PyTorch/cuBLAS/cuDNN
 * call. This header is what those call sites bind to instead: plain device pointers, sizes and a CUDA stream,
 * no torch types. Each entry point cites the reference call site (file:line under the reference root) whose
 * arithmetic it replaces. The Python mirror of the reference API lives in llark_b200/ and calls these through
 * ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions: all pointers are DEVICE pointers owned by the caller unless stated; functions never allocate,
 * never free, never synchronise; they enqueue work on `stream` and return LLARK_OK (0) or a negative LLARK_E*.
 * "bf16" buffers are raw __nv_bfloat16. Row-major everywhere. Thread-safe for distinct streams.
 */
#ifndef LLARK_B200_H
#define LLARK_B200_H

#include <cuda_runtime.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLARK_OK 0
#define LLARK_EINVAL (-1)
#define LLARK_ECUDA (-2)
#define LLARK_EDRIVER (-3)
#define LLARK_ENCCL (-4)

/* Last CUDA / driver / NCCL error code seen by the calling thread (for LLARK_ECUDA / EDRIVER / ENCCL). */
int llark_last_cuda_error(void);
/* One-time per-process set-up on the current device (TMA descriptor encoder, tile-counter ring of the GEMM scheduler).
 * The only entry point that allocates or synchronises; every other llark_* call just enqueues on the caller's stream. */
int llark_init(void);

/* ---------------------------------------------------------------------------------------------------------
 * GEMM core (tcgen05 / TMEM / TMA).  C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ R[M,N]), fp32 accumulate.
 *   a_mn_major = 0: A stored [M][lda] (K contiguous);  1: A stored [K][lda] (M contiguous).
 *   b_mn_major = 0: B stored [N][ldb] (K contiguous);  1: B stored [K][ldb] (N contiguous).
 *   in_fp32 = 0: A,B bf16 (kind::f16);  1: A,B fp32 consumed as TF32 (kind::tf32, K-major operands only).
 *   c_fp32: C (and R) fp32 instead of bf16.  bias: bf16[N] (fp32[N] if bias_fp32) or NULL.  R may alias C.
 *   max_ctas: 0 = one CTA per SM; >0 caps the persistent grid (tests / overlap with a comm stream).
 * Replaces: nn.Linear / addmm at m2t/models/llamav2.py:133 (mm_projector), :224 (LlamaModel q/k/v/o/gate/up/
 * down projections, via transformers), :312 (lm_head), their autograd dgrad/wgrad, and the prior's Conv1D
 * addmm in openai/jukebox reached from jukebox/main.py:108.
 */
int llark_gemm(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, void* C,
               long long ldc, int c_fp32, const void* bias, int bias_fp32, const void* R, long long ldr, int M, int N,
               int K, int in_fp32, int max_ctas, cudaStream_t stream);
/* Fused epilogues of the GEMM core (llark_gemm_ex; ep == NULL or kind 0 behaves exactly like llark_gemm):
 *   LLARK_EPI_QUICK_GELU  C = quick_gelu(acc + bias)            -- prior MLP `fc` (openai/jukebox MLP, reached from
 *                                                                  jukebox/main.py:108); replaces llark_quick_gelu (+ cast)
 *   LLARK_EPI_ROPE        columns [0, i1) of C (q | k heads of the fused QKV projection, head_dim 128) are rotated
 *                         with a = cos, b = sin (bf16 [S,128]), position = row % i0 -- transformers
 *                         apply_rotary_pos_emb reached from m2t/models/llamav2.py:224; replaces llark_rope(inverse=0)
 *   LLARK_EPI_SWIGLU_BWD  acc = d(act) of the down projection's dgrad; a = gu bf16 [M, ld] (gate | up, N apart);
 *                         C[:, c] = dgate, C[:, N + c] = dup (C is dgu [M, ldc >= 2N]); replaces llark_swiglu_bwd
 * Results are bit-identical to the unfused sequences (same bf16 rounding points). */
enum { LLARK_EPI_NONE = 0, LLARK_EPI_QUICK_GELU = 1, LLARK_EPI_ROPE = 2, LLARK_EPI_SWIGLU_BWD = 3 };
typedef struct LlarkGemmEpilogue {
    int kind;
    const void* a;
    const void* b;
    long long ld;
    int i0, i1;
} LlarkGemmEpilogue;
int llark_gemm_ex(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, void* C,
                  long long ldc, int c_fp32, const void* bias, int bias_fp32, const void* R, long long ldr, int M, int N,
                  int K, int in_fp32, int max_ctas, const LlarkGemmEpilogue* ep, cudaStream_t stream);
void llark_gemm_clear_cache(void);
/* Kernel variant: 1 = CTA-pair kernel (cta_group::2, 256x256 tile per 2-CTA cluster; default), 0 = single-CTA
 * kernel (128x256 tile). Default comes from the environment variable LLARK_GEMM_2CTA (unset or != "0" -> 1). */
void llark_gemm_set_mode(int two_cta);
int llark_gemm_get_mode(void);

/* ---------------------------------------------------------------------------------------------------------
 * Embedding gather + audio splice (m2t/models/llamav2.py:124 embed_tokens, :141-222 splice loop).
 * llark_splice_scan finds each sample's <audio_start> position on the device and raises the reference's two
 * ValueError conditions as bits in *flags (1: start/end count mismatch :146-152; 2: end token not at
 * start+frames+1 :169-175; 4: more than one audio segment, unsupported) instead of a host sync per sample.
 *   ids int64 [B,S]; frames int32 [B] (audio frames of sample b); start_pos int32 [B] out (-1 = no audio).
 * llark_embed_splice_fwd: out[b,s,:] = audio[audio_off[b] + s-pos-1, :] inside the audio block else table[ids].
 *   table bf16 [vocab,H]; audio bf16 [sum(frames),H] (projected features); audio_off int64 [B] row offsets.
 * llark_embed_splice_bwd: daudio <- dout audio rows; dtable[ids] += dout for the two marker rows only when
 *   detach_text != 0 (the orig_embeds_params path, :176-198), else for every text row (:199-207).
 */
int llark_splice_scan(const long long* ids, int B, int S, long long start_id, long long end_id, const int* frames,
                      int* start_pos, int* flags, cudaStream_t stream);
int llark_embed_splice_fwd(const long long* ids, const void* table, const void* audio, const long long* audio_off,
                           const int* frames, const int* start_pos, void* out, int B, int S, int H, long long vocab,
                           cudaStream_t stream);
int llark_embed_splice_bwd(const long long* ids, const void* dout, void* daudio, const long long* audio_off,
                           const int* frames, const int* start_pos, void* dtable, int B, int S, int H,
                           long long vocab, int detach_text, cudaStream_t stream);

/* dtype casts (autocast of fp32 audio_encodings to bf16 before mm_projector, llamav2.py:133). */
int llark_cast_f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream);
int llark_cast_bf16_to_f32(const void* in, float* out, long long n, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Llama decoder pieces (transformers LlamaRMSNorm / apply_rotary_pos_emb / LlamaMLP, reached at
 * m2t/models/llamav2.py:224). H must be a multiple of 256 (<= 16384).
 * rmsnorm_fwd: y = w * bf16(x * rsqrt(mean(x^2) + eps)); rstd fp32 [rows] saved for backward (may be NULL).
 * rmsnorm_bwd: dx = (dres ? dres : 0) + norm-path gradient; dw (+)= sum_rows dy * bf16(xhat).
 */
int llark_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long rows, int H, float eps,
                      cudaStream_t stream);
long long llark_rmsnorm_bwd_workspace(long long rows, int H);
int llark_rmsnorm_bwd(const void* x, const void* w, const void* dy, const float* rstd, const void* dres, void* dx,
                      void* dw, int accumulate_dw, long long rows, int H, void* workspace, cudaStream_t stream);
/* RoPE in place on the q and k heads of a packed [rows, row_stride] buffer (q heads | k heads | v heads),
 * cos/sin bf16 [S, D] tables (HF LlamaRotaryEmbedding values rounded to bf16), position = row % S.
 * inverse != 0 applies the transpose (backward). */
int llark_rope(void* qkv, const void* cos_t, const void* sin_t, long long rows, int S, int n_heads, int D,
               long long row_stride, int inverse, cudaStream_t stream);
/* SwiGLU on the fused gate|up projection gu [rows, 2I]: act = bf16(silu(gate)) * up. */
int llark_swiglu_fwd(const void* gu, void* act, long long rows, int I, cudaStream_t stream);
int llark_swiglu_bwd(const void* gu, const void* dact, void* dgu, long long rows, int I, cudaStream_t stream);

/* Causal attention with key-padding mask, head_dim 128. qkv bf16 [B*S, 3*n_heads*128] (RoPE applied),
 * key_mask uint8/bool [B,S] or NULL, out bf16 [B*S, n_heads*128], lse fp32 [B,n_heads,S].
 * Backward writes dq|dk|dv into dqkv (same packing). workspace: llark_attn_bwd_workspace bytes. */
int llark_attn_fwd(const void* qkv, const unsigned char* key_mask, void* out, float* lse, int B, int S, int n_heads,
                   int head_dim, float scale, cudaStream_t stream);
/* ALiBi variant of the causal attention for the CLAP + MPT-1B path (m2t/llava/model/mpt/attention.py:31-83,458-484):
 * scores = q.k * scale + slopes[h] * (key - (S - 1)); packed qkv WITHOUT RoPE; slopes fp32 [n_heads]; otherwise the contract of
 * llark_attn_fwd / llark_attn_bwd (workspace: llark_attn_bwd_workspace bytes). S <= 4096. */
int llark_attn_alibi_fwd(const void* qkv, const unsigned char* key_mask, void* out, float* lse, const float* slopes,
                         int B, int S, int n_heads, int head_dim, float scale, cudaStream_t stream);
int llark_attn_alibi_bwd(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout,
                         const float* lse, void* dqkv, const float* slopes, int B, int S, int n_heads, int head_dim,
                         float scale, void* workspace, cudaStream_t stream);
/* Bias-free LayerNorm over bf16 rows with row strides (elements) ldx / ldy: y = bf16(((x - mean) * rstd) * w), fp32
 * statistics written to mean / rstd (may be NULL). LPLayerNorm (m2t/llava/model/mpt/norm.py:26-52) and the q_ln / k_ln
 * of attention.py:289-292. Backward: dx = [dres +] rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w; dw (bf16
 * [W], may be NULL) receives (or accumulates) sum_rows dy * xhat; dres is a dense [rows, W] tensor or NULL. */
int llark_ln_fwd(const void* x, long long ldx, const void* w, void* y, long long ldy, float* mean, float* rstd,
                 long long rows, int W, float eps, cudaStream_t stream);
long long llark_ln_bwd_workspace(long long rows, int W);
int llark_ln_bwd(const void* x, long long ldx, const void* w, const void* dy, long long lddy, const float* mean,
                 const float* rstd, const void* dres, void* dx, long long lddx, void* dw, int accumulate_dw, long long rows,
                 int W, void* workspace, cudaStream_t stream);
/* Exact GELU (nn.GELU(approximate="none"), m2t/llava/model/mpt/blocks.py:14) and its backward du = dh * (Phi(u) +
 * u * phi(u)); bf16, n a multiple of 8. */
int llark_gelu_fwd(const void* u, void* h, long long n, cudaStream_t stream);
int llark_gelu_bwd(const void* u, const void* dh, void* du, long long n, cudaStream_t stream);
/* Skinny linear layer of the decode step: C[M,N] = A[M,K] . W[N,K]^T (+ bias[N]) (+ R[M,N]), 1 <= M <= 8, bf16, fp32
 * accumulation; K and the leading dimensions (elements) multiples of 8. Streams W once (HBM-bound SIMT kernel; the tcgen05
 * GEMM core would run 128-row tiles of padding). Same arithmetic contract as llark_gemm with these arguments
 * (transformers LlamaDecoderLayer linears reached from m2t/models/llamav2.py:224 during generate, m2t/infer.py:146-152). */
int llark_gemv(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, const void* bias,
               const void* R, long long ldr, int M, int N, int K, cudaStream_t stream);
/* Single-token decode step against a KV cache (transformers' cached LlamaAttention behind m2t/infer.py's
 * model.generate): q [B, n_heads*128] (RoPE applied), k_cache / v_cache [B, s_max, n_heads*128] bf16, lens int32 [B] =
 * valid cached positions per row including the new token (<= s_max <= 4096), out [B, n_heads*128]. */
int llark_attn_decode(const void* q, const void* k_cache, const void* v_cache, const int* lens, void* out, int B,
                      int s_max, int n_heads, int head_dim, float scale, cudaStream_t stream);
long long llark_attn_bwd_workspace(int B, int S, int n_heads);
int llark_attn_bwd(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout, const float* lse,
                   void* dqkv, int B, int S, int n_heads, int head_dim, float scale, void* workspace,
                   cudaStream_t stream);
/* llark_attn_bwd with the inverse rotary embedding of the dq | dk heads fused into the backward kernels' epilogues
 * (rope_cos / rope_sin: bf16 [S, 128], position = row % S): dqkv is then the gradient of the PRE-RoPE QKV projection
 * output, bit-identical to llark_attn_bwd followed by llark_rope(inverse = 1) -- one pass over dqkv less per layer. */
int llark_attn_bwd_rope(const void* qkv, const unsigned char* key_mask, const void* out, const void* dout,
                        const float* lse, void* dqkv, int B, int S, int n_heads, int head_dim, float scale,
                        const void* rope_cos, const void* rope_sin, void* workspace, cudaStream_t stream);

/* Column sum of a bf16 [rows, H] matrix into bf16 [H] (+= if accumulate): the mm_projector bias gradient. */
int llark_colsum(const void* x, void* out, long long rows, int H, int accumulate, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Shifted cross-entropy on bf16 logits (m2t/models/llamav2.py:314-325): target of row (b,s) is labels[b,s+1],
 * ignore_index -100, mean over valid targets. logits bf16 [B*S, ldv] (ldv >= V, multiple of 8).
 * fwd: row_loss,row_lse fp32 [B*S]; loss_out fp32[2] = {mean loss, n_valid}.
 * bwd: dlogits = (gscale or *gscale_ptr) / n_valid * (softmax - onehot); rows without target and pad columns 0.
 */
int llark_ce_fwd(const void* logits, long long ldv, int V, const long long* labels, int B, int S, float* row_loss,
                 float* row_lse, float* loss_out, cudaStream_t stream);
int llark_ce_bwd(const void* logits, long long ldv, int V, const long long* labels, int B, int S, const float* row_lse,
                 const float* loss_out, const float* gscale_ptr, float gscale, void* dlogits, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimizer over the flat bf16 parameter / gradient buckets (HF Trainer: clip_grad_norm_(1.0) + torch AdamW on
 * bf16 tensors, m2t/arguments.py:84, m2t/train.py:101-103). sumsq: fp32[1] device scalar = sum g^2 (before
 * grad_scale). Step applies g*grad_scale, the global-norm clip (max_norm <= 0 disables), then AdamW with a bf16
 * rounding after every elementwise op as torch's foreach implementation does. n % 8 == 0.
 */
long long llark_sumsq_workspace(void);
int llark_grad_sumsq(const void* g, long long n, float* out, int accumulate, void* workspace, cudaStream_t stream);
int llark_adamw_step(void* p, void* g, void* m, void* v, long long n, const float* sumsq, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale, float max_norm,
                     int zero_grad, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Gradient-bucket all-reduce (DDP, scripts/training/train_llark.sh:20-22,44): one ncclAllReduce(sum, bf16) over
 * the flat gradient bucket. The communicator is created from a 128-byte ncclUniqueId that rank 0 obtains with
 * llark_comm_unique_id and shares out of band (torch.distributed broadcast).
 */
int llark_comm_unique_id(void* id128_host);
int llark_comm_init(void** comm_out, const void* id128_host, int world, int rank);
int llark_comm_destroy(void* comm);
int llark_grad_bucket_allreduce(void* comm, void* grads_bf16, long long n, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Jukebox front end (jukebox/main.py:54-68 get_z, :101-110 get_final_activations, :113-130 windowed_average;
 * arithmetic of openai/jukebox@08efbbc). All fp32 activations unless noted.
 */
/* Strided down-conv Conv1d(c_in -> c_out, k=4, s=2, p=1) + bias. x [N, c_in, T] -> y [N, c_out, T/2]. */
int llark_jb_conv_down(const float* x, const float* w, const float* bias, float* y, int N, int c_in, int c_out, int T,
                       cudaStream_t stream);
/* Fused ResConv1DBlock: y = x + W2 * relu(W1 (*dil) relu(x) + b1) + b2, width 32 channels. x,y [N,32,T]. */
int llark_jb_resblock(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y,
                      int N, int T, int dilation, cudaStream_t stream);
/* Conv1d(c_in -> c_out, k=3, s=1, p=1) + bias (block output projection 32 -> 64). */
int llark_jb_conv_k3(const float* x, const float* w, const float* bias, float* y, int N, int c_in, int c_out, int T,
                     cudaStream_t stream);
/* Nearest-code search: x [N,64,T] (channels-first), codebook [n_codes,64]; codes int64 [N,T].
 * dist = |x|^2 - 2 x.k + |k|^2 in fp32, first minimum wins (torch.min semantics). */
int llark_jb_vq_encode(const float* x, const float* codebook, long long* codes, int N, int T, int n_codes,
                       cudaStream_t stream);
/* LayerNorm over the last dim (eps 1e-5, affine), fp32 in, bf16 or fp32 out. */
int llark_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_bf16, long long rows,
                        int W, float eps, cudaStream_t stream);
/* x0[t,:] = (t == 0 ? y_cond : emb[codes[t-1]]) + pos_emb[t] + x_cond[t]   (shift-right + conditioning). */
int llark_jb_embed_tokens(const long long* codes, const float* emb, const float* y_cond, const float* pos_emb,
                          const float* x_cond, float* out, int N, int T, int W, cudaStream_t stream);
/* Jukebox factored attention on qkv [N*T, 3*heads*dh] (fp32): pattern 0 = block (causal inside blocks of
 * `block`), 1 = transposed block (causal across blocks at equal offset), 2 = previous block (full, block 0 sees
 * nothing -> zeros). out fp32 [N*T, heads*dh]. */
int llark_jb_sparse_attn(const float* qkv, float* out, int N, int T, int heads, int dh, int block, int pattern,
                         cudaStream_t stream);
/* Same contract on tensor cores (mma.sync m16n8k8: operands rounded to TF32, fp32 accumulate and softmax): groups of
 * 64 or 128 rows, even dh <= 152. out: fp32, or bf16 when out_bf16 != 0 (the bf16 prior's cast fused in). Returns
 * LLARK_EINVAL for a geometry outside that envelope (use llark_jb_sparse_attn then). */
int llark_jb_sparse_attn_tc(const float* qkv, void* out, int out_bf16, int N, int T, int heads, int dh, int block,
                            int pattern, cudaStream_t stream);
/* y = x * sigmoid(1.702 x) in place-safe form. */
int llark_quick_gelu(const float* x, float* y, long long n, cudaStream_t stream);
/* out = a + b + c (c may be NULL). */
int llark_add3(const float* a, const float* b, const float* c, float* out, long long n, cudaStream_t stream);
/* AvgPool1d(frame, stride frame, floor) over time: acts [T, W] -> out [T/frame, W] (jukebox/main.py:113-130). */
int llark_jb_pool(const float* acts, float* out, int T_valid, int W, int frame, cudaStream_t stream);
/* Timing conditioning x_cond of the top-level prior (upstream LabelConditioner's pos_emb, reached from
 * jukebox/main.py:93): out[t,:] = (total_row + abs_table[abs_bins[t],:]) + rel_table[rel_bins[t],:], fp32 [T, W]. */
int llark_jb_label_cond(const float* total_row, const float* abs_table, const float* rel_table, const int* abs_bins,
                        const int* rel_bins, float* out, int T, int W, cudaStream_t stream);
/* the same pool over N clips in one launch: acts [N, clip_stride/W rows, W], out [N, T_valid/frame, W] */
int llark_jb_pool_batched(const float* acts, float* out, int N, long long clip_stride, int T_valid, int W, int frame,
                          cudaStream_t stream);

/* ---- fp32 verification path of the Llama forward (csrc/verify_fp32.cu): what llark_b200/verify.py runs to check
 * north_star's "fp32 logits within 1e-3 relative" against the reference's fp32 run of m2t/models/llamav2.py:259-337
 * (transformers LlamaModel: LlamaRMSNorm, rotate_half RoPE, eager softmax attention, SiLU(gate)*up, shifted mean CE).
 * All tensors fp32. The linears are llark_gemm(in_fp32=1) composed as 3xTF32 from llark_v32_split_tf32's halves. */
int llark_v32_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t stream);
int llark_v32_embed_splice_fwd(const long long* ids, const float* table, const float* audio, const long long* audio_off,
                               const int* frames, const int* start_pos, float* out, int B, int S, int H, long long vocab,
                               cudaStream_t stream);
int llark_v32_rmsnorm_fwd(const float* x, const float* w, float* y, long long rows, int H, float eps, cudaStream_t stream);
int llark_v32_rope(float* qkv, const float* cos_t, const float* sin_t, long long rows, int S, int n_rot_heads, int D,
                   long long ld, cudaStream_t stream);
int llark_v32_swiglu_fwd(const float* gu, float* act, long long rows, int I, cudaStream_t stream);
int llark_v32_attn_fwd(const float* qkv, const unsigned char* key_mask, float* out, int B, int S, int n_heads, int D,
                       float scale, cudaStream_t stream);
int llark_v32_ce_fwd(const float* logits, long long ldv, int V, const long long* labels, int B, int S, float* row_loss,
                     float* loss_out, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LLARK_B200_H */
