"""Thin torch-tensor wrappers over the C-ABI (include/llark_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every op below passes raw device
pointers + the stream to libllark_b200.so. There is no eager/PyTorch fallback: on a machine without the built
extension or without a GPU these raise.
"""
import ctypes
import math

import torch

from ._lib import EPI_QUICK_GELU, EPI_ROPE, EPI_SWIGLU_BWD, GemmEpilogue, check, lib

BF16 = torch.bfloat16

# bench.py sets this to a list: every gemm() launch then appends (start_event, end_event, flops, shape) recorded on the
# launching stream, which is how the roofline of the dominant kernel is measured live inside the timed region.
GEMM_PROFILE = None


_INITED = False


def init():
    """llark_init(): the one-time allocation + synchronisation of the library (tile-counter ring, TMA encoder), on the
    current device. Called once before the first GEMM so that no op of the hot loop allocates or synchronises."""
    global _INITED
    if not _INITED:
        check(lib().llark_init(), "llark_init")
        _INITED = True


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype=None):
    if not t.is_cuda:
        raise RuntimeError("llark_b200 ops need CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t


def gemm(a, b, *, a_mn=False, b_mn=False, M=None, N=None, K=None, out=None, out_dtype=None, bias=None,
         residual=None, lda=None, ldb=None, ldc=None, ldr=None, max_ctas=0, quick_gelu=False, rope=None,
         swiglu_bwd=None):
    """C[M,N] = A[M,K] @ B[N,K]^T (+bias) (+residual), optionally with one fused epilogue (llark_gemm_ex):

    quick_gelu=True            C = quick_gelu(acc + bias)
    rope=(cos, sin, S, cols)   columns [0, cols) rotated per 128-wide head, position = row % S (fused-QKV projection)
    swiglu_bwd=gu              acc is d(act) [M, I]; `out` must be the dgu buffer [M, 2I]: out = (dgate | dup)

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] (b_mn=False) or [K,N] (b_mn=True); row-major with
    leading dimension = last-dim stride unless lda/ldb given. bf16 inputs use kind::f16, fp32 inputs kind::tf32.
    """
    _req(a), _req(b)
    if not _INITED:
        init()
    in_fp32 = a.dtype == torch.float32
    if a.dtype != b.dtype or a.dtype not in (BF16, torch.float32):
        raise TypeError("gemm inputs must both be bf16 or both fp32")
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    lda = lda if lda is not None else a.stride(0)
    ldb = ldb if ldb is not None else b.stride(0)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype or (torch.float32 if in_fp32 else BF16))
    ldc = ldc if ldc is not None else out.stride(0)
    c_fp32 = out.dtype == torch.float32
    if residual is not None:
        if residual.dtype != out.dtype:
            raise TypeError("residual dtype must equal output dtype")
        ldr = ldr if ldr is not None else residual.stride(0)
    else:
        ldr = 0
    bias_fp32 = bias is not None and bias.dtype == torch.float32
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    ep = None
    if quick_gelu:
        ep = GemmEpilogue(EPI_QUICK_GELU, None, None, 0, 0, 0)
    elif rope is not None:
        cos_t, sin_t, S_rope, cols = rope
        ep = GemmEpilogue(EPI_ROPE, cos_t.data_ptr(), sin_t.data_ptr(), 0, int(S_rope), int(cols))
    elif swiglu_bwd is not None:
        ep = GemmEpilogue(EPI_SWIGLU_BWD, swiglu_bwd.data_ptr(), None, swiglu_bwd.stride(0), 0, 0)
    if ep is None:
        rc = lib().llark_gemm(_p(a), int(a_mn), lda, _p(b), int(b_mn), ldb, _p(out), ldc, int(c_fp32), _p(bias),
                              int(bias_fp32), _p(residual), ldr, M, N, K, int(in_fp32), max_ctas, _stream())
    else:
        rc = lib().llark_gemm_ex(_p(a), int(a_mn), lda, _p(b), int(b_mn), ldb, _p(out), ldc, int(c_fp32), _p(bias),
                                 int(bias_fp32), _p(residual), ldr, M, N, K, int(in_fp32), max_ctas, ctypes.byref(ep),
                                 _stream())
    check(rc, f"llark_gemm(M={M},N={N},K={K})")
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K, (M, N, K, int(a_mn), int(b_mn))))
    return out


def cast_f32_to_bf16(x, out=None):
    _req(x, torch.float32)
    x = x.contiguous()
    out = out if out is not None else torch.empty(x.shape, device=x.device, dtype=BF16)
    check(lib().llark_cast_f32_to_bf16(_p(x), _p(out), x.numel(), _stream()), "cast_f32_to_bf16")
    return out


def cast_bf16_to_f32(x, out=None):
    _req(x, BF16)
    x = x.contiguous()
    out = out if out is not None else torch.empty(x.shape, device=x.device, dtype=torch.float32)
    check(lib().llark_cast_bf16_to_f32(_p(x), _p(out), x.numel(), _stream()), "cast_bf16_to_f32")
    return out


def splice_scan(ids, start_id, end_id, frames, start_pos, flags):
    B, S = ids.shape
    check(lib().llark_splice_scan(_p(ids), B, S, start_id, end_id, _p(frames), _p(start_pos), _p(flags), _stream()),
          "splice_scan")


def embed_splice_fwd(ids, table, audio, audio_off, frames, start_pos, out):
    B, S = ids.shape
    H = table.shape[1]
    check(lib().llark_embed_splice_fwd(_p(ids), _p(table), _p(audio), _p(audio_off), _p(frames), _p(start_pos),
                                       _p(out), B, S, H, table.shape[0], _stream()), "embed_splice_fwd")
    return out


def embed_splice_bwd(ids, dout, daudio, audio_off, frames, start_pos, dtable, vocab, detach_text):
    B, S = ids.shape
    H = dout.shape[-1]
    check(lib().llark_embed_splice_bwd(_p(ids), _p(dout), _p(daudio), _p(audio_off), _p(frames), _p(start_pos),
                                       _p(dtable), B, S, H, vocab, int(detach_text), _stream()), "embed_splice_bwd")


def rmsnorm_fwd(x, w, eps, y=None, rstd=None):
    rows, H = x.shape
    y = y if y is not None else torch.empty_like(x)
    check(lib().llark_rmsnorm_fwd(_p(x), _p(w), _p(y), _p(rstd), rows, H, eps, _stream()), "rmsnorm_fwd")
    return y


def rmsnorm_bwd(x, w, dy, rstd, dres, dx, dw, accumulate_dw, workspace):
    rows, H = x.shape
    need = lib().llark_rmsnorm_bwd_workspace(rows, H)
    if workspace.numel() * workspace.element_size() < need:
        raise RuntimeError("rmsnorm_bwd workspace too small")
    check(lib().llark_rmsnorm_bwd(_p(x), _p(w), _p(dy), _p(rstd), _p(dres), _p(dx), _p(dw), int(accumulate_dw), rows,
                                  H, _p(workspace), _stream()), "rmsnorm_bwd")
    return dx


def rope_(qkv, cos_t, sin_t, S, n_heads, D, inverse=False):
    rows = qkv.shape[0]
    check(lib().llark_rope(_p(qkv), _p(cos_t), _p(sin_t), rows, S, n_heads, D, qkv.stride(0), int(inverse), _stream()),
          "rope")
    return qkv


def swiglu_fwd(gu, act=None):
    rows, I2 = gu.shape
    act = act if act is not None else torch.empty((rows, I2 // 2), device=gu.device, dtype=BF16)
    check(lib().llark_swiglu_fwd(_p(gu), _p(act), rows, I2 // 2, _stream()), "swiglu_fwd")
    return act


def swiglu_bwd(gu, dact, dgu=None):
    rows, I2 = gu.shape
    dgu = dgu if dgu is not None else torch.empty_like(gu)
    check(lib().llark_swiglu_bwd(_p(gu), _p(dact), _p(dgu), rows, I2 // 2, _stream()), "swiglu_bwd")
    return dgu


def attn_fwd(qkv, key_mask, B, S, n_heads, head_dim, out=None, lse=None):
    out = out if out is not None else torch.empty((B * S, n_heads * head_dim), device=qkv.device, dtype=BF16)
    lse = lse if lse is not None else torch.empty((B, n_heads, S), device=qkv.device, dtype=torch.float32)
    check(lib().llark_attn_fwd(_p(qkv), _p(key_mask), _p(out), _p(lse), B, S, n_heads, head_dim,
                               1.0 / math.sqrt(head_dim), _stream()), "attn_fwd")
    return out, lse


def attn_decode(q, k_cache, v_cache, lens, n_heads, head_dim, out=None):
    """One query token per row against the KV cache: q [B, H], caches [B, s_max, H], lens int32 [B] -> out [B, H]."""
    B, s_max = k_cache.shape[0], k_cache.shape[1]
    out = out if out is not None else torch.empty((B, n_heads * head_dim), device=q.device, dtype=BF16)
    check(lib().llark_attn_decode(_p(q), _p(k_cache), _p(v_cache), _p(lens), _p(out), B, s_max, n_heads, head_dim,
                                  1.0 / math.sqrt(head_dim), _stream()), "attn_decode")
    return out


def attn_bwd(qkv, key_mask, out, dout, lse, B, S, n_heads, head_dim, dqkv=None, workspace=None, rope=None):
    """dqkv = (dq | dk | dv). rope=(cos, sin) (bf16 [>= S, head_dim]): the inverse rotary embedding of the dq | dk heads is
    fused into the backward kernels (llark_attn_bwd_rope), bit-identical to attn_bwd followed by rope_(inverse=True)."""
    dqkv = dqkv if dqkv is not None else torch.empty_like(qkv)
    need = lib().llark_attn_bwd_workspace(B, S, n_heads)
    if workspace is None:
        workspace = torch.empty(need, device=qkv.device, dtype=torch.uint8)
    if rope is not None:
        cos_t, sin_t = rope
        check(lib().llark_attn_bwd_rope(_p(qkv), _p(key_mask), _p(out), _p(dout), _p(lse), _p(dqkv), B, S, n_heads,
                                        head_dim, 1.0 / math.sqrt(head_dim), _p(_req(cos_t, BF16)), _p(_req(sin_t, BF16)),
                                        _p(workspace), _stream()), "attn_bwd_rope")
        return dqkv
    check(lib().llark_attn_bwd(_p(qkv), _p(key_mask), _p(out), _p(dout), _p(lse), _p(dqkv), B, S, n_heads, head_dim,
                               1.0 / math.sqrt(head_dim), _p(workspace), _stream()), "attn_bwd")
    return dqkv


def gemv(x, w, bias=None, residual=None, out=None, N=None):
    """y[M, N] = x[M, K] @ w[N, K]^T (+ bias) (+ residual) for M <= 8 rows: the weight-streaming kernel of the decode step
    (csrc/gemv.cu). `out` may be wider than N (padded logits buffer)."""
    M, K = x.shape
    N = int(N) if N is not None else w.shape[0]
    out = out if out is not None else torch.empty((M, N), device=x.device, dtype=BF16)
    check(lib().llark_gemv(_p(x), x.stride(0), _p(w), w.stride(0), _p(out), out.stride(0), _p(bias), _p(residual),
                           residual.stride(0) if residual is not None else 0, M, N, K, _stream()), "gemv")
    return out


# ------------------------------------------------------------------------------------------------ MPT variant (staged)
def attn_alibi_fwd(qkv, key_mask, B, S, n_heads, head_dim, slopes, out=None, lse=None):
    out = out if out is not None else torch.empty((B * S, n_heads * head_dim), device=qkv.device, dtype=BF16)
    lse = lse if lse is not None else torch.empty((B, n_heads, S), device=qkv.device, dtype=torch.float32)
    check(lib().llark_attn_alibi_fwd(_p(qkv), _p(key_mask), _p(out), _p(lse), _p(slopes), B, S, n_heads, head_dim,
                                     1.0 / math.sqrt(head_dim), _stream()), "attn_alibi_fwd")
    return out, lse


def attn_alibi_bwd(qkv, key_mask, out, dout, lse, B, S, n_heads, head_dim, slopes, dqkv=None):
    dqkv = dqkv if dqkv is not None else torch.empty_like(qkv)
    ws = torch.empty(lib().llark_attn_bwd_workspace(B, S, n_heads), device=qkv.device, dtype=torch.uint8)
    check(lib().llark_attn_alibi_bwd(_p(qkv), _p(key_mask), _p(out), _p(dout), _p(lse), _p(dqkv), _p(slopes), B, S,
                                     n_heads, head_dim, 1.0 / math.sqrt(head_dim), _p(ws), _stream()), "attn_alibi_bwd")
    return dqkv


def ln_fwd(x, w, eps, mean=None, rstd=None, out=None, out_dtype=BF16):
    """Bias-free LayerNorm of the rows of a [rows, W] bf16 view (row stride = x.stride(0)); out may be a strided view."""
    rows, W = x.shape
    out = out if out is not None else torch.empty((rows, W), device=x.device, dtype=BF16)
    check(lib().llark_ln_fwd(_p(x), x.stride(0), _p(w), _p(out), out.stride(0), _p(mean), _p(rstd), rows, W, eps,
                             _stream()), "ln_fwd")
    return out


def ln_bwd(x, w, dy, mean, rstd, dres, dx, dw, accumulate_dw):
    rows, W = x.shape
    ws = torch.empty(lib().llark_ln_bwd_workspace(rows, W), device=x.device, dtype=torch.uint8) if dw is not None else None
    check(lib().llark_ln_bwd(_p(x), x.stride(0), _p(w), _p(dy), dy.stride(0), _p(mean), _p(rstd), _p(dres), _p(dx),
                             dx.stride(0), _p(dw), int(accumulate_dw), rows, W, _p(ws), _stream()), "ln_bwd")
    return dx


def gelu_fwd(u, out=None):
    out = out if out is not None else torch.empty_like(u)
    check(lib().llark_gelu_fwd(_p(u), _p(out), u.numel(), _stream()), "gelu_fwd")
    return out


def gelu_bwd(u, dh, out=None):
    out = out if out is not None else torch.empty_like(u)
    check(lib().llark_gelu_bwd(_p(u), _p(dh), _p(out), u.numel(), _stream()), "gelu_bwd")
    return out


def colsum(x, out, accumulate=False):
    rows, H = x.shape
    check(lib().llark_colsum(_p(x), _p(out), rows, H, int(accumulate), _stream()), "colsum")
    return out


def ce_fwd(logits, ldv, V, labels, B, S, row_loss, row_lse, loss_out):
    check(lib().llark_ce_fwd(_p(logits), ldv, V, _p(labels), B, S, _p(row_loss), _p(row_lse), _p(loss_out), _stream()),
          "ce_fwd")


def ce_bwd(logits, ldv, V, labels, B, S, row_lse, loss_out, gscale_ptr, gscale, dlogits):
    check(lib().llark_ce_bwd(_p(logits), ldv, V, _p(labels), B, S, _p(row_lse), _p(loss_out), _p(gscale_ptr),
                             float(gscale), _p(dlogits), _stream()), "ce_bwd")


def grad_sumsq(g, out, accumulate, workspace):
    check(lib().llark_grad_sumsq(_p(g), g.numel(), _p(out), int(accumulate), _p(workspace), _stream()), "grad_sumsq")


def adamw_step(p, g, m, v, sumsq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, max_norm=0.0,
               zero_grad=False):
    check(lib().llark_adamw_step(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(sumsq), lr, beta1, beta2, eps, weight_decay,
                                 step, grad_scale, max_norm, int(zero_grad), _stream()), "adamw_step")


# ------------------------------------------------------------------------------------------------ Jukebox front end
def jb_conv_down(x, w, b):
    N, c_in, T = x.shape
    c_out = w.shape[0]
    y = torch.empty((N, c_out, T // 2), device=x.device, dtype=torch.float32)
    check(lib().llark_jb_conv_down(_p(x), _p(w), _p(b), _p(y), N, c_in, c_out, T, _stream()), "jb_conv_down")
    return y


def jb_resblock(x, w1, b1, w2, b2, dilation):
    N, C, T = x.shape
    y = torch.empty_like(x)
    check(lib().llark_jb_resblock(_p(x), _p(w1), _p(b1), _p(w2), _p(b2), _p(y), N, T, dilation, _stream()),
          "jb_resblock")
    return y


def jb_conv_k3(x, w, b):
    N, c_in, T = x.shape
    c_out = w.shape[0]
    y = torch.empty((N, c_out, T), device=x.device, dtype=torch.float32)
    check(lib().llark_jb_conv_k3(_p(x), _p(w), _p(b), _p(y), N, c_in, c_out, T, _stream()), "jb_conv_k3")
    return y


def jb_vq_encode(x, codebook):
    N, E, T = x.shape
    codes = torch.empty((N, T), device=x.device, dtype=torch.int64)
    check(lib().llark_jb_vq_encode(_p(x), _p(codebook), _p(codes), N, T, codebook.shape[0], _stream()), "jb_vq_encode")
    return codes


def layernorm_fwd(x, gamma, beta, eps=1e-5, out_dtype=torch.float32, out=None):
    rows, W = x.shape
    out = out if out is not None else torch.empty((rows, W), device=x.device, dtype=out_dtype)
    check(lib().llark_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(out), int(out.dtype == BF16), rows, W, eps,
                                    _stream()), "layernorm_fwd")
    return out


def jb_embed_tokens(codes, emb, y_cond, pos_emb, x_cond):
    N, T = codes.shape
    W = emb.shape[1]
    out = torch.empty((N * T, W), device=codes.device, dtype=torch.float32)
    check(lib().llark_jb_embed_tokens(_p(codes), _p(emb), _p(y_cond), _p(pos_emb), _p(x_cond), _p(out), N, T, W,
                                      _stream()), "jb_embed_tokens")
    return out


def jb_sparse_attn(qkv, N, T, heads, dh, block, pattern, out=None, out_dtype=torch.float32, exact=False):
    """Factored attention of the prior. Default: the tensor-core kernel (TF32 operands, fp32 accumulate, fp32 or bf16
    output); exact=True (or a geometry outside that kernel's envelope) runs the fp32 SIMT kernel (fp32 output)."""
    own = out is None
    if own:
        out = torch.empty((N * T, heads * dh), device=qkv.device, dtype=torch.float32 if exact else out_dtype)
    if not exact:
        rc = lib().llark_jb_sparse_attn_tc(_p(qkv), _p(out), int(out.dtype == BF16), N, T, heads, dh, block, pattern,
                                           _stream())
        if rc != -1:  # LLARK_EINVAL = outside the envelope -> SIMT kernel below
            check(rc, "jb_sparse_attn_tc")
            return out
        if out.dtype != torch.float32:
            if not own:
                raise TypeError("the SIMT factored-attention kernel writes fp32")
            out = torch.empty((N * T, heads * dh), device=qkv.device, dtype=torch.float32)  # caller casts
    check(lib().llark_jb_sparse_attn(_p(qkv), _p(out), N, T, heads, dh, block, pattern, _stream()), "jb_sparse_attn")
    return out


def quick_gelu(x, out=None):
    out = out if out is not None else torch.empty_like(x)
    check(lib().llark_quick_gelu(_p(x), _p(out), x.numel(), _stream()), "quick_gelu")
    return out


def add3(a, b, c=None, out=None):
    out = out if out is not None else torch.empty_like(a)
    check(lib().llark_add3(_p(a), _p(b), _p(c), _p(out), a.numel(), _stream()), "add3")
    return out


def jb_label_cond(total_row, abs_table, rel_table, abs_bins, rel_bins):
    """x_cond [T, W] = (total_row + abs_table[abs_bins]) + rel_table[rel_bins] (fp32 tables, int32 bins on the device)."""
    T, W = abs_bins.numel(), abs_table.shape[1]
    out = torch.empty((T, W), device=abs_table.device, dtype=torch.float32)
    check(lib().llark_jb_label_cond(_p(_req(total_row, torch.float32)), _p(_req(abs_table, torch.float32)),
                                    _p(_req(rel_table, torch.float32)), _p(_req(abs_bins, torch.int32)),
                                    _p(_req(rel_bins, torch.int32)), _p(out), T, W, _stream()), "jb_label_cond")
    return out


def jb_pool_batched(acts, T_valid, frame):
    """acts fp32 [N, T, W] (contiguous) -> [N, T_valid // frame, W]: windowed mean over time of every clip, one launch."""
    _req(acts, torch.float32)
    N, T, W = acts.shape
    if not acts.is_contiguous() or T_valid > T:
        raise ValueError("jb_pool_batched needs a contiguous [N, T, W] tensor with T >= T_valid")
    out = torch.empty((N, T_valid // frame, W), device=acts.device, dtype=torch.float32)
    check(lib().llark_jb_pool_batched(_p(acts), _p(out), N, T * W, T_valid, W, frame, _stream()), "jb_pool_batched")
    return out


def jb_pool(acts, T_valid, frame):
    W = acts.shape[1]
    out = torch.empty((T_valid // frame, W), device=acts.device, dtype=torch.float32)
    check(lib().llark_jb_pool(_p(acts), _p(out), T_valid, W, frame, _stream()), "jb_pool")
    return out


# ------------------------------------------------------------------------------------------------ fp32 verification path
def _f32(t):
    return _req(t, torch.float32)


def v32_split_tf32(x):
    x = _f32(x).contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    check(lib().llark_v32_split_tf32(_p(x), _p(hi), _p(lo), x.numel(), _stream()), "v32_split_tf32")
    return hi, lo


def v32_embed_splice_fwd(ids, table, audio, audio_off, frames, start_pos, out):
    B, S = ids.shape
    H = table.shape[1]
    check(lib().llark_v32_embed_splice_fwd(_p(ids), _p(_f32(table)), _p(audio), _p(audio_off), _p(frames), _p(start_pos),
                                           _p(_f32(out)), B, S, H, table.shape[0], _stream()), "v32_embed_splice_fwd")
    return out


def v32_rmsnorm_fwd(x, w, eps):
    rows, H = x.shape
    y = torch.empty_like(_f32(x))
    check(lib().llark_v32_rmsnorm_fwd(_p(x), _p(_f32(w)), _p(y), rows, H, eps, _stream()), "v32_rmsnorm_fwd")
    return y


def v32_rope_(qkv, cos_t, sin_t, S, n_rot_heads, D):
    check(lib().llark_v32_rope(_p(_f32(qkv)), _p(_f32(cos_t)), _p(_f32(sin_t)), qkv.shape[0], S, n_rot_heads, D,
                               qkv.stride(0), _stream()), "v32_rope")
    return qkv


def v32_swiglu_fwd(gu):
    rows, I2 = gu.shape
    act = torch.empty((rows, I2 // 2), device=gu.device, dtype=torch.float32)
    check(lib().llark_v32_swiglu_fwd(_p(_f32(gu)), _p(act), rows, I2 // 2, _stream()), "v32_swiglu_fwd")
    return act


def v32_attn_fwd(qkv, key_mask, B, S, n_heads, head_dim):
    out = torch.empty((B * S, n_heads * head_dim), device=qkv.device, dtype=torch.float32)
    check(lib().llark_v32_attn_fwd(_p(_f32(qkv)), _p(key_mask), _p(out), B, S, n_heads, head_dim,
                                   1.0 / math.sqrt(head_dim), _stream()), "v32_attn_fwd")
    return out


def v32_ce_fwd(logits, V, labels, B, S):
    row_loss = torch.empty(B * S, device=logits.device, dtype=torch.float32)
    loss_out = torch.empty(2, device=logits.device, dtype=torch.float32)
    check(lib().llark_v32_ce_fwd(_p(_f32(logits)), logits.stride(0), V, _p(labels), B, S, _p(row_loss), _p(loss_out),
                                 _stream()), "v32_ce_fwd")
    return loss_out[0]
