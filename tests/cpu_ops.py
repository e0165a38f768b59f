"""TEST INFRASTRUCTURE ONLY: a plain-PyTorch CPU stand-in for the forward subset of `llark_b200.ops` (the ctypes layer
over libllark_b200.so), with the same call signatures. It exists so that the HOST logic of `llark_b200.engine` --
flat-bucket views, the order of the launches, splice bookkeeping, the KV-cache indexing and the position-offset RoPE
trick of the decode step -- can be exercised by the CPU test suite, which has no GPU. The product never imports this
module (tests/test_dist_cpu.py::test_no_cpu_fallback); the CUDA kernels themselves are tested in the `-m gpu` suite.
Arithmetic: fp32 math on bf16 tensors, rounded to bf16 where the kernels round."""
import math

import torch

BF16 = torch.bfloat16
GEMM_PROFILE = None


def lib():
    raise RuntimeError("the CPU ops stand-in has no C library")


def _round(x):
    return x.to(BF16)


def cast_f32_to_bf16(x, out=None):
    return x.to(BF16)


def gemm(a, b, *, a_mn=False, b_mn=False, M=None, N=None, K=None, out=None, out_dtype=None, bias=None, residual=None,
         lda=None, ldb=None, ldc=None, ldr=None, max_ctas=0, quick_gelu=False, rope=None, swiglu_bwd=None):
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    if M is not None:
        A = A[:M]
    if K is not None:
        A, Bm = A[:, :K], Bm[:K]
    if N is not None:
        Bm = Bm[:, :N]
    c = A @ Bm
    if bias is not None:
        c = c + bias.float()
    if quick_gelu:
        c = c / (1 + torch.exp(-1.702 * c))
    if rope is not None:
        cos_t, sin_t, S, cols = rope
        x = _round(c)
        rows = torch.arange(x.shape[0]) % S
        _rope_rows(x[:, :cols], cos_t[rows], sin_t[rows])
        c = x.float()
    if residual is not None:
        c = c + residual.float()[:, : c.shape[1]]
    dt = out.dtype if out is not None else (out_dtype or (torch.float32 if a.dtype == torch.float32 else BF16))
    c = c.to(dt)
    if out is not None:
        out[:, : c.shape[1]] = c
        return out
    return c


def _rope_rows(x, cos, sin):
    """x [rows, n_heads*128] (q | k heads), cos / sin [rows, 128]: HF apply_rotary_pos_emb in bf16, in place."""
    rows, W = x.shape
    v = x.view(rows, W // 128, 128)
    c, s = cos[:, None, :].to(BF16), sin[:, None, :].to(BF16)
    rot = torch.cat((-v[..., 64:], v[..., :64]), dim=-1)
    v.copy_((v * c) + (rot * s))


def rope_(qkv, cos_t, sin_t, S, n_heads, D, inverse=False):
    assert not inverse and D == 128
    rows = torch.arange(qkv.shape[0]) % S
    _rope_rows(qkv[:, : 2 * n_heads * D], cos_t[rows], sin_t[rows])
    return qkv


def splice_scan(ids, start_id, end_id, frames, start_pos, flags):
    for b in range(ids.shape[0]):
        st = (ids[b] == start_id).nonzero().flatten()
        en = (ids[b] == end_id).nonzero().flatten()
        if len(st) != len(en):
            flags[0] |= 1
        if len(st) == 0:
            start_pos[b] = -1
            continue
        if len(st) > 1:
            flags[0] |= 4
        p = int(st[-1])
        start_pos[b] = p
        e = p + int(frames[b]) + 1
        if e >= ids.shape[1] or ids[b, e] != end_id:
            flags[0] |= 2


def embed_splice_fwd(ids, table, audio, audio_off, frames, start_pos, out):
    B, S = ids.shape
    out.copy_(table[ids.clamp(0, table.shape[0] - 1)])
    if audio is not None:
        for b in range(B):
            p = int(start_pos[b])
            if p >= 0:
                F_ = int(frames[b])
                o = int(audio_off[b])
                out[b, p + 1: p + 1 + F_] = audio[o: o + F_]
    return out


def rmsnorm_fwd(x, w, eps, y=None, rstd=None):
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    if rstd is not None:
        rstd.copy_(r.flatten())
    return (w.float() * _round(xf * r).float()).to(BF16)


def attn_fwd(qkv, key_mask, B, S, n_heads, head_dim, out=None, lse=None):
    H = n_heads * head_dim
    q, k, v = qkv.float().view(B, S, 3, n_heads, head_dim).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(head_dim)
    allow = torch.ones(S, S, dtype=torch.bool).tril()[None, None]
    if key_mask is not None:
        allow = allow & key_mask.view(B, 1, 1, S).bool()
    sc = sc.masked_fill(~allow, float("-inf"))
    p = torch.softmax(sc, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * S, H)
    return o.to(BF16), torch.logsumexp(sc, dim=-1)


def attn_decode(q, k_cache, v_cache, lens, n_heads, head_dim, out=None):
    B, s_max, H = k_cache.shape
    res = torch.empty((B, H), dtype=BF16)
    for b in range(B):
        n = int(lens[b])
        qf = q[b].float().view(n_heads, head_dim)
        kf = k_cache[b, :n].float().view(n, n_heads, head_dim).permute(1, 0, 2)
        vf = v_cache[b, :n].float().view(n, n_heads, head_dim).permute(1, 0, 2)
        p = torch.softmax((kf @ qf[:, :, None]).squeeze(-1) / math.sqrt(head_dim), dim=-1)
        res[b] = (p[:, None, :] @ vf).squeeze(1).reshape(-1).to(BF16)
    return res


def swiglu_fwd(gu, act=None):
    I = gu.shape[1] // 2
    g, u = gu[:, :I].float(), gu[:, I:].float()
    return (_round(g * torch.sigmoid(g)).float() * u).to(BF16)
