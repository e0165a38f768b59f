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
                F_ = min(int(frames[b]), S - p - 1)  # the kernel only writes positions inside the sequence
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


# ------------------------------------------------------------------------------------------------ backward / optimizer
class _Lib:
    """Workspace-size queries of the C-ABI the engine makes through ops.lib()."""

    @staticmethod
    def llark_rmsnorm_bwd_workspace(rows, H):
        return 16

    @staticmethod
    def llark_attn_bwd_workspace(B, S, n_heads):
        return 16

    @staticmethod
    def llark_sumsq_workspace():
        return 16


def lib():  # noqa: F811  (replaces the placeholder above)
    return _Lib


_gemm_fwd = gemm


def gemm(a, b, **kw):  # noqa: F811
    """Adds the SwiGLU-backward epilogue and residual-into-out accumulation to the forward stand-in."""
    gu = kw.pop("swiglu_bwd", None)
    if gu is None:
        return _gemm_fwd(a, b, **kw)
    out = kw.pop("out")
    N = kw.get("N")
    d = _gemm_fwd(a, b, **kw).float()[:, :N]  # d(act), rounded to bf16 as the unfused GEMM stores it
    out.copy_(swiglu_bwd(gu, d.to(BF16)))
    return out


def swiglu_bwd(gu, dact, dgu=None):
    I = gu.shape[1] // 2
    g, u, d = gu[:, :I].float(), gu[:, I:].float(), dact.float()
    sg = torch.sigmoid(g)
    du = d * _round(g * sg).float()
    dg = d * u * (sg * (1 + g * (1 - sg)))
    return torch.cat([dg, du], dim=1).to(BF16)


def ce_fwd(logits, ldv, V, labels, B, S, row_loss, row_lse, loss_out):
    lg = logits.float().view(B, S, -1)[:, :, :V]
    tgt = torch.full((B, S), -100, dtype=torch.int64)
    tgt[:, :-1] = labels[:, 1:]
    lse = torch.logsumexp(lg, dim=-1)
    valid = (tgt >= 0) & (tgt < V)
    picked = torch.gather(lg, 2, tgt.clamp(0, V - 1)[..., None]).squeeze(-1)
    loss_rows = torch.where(valid, lse - picked, torch.zeros_like(lse))
    row_loss.copy_(loss_rows.flatten())
    row_lse.copy_(torch.where(valid, lse, torch.full_like(lse, float("inf"))).flatten())
    n = valid.sum().clamp(min=1).float()
    loss_out[0] = loss_rows.sum() / n
    loss_out[1] = n


def ce_bwd(logits, ldv, V, labels, B, S, row_lse, loss_out, gscale_ptr, gscale, dlogits):
    lg = logits.float().view(B * S, -1)
    tgt = torch.full((B, S), -100, dtype=torch.int64)
    tgt[:, :-1] = labels[:, 1:]
    tgt = tgt.flatten()
    scale = (gscale_ptr[0].item() if gscale_ptr is not None else gscale) / loss_out[1].item()
    out = torch.zeros_like(lg)
    valid = (tgt >= 0) & (tgt < V)
    p = torch.softmax(lg[:, :V], dim=-1)
    p[torch.arange(B * S)[valid], tgt[valid]] -= 1.0
    out[:, :V] = torch.where(valid[:, None], p * scale, torch.zeros_like(p))
    dlogits.copy_(out.to(BF16).view_as(dlogits))


def rmsnorm_bwd(x, w, dy, rstd, dres, dx, dw, accumulate_dw, workspace):
    xf, dyf, wf = x.float(), dy.float(), w.float()
    r = rstd.float()[:, None]
    xhat = xf * r
    dxhat = dyf * wf
    g = r * (dxhat - xhat * (dxhat * xhat).mean(-1, keepdim=True))
    if dres is not None:
        g = g + dres.float()
    dx.copy_(g.to(BF16))
    if dw is not None:
        d = (dyf * _round(xhat).float()).sum(0)
        dw.copy_(((dw.float() if accumulate_dw else 0) + d).to(BF16))
    return dx


def attn_bwd(qkv, key_mask, out, dout, lse, B, S, n_heads, head_dim, dqkv=None, workspace=None, rope=None):
    if rope is not None:  # fused form: inverse RoPE of dq | dk applied on the way out
        g = attn_bwd(qkv, key_mask, out, dout, lse, B, S, n_heads, head_dim)
        return rope_(g, rope[0], rope[1], S, n_heads, head_dim, inverse=True)
    x = qkv.float().clone().requires_grad_(True)
    with torch.enable_grad():
        q, k, v = x.view(B, S, 3, n_heads, head_dim).permute(2, 0, 3, 1, 4)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(head_dim)
        allow = torch.ones(S, S, dtype=torch.bool).tril()[None, None]
        if key_mask is not None:
            allow = allow & key_mask.view(B, 1, 1, S).bool()
        p = torch.softmax(sc.masked_fill(~allow, float("-inf")), dim=-1)
        o = (p @ v).permute(0, 2, 1, 3).reshape(B * S, n_heads * head_dim)
        o.backward(dout.float())
    return x.grad.to(BF16)


_rope_fwd = rope_


def rope_(qkv, cos_t, sin_t, S, n_heads, D, inverse=False):  # noqa: F811
    if not inverse:
        return _rope_fwd(qkv, cos_t, sin_t, S, n_heads, D)
    rows = torch.arange(qkv.shape[0]) % S
    W = 2 * n_heads * D
    v = qkv[:, :W].view(qkv.shape[0], W // 128, 128)
    c, s = cos_t[rows][:, None, :].to(BF16), sin_t[rows][:, None, :].to(BF16)
    y = v * s
    v.copy_((v * c) + torch.cat((y[..., 64:], -y[..., :64]), dim=-1))  # transpose of x*cos + rotate_half(x)*sin
    return qkv


def embed_splice_bwd(ids, dout, daudio, audio_off, frames, start_pos, dtable, vocab, detach_text):
    B, S = ids.shape
    for b in range(B):
        p = int(start_pos[b]) if start_pos is not None else -1
        F_ = int(frames[b]) if (frames is not None and p >= 0) else 0
        for s_ in range(S):
            if p >= 0 and p < s_ <= p + F_:
                if daudio is not None:
                    daudio[int(audio_off[b]) + s_ - p - 1] = dout[b, s_]
                continue
            if dtable is None:
                continue
            marker = p >= 0 and (s_ == p or s_ == p + F_ + 1)
            if detach_text and not marker and p >= 0:  # an audio-less sample stays undetached (llamav2.py:162-208)
                continue
            t = int(ids[b, s_])
            dtable[t] = (dtable[t].float() + dout[b, s_].float()).to(BF16)


def colsum(x, out, accumulate=False):
    out.copy_(((out.float() if accumulate else 0) + x.float().sum(0)).to(BF16))
    return out


def grad_sumsq(g, out, accumulate, workspace):
    out[0] = (out[0] if accumulate else 0) + g.float().pow(2).sum()


def adamw_step(p, g, m, v, sumsq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, max_norm=0.0,
               zero_grad=False):
    clip = 1.0
    if max_norm > 0:
        clip = min(1.0, max_norm / (math.sqrt(sumsq[0].item()) * grad_scale + 1e-6))
    gf = _round(g.float() * (grad_scale * clip)).float()
    pf = _round(p.float() * (1 - lr * weight_decay)).float()
    mf = _round(torch.lerp(m.float(), gf, 1 - beta1)).float()
    vf = _round(_round(v.float() * beta2).float() + (1 - beta2) * gf * gf).float()
    bc1, bc2s = 1 - beta1 ** step, math.sqrt(1 - beta2 ** step)
    den = _round(_round(_round(vf.sqrt()).float() / bc2s).float() + eps).float()
    p.copy_(_round(pf - (lr / bc1) * mf / den))
    m.copy_(mf.to(BF16))
    v.copy_(vf.to(BF16))
    if zero_grad:
        g.zero_()


# ------------------------------------------------------------------------------------------------ Jukebox front end
import torch.nn.functional as _F  # noqa: E402


def jb_conv_down(x, w, b):
    return _F.conv1d(x, w, b, stride=2, padding=1)


def jb_resblock(x, w1, b1, w2, b2, dilation):
    h = _F.conv1d(torch.relu(x), w1, b1, padding=dilation, dilation=dilation)
    return x + _F.conv1d(torch.relu(h), w2, b2)


def jb_conv_k3(x, w, b):
    return _F.conv1d(x, w, b, padding=1)


def jb_vq_encode(x, codebook):
    N, E, T = x.shape
    flat = x.permute(0, 2, 1).reshape(N * T, E)
    d = (flat * flat).sum(1, keepdim=True) - 2 * flat @ codebook.t() + (codebook * codebook).sum(1)[None]
    return d.argmin(dim=1).view(N, T)


def layernorm_fwd(x, gamma, beta, eps=1e-5, out_dtype=torch.float32, out=None):
    return _F.layer_norm(x.float(), x.shape[-1:], gamma, beta, eps).to(out_dtype)


def jb_embed_tokens(codes, emb, y_cond, pos_emb, x_cond):
    N, T = codes.shape
    x = emb[codes]
    x = torch.roll(x, 1, dims=1)
    x[:, 0] = y_cond
    return ((x + pos_emb[None, :T]) + x_cond[None, :T]).reshape(N * T, -1)


def jb_sparse_attn(qkv, N, T, heads, dh, block, pattern, out=None, out_dtype=torch.float32, exact=False):
    A = heads * dh
    q, k, v = qkv.float().view(N, T, 3, heads, dh).permute(2, 0, 3, 1, 4)  # [N, heads, T, dh]
    nb = T // block

    def dense(q_, k_, v_, causal):
        s = (q_ @ k_.transpose(-1, -2)) / math.sqrt(dh)
        if causal:
            n = s.shape[-1]
            s = s.masked_fill(~torch.ones(n, n, dtype=torch.bool).tril(), float("-inf"))
        return torch.softmax(s, dim=-1) @ v_

    if pattern == 0:
        f = lambda t: t.reshape(N, heads, nb, block, dh)  # noqa: E731
        a = dense(f(q), f(k), f(v), True).reshape(N, heads, T, dh)
    elif pattern == 1:
        f = lambda t: t.reshape(N, heads, nb, block, dh).transpose(2, 3)  # noqa: E731
        a = dense(f(q), f(k), f(v), True).transpose(2, 3).reshape(N, heads, T, dh)
    else:
        qb = q.reshape(N, heads, nb, block, dh)
        kb = k.reshape(N, heads, nb, block, dh)
        vb = v.reshape(N, heads, nb, block, dh)
        a = torch.zeros_like(qb)
        a[:, :, 1:] = dense(qb[:, :, 1:], kb[:, :, :-1], vb[:, :, :-1], False)
        a = a.reshape(N, heads, T, dh)
    return a.permute(0, 2, 1, 3).reshape(N * T, A).to(out_dtype)


def quick_gelu(x, out=None):
    return x / (1 + torch.exp(-1.702 * x))


def add3(a, b, c=None, out=None):
    return (a + b) + c if c is not None else a + b


def jb_pool(acts, T_valid, frame):
    n = T_valid // frame
    return acts[: n * frame].view(n, frame, -1).mean(1)


def jb_label_cond(total_row, abs_table, rel_table, abs_bins, rel_bins):
    return (total_row[None] + abs_table[abs_bins.long()]) + rel_table[rel_bins.long()]


def jb_pool_batched(acts, T_valid, frame):
    n = T_valid // frame
    return acts[:, : n * frame].reshape(acts.shape[0], n, frame, -1).mean(2)


# ------------------------------------------------------------------------------------------------ MPT variant
def gemv(x, w, bias=None, residual=None, out=None, N=None):
    N = int(N) if N is not None else w.shape[0]
    y = x.float() @ w[:N].float().t()
    if bias is not None:
        y = y + bias.float()
    if residual is not None:
        y = y + residual.float()
    y = y.to(BF16)
    if out is not None:
        out[:, :N].copy_(y)
        return out
    return y


def ln_fwd(x, w, eps, mean=None, rstd=None, out=None, out_dtype=BF16):
    """Bias-free LayerNorm over the last dim of a (possibly strided) [rows, W] view; fp32 statistics."""
    xf = x.float()
    mu = xf.mean(-1, keepdim=True)
    r = torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + eps)
    if mean is not None:
        mean.copy_(mu.flatten())
    if rstd is not None:
        rstd.copy_(r.flatten())
    y = (((xf - mu) * r) * w.float()).to(out_dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def ln_bwd(x, w, dy, mean, rstd, dres, dx, dw, accumulate_dw):
    xf, dyf, wf = x.float(), dy.float(), w.float()
    xhat = (xf - mean.float()[:, None]) * rstd.float()[:, None]
    g = dyf * wf
    gx = rstd.float()[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    if dres is not None:
        gx = gx + dres.float()
    dx.copy_(gx.to(dx.dtype))
    if dw is not None:
        d = (dyf * xhat).sum(0)
        dw.copy_(((dw.float() if accumulate_dw else 0) + d).to(dw.dtype))
    return dx


def gelu_fwd(u):
    return _F.gelu(u.float()).to(BF16)


def gelu_bwd(u, dh):
    uf = u.float()
    cdf = 0.5 * (1 + torch.erf(uf / math.sqrt(2)))
    pdf = torch.exp(-0.5 * uf * uf) / math.sqrt(2 * math.pi)
    return (dh.float() * (cdf + uf * pdf)).to(BF16)


def _alibi_attention(x, key_mask, B, S, n_heads, head_dim, slopes):
    q, k, v = x.view(B, S, 3, n_heads, head_dim).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(head_dim)
    sc = sc + (torch.arange(1 - S, 1, dtype=torch.float32).view(1, 1, 1, S) * slopes.float().view(1, n_heads, 1, 1))
    allow = torch.ones(S, S, dtype=torch.bool).tril()[None, None]
    if key_mask is not None:
        allow = allow & key_mask.view(B, 1, 1, S).bool()
    sc = sc.masked_fill(~allow, float("-inf"))
    o = (torch.softmax(sc, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * S, n_heads * head_dim)
    return o, sc


def attn_alibi_fwd(qkv, key_mask, B, S, n_heads, head_dim, slopes):
    o, sc = _alibi_attention(qkv.float(), key_mask, B, S, n_heads, head_dim, slopes)
    return o.to(BF16), torch.logsumexp(sc, dim=-1)


def attn_alibi_bwd(qkv, key_mask, out, dout, lse, B, S, n_heads, head_dim, slopes):
    x = qkv.float().clone().requires_grad_(True)
    with torch.enable_grad():
        o, _ = _alibi_attention(x, key_mask, B, S, n_heads, head_dim, slopes)
        o.backward(dout.float())
    return x.grad.to(BF16)


# ------------------------------------------------------------------------------------------------ fp32 verification path
# stand-ins for llark_b200.ops.v32_* (csrc/verify_fp32.cu): plain fp32 PyTorch, so tests/test_verify_host_cpu.py can check
# the SCHEDULE of llark_b200/verify.py (what is composed from what, in which order) against the reference's fp32 golden
def v32_split_tf32(x):
    hi = (x.contiguous().view(torch.int32) & -8192).view(torch.float32)   # 13 low mantissa bits cleared (0xffffe000)
    return hi, x - hi


def v32_embed_splice_fwd(ids, table, audio, audio_off, frames, start_pos, out):
    B, S = ids.shape
    for b in range(B):
        out[b] = table[ids[b].clamp(0, table.shape[0] - 1)]
        p = int(start_pos[b]) if start_pos is not None else -1
        if p >= 0:
            F_ = int(frames[b])
            out[b, p + 1: p + 1 + F_] = audio[int(audio_off[b]): int(audio_off[b]) + F_]
    return out


def v32_rmsnorm_fwd(x, w, eps):
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def v32_rope_(qkv, cos_t, sin_t, S, n_rot_heads, D):
    rows = torch.arange(qkv.shape[0]) % S
    v = qkv[:, : n_rot_heads * D].view(qkv.shape[0], n_rot_heads, D)
    c, s = cos_t[rows][:, None, :], sin_t[rows][:, None, :]
    rot = torch.cat((-v[..., D // 2:], v[..., : D // 2]), dim=-1)
    v.copy_(v * c + rot * s)
    return qkv


def v32_swiglu_fwd(gu):
    I = gu.shape[1] // 2
    return torch.nn.functional.silu(gu[:, :I]) * gu[:, I:]


def v32_attn_fwd(qkv, key_mask, B, S, n_heads, head_dim):
    q, k, v = qkv.view(B, S, 3, n_heads, head_dim).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(head_dim)
    allow = torch.ones(S, S, dtype=torch.bool).tril()[None, None]
    if key_mask is not None:
        allow = allow & key_mask.view(B, 1, 1, S).bool()
    p = torch.softmax(sc.masked_fill(~allow, float("-inf")), dim=-1).nan_to_num(0.0)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, n_heads * head_dim)


def v32_ce_fwd(logits, V, labels, B, S):
    lg = logits.view(B, S, -1)[:, :, :V]
    return torch.nn.functional.cross_entropy(lg[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=-100)
