"""GPU parity tests of every C-ABI kernel against a plain PyTorch fp32 restatement of the same op.

These are kernel-level tests (one op at a time, through the C-ABI via llark_b200.ops). The model-level parity
tests against the oracle / golden fixtures live in test_model_gpu.py.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _ops():
    from llark_b200 import ops

    return ops


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.fixture(params=[1, 0], ids=["cta_pair", "single_cta"])
def gemm_mode(request):
    """Run every GEMM test on both kernel variants (cta_group::2 pair kernel and the single-CTA kernel)."""
    from llark_b200._lib import lib

    old = lib().llark_gemm_get_mode()
    lib().llark_gemm_set_mode(request.param)
    yield request.param
    lib().llark_gemm_set_mode(old)


GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 256), (256, 512, 4096), (240, 4096, 4800), (512, 1000, 256), (384, 768, 1000),
    (100, 72, 136), (1024, 4096, 4096),
]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bf16(M, N, K, a_mn, b_mn, gemm_mode):
    ops = _ops()
    torch.manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda") * 0.5
    B = torch.randn(N, K, device="cuda") * 0.5
    a = A.to(BF16)
    b = B.to(BF16)
    ref = a.float() @ b.float().t()
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    if (a_in.stride(0) % 8) or (b_in.stride(0) % 8):
        pytest.skip("leading dimension not 16-byte aligned for TMA")
    out = ops.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn)
    assert out.shape == (M, N)
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert _rel(out, ref) < 5e-3, (err, scale)


@pytest.mark.parametrize("M,N,K", [(1, 12288, 4096), (2, 2048, 512), (17, 4096, 11008), (32, 50372, 2048)])
def test_gemm_very_few_rows(M, N, K, gemm_mode):
    """Fewer A rows than one TMA box (the box reads past the tensor and is zero-filled): the decode step of the KV-cache
    path (M = batch), the CLAP projector (one frame per sample) and the MPT head at decode time."""
    ops = _ops()
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(BF16)
    b = (torch.randn(N, K, device="cuda") * 0.5).to(BF16)
    ldn = (N + 63) // 64 * 64
    buf = torch.zeros(M, ldn, device="cuda", dtype=BF16)
    ops.gemm(a, b, out=buf, N=N)
    assert _rel(buf[:, :N], a.float() @ b.float().t()) < 5e-3
    assert buf[:, N:].abs().sum().item() == 0


@pytest.mark.parametrize("M,N,K", [(1, 12288, 4096), (1, 4096, 11008), (3, 22016, 4096), (8, 32004, 4096), (2, 72, 136)])
def test_gemv_decode_linear(M, N, K):
    """llark_gemv (weight-streaming linear of the decode step) against fp32 math and against the GEMM core it replaces."""
    ops = _ops()
    torch.manual_seed(M * 5 + N + K)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.5).to(BF16)
    bias = torch.randn(N, device="cuda").to(BF16)
    res = torch.randn(M, N, device="cuda").to(BF16)
    ref = x.float() @ w.float().t()
    assert _rel(ops.gemv(x, w), ref) < 4e-3
    assert _rel(ops.gemv(x, w, bias=bias, residual=res), ref + bias.float() + res.float()) < 4e-3
    ldn = (N + 63) // 64 * 64
    buf = torch.zeros(M, ldn, device="cuda", dtype=BF16)
    ops.gemv(x, w, out=buf, N=N)
    assert _rel(buf[:, :N], ref) < 4e-3 and buf[:, N:].abs().sum().item() == 0
    # strided input rows (a column slice of a wider buffer), as qkv[:, :H] would be
    wide = torch.zeros(M, K + 64, device="cuda", dtype=BF16)
    wide[:, :K] = x
    assert torch.equal(ops.gemv(wide[:, :K], w), ops.gemv(x, w))
    if K % 64 == 0 and N % 8 == 0:
        g = ops.gemm(x, w)                                  # both round fp32 sums once: equal up to summation order
        assert _rel(ops.gemv(x, w), g.float()) < 4e-3


def test_gemm_bias_residual_and_fp32_out(gemm_mode):
    ops = _ops()
    torch.manual_seed(0)
    M, N, K = 300, 520, 192
    a = torch.randn(M, K, device="cuda").to(BF16)
    b = torch.randn(N, K, device="cuda").to(BF16)
    bias = torch.randn(N, device="cuda").to(BF16)
    res = torch.randn(M, N, device="cuda").to(BF16)
    ref = a.float() @ b.float().t() + bias.float() + res.float()
    out = ops.gemm(a, b, bias=bias, residual=res)
    assert _rel(out, ref) < 5e-3
    # accumulate in place (wgrad accumulation): R aliases C
    acc = res.clone()
    ops.gemm(a, b, out=acc, residual=acc)
    assert _rel(acc, a.float() @ b.float().t() + res.float()) < 5e-3
    out32 = ops.gemm(a, b, out_dtype=torch.float32)
    assert out32.dtype == torch.float32
    assert _rel(out32, a.float() @ b.float().t()) < 1e-5


def test_gemm_tf32_mn_major_rejected():
    ops = _ops()
    from llark_b200._lib import LlarkError

    a = torch.randn(256, 128, device="cuda")
    with pytest.raises(LlarkError):
        ops.gemm(a, a.clone(), b_mn=True)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False)])
def test_gemm_tf32(a_mn, b_mn, gemm_mode):
    ops = _ops()
    torch.manual_seed(1)
    M, N, K = 256, 1200, 4800
    a = torch.randn(M, K, device="cuda") * 0.3
    b = torch.randn(N, K, device="cuda") * 0.3
    bias = torch.randn(N, device="cuda")
    ref = a.double() @ b.double().t() + bias.double()
    a_in = a.t().contiguous() if a_mn else a
    b_in = b.t().contiguous() if b_mn else b
    out = ops.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn, bias=bias)
    assert out.dtype == torch.float32
    assert _rel(out, ref) < 2e-3


def test_gemm_persistent_many_tiles_and_ld(gemm_mode):
    """More tiles than SMs (several waves per CTA), padded leading dimensions, partial N tile."""
    ops = _ops()
    torch.manual_seed(2)
    M, N, K = 2048, 32004, 512
    ldv = 32064
    a = (torch.randn(M, K, device="cuda") * 0.2).to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.2).to(BF16)
    out = torch.zeros(M, ldv, device="cuda", dtype=BF16)
    ops.gemm(a, w, out=out, N=N)
    ref = a.float() @ w.float().t()
    assert _rel(out[:, :N], ref) < 5e-3
    assert out[:, N:].abs().max().item() == 0.0
    # dgrad through the padded buffer: dH = dlogits[M, V(ld 32064)] @ W[V, K]
    dl = (torch.randn(M, ldv, device="cuda") * 0.1).to(BF16)
    dl[:, N:] = float("nan")  # padding must never be read as data (TMA clips at V)
    dh = ops.gemm(dl, w, b_mn=True, M=M, N=K, K=N, lda=ldv)
    ref2 = dl[:, :N].float() @ w.float()
    assert torch.isfinite(dh.float()).all()
    assert _rel(dh, ref2) < 5e-3


# ------------------------------------------------------------------------------------------------ elementwise
def test_rmsnorm_fwd_bwd():
    ops = _ops()
    torch.manual_seed(3)
    rows, H = 257, 4096
    x = torch.randn(rows, H, device="cuda").to(BF16)
    w = (1 + 0.1 * torch.randn(H, device="cuda")).to(BF16)
    rstd = torch.empty(rows, device="cuda")
    y = ops.rmsnorm_fwd(x, w, 1e-5, rstd=rstd)

    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    var = xf.pow(2).mean(-1, keepdim=True)
    ref = wf * (xf * torch.rsqrt(var + 1e-5))
    assert _rel(y, ref) < 4e-3
    dy = torch.randn(rows, H, device="cuda").to(BF16)
    dres = torch.randn(rows, H, device="cuda").to(BF16)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dw = torch.zeros(H, device="cuda", dtype=BF16)
    ws = torch.empty(296 * H, device="cuda")
    ops.rmsnorm_bwd(x, w, dy, rstd, dres, dx, dw, False, ws)
    assert _rel(dx, xf.grad + dres.float()) < 6e-3
    assert _rel(dw, wf.grad) < 6e-3
    ops.rmsnorm_bwd(x, w, dy, rstd, None, dx, dw, True, ws)
    assert _rel(dx, xf.grad) < 6e-3
    assert _rel(dw, 2 * wf.grad) < 1e-2


def _hf_rope_tables(S, D, theta=10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, device="cuda", dtype=torch.float32) / D))
    pos = torch.arange(S, device="cuda", dtype=torch.float32)
    freqs = pos[:, None] * inv[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(BF16), emb.sin().to(BF16)


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def test_rope_fwd_and_inverse():
    ops = _ops()
    torch.manual_seed(4)
    B, S, NH, D = 2, 96, 4, 128
    qkv = torch.randn(B * S, 3 * NH * D, device="cuda").to(BF16)
    cos, sin = _hf_rope_tables(S, D)
    ref = qkv.clone().view(B, S, 3, NH, D)
    c = cos[None, :, None, :]
    s = sin[None, :, None, :]
    for i in (0, 1):
        t = ref[:, :, i]
        ref[:, :, i] = (t * c) + (_rotate_half(t) * s)  # bf16 arithmetic exactly as HF does it
    got = qkv.clone()
    ops.rope_(got, cos, sin, S, NH, D)
    assert torch.equal(got.view(B, S, 3, NH, D)[:, :, 2], qkv.view(B, S, 3, NH, D)[:, :, 2])
    assert (got.float() - ref.view(B * S, -1).float()).abs().max().item() <= 0.04
    assert _rel(got, ref.view(B * S, -1)) < 3e-3
    # inverse is the transpose: <rope(x), y> == <x, rope^T(y)>
    y = torch.randn_like(qkv.float()).to(BF16)
    yi = y.clone()
    ops.rope_(yi, cos, sin, S, NH, D, inverse=True)
    sel = slice(0, 2 * NH * D)
    lhs = (got.float()[:, sel] * y.float()[:, sel]).sum().item()
    rhs = (qkv.float()[:, sel] * yi.float()[:, sel]).sum().item()
    assert abs(lhs - rhs) / (abs(lhs) + 1e-6) < 2e-2


def test_swiglu_fwd_bwd():
    ops = _ops()
    torch.manual_seed(5)
    rows, I = 130, 11008
    gu = torch.randn(rows, 2 * I, device="cuda").to(BF16)
    act = ops.swiglu_fwd(gu)
    g = gu[:, :I].float().requires_grad_(True)
    u = gu[:, I:].float().requires_grad_(True)
    ref = torch.nn.functional.silu(g) * u
    assert _rel(act, ref) < 4e-3
    d = torch.randn(rows, I, device="cuda").to(BF16)
    ref.backward(d.float())
    dgu = ops.swiglu_bwd(gu, d)
    assert _rel(dgu[:, :I], g.grad) < 5e-3
    assert _rel(dgu[:, I:], u.grad) < 5e-3


# ------------------------------------------------------------------------------------------------ fused GEMM epilogues
def test_gemm_rope_epilogue_is_bit_identical_to_gemm_then_rope(gemm_mode):
    """llark_gemm_ex(LLARK_EPI_ROPE) == llark_gemm followed by llark_rope, bit for bit (q | k rotated, v untouched)."""
    ops = _ops()
    torch.manual_seed(21)
    B, S, NH, D, K = 3, 100, 2, 128, 320
    H = NH * D
    x = (torch.randn(B * S, K, device="cuda") * 0.5).to(BF16)
    w = (torch.randn(3 * H, K, device="cuda") * 0.1).to(BF16)
    cos, sin = _hf_rope_tables(S, D)
    ref = ops.gemm(x, w)
    ops.rope_(ref, cos, sin, S, NH, D)
    got = ops.gemm(x, w, rope=(cos, sin, S, 2 * H))
    assert torch.equal(got, ref)


@pytest.mark.parametrize("M,I,H", [(300, 344, 256), (256, 11008, 128)])
def test_gemm_swiglu_bwd_epilogue_is_bit_identical(M, I, H, gemm_mode):
    """dgrad of the down projection with the SwiGLU backward in its epilogue == gemm -> llark_swiglu_bwd."""
    ops = _ops()
    torch.manual_seed(22)
    dx = (torch.randn(M, H, device="cuda") * 0.5).to(BF16)
    wd = (torch.randn(H, I, device="cuda") * 0.1).to(BF16)  # down_proj.weight [H, I]
    gu = torch.randn(M, 2 * I, device="cuda").to(BF16)
    dact = ops.gemm(dx, wd, b_mn=True)
    ref = ops.swiglu_bwd(gu, dact)
    got = torch.full_like(gu, float("nan"))
    ops.gemm(dx, wd, b_mn=True, out=got, N=I, swiglu_bwd=gu)
    assert torch.equal(got, ref)


@pytest.mark.parametrize("dt", [torch.float32, BF16])
def test_gemm_quick_gelu_epilogue_is_bit_identical(dt, gemm_mode):
    ops = _ops()
    torch.manual_seed(23)
    M, N, K = 200, 520, 192
    x = (torch.randn(M, K, device="cuda") * 0.5).to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.2).to(BF16)
    bias = torch.randn(N, device="cuda")
    ref = ops.quick_gelu(ops.gemm(x, w, bias=bias, out_dtype=torch.float32))
    if dt == BF16:
        ref = ops.cast_f32_to_bf16(ref)
    got = ops.gemm(x, w, bias=bias, out_dtype=dt, quick_gelu=True)
    assert torch.equal(got, ref)


def _attn_ref(qkv, mask, B, S, NH, D):
    q, k, v = qkv.float().view(B, S, 3, NH, D).permute(2, 0, 3, 1, 4)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(D)
    causal = torch.ones(S, S, device=qkv.device, dtype=torch.bool).tril()
    allow = causal[None, None]
    if mask is not None:
        allow = allow & mask[:, None, None, :].bool()
    sc = sc.masked_fill(~allow, float("-inf"))
    p = torch.softmax(sc, dim=-1)
    o = p @ v
    return o.permute(0, 2, 1, 3).reshape(B * S, NH * D)


@pytest.mark.parametrize("B,S,NH,pad", [(2, 512, 4, False), (2, 200, 2, True), (1, 64, 1, False), (3, 333, 2, True)])
def test_attention_fwd_bwd(B, S, NH, pad):
    ops = _ops()
    torch.manual_seed(6 + S)
    D = 128
    qkv = (torch.randn(B * S, 3 * NH * D, device="cuda") * 0.8).to(BF16)
    mask = None
    if pad:
        lens = torch.randint(S // 2, S + 1, (B,), device="cuda")
        mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None])
        mask_u8 = mask.to(torch.uint8).contiguous()
    else:
        mask_u8 = None
    out, lse = ops.attn_fwd(qkv, mask_u8, B, S, NH, D)
    qf = qkv.float().requires_grad_(True)
    ref = _attn_ref(qf, mask, B, S, NH, D)
    assert _rel(out, ref) < 6e-3
    dout = torch.randn(B * S, NH * D, device="cuda").to(BF16)
    if pad:
        dout = dout * mask.reshape(B * S, 1).to(BF16)  # padded query rows carry no loss gradient
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, mask_u8, out, dout, lse, B, S, NH, D)
    g = qf.grad.view(B, S, 3, NH, D)
    got = dqkv.float().view(B, S, 3, NH, D)
    for i, name in enumerate("qkv"):
        assert _rel(got[:, :, i], g[:, :, i]) < 1.5e-2, name


@pytest.mark.parametrize("B,S,NH,pad", [(2, 512, 4, True), (1, 200, 2, False), (1, 4160, 1, True)])
def test_attention_bwd_with_fused_inverse_rope_is_bit_identical(B, S, NH, pad):
    """llark_attn_bwd_rope == llark_attn_bwd followed by llark_rope(inverse=1), bit for bit: on the tcgen05 kernels (rotation
    in the dq / dk epilogues) and beyond their envelope (S = 4160: mma.sync kernels + the rope kernel inside the call)."""
    ops = _ops()
    D, H = 128, NH * 128
    g = torch.Generator(device="cuda").manual_seed(S + NH)
    qkv = (torch.randn(B * S, 3 * H, device="cuda", generator=g) * 0.7).to(BF16)
    dout = torch.randn(B * S, H, device="cuda", generator=g).to(BF16)
    mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
    if pad:
        mask[-1, S - 37:] = False
    mask_u8 = mask.view(torch.uint8)
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, device="cuda").float() / D))
    emb = torch.cat((torch.arange(S, device="cuda").float()[:, None] * inv[None],) * 2, dim=-1)
    cos, sin = emb.cos().to(BF16).contiguous(), emb.sin().to(BF16).contiguous()
    out, lse = ops.attn_fwd(qkv, mask_u8, B, S, NH, D)
    want = ops.attn_bwd(qkv, mask_u8, out, dout, lse, B, S, NH, D)
    ops.rope_(want, cos, sin, S, NH, D, inverse=True)
    got = ops.attn_bwd(qkv, mask_u8, out, dout, lse, B, S, NH, D, rope=(cos, sin))
    valid = mask.view(-1)
    assert torch.equal(got[valid], want[valid])


@pytest.mark.parametrize("B,S,NH,pad", [(1, 2048, 2, True),     # the reference's --model_max_length (train_llark.sh)
                                        (1, 4096, 1, False),    # upper edge of the tcgen05 kernels' envelope
                                        (1, 4160, 1, True)])    # beyond it: llark_attn_* must fall back to mma.sync
def test_attention_fwd_bwd_long_sequences(B, S, NH, pad):
    """Maximum sizes: same check as test_attention_fwd_bwd at the reference's model_max_length, at the largest S the
    tcgen05 kernels accept (MAX_S in attention_tc.cu) and one tile past it (the mma.sync kernels take over)."""
    test_attention_fwd_bwd(B, S, NH, pad)


def test_embed_splice_fwd_bwd_and_flags():
    ops = _ops()
    torch.manual_seed(7)
    B, S, H, V, F = 3, 40, 256, 100, 5
    start_id, end_id, patch_id = 98, 99, 97
    ids = torch.randint(3, 90, (B, S), device="cuda")
    pos = [4, 10, -1]
    for b, p in enumerate(pos):
        if p >= 0:
            ids[b, p] = start_id
            ids[b, p + 1: p + 1 + F] = patch_id
            ids[b, p + F + 1] = end_id
    table = torch.randn(V, H, device="cuda").to(BF16)
    audio = torch.randn(B * F, H, device="cuda").to(BF16)
    frames = torch.full((B,), F, device="cuda", dtype=torch.int32)
    audio_off = (torch.arange(B, device="cuda") * F).to(torch.int64)
    start_pos = torch.empty(B, device="cuda", dtype=torch.int32)
    flags = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.splice_scan(ids, start_id, end_id, frames, start_pos, flags)
    assert start_pos.tolist() == pos
    assert flags.item() == 0
    out = torch.empty(B, S, H, device="cuda", dtype=BF16)
    ops.embed_splice_fwd(ids, table, audio, audio_off, frames, start_pos, out)
    ref = table[ids]
    for b, p in enumerate(pos):
        if p >= 0:
            ref[b, p + 1: p + 1 + F] = audio[b * F: (b + 1) * F]
    assert torch.equal(out, ref)
    # backward, detach_text path: samples WITH an audio segment give gradient to the marker rows + audio rows only; the
    # sample without <audio_start> keeps the reference's undetached embeddings (llamav2.py:162-208 never runs for it), so
    # every one of its token rows receives gradient
    dout = torch.randn(B, S, H, device="cuda").to(BF16)
    daudio = torch.zeros(B * F, H, device="cuda", dtype=BF16)
    dtable = torch.zeros(V, H, device="cuda", dtype=BF16)
    ops.embed_splice_bwd(ids, dout, daudio, audio_off, frames, start_pos, dtable, V, True)
    ref_da = torch.zeros_like(daudio)
    ref_dt = torch.zeros(V, H, device="cuda")
    for b, p in enumerate(pos):
        if p >= 0:
            ref_da[b * F: (b + 1) * F] = dout[b, p + 1: p + 1 + F]
            ref_dt[start_id] += dout[b, p].float()
            ref_dt[end_id] += dout[b, p + F + 1].float()
        else:
            ref_dt.index_add_(0, ids[b], dout[b].float())
    assert torch.equal(daudio, ref_da)
    nz = dtable.float().abs().sum(-1).nonzero().flatten().tolist()
    assert nz == sorted(set(ids[2].tolist()) | {start_id, end_id})
    assert _rel(dtable, ref_dt) < 2e-2   # bf16 atomics: rows hit by several tokens round after every add
    # flag: end token misplaced / counts differ
    bad = ids.clone()
    bad[0, pos[0] + F + 1] = 5
    flags.zero_()
    ops.splice_scan(bad, start_id, end_id, frames, start_pos, flags)
    assert flags.item() & 1 and flags.item() & 2


def test_ce_fwd_bwd():
    ops = _ops()
    torch.manual_seed(8)
    B, S, V, ldv = 2, 64, 32004, 32064
    logits = torch.zeros(B * S, ldv, device="cuda", dtype=BF16)
    logits[:, :V] = (torch.randn(B * S, V, device="cuda") * 2).to(BF16)
    logits[:, V:] = 77.0  # padding columns must be ignored
    labels = torch.randint(0, V, (B, S), device="cuda")
    labels[:, : S // 2] = -100
    row_loss = torch.empty(B * S, device="cuda")
    row_lse = torch.empty(B * S, device="cuda")
    loss_out = torch.empty(2, device="cuda")
    ops.ce_fwd(logits, ldv, V, labels, B, S, row_loss, row_lse, loss_out)
    lf = logits[:, :V].float().view(B, S, V).requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1), ignore_index=-100)
    assert abs(loss_out[0].item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    assert loss_out[1].item() == (labels[:, 1:] != -100).sum().item()
    ref.backward()
    dl = torch.empty_like(logits)
    ops.ce_bwd(logits, ldv, V, labels, B, S, row_lse, loss_out, None, 1.0, dl)
    assert dl[:, V:].abs().max().item() == 0
    assert _rel(dl[:, :V].view(B, S, V), lf.grad) < 5e-3


def test_adamw_matches_torch_bf16():
    ops = _ops()
    torch.manual_seed(9)
    n = 1 << 16
    p0 = torch.randn(n, device="cuda").to(BF16)
    g0 = (torch.randn(n, device="cuda") * 0.01).to(BF16)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, foreach=True)
    p = p0.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    sumsq = torch.zeros(1, device="cuda")
    ws = torch.empty(4096, device="cuda")
    for step in range(1, 4):
        g = (g0.float() * step).to(BF16)
        p_ref.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        gg = g.clone()
        ops.grad_sumsq(gg, sumsq, False, ws)
        assert abs(sumsq.sqrt().item() - total.item()) < 2e-2 * total.item()
        ops.adamw_step(p, gg, m, v, sumsq, 5e-5, 0.9, 0.999, 1e-8, 0.0, step, 1.0, 1.0, True)
        assert gg.abs().max().item() == 0
    diff = (p.float() - p_ref.detach().float()).abs()
    ulp = p_ref.detach().float().abs().clamp_min(1e-3) * 2 ** -7
    assert (diff <= 2 * ulp).float().mean().item() > 0.999


def test_cast_and_colsum():
    ops = _ops()
    x = torch.randn(1000, 257, device="cuda")
    b = ops.cast_f32_to_bf16(x)
    assert torch.equal(b, x.to(BF16))
    assert torch.equal(ops.cast_bf16_to_f32(b), b.float())
    xb = torch.randn(300, 4096, device="cuda").to(BF16)
    out = torch.zeros(4096, device="cuda", dtype=BF16)
    ops.colsum(xb, out)
    assert _rel(out, xb.float().sum(0)) < 4e-3


@pytest.mark.parametrize("B,s_max,lens,NH", [(1, 64, [17], 2), (3, 300, [1, 300, 129], 4)])
def test_attention_decode_step(B, s_max, lens, NH):
    ops = _ops()
    torch.manual_seed(31)
    D = 128
    q = (torch.randn(B, NH * D, device="cuda") * 0.8).to(BF16)
    kc = (torch.randn(B, s_max, NH * D, device="cuda") * 0.8).to(BF16)
    vc = (torch.randn(B, s_max, NH * D, device="cuda") * 0.8).to(BF16)
    lens_t = torch.tensor(lens, device="cuda", dtype=torch.int32)
    out = ops.attn_decode(q, kc, vc, lens_t, NH, D)
    for b in range(B):
        n = lens[b]
        qf = q[b].float().view(NH, D)
        kf = kc[b, :n].float().view(n, NH, D).permute(1, 0, 2)
        vf = vc[b, :n].float().view(n, NH, D).permute(1, 0, 2)
        p = torch.softmax((kf @ qf[:, :, None]).squeeze(-1) / math.sqrt(D), dim=-1)  # [NH, n]
        ref = (p[:, None, :] @ vf).squeeze(1).reshape(-1)
        assert _rel(out[b], ref) < 6e-3
