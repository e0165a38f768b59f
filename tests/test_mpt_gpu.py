"""GPU tests of the MPT variant's kernels: bias-free LayerNorm fwd/bwd, exact GELU fwd/bwd, the
ALiBi instantiation of the tcgen05 attention, and the whole MPTEngine step against oracle/mpt_oracle.py (which
tests/test_mpt_oracle_cpu.py pins to the reference's WrappedMPTForCausalLM). First run on a B200 in round 2
(profiles/r2a_staged_pytest.txt: all green)."""
import math
import os
import sys

import pytest
import torch

from llark_b200.testing import check_close  # noqa: E402

sys.path.insert(0, os.path.dirname(__file__))
BF16 = torch.bfloat16
pytestmark = [pytest.mark.gpu]


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


@pytest.mark.parametrize("rows,W,strided", [(5, 256, False), (300, 2048, False), (300, 2048, True), (129, 4096, False),
                                            (64, 768, True)])
def test_layernorm_forward_backward(rows, W, strided):
    import cpu_ops
    from llark_b200 import ops

    g = torch.Generator().manual_seed(rows + W)
    full_w = 3 * W if strided else W
    x_full = (torch.randn((rows, full_w), generator=g) * 2 + 0.5).to(BF16)
    dy_full = torch.randn((rows, full_w), generator=g).to(BF16)
    w = (1 + 0.1 * torch.randn((W,), generator=g)).to(BF16)
    dres = torch.randn((rows, W), generator=g).to(BF16)
    sl = slice(W, 2 * W) if strided else slice(0, W)
    # reference (fp32 math on the same bf16 inputs)
    mu_r, rs_r = torch.empty(rows), torch.empty(rows)
    y_r = cpu_ops.ln_fwd(x_full[:, sl], w, 1e-5, mean=mu_r, rstd=rs_r)
    dx_r, dw_r = torch.empty((rows, W), dtype=BF16), torch.full((W,), 0.25, dtype=BF16)
    cpu_ops.ln_bwd(x_full[:, sl], w, dy_full[:, sl], mu_r, rs_r, dres, dx_r, dw_r, True)

    xd, dyd, wd, dresd = x_full.cuda(), dy_full.cuda(), w.cuda(), dres.cuda()
    mu = torch.empty(rows, device="cuda")
    rs = torch.empty(rows, device="cuda")
    y_full = torch.zeros((rows, full_w), device="cuda", dtype=BF16)
    ops.ln_fwd(xd[:, sl], wd, 1e-5, mean=mu, rstd=rs, out=y_full[:, sl])
    assert torch.allclose(mu.cpu(), mu_r, rtol=1e-4, atol=1e-4)
    assert torch.allclose(rs.cpu(), rs_r, rtol=1e-4, atol=1e-5)
    check_close("test_layernorm_forward_backward:_rel_y_full_sl_cpu_y_r_", _rel(y_full[:, sl].cpu(), y_r), 4e-3)
    if strided:   # neighbours of the view untouched
        assert y_full[:, :W].abs().sum().item() == 0 and y_full[:, 2 * W:].abs().sum().item() == 0
    dx_full = torch.zeros((rows, full_w), device="cuda", dtype=BF16)
    dw = torch.full((W,), 0.25, device="cuda", dtype=BF16)
    ops.ln_bwd(xd[:, sl], wd, dyd[:, sl], mu, rs, dresd, dx_full[:, sl], dw, True)
    check_close("test_layernorm_forward_backward:_rel_dx_full_sl_cpu_dx_r_", _rel(dx_full[:, sl].cpu(), dx_r), 6e-3)
    check_close("test_layernorm_forward_backward:_rel_dw_cpu_dw_r_", _rel(dw.cpu(), dw_r), 1e-2)
    dw2 = torch.full((W,), 7.0, device="cuda", dtype=BF16)
    ops.ln_bwd(xd[:, sl], wd, dyd[:, sl], mu, rs, None, dx_full[:, sl], dw2, False)   # overwrite, no residual gradient
    dx_r2, dw_r2 = torch.empty((rows, W), dtype=BF16), torch.empty((W,), dtype=BF16)
    cpu_ops.ln_bwd(x_full[:, sl], w, dy_full[:, sl], mu_r, rs_r, None, dx_r2, dw_r2, False)
    check_close("test_layernorm_forward_backward:_rel_dx_full_sl_cpu_dx_r2_6e_3_and__rel_", _rel(dx_full[:, sl].cpu(), dx_r2) < 6e-3 and _rel(dw2.cpu(), dw_r2), 1e-2)
    ops.ln_bwd(xd[:, sl], wd, dyd[:, sl], mu, rs, None, dx_full[:, sl], None, False)   # dw not wanted
    check_close("test_layernorm_forward_backward:_rel_dx_full_sl_cpu_dx_r2_", _rel(dx_full[:, sl].cpu(), dx_r2), 6e-3)


@pytest.mark.parametrize("n", [8, 4096, 300 * 8192])
def test_exact_gelu_forward_backward(n):
    import cpu_ops
    from llark_b200 import ops

    g = torch.Generator().manual_seed(n)
    u = (torch.randn((n,), generator=g) * 2.5).to(BF16)
    dh = torch.randn((n,), generator=g).to(BF16)
    h = ops.gelu_fwd(u.cuda())
    du = ops.gelu_bwd(u.cuda(), dh.cuda())
    h_r, du_r = cpu_ops.gelu_fwd(u), cpu_ops.gelu_bwd(u, dh)
    # same fp32 formula, one bf16 rounding: allow one ulp on a handful of elements
    check_close("test_exact_gelu_forward_backward:_rel_h_cpu_h_r_3e_3_and__rel_du_cpu_du_r", _rel(h.cpu(), h_r) < 3e-3 and _rel(du.cpu(), du_r), 3e-3)
    assert (h.cpu().float() - h_r.float()).abs().max().item() <= 2 ** -7 * h_r.float().abs().max().item()


@pytest.mark.parametrize("B,S,NH,lens", [(1, 128, 2, None), (2, 200, 4, [200, 77]), (2, 640, 16, [640, 513])])
def test_alibi_attention_forward_backward(B, S, NH, lens):
    import cpu_ops
    from llark_b200 import ops
    from llark_b200.engine_mpt import alibi_slopes

    D = 128
    g = torch.Generator().manual_seed(B * S + NH)
    qkv = (torch.randn((B * S, 3 * NH * D), generator=g) * 0.7).to(BF16)
    dout = torch.randn((B * S, NH * D), generator=g).to(BF16)
    mask = torch.ones((B, S), dtype=torch.uint8)
    if lens:
        for b, n in enumerate(lens):
            mask[b, n:] = 0
    slopes = alibi_slopes(NH).float()
    o_r, lse_r = cpu_ops.attn_alibi_fwd(qkv, mask, B, S, NH, D, slopes)
    dqkv_r = cpu_ops.attn_alibi_bwd(qkv, mask, o_r, dout, lse_r, B, S, NH, D, slopes)
    o, lse = ops.attn_alibi_fwd(qkv.cuda(), mask.cuda(), B, S, NH, D, slopes.cuda())
    valid = mask.bool().view(-1)
    check_close("test_alibi_attention_forward_backward:_rel_o_cpu_valid_o_r_valid_", _rel(o.cpu()[valid], o_r[valid]), 6e-3)
    lv = mask.bool()[:, None, :].expand(B, NH, S)
    assert (lse.cpu()[lv] - lse_r[lv]).abs().max().item() < 2e-2
    dqkv = ops.attn_alibi_bwd(qkv.cuda(), mask.cuda(), o, dout.cuda(), lse, B, S, NH, D, slopes.cuda())
    dd = dout.clone()
    dd[~valid] = 0   # rows of padded queries carry no loss in the engine; compare on what matters
    dqkv_r = cpu_ops.attn_alibi_bwd(qkv, mask, o_r, dd, lse_r, B, S, NH, D, slopes)
    dqkv_v = ops.attn_alibi_bwd(qkv.cuda(), mask.cuda(), o, dd.cuda(), lse, B, S, NH, D, slopes.cuda())
    check_close("test_alibi_attention_forward_backward:_rel_dqkv_v_cpu_valid_dqkv_r_valid_", _rel(dqkv_v.cpu()[valid], dqkv_r[valid]), 1.5e-2)
    # slope 0 reduces to the plain causal kernel (bit-identical code path apart from one FFMA per score)
    o0, lse0 = ops.attn_alibi_fwd(qkv.cuda(), mask.cuda(), B, S, NH, D, torch.zeros(NH, device="cuda"))
    o1, lse1 = ops.attn_fwd(qkv.cuda(), mask.cuda(), B, S, NH, D)
    check_close("test_alibi_attention_forward_backward:_rel_o0_valid_cuda_o1_valid_cuda_", _rel(o0[valid.cuda()], o1[valid.cuda()]), 1e-3)


@pytest.mark.parametrize("F_frames,detach", [(1, False), (4, True)])
def test_mpt_engine_step_matches_oracle_on_gpu(F_frames, detach):
    from oracle import mpt_oracle as mo
    from test_mpt_host_cpu import _build

    cfg = mo.MPTCfg(d_model=256, n_heads=2)
    params = mo.init_params(cfg, seed=0)
    batch = mo.make_batch(cfg, B=2, S=40, F_frames=F_frames, seed=F_frames)
    model = _build(cfg, params).cuda()
    if detach:
        model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    dev = {k: (v.cuda() if torch.is_tensor(v) else [e.cuda() for e in v] if isinstance(v, list) else v)
           for k, v in batch.items()}
    out = model(**dev)
    out.loss.backward()
    pb = {k: v.to(BF16) for k, v in params.items()}
    loss_o, logits_o, grads_o = mo.forward_backward(pb, cfg, batch, detach_text=detach, autocast_dtype=BF16)
    assert abs(out.loss.item() - loss_o.item()) < 3e-3 * loss_o.item()
    valid = batch["attention_mask"]
    check_close("test_mpt_engine_step_matches_oracle_on_gpu:_rel_out_logits_float_cpu_valid_logits_o", _rel(out.logits.float().cpu()[valid], logits_o.float()[valid]), 2e-2)
    named = dict(model.named_parameters())
    for name, g_ref in grads_o.items():
        assert named[name].grad is not None, name
        check_close(f"test_mpt_engine_step_matches_oracle_on_gpu:_rel_named_name_grad_cpu_g_ref_:{name}", _rel(named[name].grad.cpu(), g_ref), 4e-2)
