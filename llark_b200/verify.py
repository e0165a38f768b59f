"""fp32 verification mode of the Llama forward (`LLARK_VERIFY=fp32`).

north_star asks for "fp32 logits within 1e-3 relative" of the reference's own PyTorch path. The training path is the
reference's bf16 regime (two bf16 pipelines differ by ~1e-2 after a few layers, whoever computes them), so the
criterion is checked here: the same forward schedule as engine.LlamaEngine._forward -- embed + projector + splice, L x
(RMSNorm, fused QKV, RoPE, causal attention, O + residual, RMSNorm, gate|up, SwiGLU, down + residual), final norm,
lm_head, shifted CE (m2t/models/llamav2.py:95-337) -- with fp32 tensors end to end. Linears run on the tcgen05 GEMM core
in kind::tf32; `LLARK_VERIFY_GEMM=3xtf32` (default) composes every product from TF32 halves so the result carries ~fp32
accuracy, `tf32` is the plain single-pass TF32 product. Everything else is csrc/verify_fp32.cu. Forward only, no
autograd, tiny throughput: a checker that still runs every FLOP through the C-ABI (no PyTorch math, no CPU path).
"""
import os

import torch

from . import ops

F32 = torch.float32


def _linear(x, w, bias=None, residual=None, mode="3xtf32"):
    """y[M,N] = x[M,K] @ w[N,K]^T (+ bias) (+ residual) in fp32 storage on the TF32 tensor-core path."""
    if mode == "tf32":
        return ops.gemm(x, w, bias=bias, residual=residual, out_dtype=F32)
    x_hi, x_lo = ops.v32_split_tf32(x)
    w_hi, w_lo = ops.v32_split_tf32(w)
    y = ops.gemm(x_lo, w_hi, residual=residual, out_dtype=F32)      # small terms first, big term last
    y = ops.gemm(x_hi, w_lo, residual=y, out=y)
    return ops.gemm(x_hi, w_hi, residual=y, bias=bias, out=y)


def _require_cuda(dev):
    if dev.type != "cuda":
        raise RuntimeError("llark_b200 has no CPU path: move the model to a CUDA device (B200) first")


@torch.no_grad()
def forward_fp32(lm, input_ids, attention_mask=None, labels=None, audio_encodings=None):
    """(loss or None, logits fp32 [B, S, V]) of WrappedLlamav2ForCausalLM `lm` computed in fp32 on the GPU."""
    mode = os.environ.get("LLARK_VERIFY_GEMM", "3xtf32")
    if mode not in ("3xtf32", "tf32"):
        raise ValueError("LLARK_VERIFY_GEMM must be 3xtf32 or tf32")
    cfg = lm.config
    P = {n: p.detach().to(F32).contiguous() for n, p in lm.named_parameters()}
    dev = P["model.embed_tokens.weight"].device
    _require_cuda(dev)
    ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
    B, S = ids.shape
    H, L, NH = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads
    D = H // NH
    V = P["lm_head.weight"].shape[0]
    M = B * S
    eps = cfg.rms_norm_eps
    mask_u8 = None
    if attention_mask is not None:
        mask_u8 = attention_mask.to(dev).to(torch.bool).contiguous().view(torch.uint8)
    table = P["model.embed_tokens.weight"]
    x = torch.empty((M, H), device=dev, dtype=F32)
    model = lm.model
    aec = model.audio_encoder_config
    if audio_encodings is not None and getattr(cfg, "use_mm_proj", False):
        if not aec.use_audio_start_end:
            raise NotImplementedError("audio_encoder_config.use_audio_start_end=False is not implemented.")
        if isinstance(audio_encodings, (list, tuple)):
            frames_list = [int(e.shape[0]) for e in audio_encodings]
            enc = torch.cat([e.reshape(e.shape[0], -1) for e in audio_encodings], dim=0)
        else:
            frames_list = [int(audio_encodings.shape[1])] * audio_encodings.shape[0]
            enc = audio_encodings.reshape(-1, audio_encodings.shape[-1])
        enc = enc.to(device=dev, dtype=F32).contiguous()
        feats = _linear(enc, P["model.mm_projector.weight"], bias=P["model.mm_projector.bias"], mode=mode)
        frames = torch.tensor(frames_list, dtype=torch.int32, device=dev)
        offs = [0]
        for f in frames_list[:-1]:
            offs.append(offs[-1] + f)
        audio_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        start_pos = torch.empty(B, device=dev, dtype=torch.int32)
        flags = torch.zeros(1, device=dev, dtype=torch.int32)
        ops.splice_scan(ids, int(aec.audio_start_token), int(aec.audio_end_token), frames, start_pos, flags)
        f = int(flags.item())
        if f & 1:
            raise ValueError("The number of image start tokens and image end tokens should be the same.")
        if f & 2:
            raise ValueError("The image end token should follow the image start token.")
        ops.v32_embed_splice_fwd(ids, table, feats, audio_off, frames, start_pos, x.view(B, S, H))
    else:
        ops.v32_embed_splice_fwd(ids, table, None, None, None, None, x.view(B, S, H))

    theta = getattr(cfg, "rope_theta", None)
    if theta is None:
        theta = (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0)
    # HF LlamaRotaryEmbedding: fp32 inv_freq, freqs, cos/sin (one-time table set-up, not part of the forward math)
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).to(device=dev, dtype=F32) / D))
    pos = torch.arange(S, device=dev, dtype=F32)
    emb = torch.cat((pos[:, None] * inv[None, :],) * 2, dim=-1)
    cos, sin = emb.cos().contiguous(), emb.sin().contiguous()

    for l in range(L):
        pfx = f"model.layers.{l}."
        Wqkv = torch.cat([P[pfx + f"self_attn.{n}_proj.weight"] for n in ("q", "k", "v")], dim=0)
        Wgu = torch.cat([P[pfx + "mlp.gate_proj.weight"], P[pfx + "mlp.up_proj.weight"]], dim=0)
        xn1 = ops.v32_rmsnorm_fwd(x, P[pfx + "input_layernorm.weight"], eps)
        qkv = _linear(xn1, Wqkv, mode=mode)
        ops.v32_rope_(qkv, cos, sin, S, 2 * NH, D)
        attn = ops.v32_attn_fwd(qkv, mask_u8, B, S, NH, D)
        x_mid = _linear(attn, P[pfx + "self_attn.o_proj.weight"], residual=x, mode=mode)
        xn2 = ops.v32_rmsnorm_fwd(x_mid, P[pfx + "post_attention_layernorm.weight"], eps)
        act = ops.v32_swiglu_fwd(_linear(xn2, Wgu, mode=mode))
        x = _linear(act, P[pfx + "mlp.down_proj.weight"], residual=x_mid, mode=mode)
    hn = ops.v32_rmsnorm_fwd(x, P["model.norm.weight"], eps)
    ldv = (V + 3) // 4 * 4
    logits_buf = torch.empty((M, ldv), device=dev, dtype=F32)
    lm_w = P["lm_head.weight"]
    if mode == "tf32":
        ops.gemm(hn, lm_w, out=logits_buf, N=V)
    else:
        h_hi, h_lo = ops.v32_split_tf32(hn)
        w_hi, w_lo = ops.v32_split_tf32(lm_w)
        ops.gemm(h_lo, w_hi, out=logits_buf, N=V)
        ops.gemm(h_hi, w_lo, out=logits_buf, N=V, residual=logits_buf)
        ops.gemm(h_hi, w_hi, out=logits_buf, N=V, residual=logits_buf)
    logits = logits_buf.view(B, S, ldv)[:, :, :V]
    loss = None
    if labels is not None:
        lab = labels.to(device=dev, dtype=torch.int64).contiguous()
        loss = ops.v32_ce_fwd(logits_buf, V, lab, B, S)
    return loss, logits
