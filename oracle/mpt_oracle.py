"""ORACLE (test infrastructure only; never imported by the product): CPU restatement, in plain PyTorch, of the
reference's CLAP + MPT-1B variant of the hot path (BASELINE configs[4]) -- what
`m2t/models/mpt.py` WrappedMPTForCausalLM computes on top of m2t/llava/model/mpt/{modeling_mpt,blocks,attention,norm}.py:

    embed (wte) + mm_projector + splice            m2t/models/mpt.py:104-236 (same splice rules as the Llama model)
    n_layers x MPTBlock                            blocks.py:67-86: x += attn(norm_1(x)); x += ffn(norm_2(x))
      LPLayerNorm (no bias with no_bias=True)      norm.py:26-52
      fused Wqkv, LayerNorm over the FULL d_model
      of q and of k (qk_ln), ALiBi, causal softmax attention.py:275-313, :31-83, :458-484
      GELU(exact) MLP, expansion 4                 blocks.py:10-20
    norm_f, logits = h @ wte^T (tied), rolled CE   m2t/models/mpt.py:311-330

Pinned by tests/golden/mpt_tiny_ref.pt = outputs of the reference's own classes imported in place
(oracle/make_golden_mpt.py); tests/test_mpt_oracle_cpu.py. No kernels exist for this row yet: this file and its fixture
are the first step (oracle + boundary) of building it."""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class MPTCfg:
    d_model: int = 64
    n_heads: int = 2
    n_layers: int = 2
    expansion_ratio: int = 4
    max_seq_len: int = 128
    vocab_size: int = 304
    mm_hidden_size: int = 16
    alibi_bias_max: int = 8
    ln_eps: float = 1e-5
    audio_start_token: int = 302
    audio_end_token: int = 303
    audio_patch_token: int = 301
    pad_token: int = 300

    @property
    def head_dim(self):
        return self.d_model // self.n_heads


def param_names(cfg):
    names = ["transformer.wte.weight"]
    for i in range(cfg.n_layers):
        p = f"transformer.blocks.{i}."
        names += [p + "norm_1.weight", p + "attn.Wqkv.weight", p + "attn.q_ln.weight", p + "attn.k_ln.weight",
                  p + "attn.out_proj.weight", p + "norm_2.weight", p + "ffn.up_proj.weight", p + "ffn.down_proj.weight"]
    names += ["transformer.norm_f.weight", "transformer.mm_projector.weight", "transformer.mm_projector.bias"]
    return names


def param_shape(cfg, name):
    d, e = cfg.d_model, cfg.expansion_ratio
    if name.endswith("wte.weight"):
        return (cfg.vocab_size, d)
    if name.endswith("Wqkv.weight"):
        return (3 * d, d)
    if name.endswith("out_proj.weight"):
        return (d, d)
    if name.endswith("up_proj.weight"):
        return (e * d, d)
    if name.endswith("down_proj.weight"):
        return (d, e * d)
    if name.endswith("mm_projector.weight"):
        return (d, cfg.mm_hidden_size)
    return (d,)  # norms, q_ln / k_ln, projector bias


def init_params(cfg, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for n in param_names(cfg):
        shp = param_shape(cfg, n)
        if len(shp) == 1 and not n.endswith("mm_projector.bias"):
            out[n] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
        else:
            out[n] = (torch.randn(shp, generator=g) * 0.05).to(dtype)
    return out


def make_batch(cfg, B=2, S=40, F_frames=1, seed=0, pad_last=5):
    """ids with <audio_start> <audio_patch>*F <audio_end> at position 3, labels on the tail, right padding on the last
    row; audio_encodings [B, F, mm_hidden] (CLAP gives F = 1: one 512-d embedding per clip)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 300, (B, S), generator=g)
    pos = 3
    ids[:, pos] = cfg.audio_start_token
    ids[:, pos + 1: pos + 1 + F_frames] = cfg.audio_patch_token
    ids[:, pos + 1 + F_frames] = cfg.audio_end_token
    labels = torch.full((B, S), -100, dtype=torch.int64)
    labels[:, S // 2:] = ids[:, S // 2:]
    mask = torch.ones(B, S, dtype=torch.bool)
    if pad_last:
        ids[-1, S - pad_last:] = cfg.pad_token
        labels[-1, S - pad_last:] = -100
        mask[-1, S - pad_last:] = False
    enc = torch.randn(B, F_frames, cfg.mm_hidden_size, generator=g) * 0.5
    return dict(input_ids=ids, attention_mask=mask, labels=labels, audio_encodings=enc)


def gen_slopes(n_heads, alibi_bias_max=8):
    """attention.py:458-465."""
    _n = 2 ** math.ceil(math.log2(n_heads))
    m = torch.arange(1, _n + 1, dtype=torch.float32) * (alibi_bias_max / _n)
    slopes = 1.0 / torch.pow(2, m)
    if _n != n_heads:
        slopes = torch.concat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes


def layernorm(x, w, eps, low_precision):
    """LPLayerNorm (norm.py:26-52): under autocast the input and weight are cast DOWN to the autocast dtype and
    F.layer_norm runs there; q_ln / k_ln are plain nn.LayerNorm, which autocast runs in fp32 (low_precision=False)."""
    if low_precision:
        return F.layer_norm(x, x.shape[-1:], w.to(x.dtype), None, eps)
    return F.layer_norm(x.float(), x.shape[-1:], w.float(), None, eps)


def attention(q, k, v, cfg, attention_mask):
    """scaled_multihead_dot_product_attention (attention.py:31-83) with the ALiBi bias of build_alibi_bias (:468-484,
    causal form: slope_h * (key - (S-1))), key-padding mask folded into the bias with finfo.min (modeling_mpt.py:
    133-147) and the causal mask applied with finfo.min on top."""
    B, S, d = q.shape
    H, D = cfg.n_heads, cfg.head_dim
    qh = q.view(B, S, H, D).transpose(1, 2)
    kh = k.view(B, S, H, D).permute(0, 2, 3, 1)
    vh = v.view(B, S, H, D).transpose(1, 2)
    w = qh.matmul(kh) * (1.0 / math.sqrt(D))
    min_val = torch.finfo(q.dtype).min
    bias = (torch.arange(1 - S, 1, dtype=torch.int32).view(1, 1, 1, S) * gen_slopes(H, cfg.alibi_bias_max).view(1, H, 1, 1)).to(q.dtype)
    if attention_mask is not None:
        bias = bias.expand(B, H, 1, S).masked_fill(~attention_mask.view(B, 1, 1, S).bool(), min_val)
    w = w + bias
    causal = ~torch.ones(S, S, dtype=torch.bool).tril()
    w = w.masked_fill(causal.view(1, 1, S, S), min_val)
    p = torch.softmax(w, dim=-1)
    return p.matmul(vh).transpose(1, 2).reshape(B, S, d)


def splice(cfg, input_ids, inputs_embeds, audio_features, detach_text):
    """m2t/models/mpt.py:117-236, use_audio_start_end=True branch (identical rules to the Llama wrapper)."""
    new = []
    idx = 0
    for cur_ids, cur_emb in zip(input_ids, inputs_embeds):
        n_start = int((cur_ids == cfg.audio_start_token).sum())
        if n_start != int((cur_ids == cfg.audio_end_token).sum()):
            raise ValueError("The number of image start tokens and image end tokens should be the same.")
        starts = torch.where(cur_ids == cfg.audio_start_token)[0]
        cur_new = cur_emb
        for pos in starts:
            pos = int(pos)
            feats = audio_features[idx]
            nf = feats.shape[0]
            if cur_ids[pos + nf + 1] != cfg.audio_end_token:
                raise ValueError("The image end token should follow the image start token.")
            if detach_text:
                cur_new = torch.cat((cur_emb[:pos].detach(), cur_emb[pos: pos + 1], feats,
                                     cur_emb[pos + nf + 1: pos + nf + 2], cur_emb[pos + nf + 2:].detach()), dim=0)
            else:
                cur_new = torch.cat((cur_emb[: pos + 1], feats, cur_emb[pos + nf + 1:]), dim=0)
            idx += 1
        new.append(cur_new)
    return torch.stack(new, dim=0)


def forward(params, cfg, input_ids, attention_mask=None, labels=None, audio_encodings=None, detach_text=False,
            autocast_dtype=None):
    """WrappedMPTForCausalLM.forward. params fp32 (or bf16). autocast_dtype=torch.bfloat16 reproduces HF Trainer's
    bf16 autocast over fp32/bf16 weights: Linear / matmul in bf16, LPLayerNorm in bf16, q_ln / k_ln, softmax and the
    cross entropy in fp32. Returns (loss or None, logits)."""
    ac = autocast_dtype

    def lin(x, w, b=None):
        if ac is not None:
            return F.linear(x.to(ac), w.to(ac), None if b is None else b.to(ac))
        return F.linear(x, w, b)

    x = F.embedding(input_ids, params["transformer.wte.weight"])
    if audio_encodings is not None:
        W, b = params["transformer.mm_projector.weight"], params["transformer.mm_projector.bias"]
        if isinstance(audio_encodings, (list, tuple)):
            feats = [lin(a.to(x.dtype), W, b) for a in audio_encodings]
        else:
            feats = lin(audio_encodings.to(x.dtype), W, b)
        feats = [f.to(x.dtype) for f in feats] if isinstance(feats, list) else feats.to(x.dtype)
        x = splice(cfg, input_ids, x, feats, detach_text)
    d = cfg.d_model
    for i in range(cfg.n_layers):
        p = f"transformer.blocks.{i}."
        a = layernorm(x.to(ac) if ac is not None else x, params[p + "norm_1.weight"], cfg.ln_eps, True)
        qkv = lin(a, params[p + "attn.Wqkv.weight"])
        q, k, v = qkv.chunk(3, dim=2)
        dt = q.dtype
        q = layernorm(q, params[p + "attn.q_ln.weight"], cfg.ln_eps, ac is None).to(dt)
        k = layernorm(k, params[p + "attn.k_ln.weight"], cfg.ln_eps, ac is None).to(dt)
        ctx = attention(q, k, v, cfg, attention_mask)
        x = x + lin(ctx, params[p + "attn.out_proj.weight"])
        m = layernorm(x.to(ac) if ac is not None else x, params[p + "norm_2.weight"], cfg.ln_eps, True)
        h = F.gelu(lin(m, params[p + "ffn.up_proj.weight"]))
        x = x + lin(h, params[p + "ffn.down_proj.weight"])
    x = layernorm(x.to(ac) if ac is not None else x, params["transformer.norm_f.weight"], cfg.ln_eps, True)
    logits = lin(x, params["transformer.wte.weight"])
    loss = None
    if labels is not None:
        lab = torch.roll(labels, shifts=-1)
        lab[:, -1] = -100
        loss = F.cross_entropy(logits.view(-1, logits.size(-1)).float(), lab.view(-1))
    return loss, logits


def forward_backward(params, cfg, batch, detach_text=False, autocast_dtype=None):
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss, logits = forward(ps, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"],
                           batch["audio_encodings"], detach_text, autocast_dtype)
    loss.backward()
    return loss.detach(), logits.detach(), {k: (v.grad.detach() if v.grad is not None else None) for k, v in ps.items()}
