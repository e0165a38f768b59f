"""ORACLE (test infrastructure only -- never imported by the product path).

CPU/PyTorch restatement of the LLark model half of the hot path: embedding gather, mm_projector, audio-token
splice, the Llama decoder stack, lm_head and the shifted cross-entropy.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

What it follows (file:line under the reference root /root/reference, or the installed transformers 5.5.0 for the
arithmetic the reference delegates to `transformers.LlamaModel`, which is NOT vendored in the reference):

  embed + projector + splice   m2t/models/llamav2.py:124-222
  decoder stack                m2t/models/llamav2.py:224 -> transformers/models/llama/modeling_llama.py
                               (LlamaRMSNorm:53-68, LlamaRotaryEmbedding:124-135, apply_rotary_pos_emb:138-168,
                                LlamaMLP:182-184, eager_attention_forward:199-222, LlamaDecoderLayer:303-333)
  lm_head + shifted CE         m2t/models/llamav2.py:311-325

Pinning: tests/golden/llama_tiny_ref.pt holds outputs of the reference's own WrappedLlamav2ForCausalLM (imported
from /root/reference in the build container by oracle/make_golden.py) on parameters produced by `init_params`
below; tests/test_oracle_cpu.py checks this restatement against them (fp32: 1e-5; bf16: 2e-2 rel-Frobenius).
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class LlamaCfg:
    hidden_size: int = 256
    intermediate_size: int = 512
    num_layers: int = 2
    num_heads: int = 2
    vocab_size: int = 304
    mm_hidden_size: int = 64
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    audio_start_token: int = 302
    audio_end_token: int = 303
    audio_patch_token: int = 301
    pad_token: int = 300

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads


def param_names(cfg):
    """Parameter names in the reference's state_dict order (SURVEY Appendix C probe)."""
    names = ["model.embed_tokens.weight"]
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        names += [p + "self_attn.q_proj.weight", p + "self_attn.k_proj.weight", p + "self_attn.v_proj.weight",
                  p + "self_attn.o_proj.weight", p + "mlp.gate_proj.weight", p + "mlp.up_proj.weight",
                  p + "mlp.down_proj.weight", p + "input_layernorm.weight", p + "post_attention_layernorm.weight"]
    names += ["model.norm.weight", "model.mm_projector.weight", "model.mm_projector.bias", "lm_head.weight"]
    return names


def param_shape(cfg, name):
    H, I, V, A = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.mm_hidden_size
    if name.endswith("embed_tokens.weight") or name == "lm_head.weight":
        return (V, H)
    if name.endswith(("q_proj.weight", "k_proj.weight", "v_proj.weight", "o_proj.weight")):
        return (H, H)
    if name.endswith(("gate_proj.weight", "up_proj.weight")):
        return (I, H)
    if name.endswith("down_proj.weight"):
        return (H, I)
    if name.endswith("mm_projector.weight"):
        return (H, A)
    if name.endswith("layernorm.weight") or name.endswith("norm.weight") or name.endswith("mm_projector.bias"):
        return (H,)
    raise KeyError(name)


def init_params(cfg, seed=0, dtype=torch.float32):
    """Seeded synthetic parameters (bf16-representable values so fp32 and bf16 runs start identical)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for n in param_names(cfg):
        shp = param_shape(cfg, n)
        if len(shp) == 1 and not n.endswith("bias"):
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif n.endswith("bias"):
            w = 0.02 * torch.randn(shp, generator=g)
        elif "embed_tokens" in n:
            w = torch.randn(shp, generator=g) * 0.5
        elif "lm_head" in n:
            w = torch.randn(shp, generator=g) * 0.05
        else:
            w = torch.randn(shp, generator=g) * (0.5 / math.sqrt(shp[1]))
        out[n] = w.to(torch.bfloat16).to(dtype)
    return out


def make_batch(cfg, B=2, S=48, F_frames=5, seed=0, pad_last=6):
    """Synthetic collated batch shaped like DataCollatorForSupervisedDataset output (m2t/data_modules.py:195-222)."""
    g = torch.Generator().manual_seed(1000 + seed)
    ids = torch.randint(3, cfg.pad_token, (B, S), generator=g)
    labels = ids.clone()
    for b in range(B):
        pos = 4 + 3 * b
        ids[b, pos] = cfg.audio_start_token
        ids[b, pos + 1: pos + 1 + F_frames] = cfg.audio_patch_token
        ids[b, pos + 1 + F_frames] = cfg.audio_end_token
        labels[b] = ids[b]
        labels[b, : pos + F_frames + 8] = -100
    if pad_last and B > 1:
        ids[-1, S - pad_last:] = cfg.pad_token
        labels[-1, S - pad_last:] = -100
    attention_mask = ids.ne(cfg.pad_token)
    audio = torch.randn(B, F_frames, cfg.mm_hidden_size, generator=g) * 0.5
    return dict(input_ids=ids, labels=labels, attention_mask=attention_mask, audio_encodings=audio)


# ------------------------------------------------------------------------------------------------ building blocks
def rmsnorm(x, w, eps):
    """transformers LlamaRMSNorm.forward (modeling_llama.py:53-68): fp32 variance, weight * x.to(input_dtype)."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rope_tables(S, D, theta, device, dtype):
    """LlamaRotaryEmbedding.forward (modeling_llama.py:124-135): fp32 cos/sin, cast to the activation dtype."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / D))
    pos = torch.arange(S, device=device, dtype=torch.float32)
    freqs = pos[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def attention(q, k, v, attention_mask):
    """Causal + key-padding attention, eager_attention_forward semantics (modeling_llama.py:199-222):
    scores scaled by 1/sqrt(d), additive mask, softmax in fp32, cast back, P @ V."""
    B, NH, S, D = q.shape
    scores = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(D)
    allow = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()[None, None]
    if attention_mask is not None:
        allow = allow & attention_mask[:, None, None, :].to(torch.bool)
    scores = scores.masked_fill(~allow, torch.finfo(scores.dtype).min)
    p = F.softmax(scores, dim=-1, dtype=torch.float32).to(q.dtype)
    return torch.matmul(p, v)


def splice(cfg, input_ids, inputs_embeds, audio_features, detach_text):
    """m2t/models/llamav2.py:141-222 for use_audio_start_end=True. audio_features: tensor [B,F,H] or list."""
    new = []
    cur_audio_idx = 0
    for cur_ids, cur_emb in zip(input_ids, inputs_embeds):
        n_start = int((cur_ids == cfg.audio_start_token).sum())
        n_end = int((cur_ids == cfg.audio_end_token).sum())
        if n_start != n_end:
            raise ValueError("The number of image start tokens and image end tokens should be the same.")
        starts = torch.where(cur_ids == cfg.audio_start_token)[0]
        if len(starts) == 0:
            new.append(cur_emb)
            continue
        for pos in starts:
            pos = int(pos)
            feats = audio_features[cur_audio_idx].to(cur_emb.device)
            nf = feats.shape[0]
            if cur_ids[pos + nf + 1] != cfg.audio_end_token:
                raise ValueError("The image end token should follow the image start token.")
            if detach_text:
                cur_new = torch.cat((cur_emb[:pos].detach(), cur_emb[pos: pos + 1], feats,
                                     cur_emb[pos + nf + 1: pos + nf + 2], cur_emb[pos + nf + 2:].detach()), dim=0)
            else:
                cur_new = torch.cat((cur_emb[: pos + 1], feats, cur_emb[pos + nf + 1:]), dim=0)
            cur_audio_idx += 1
        new.append(cur_new)
    return torch.stack(new, dim=0)


def forward(params, cfg, input_ids, attention_mask=None, labels=None, audio_encodings=None, detach_text=True,
            return_hidden=False):
    """Full forward of WrappedLlamav2ForCausalLM (llamav2.py:259-337). params: name -> tensor (fp32 or bf16).
    Returns (loss or None, logits)."""
    dt = params["model.embed_tokens.weight"].dtype
    H, NH, D = cfg.hidden_size, cfg.num_heads, cfg.head_dim
    B, S = input_ids.shape
    x = F.embedding(input_ids, params["model.embed_tokens.weight"])  # llamav2.py:124
    if audio_encodings is not None:
        W, b = params["model.mm_projector.weight"], params["model.mm_projector.bias"]
        if isinstance(audio_encodings, (list, tuple)):
            feats = [F.linear(a.to(dt), W, b) for a in audio_encodings]  # llamav2.py:128-131
        else:
            feats = F.linear(audio_encodings.to(dt), W, b)  # llamav2.py:133
        x = splice(cfg, input_ids, x, feats, detach_text)
    cos, sin = rope_tables(S, D, cfg.rope_theta, x.device, dt)
    cos, sin = cos[None, None], sin[None, None]
    hiddens = []
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        res = x
        h = rmsnorm(x, params[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = F.linear(h, params[p + "self_attn.q_proj.weight"]).view(B, S, NH, D).transpose(1, 2)
        k = F.linear(h, params[p + "self_attn.k_proj.weight"]).view(B, S, NH, D).transpose(1, 2)
        v = F.linear(h, params[p + "self_attn.v_proj.weight"]).view(B, S, NH, D).transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)
        k = (k * cos) + (rotate_half(k) * sin)
        a = attention(q, k, v, attention_mask).transpose(1, 2).reshape(B, S, H)
        x = res + F.linear(a, params[p + "self_attn.o_proj.weight"])
        res = x
        h = rmsnorm(x, params[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        m = F.linear(F.silu(F.linear(h, params[p + "mlp.gate_proj.weight"])) *
                     F.linear(h, params[p + "mlp.up_proj.weight"]), params[p + "mlp.down_proj.weight"])
        x = res + m
        hiddens.append(x)
    x = rmsnorm(x, params["model.norm.weight"], cfg.rms_norm_eps)
    logits = F.linear(x, params["lm_head.weight"])  # llamav2.py:312
    loss = None
    if labels is not None:  # llamav2.py:314-325 (CrossEntropyLoss upcasts bf16 logits to fp32 under autocast)
        shift_logits = logits[..., :-1, :].contiguous().view(-1, cfg.vocab_size).float()
        shift_labels = labels[..., 1:].contiguous().view(-1)
        loss = F.cross_entropy(shift_logits, shift_labels, ignore_index=-100)
    if return_hidden:
        return loss, logits, hiddens
    return loss, logits


def forward_backward(params, cfg, batch, detach_text=True, freeze_lm_head=True):
    """Loss, logits and gradients (dict name -> grad or None) with the reference's trainable set."""
    ps = {}
    for n, t in params.items():
        t = t.detach().clone()
        t.requires_grad_(not (freeze_lm_head and n == "lm_head.weight"))
        ps[n] = t
    loss, logits = forward(ps, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"],
                           batch["audio_encodings"], detach_text=detach_text)
    loss.backward()
    grads = {n: (t.grad if t.grad is not None else None) for n, t in ps.items()}
    return loss.detach(), logits.detach(), grads


def adamw_reference(p, g, m, v, step, lr=5e-5, beta1=0.9, beta2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.AdamW single-tensor update in the tensor's own dtype (bf16 -> rounding after each op), the
    regime the reference trains in (m2t/train.py:101-103 + optim=adamw_torch, m2t/arguments.py:84)."""
    p = p * (1 - lr * wd) if wd else p
    m = torch.lerp(m, g, 1 - beta1)
    v = v * beta2
    v = torch.addcmul(v, g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = torch.addcdiv(p, m, denom, value=-(lr / bc1))
    return p, m, v
