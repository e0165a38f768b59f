"""Execution schedule of the CLAP + MPT-1B variant (BASELINE configs[4]) on the engine's flat
buckets: what m2t/models/mpt.py WrappedMPTForCausalLM + m2t/llava/model/mpt/{modeling_mpt,blocks,attention,norm}.py
compute, as a manual forward / backward over C-ABI ops.

STATUS: runs on the B200 since round 2 (tests/test_mpt_gpu.py in `pytest -m gpu`, tools/bench_mpt.py: 392 samples/s on one
GPU, 757 on two). The schedule is verified on CPU against oracle/mpt_oracle.py (itself pinned to the reference's classes)
with the test-only ops stand-in (tests/test_mpt_host_cpu.py). The ops only this variant calls -- `ln_fwd`, `ln_bwd`
(bias-free LayerNorm, also over the q / k thirds of the packed qkv rows), `gelu_fwd`, `gelu_bwd` (csrc/mpt_ops.cu) and
`attn_alibi_fwd/bwd` (csrc/attention_tc_alibi.cu: the tcgen05 attention with the per-head slope * (key - (S-1)) bias and
no RoPE) -- have their per-op parity tests against fp32 references in tests/test_mpt_gpu.py. There is no fallback: a
failing kernel fails the step.

Per block (blocks.py:67-86), bf16 weights under bf16 autocast as m2t/train.py:101-103 sets up:
    a    = LN(x; norm_1)                         low-precision LayerNorm (norm.py:26-52), no bias (no_bias=True)
    qkv  = a @ Wqkv^T                            one fused [3d, d] projection (attention.py:245-247)
    q, k = LN_fp32(q; q_ln), LN_fp32(k; k_ln)    LayerNorm over the FULL d_model of q and of k (attention.py:289-292)
    ctx  = softmax(q k^T / sqrt(D) + slope_h * (key - (S-1)), causal, key padding) v     (attention.py:31-83, 458-484)
    x    = x + ctx @ Wo^T ;  m = LN(x; norm_2) ;  x = x + GELU(m @ Wup^T) @ Wdown^T      (blocks.py:10-20)
Head: LN(x; norm_f) @ wte^T (tied), loss = CE over labels rolled by one (m2t/models/mpt.py:311-330) -- the same shifted
cross entropy as the Llama path, so `llark_ce_fwd/bwd` serve unchanged. The tied table receives two gradient
contributions: the head's wgrad (every row) and the embedding scatter (marker rows only when text rows are detached)."""
import math

import torch

from . import ops
from .engine import BF16, LlamaEngine


def alibi_slopes(n_heads, alibi_bias_max=8):
    """gen_slopes of attention.py:458-465 (fp32 [n_heads])."""
    n = 2 ** math.ceil(math.log2(n_heads))
    m = torch.arange(1, n + 1, dtype=torch.float32) * (alibi_bias_max / n)
    slopes = 1.0 / torch.pow(2, m)
    if n != n_heads:
        slopes = torch.concat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes


class MPTEngine(LlamaEngine):
    """Shares the flat-bucket machinery, the autograd anchor and the flag word of LlamaEngine; the schedule differs."""

    def _dims(self):
        cfg = self.cfg
        return cfg.d_model, cfg.n_heads, cfg.n_layers, cfg.expansion_ratio * cfg.d_model

    def _slopes(self, dev):
        if getattr(self, "_alibi", None) is None or self._alibi.device != dev:
            self._alibi = alibi_slopes(self.cfg.n_heads, self.cfg.attn_config.get("alibi_bias_max", 8)).to(dev)
        return self._alibi

    # ------------------------------------------------------------------------------------------ forward
    def _forward(self, batch, need_grad, kv_sink=None):
        if kv_sink is not None:
            raise NotImplementedError("KV-cache decoding is not built for the MPT variant")
        cfg, flat = self.cfg, self.flat
        d, NH, L, E = self._dims()
        D = d // NH
        eps = 1e-5
        ids = batch["input_ids"].to(torch.int64).contiguous()
        dev = ids.device
        B, S = ids.shape
        M = B * S
        V = self.lm.transformer.wte.weight.shape[0]
        mask = batch["attention_mask"]
        mask_u8 = mask.to(torch.bool).contiguous().view(torch.uint8) if mask is not None else None
        table = flat.pview("transformer.wte.weight", (V, d))
        x = torch.empty((M, d), device=dev, dtype=BF16)
        model = self.lm.transformer
        aec = model.audio_encoder_config
        enc = batch["audio_encodings"]
        saved = {"B": B, "S": S, "ids": ids, "mask": mask_u8, "layers": []}
        if enc is not None and getattr(cfg, "use_mm_proj", False):
            if not aec.use_audio_start_end:
                raise NotImplementedError("audio_encoder_config.use_audio_start_end=False is not implemented.")
            if isinstance(enc, (list, tuple)):
                frames_list = [int(e.shape[0]) for e in enc]
                enc_cat = torch.cat([e.reshape(e.shape[0], -1) for e in enc], dim=0)
            else:
                frames_list = [int(enc.shape[1])] * enc.shape[0]
                enc_cat = enc.reshape(-1, enc.shape[-1])
            if len(frames_list) != B:
                raise ValueError("one audio encoding per sample is required")
            enc_cat = enc_cat.to(dev)
            enc_bf = ops.cast_f32_to_bf16(enc_cat.contiguous()) if enc_cat.dtype == torch.float32 else enc_cat.to(BF16)
            A = enc_bf.shape[1]
            feats = ops.gemm(enc_bf, flat.pview("transformer.mm_projector.weight", (d, A)),
                             bias=flat.pview("transformer.mm_projector.bias", (d,)))
            frames = torch.tensor(frames_list, dtype=torch.int32).to(dev)
            offs = [0]
            for f in frames_list[:-1]:
                offs.append(offs[-1] + f)
            audio_off = torch.tensor(offs, dtype=torch.int64).to(dev)
            start_pos = torch.empty(B, device=dev, dtype=torch.int32)
            if self.flags is None or self.flags.device != dev:
                self.flags = torch.zeros(1, device=dev, dtype=torch.int32)
            ops.splice_scan(ids, int(aec.audio_start_token), int(aec.audio_end_token), frames, start_pos, self.flags)
            ops.embed_splice_fwd(ids, table, feats, audio_off, frames, start_pos, x.view(B, S, d))
            saved.update(enc_bf=enc_bf, frames=frames, audio_off=audio_off, start_pos=start_pos, n_feat=feats.shape[0])
        else:
            ops.embed_splice_fwd(ids, table, None, None, None, None, x.view(B, S, d))
            saved.update(enc_bf=None, frames=None, audio_off=None, start_pos=None, n_feat=0)

        slopes = self._slopes(dev)

        def stats():
            return (torch.empty(M, device=dev, dtype=torch.float32), torch.empty(M, device=dev, dtype=torch.float32))

        for l in range(L):
            pfx = f"transformer.blocks.{l}."
            mu1, rs1 = stats()
            a = ops.ln_fwd(x, flat.pview(pfx + "norm_1.weight"), eps, mean=mu1, rstd=rs1)
            qkv_raw = ops.gemm(a, flat.pview(pfx + "attn.Wqkv.weight", (3 * d, d)))
            qkv = qkv_raw.clone()  # v is used as projected; q and k are replaced by their LayerNorms
            muq, rsq = stats()
            muk, rsk = stats()
            ops.ln_fwd(qkv_raw[:, :d], flat.pview(pfx + "attn.q_ln.weight"), eps, mean=muq, rstd=rsq, out=qkv[:, :d])
            ops.ln_fwd(qkv_raw[:, d:2 * d], flat.pview(pfx + "attn.k_ln.weight"), eps, mean=muk, rstd=rsk,
                       out=qkv[:, d:2 * d])
            ctx, lse = ops.attn_alibi_fwd(qkv, mask_u8, B, S, NH, D, slopes)
            x_mid = ops.gemm(ctx, flat.pview(pfx + "attn.out_proj.weight", (d, d)), residual=x)
            mu2, rs2 = stats()
            m = ops.ln_fwd(x_mid, flat.pview(pfx + "norm_2.weight"), eps, mean=mu2, rstd=rs2)
            u = ops.gemm(m, flat.pview(pfx + "ffn.up_proj.weight", (E, d)))
            h = ops.gelu_fwd(u)
            x_out = ops.gemm(h, flat.pview(pfx + "ffn.down_proj.weight", (d, E)), residual=x_mid)
            if need_grad:
                saved["layers"].append((x, mu1, rs1, a, qkv_raw, muq, rsq, muk, rsk, qkv, ctx, lse, x_mid, mu2, rs2, m, u, h))
            x = x_out

        muf = torch.empty(M, device=dev, dtype=torch.float32)
        rsf = torch.empty(M, device=dev, dtype=torch.float32)
        hn = ops.ln_fwd(x, flat.pview("transformer.norm_f.weight"), eps, mean=muf, rstd=rsf)
        ldv = (V + 63) // 64 * 64
        logits_buf = torch.empty((M, ldv), device=dev, dtype=BF16)
        ops.gemm(hn, table, out=logits_buf, N=V)  # tied head: logits = h @ wte^T
        logits = logits_buf.view(B, S, ldv)[:, :, :V]
        loss = None
        labels = batch["labels"]
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.int64).contiguous()
            row_loss = torch.empty(M, device=dev, dtype=torch.float32)
            row_lse = torch.empty(M, device=dev, dtype=torch.float32)
            loss_out = torch.empty(2, device=dev, dtype=torch.float32)
            ops.ce_fwd(logits_buf, ldv, V, labels, B, S, row_loss, row_lse, loss_out)
            loss = loss_out[0]
            if need_grad:
                saved.update(x_final=x, muf=muf, rsf=rsf, hn=hn, logits_buf=logits_buf, ldv=ldv, labels=labels,
                             row_lse=row_lse, loss_out=loss_out)
        self._saved = saved if need_grad else None
        return loss, logits

    # ------------------------------------------------------------------------------------------ backward
    def _backward(self, dloss):
        flat, sv = self.flat, self._saved
        if sv is None:
            raise RuntimeError("backward called without a saved forward (double backward is not supported)")
        self._saved = None
        d, NH, L, E = self._dims()
        D = d // NH
        B, S = sv["B"], sv["S"]
        M = B * S
        V = self.lm.transformer.wte.weight.shape[0]
        dev = sv["ids"].device
        acc = flat.grads_live()
        slopes = self._slopes(dev)

        def wgrad(name, dy, xin, shape, force_acc=False, **kw):
            g = flat.gview(name, shape)
            if g is None:
                return
            ops.gemm(dy, xin, a_mn=True, b_mn=True, out=g, residual=g if (acc or force_acc) else None, **kw)

        ldv = sv["ldv"]
        dlogits = torch.empty_like(sv["logits_buf"])
        dl = dloss.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        ops.ce_bwd(sv["logits_buf"], ldv, V, sv["labels"], B, S, sv["row_lse"], sv["loss_out"], dl, 1.0, dlogits)
        table = flat.pview("transformer.wte.weight", (V, d))
        dhn = ops.gemm(dlogits, table, b_mn=True, M=M, N=d, K=V, lda=ldv)
        # tied head: d(wte) = dlogits^T @ hn  (every row); the embedding scatter is added at the end
        wgrad("transformer.wte.weight", dlogits, sv["hn"], (V, d), M=V, N=d, K=M, lda=ldv)
        del dlogits
        dx = torch.empty((M, d), device=dev, dtype=BF16)
        ops.ln_bwd(sv["x_final"], flat.pview("transformer.norm_f.weight"), dhn, sv["muf"], sv["rsf"], None, dx,
                   flat.gview("transformer.norm_f.weight"), acc)

        for l in reversed(range(L)):
            pfx = f"transformer.blocks.{l}."
            (x, mu1, rs1, a, qkv_raw, muq, rsq, muk, rsk, qkv, ctx, lse, x_mid, mu2, rs2, m, u, h) = sv["layers"][l]
            sv["layers"][l] = None
            Wqkv = flat.pview(pfx + "attn.Wqkv.weight", (3 * d, d))
            Wo = flat.pview(pfx + "attn.out_proj.weight", (d, d))
            Wup = flat.pview(pfx + "ffn.up_proj.weight", (E, d))
            Wdn = flat.pview(pfx + "ffn.down_proj.weight", (d, E))
            # MLP: x_out = x_mid + GELU(m Wup^T) Wdn^T
            dh = ops.gemm(dx, Wdn, b_mn=True)
            wgrad(pfx + "ffn.down_proj.weight", dx, h, (d, E))
            du = ops.gelu_bwd(u, dh)
            del dh
            dm = ops.gemm(du, Wup, b_mn=True)
            wgrad(pfx + "ffn.up_proj.weight", du, m, (E, d))
            del du
            dx_mid = torch.empty((M, d), device=dev, dtype=BF16)
            ops.ln_bwd(x_mid, flat.pview(pfx + "norm_2.weight"), dm, mu2, rs2, dx, dx_mid,
                       flat.gview(pfx + "norm_2.weight"), acc)
            # attention: x_mid = x + ctx Wo^T
            dctx = ops.gemm(dx_mid, Wo, b_mn=True)
            wgrad(pfx + "attn.out_proj.weight", dx_mid, ctx, (d, d))
            dqkv = ops.attn_alibi_bwd(qkv, sv["mask"], ctx, dctx, lse, B, S, NH, D, slopes)
            # through the q / k LayerNorms back to the raw projection (v passes straight through)
            dqkv_raw = dqkv.clone()
            ops.ln_bwd(qkv_raw[:, :d], flat.pview(pfx + "attn.q_ln.weight"), dqkv[:, :d], muq, rsq, None,
                       dqkv_raw[:, :d], flat.gview(pfx + "attn.q_ln.weight"), acc)
            ops.ln_bwd(qkv_raw[:, d:2 * d], flat.pview(pfx + "attn.k_ln.weight"), dqkv[:, d:2 * d], muk, rsk, None,
                       dqkv_raw[:, d:2 * d], flat.gview(pfx + "attn.k_ln.weight"), acc)
            da = ops.gemm(dqkv_raw, Wqkv, b_mn=True)
            wgrad(pfx + "attn.Wqkv.weight", dqkv_raw, a, (3 * d, d))
            ops.ln_bwd(x, flat.pview(pfx + "norm_1.weight"), da, mu1, rs1, dx_mid, dx,
                       flat.gview(pfx + "norm_1.weight"), acc)
            if self.grad_ready_hook is not None and l % self.comm_chunk_layers == 0:
                # blocks [l, hi) are final and contiguous in the bucket: their all-reduce can start now (the tied table,
                # final norm and projector are reduced by the optimizer's step())
                lo_off = flat.offsets[f"transformer.blocks.{l}.norm_1.weight"][0]
                hi_l = min(l + self.comm_chunk_layers, L)
                hi_off = (flat.offsets[f"transformer.blocks.{hi_l}.norm_1.weight"][0] if hi_l < L
                          else flat.offsets["transformer.norm_f.weight"][0])
                if lo_off < flat.n_train and hi_off <= flat.n_train and hi_off > lo_off:
                    self.grad_ready_hook(lo_off, hi_off)

        # splice + projector + embedding scatter into the (already written) tied-table gradient
        g_table = flat.gview("transformer.wte.weight", (V, d))
        detach_text = getattr(self.lm.transformer, "orig_embeds_params", None) is not None
        n_feat = sv["n_feat"]
        daudio = torch.empty((max(n_feat, 1), d), device=dev, dtype=BF16) if n_feat else None
        ops.embed_splice_bwd(sv["ids"], dx.view(B, S, d), daudio, sv["audio_off"], sv["frames"], sv["start_pos"], g_table,
                             V, detach_text)
        if n_feat:
            A = sv["enc_bf"].shape[1]
            wgrad("transformer.mm_projector.weight", daudio, sv["enc_bf"], (d, A))
            g_b = flat.gview("transformer.mm_projector.bias")
            if g_b is not None:
                ops.colsum(daudio, g_b, accumulate=acc)
        elif not acc:
            for nm in ("transformer.mm_projector.weight", "transformer.mm_projector.bias"):
                if nm in flat.offsets and flat.gview(nm) is not None:
                    flat.gview(nm).zero_()
        flat.attach_grads()
