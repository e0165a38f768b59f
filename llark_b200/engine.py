"""Execution engine of the LLark/Llama training step on B200.

Host-side orchestration only: every FLOP is a C-ABI call (llark_b200.ops -> libllark_b200.so). The engine owns

  * the flat bf16 parameter bucket (all nn.Parameters of the model are re-pointed to views of it, trainable
    parameters first) and the flat bf16 gradient bucket (`.grad` of every trainable parameter is a view of it):
    one ncclAllReduce + one fused AdamW kernel cover the whole model, q/k/v and gate/up projections are adjacent
    so each is ONE fused GEMM without copying weights, and wgrad GEMMs accumulate in place in their epilogue;
  * the forward/backward schedule of the decoder stack (manual backward: dgrad/wgrad GEMMs read weights and saved
    activations in place through MN-major UMMA descriptors, no transposes are materialised).

Reference semantics reproduced: pure-bf16 weights/grads, loss = per-rank mean over
non-ignored shifted tokens computed on bf16 logits, lm_head frozen when requires_grad is False, text-token
embedding rows detached when `orig_embeds_params` is set (m2t/models/llamav2.py:176-198).
"""
import os

import torch

from . import ops

BF16 = torch.bfloat16
_ALIGN = 128  # elements (256 B): keeps every view 16-byte aligned for TMA and fused views gap-free


class _Nvtx:
    """NVTX ranges around the phases of a step (LLARK_NVTX=1; SURVEY section 5 tracing): `llark.forward`, `llark.backward`,
    `llark.layer_fwd[l]` / `llark.layer_bwd[l]`, and `llark.optimizer` (optim.py), so `ncu --nvtx --nvtx-include` and any
    timeline tool can cut a capture down to one phase or one layer. Off by default: two calls per range otherwise."""
    on = os.environ.get("LLARK_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.on and torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if _Nvtx.on and torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


class FlatBuckets:
    """Packs a list of (name, nn.Parameter) into one flat bf16 buffer, trainable parameters first."""

    def __init__(self, named_params, device, cast_to_bf16=False):
        self.device = device
        self.entries = []  # (name, param, offset, numel)
        train = [(n, p) for n, p in named_params if p.requires_grad]
        frozen = [(n, p) for n, p in named_params if not p.requires_grad]
        off = 0
        for group in (train, frozen):
            for n, p in group:
                if p.dtype != BF16 and not cast_to_bf16:
                    raise TypeError(
                        f"{n} is {p.dtype}: llark_b200 runs the reference's pure-bf16 regime (m2t/train.py:101-103); "
                        "cast the model with .to(torch.bfloat16) first")
                self.entries.append((n, p, off, p.numel()))
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if group is train:
                self.n_train = off
        self.total = off
        self.params = torch.empty(self.total, device=device, dtype=BF16)
        self.params.zero_()
        self.grads = torch.zeros(max(self.n_train, _ALIGN), device=device, dtype=BF16)
        self.signature = tuple((n, p.requires_grad, tuple(p.shape)) for n, p, _, _ in self.entries)
        self.offsets = {n: (o, k) for n, _, o, k in self.entries}
        for n, p, o, k in self.entries:
            view = self.params[o:o + k].view(p.shape)
            view.copy_(p.data)  # converts when cast_to_bf16 admitted another storage dtype
            p.data = view
        self.grads_attached = False
        # set by the fused optimizer step: the bucket's contents have been consumed, so the next backward
        # OVERWRITES instead of accumulating (saves zeroing 13 GB per optimizer step)
        self.consumed = False

    def is_current(self, named_params):
        sig = tuple((n, p.requires_grad, tuple(p.shape)) for n, p in named_params)
        if sorted(sig) != sorted(self.signature):
            return False
        base = self.params.data_ptr()
        for n, p, o, k in self.entries:
            if p.data_ptr() != base + 2 * o:
                return False
        return True

    def pview(self, name, shape=None, span=1):
        """View of `span` consecutive parameters starting at `name` (fused q|k|v, gate|up)."""
        o, k = self.offsets[name]
        if span == 1:
            t = self.params[o:o + k]
            return t.view(shape) if shape is not None else t
        names = [e[0] for e in self.entries]
        i = names.index(name)
        tot = 0
        for j in range(span):
            oj, kj = self.entries[i + j][2], self.entries[i + j][3]
            if oj != o + tot:
                raise RuntimeError(f"parameters after {name} are not contiguous in the flat bucket")
            tot += kj
        return self.params[o:o + tot].view(shape)

    def gview(self, name, shape=None, span=1):
        o, k = self.offsets[name]
        if o >= self.n_train:
            return None
        if span > 1:
            names = [e[0] for e in self.entries]
            i = names.index(name)
            k = sum(self.entries[i + j][3] for j in range(span))
        t = self.grads[o:o + k]
        return t.view(shape) if shape is not None else t

    def grads_live(self):
        """True when every trainable parameter's .grad is still our view (gradient accumulation in progress)."""
        if not self.grads_attached or self.consumed:
            return False
        gbase = self.grads.data_ptr()
        for n, p, o, k in self.entries:
            if o >= self.n_train:
                continue
            if p.grad is None or p.grad.data_ptr() != gbase + 2 * o:
                return False
        return True

    def attach_grads(self):
        for n, p, o, k in self.entries:
            if o >= self.n_train:
                continue
            p.grad = self.grads[o:o + k].view(p.shape)
        self.grads_attached = True
        self.consumed = False


class KVCache:
    """Per-layer key / value store of the generation path (what transformers hands around as `past_key_values`):
    k[l], v[l] are bf16 [B, s_max, H] in the token-major layout of the packed qkv buffer (keys post-RoPE); `length` is
    the number of cached positions, the same for every row (the reference decodes one prompt at a time)."""

    def __init__(self, n_layers, B, s_max, H, device):
        self.k = [torch.empty((B, s_max, H), device=device, dtype=BF16) for _ in range(n_layers)]
        self.v = [torch.empty((B, s_max, H), device=device, dtype=BF16) for _ in range(n_layers)]
        self.B, self.s_max, self.H = B, s_max, H
        self.length = 0
        self.lens = torch.zeros(B, device=device, dtype=torch.int32)

    def __len__(self):  # `if past_key_values:` of prepare_inputs_for_generation (m2t/models/llamav2.py:346)
        return self.length

    def get_seq_length(self, layer_idx=0):
        return self.length


class _StepFn(torch.autograd.Function):
    """Autograd anchor: forward runs the whole CUDA forward, backward the whole manual CUDA backward, which writes
    the flat gradient bucket and attaches `.grad` views (so no per-parameter autograd accumulation kernels)."""

    @staticmethod
    def forward(ctx, anchor, engine, batch):
        with _Nvtx("llark.forward"):
            loss, logits = engine._forward(batch, need_grad=True)
        ctx.engine = engine
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        with _Nvtx("llark.backward"):
            ctx.engine._backward(dloss)
        return None, None, None


class LlamaEngine:
    def __init__(self, causal_lm):
        self.lm = causal_lm
        self.cfg = causal_lm.config
        self.flat = None
        self._rope = {}
        self._saved = None
        self._ws = {}
        # data-parallel overlap: called as hook(start_elem, end_elem) when a contiguous slice of the flat gradient
        # bucket is final for this backward (set by the optimizer when world > 1; None = no overlap)
        self.grad_ready_hook = None
        self.flags = None  # device int32[1]: splice validation bits (see llark_splice_scan)
        self.launches = 0  # kernels launched by the last forward(+backward): bench.py's gpu_launches
        # layers per all-reduce slice (1.6 GB at 4 for the 7B model); the last slice is the exposed one
        self.comm_chunk_layers = max(1, int(os.environ.get("LLARK_COMM_CHUNK_LAYERS", "4")))
        # RoPE and SwiGLU-backward fused into GEMM epilogues (llark_gemm_ex); LLARK_FUSE_EPILOGUES=0 keeps the
        # separate llark_rope / llark_swiglu_bwd kernels (same results bit for bit)
        self.fuse_epilogues = os.environ.get("LLARK_FUSE_EPILOGUES", "1") != "0"
        # Activation recompute (the reference's --gradient_checkpointing True, m2t/models/llamav2.py:241,255-257):
        # "auto" = only when the model's gradient_checkpointing flag is set AND the saved activations of the step would
        # not fit comfortably in free HBM (a 7B step at S=512, B=16 keeps 35 GB and does not need it); "1" / "0" force it.
        self.recompute = os.environ.get("LLARK_RECOMPUTE", "auto")
        self.recomputed_last_step = False

    # ------------------------------------------------------------------------------------------ parameters
    def _named(self):
        return [(n, p) for n, p in self.lm.named_parameters()]

    def ensure_packed(self):
        named = self._named()
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("llark_b200 has no CPU path: move the model to a CUDA device (B200) first")
        if self.flat is None or not self.flat.is_current(named):
            # Inference scripts load the checkpoint in fp16 and run under torch.autocast(bfloat16)
            # (m2t/models/utils.py:129, scripts/inference/infer_from_webdataset.py:82): every matmul then sees the
            # weights rounded to bf16. With gradients disabled the bucket is built by that same rounding; training keeps
            # the strict check (a silent down-cast of fp32 master weights would change the optimizer's regime).
            cast = not torch.is_grad_enabled()
            if cast and any(p.dtype != BF16 for _, p in named):
                import logging

                logging.warning("llark_b200: parameters held in %s are computed (and now stored) in bfloat16, as "
                                "torch.autocast(bfloat16) computes them in the reference's inference scripts",
                                sorted({str(p.dtype) for _, p in named if p.dtype != BF16}))
            self.flat = FlatBuckets(named, dev, cast_to_bf16=cast)
        return self.flat

    def _tables(self, S, device):
        key = (S, device)
        if key not in self._rope:
            D = self.cfg.hidden_size // self.cfg.num_attention_heads
            theta = getattr(self.cfg, "rope_theta", None)
            if theta is None:
                theta = (getattr(self.cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0)
            # HF LlamaRotaryEmbedding: fp32 cos/sin cast to the activation dtype (one-time setup, not the hot path)
            inv = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).to(device=device, dtype=torch.float) / D))
            pos = torch.arange(S, device=device, dtype=torch.float32)
            emb = torch.cat((pos[:, None] * inv[None, :],) * 2, dim=-1)
            self._rope[key] = (emb.cos().to(BF16).contiguous(), emb.sin().to(BF16).contiguous())
        return self._rope[key]

    def _workspace(self, name, nbytes, device):
        t = self._ws.get(name)
        if t is None or t.numel() < nbytes or t.device != device:
            t = torch.empty(nbytes, device=device, dtype=torch.uint8)
            self._ws[name] = t
        return t

    # ------------------------------------------------------------------------------------------ public entry
    def run(self, input_ids, attention_mask, labels, audio_encodings):
        flat = self.ensure_packed()
        batch = dict(input_ids=input_ids, attention_mask=attention_mask, labels=labels,
                     audio_encodings=audio_encodings)
        need_grad = torch.is_grad_enabled() and labels is not None and flat.n_train > 0
        if need_grad:
            anchor = flat.entries[0][1]  # first trainable parameter
            return _StepFn.apply(anchor, self, batch)
        return self._forward(batch, need_grad=False)

    # ------------------------------------------------------------------------------------------ one decoder layer
    def _layer_forward(self, l, x, mask_u8, B, S, cos, sin, kv_sink=None):
        """Forward of decoder layer l on x [B*S, H]; returns (x_out, tensors the backward needs, kernels launched)."""
        cfg, flat = self.cfg, self.flat
        H, I, NH = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        D = H // NH
        M, dev, eps = B * S, x.device, cfg.rms_norm_eps
        n = 0
        pfx = f"model.layers.{l}."
        w_in = flat.pview(pfx + "input_layernorm.weight")
        w_post = flat.pview(pfx + "post_attention_layernorm.weight")
        Wqkv = flat.pview(pfx + "self_attn.q_proj.weight", (3 * H, H), span=3)
        Wo = flat.pview(pfx + "self_attn.o_proj.weight", (H, H))
        Wgu = flat.pview(pfx + "mlp.gate_proj.weight", (2 * I, H), span=2)
        Wd = flat.pview(pfx + "mlp.down_proj.weight", (H, I))
        rstd1 = torch.empty(M, device=dev, dtype=torch.float32)
        xn1 = ops.rmsnorm_fwd(x, w_in, eps, rstd=rstd1)
        if self.fuse_epilogues and D == 128:
            # RoPE of the q | k heads in the QKV projection's epilogue (bit-identical to gemm -> llark_rope)
            qkv = ops.gemm(xn1, Wqkv, rope=(cos, sin, S, 2 * H))
        else:
            qkv = ops.gemm(xn1, Wqkv)
            ops.rope_(qkv, cos, sin, S, NH, D)
            n += 1
        if kv_sink is not None:  # prefill of the generation path: keep the (post-RoPE) keys and the values
            q3 = qkv.view(B, S, 3, H)
            kv_sink.k[l][:, :S].copy_(q3[:, :, 1])
            kv_sink.v[l][:, :S].copy_(q3[:, :, 2])
        attn, lse = ops.attn_fwd(qkv, mask_u8, B, S, NH, D)
        x_mid = ops.gemm(attn, Wo, residual=x)
        rstd2 = torch.empty(M, device=dev, dtype=torch.float32)
        xn2 = ops.rmsnorm_fwd(x_mid, w_post, eps, rstd=rstd2)
        gu = ops.gemm(xn2, Wgu)
        act = ops.swiglu_fwd(gu)
        x_out = ops.gemm(act, Wd, residual=x_mid)
        n += 8
        return x_out, (x, rstd1, xn1, qkv, attn, lse, x_mid, rstd2, xn2, gu, act), n

    def _want_recompute(self, M, dev):
        if self.recompute in ("0", "1"):
            return self.recompute == "1"
        if not getattr(self.lm.model, "gradient_checkpointing", False) or dev.type != "cuda":
            return False
        cfg = self.cfg
        per_token = 2 * (8 * cfg.hidden_size + 3 * cfg.intermediate_size)  # bytes of bf16 activations kept per layer
        need = M * cfg.num_hidden_layers * per_token
        free, _ = torch.cuda.mem_get_info(dev)
        free += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)  # cached blocks are reusable
        return need > 0.6 * free

    # ------------------------------------------------------------------------------------------ forward
    def _forward(self, batch, need_grad, kv_sink=None):
        cfg, flat = self.cfg, self.flat
        ids = batch["input_ids"]
        if ids.dtype != torch.int64:
            ids = ids.to(torch.int64)
        ids = ids.contiguous()
        dev = ids.device
        B, S = ids.shape
        H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        NH = cfg.num_attention_heads
        D = H // NH
        V = self.lm.lm_head.weight.shape[0]
        M = B * S
        eps = cfg.rms_norm_eps
        mask = batch["attention_mask"]
        mask_u8 = None
        if mask is not None:
            mask_u8 = mask.to(torch.bool).contiguous().view(torch.uint8)
        n0 = self.launches = 0

        table = flat.pview("model.embed_tokens.weight", (V, H))
        x = torch.empty((M, H), device=dev, dtype=BF16)
        model = self.lm.model
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
            if enc_cat.dtype == torch.float32:
                enc_bf = ops.cast_f32_to_bf16(enc_cat.contiguous())
                n0 += 1
            else:
                enc_bf = enc_cat.to(BF16).contiguous()
            A = enc_bf.shape[1]
            Wp = flat.pview("model.mm_projector.weight", (H, A))
            bp = flat.pview("model.mm_projector.bias", (H,))
            feats = ops.gemm(enc_bf, Wp, bias=bp)  # [sumF, H]
            frames = torch.tensor(frames_list, dtype=torch.int32).to(dev, non_blocking=True)
            offs = [0]
            for f in frames_list[:-1]:
                offs.append(offs[-1] + f)
            audio_off = torch.tensor(offs, dtype=torch.int64).to(dev, non_blocking=True)
            start_pos = torch.empty(B, device=dev, dtype=torch.int32)
            if self.flags is None or self.flags.device != dev:
                self.flags = torch.zeros(1, device=dev, dtype=torch.int32)
            ops.splice_scan(ids, int(aec.audio_start_token), int(aec.audio_end_token), frames, start_pos, self.flags)
            ops.embed_splice_fwd(ids, table, feats, audio_off, frames, start_pos, x.view(B, S, H))
            n0 += 3
            saved.update(enc_bf=enc_bf, frames=frames, audio_off=audio_off, start_pos=start_pos, n_feat=feats.shape[0])
        else:
            ops.embed_splice_fwd(ids, table, None, None, None, None, x.view(B, S, H))
            n0 += 1
            saved.update(enc_bf=None, frames=None, audio_off=None, start_pos=None, n_feat=0)

        cos, sin = self._tables(S, dev)
        recompute = need_grad and self._want_recompute(M, dev)
        self.recomputed_last_step = recompute
        for l in range(L):
            with _Nvtx(f"llark.layer_fwd[{l}]"):
                x_out, keep, n = self._layer_forward(l, x, mask_u8, B, S, cos, sin, kv_sink)
            n0 += n
            if need_grad:
                # with recompute only the layer's input survives the forward; _backward regenerates the rest
                saved["layers"].append((x,) if recompute else keep)
            del keep
            x = x_out

        w_norm = flat.pview("model.norm.weight")
        rstd_f = torch.empty(M, device=dev, dtype=torch.float32)
        hn = ops.rmsnorm_fwd(x, w_norm, eps, rstd=rstd_f)
        ldv = (V + 63) // 64 * 64
        logits_buf = torch.empty((M, ldv), device=dev, dtype=BF16)
        Wlm = flat.pview("lm_head.weight", (V, H))
        ops.gemm(hn, Wlm, out=logits_buf, N=V)
        n0 += 2
        logits = logits_buf.view(B, S, ldv)[:, :, :V]
        loss = None
        labels = batch["labels"]
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.int64).contiguous()
            row_loss = torch.empty(M, device=dev, dtype=torch.float32)
            row_lse = torch.empty(M, device=dev, dtype=torch.float32)
            loss_out = torch.empty(2, device=dev, dtype=torch.float32)
            ops.ce_fwd(logits_buf, ldv, V, labels, B, S, row_loss, row_lse, loss_out)
            n0 += 2
            loss = loss_out[0]
            if need_grad:
                saved.update(x_final=x, rstd_f=rstd_f, hn=hn, logits_buf=logits_buf, ldv=ldv, labels=labels,
                             row_lse=row_lse, loss_out=loss_out)
        self._saved = saved if need_grad else None
        self.launches = n0
        return loss, logits

    # ------------------------------------------------------------------------------------------ generation (KV cache)
    # Prefill + one-token decode steps against a KV cache: the default of generate() (LLARK_KV_CACHE=0 switches it off);
    # verified on a B200 against the cache-less loop (tests/test_model_gpu.py::test_generate_with_kv_cache_matches_cacheless).
    def new_cache(self, B, s_max, device):
        cfg = self.cfg
        return KVCache(cfg.num_hidden_layers, B, s_max, cfg.hidden_size, device)

    @torch.no_grad()
    def prefill(self, input_ids, audio_encodings, cache):
        """Full forward over the prompt that also fills `cache`; returns logits [B, S, V]."""
        self.ensure_packed()
        B, S = input_ids.shape
        if S > cache.s_max or B != cache.B:
            raise ValueError("prompt does not fit the KV cache")
        batch = dict(input_ids=input_ids, attention_mask=None, labels=None, audio_encodings=audio_encodings)
        _, logits = self._forward(batch, need_grad=False, kv_sink=cache)
        cache.length = S
        cache.lens.fill_(S)
        return logits

    @torch.no_grad()
    def decode_step(self, last_ids, cache):
        """One token per row: last_ids int64 [B, 1] at position cache.length -> logits [B, 1, V]; appends to the cache."""
        cfg, flat = self.cfg, self.ensure_packed()
        dev = last_ids.device
        B = last_ids.shape[0]
        H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        NH = cfg.num_attention_heads
        D = H // NH
        V = self.lm.lm_head.weight.shape[0]
        eps = cfg.rms_norm_eps
        t = cache.length
        if t >= cache.s_max:
            raise ValueError("KV cache is full")
        ids = last_ids.to(torch.int64).contiguous()
        table = flat.pview("model.embed_tokens.weight", (V, H))
        x = torch.empty((B, H), device=dev, dtype=BF16)
        ops.embed_splice_fwd(ids, table, None, None, None, None, x.view(B, 1, H))
        cos, sin = self._tables(cache.s_max, dev)
        cos_t, sin_t = cos[t:], sin[t:]  # row 0 of these views is position t: with S = 1 every row reads it
        cache.lens.fill_(t + 1)
        rstd = torch.empty(B, device=dev, dtype=torch.float32)
        # up to 8 rows the linears are weight-streaming GEMVs (csrc/gemv.cu); LLARK_DECODE_GEMV=0 keeps the GEMM core
        skinny = B <= 8 and os.environ.get("LLARK_DECODE_GEMV", "1") != "0"
        lin = ops.gemv if skinny else ops.gemm
        for l in range(L):
            pfx = f"model.layers.{l}."
            Wqkv = flat.pview(pfx + "self_attn.q_proj.weight", (3 * H, H), span=3)
            Wo = flat.pview(pfx + "self_attn.o_proj.weight", (H, H))
            Wgu = flat.pview(pfx + "mlp.gate_proj.weight", (2 * I, H), span=2)
            Wd = flat.pview(pfx + "mlp.down_proj.weight", (H, I))
            xn1 = ops.rmsnorm_fwd(x, flat.pview(pfx + "input_layernorm.weight"), eps, rstd=rstd)
            if self.fuse_epilogues and D == 128 and not skinny:
                qkv = ops.gemm(xn1, Wqkv, rope=(cos_t, sin_t, 1, 2 * H))
            else:
                qkv = lin(xn1, Wqkv)
                ops.rope_(qkv, cos_t, sin_t, 1, NH, D)
            cache.k[l][:, t].copy_(qkv[:, H:2 * H])
            cache.v[l][:, t].copy_(qkv[:, 2 * H:])
            attn = ops.attn_decode(qkv[:, :H].contiguous(), cache.k[l], cache.v[l], cache.lens, NH, D)
            x_mid = lin(attn, Wo, residual=x)
            xn2 = ops.rmsnorm_fwd(x_mid, flat.pview(pfx + "post_attention_layernorm.weight"), eps, rstd=rstd)
            gu = lin(xn2, Wgu)
            act = ops.swiglu_fwd(gu)
            x = lin(act, Wd, residual=x_mid)
        hn = ops.rmsnorm_fwd(x, flat.pview("model.norm.weight"), eps, rstd=rstd)
        ldv = (V + 63) // 64 * 64
        logits_buf = torch.empty((B, ldv), device=dev, dtype=BF16)
        lin(hn, flat.pview("lm_head.weight", (V, H)), out=logits_buf, N=V)
        cache.length = t + 1
        return logits_buf.view(B, 1, ldv)[:, :, :V]

    # ------------------------------------------------------------------------------------------ backward
    def _backward(self, dloss):
        cfg, flat, sv = self.cfg, self.flat, self._saved
        if sv is None:
            raise RuntimeError("backward called without a saved forward (double backward is not supported)")
        self._saved = None
        B, S = sv["B"], sv["S"]
        H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
        NH = cfg.num_attention_heads
        D = H // NH
        V = self.lm.lm_head.weight.shape[0]
        M = B * S
        dev = sv["ids"].device
        n0 = self.launches
        acc = flat.grads_live()  # accumulate into existing grads (micro-batches 2..k) or overwrite (first)
        ws_norm = self._workspace("norm", ops.lib().llark_rmsnorm_bwd_workspace(M, H), dev)
        ws_attn = self._workspace("attn", ops.lib().llark_attn_bwd_workspace(B, S, NH), dev)

        def wgrad(name, dy, xin, shape, span=1):
            """dW[N,K] (+)= dy[M,N]^T @ xin[M,K] written straight into the flat gradient bucket."""
            g = flat.gview(name, shape, span=span)
            if g is None:
                return 0
            ops.gemm(dy, xin, a_mn=True, b_mn=True, out=g, residual=g if acc else None)
            return 1

        # loss -> dlogits -> d(final hidden)
        ldv = sv["ldv"]
        dlogits = torch.empty_like(sv["logits_buf"])
        dl = dloss.to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        ops.ce_bwd(sv["logits_buf"], ldv, V, sv["labels"], B, S, sv["row_lse"], sv["loss_out"], dl, 1.0, dlogits)
        Wlm = flat.pview("lm_head.weight", (V, H))
        dhn = ops.gemm(dlogits, Wlm, b_mn=True, M=M, N=H, K=V, lda=ldv)
        n0 += 2
        g_lm = flat.gview("lm_head.weight", (V, H))
        if g_lm is not None:  # only when lm_head is trainable (the reference freezes it, llamav2.py:412-415)
            ops.gemm(dlogits, sv["hn"], a_mn=True, b_mn=True, M=V, N=H, K=M, lda=ldv, out=g_lm,
                     residual=g_lm if acc else None)
            n0 += 1
        del dlogits
        dx = torch.empty((M, H), device=dev, dtype=BF16)
        g_norm = flat.gview("model.norm.weight")
        ops.rmsnorm_bwd(sv["x_final"], flat.pview("model.norm.weight"), dhn, sv["rstd_f"], None, dx, g_norm, acc,
                        ws_norm)
        n0 += 2
        cos, sin = self._tables(S, dev)

        for l in reversed(range(L)):
            nv = _Nvtx(f"llark.layer_bwd[{l}]").__enter__()
            pfx = f"model.layers.{l}."
            kept = sv["layers"][l]
            sv["layers"][l] = None
            if len(kept) == 1:  # activation recompute: same kernels on the same input -> the same bits as the forward
                _, kept, n = self._layer_forward(l, kept[0], sv["mask"], B, S, cos, sin)
                n0 += n
            x, rstd1, xn1, qkv, attn, lse, x_mid, rstd2, xn2, gu, act = kept
            del kept
            Wqkv = flat.pview(pfx + "self_attn.q_proj.weight", (3 * H, H), span=3)
            Wo = flat.pview(pfx + "self_attn.o_proj.weight", (H, H))
            Wgu = flat.pview(pfx + "mlp.gate_proj.weight", (2 * I, H), span=2)
            Wd = flat.pview(pfx + "mlp.down_proj.weight", (H, I))
            # MLP
            if self.fuse_epilogues:
                # d(act)[M,I] = dx[M,H] @ Wd[H,I] never reaches HBM: the dgrad's epilogue applies the SwiGLU backward
                # and writes dgu = (dgate | dup) directly (bit-identical to gemm -> llark_swiglu_bwd)
                dgu = torch.empty_like(gu)
                ops.gemm(dx, Wd, b_mn=True, out=dgu, N=I, swiglu_bwd=gu)
                n0 += 1 + wgrad(pfx + "mlp.down_proj.weight", dx, act, (H, I))
            else:
                dact = ops.gemm(dx, Wd, b_mn=True)  # [M,I] = dx[M,H] @ Wd[H,I]
                n0 += 2 + wgrad(pfx + "mlp.down_proj.weight", dx, act, (H, I))
                dgu = ops.swiglu_bwd(gu, dact)
                del dact
            dxn2 = ops.gemm(dgu, Wgu, b_mn=True)  # [M,H] = dgu[M,2I] @ Wgu[2I,H]
            n0 += 1 + wgrad(pfx + "mlp.gate_proj.weight", dgu, xn2, (2 * I, H), span=2)
            del dgu
            dx_mid = torch.empty((M, H), device=dev, dtype=BF16)
            ops.rmsnorm_bwd(x_mid, flat.pview(pfx + "post_attention_layernorm.weight"), dxn2, rstd2, dx, dx_mid,
                            flat.gview(pfx + "post_attention_layernorm.weight"), acc, ws_norm)
            # attention
            dattn = ops.gemm(dx_mid, Wo, b_mn=True)
            n0 += 3 + wgrad(pfx + "self_attn.o_proj.weight", dx_mid, attn, (H, H))
            if self.fuse_epilogues and D == 128:
                # inverse RoPE of dq | dk in the attention backward's epilogues (bit-identical to attn_bwd -> llark_rope)
                dqkv = ops.attn_bwd(qkv, sv["mask"], attn, dattn, lse, B, S, NH, D, workspace=ws_attn, rope=(cos, sin))
            else:
                dqkv = ops.attn_bwd(qkv, sv["mask"], attn, dattn, lse, B, S, NH, D, workspace=ws_attn)
                ops.rope_(dqkv, cos, sin, S, NH, D, inverse=True)
                n0 += 1
            dxn1 = ops.gemm(dqkv, Wqkv, b_mn=True)
            n0 += 4 + wgrad(pfx + "self_attn.q_proj.weight", dqkv, xn1, (3 * H, H), span=3)
            ops.rmsnorm_bwd(x, flat.pview(pfx + "input_layernorm.weight"), dxn1, rstd1, dx_mid, dx,
                            flat.gview(pfx + "input_layernorm.weight"), acc, ws_norm)
            n0 += 2
            if self.grad_ready_hook is not None and l in self._comm_slices(L):
                # layers [l, hi) are final: their gradients form one contiguous slice of the bucket
                lo_off = flat.offsets[f"model.layers.{l}.self_attn.q_proj.weight"][0]
                hi_l = self._comm_slices(L)[l]
                hi_off = (flat.offsets[f"model.layers.{hi_l}.self_attn.q_proj.weight"][0] if hi_l < L
                          else flat.offsets["model.norm.weight"][0])
                if lo_off < flat.n_train and hi_off <= flat.n_train and hi_off > lo_off:
                    self.grad_ready_hook(lo_off, hi_off)
            nv.__exit__()

        # splice + projector
        g_table = flat.gview("model.embed_tokens.weight", (V, H))
        detach_text = getattr(self.lm.model, "orig_embeds_params", None) is not None
        if not acc:  # overwrite mode: buckets that are scatter-accumulated or may see no work start from zero
            if g_table is not None:
                g_table.zero_()
            if sv["n_feat"] == 0:
                for nm in ("model.mm_projector.weight", "model.mm_projector.bias"):
                    if nm in flat.offsets and flat.gview(nm) is not None:
                        flat.gview(nm).zero_()
        n_feat = sv["n_feat"]
        daudio = torch.empty((max(n_feat, 1), H), device=dev, dtype=BF16) if n_feat else None
        ops.embed_splice_bwd(sv["ids"], dx.view(B, S, H), daudio, sv["audio_off"], sv["frames"], sv["start_pos"],
                             g_table, V, detach_text)
        n0 += 1
        if n_feat:
            A = sv["enc_bf"].shape[1]
            n0 += wgrad("model.mm_projector.weight", daudio, sv["enc_bf"], (H, A))
            g_b = flat.gview("model.mm_projector.bias")
            if g_b is not None:
                ops.colsum(daudio, g_b, accumulate=acc)
                n0 += 1
        flat.attach_grads()
        self.launches = n0

    def _comm_slices(self, L):
        """{first layer of a slice: one past its last layer}. Slices are `comm_chunk_layers` layers wide, except that the
        slice holding layers [0, chunk) -- final only when the backward is almost over, so its all-reduce cannot hide
        behind any GEMM -- is split into single layers: the exposed tail is then one layer (0.4 GB at 7B) plus the
        embedding / projector remainder instead of `chunk` layers (1.6 GB). LLARK_COMM_TAIL_SPLIT=0 restores equal slices."""
        key = (L, self.comm_chunk_layers)
        if getattr(self, "_slices_key", None) != key:
            c = self.comm_chunk_layers
            starts = set(range(0, L, c))
            if os.environ.get("LLARK_COMM_TAIL_SPLIT", "1") != "0":
                starts |= set(range(0, min(c, L)))
            starts = sorted(starts)
            self._slices = {a: (starts[i + 1] if i + 1 < len(starts) else L) for i, a in enumerate(starts)}
            self._slices_key = key
        return self._slices

    # ------------------------------------------------------------------------------------------ validation
    def check_flags(self):
        """Raise the reference's ValueErrors (llamav2.py:146-152, :169-175) from the device flag word (one sync)."""
        if self.flags is None:
            return
        f = int(self.flags.item())
        if f:
            self.flags.zero_()
        if f & 1:
            raise ValueError("The number of image start tokens and image end tokens should be the same.")
        if f & 2:
            raise ValueError("The image end token should follow the image start token.")
        if f & 4:
            raise ValueError("llark_b200 supports one audio segment per sample (the reference keeps only the last).")
