"""Host-logic tests of llark_b200.engine on CPU: the engine's forward schedule, the cache-less generate loop and the
KV-cache path (prefill + decode steps) are run with tests/cpu_ops.py standing in for the CUDA
ops (test infrastructure, see its header) and checked against the oracle. What this pins: flat-bucket views, splice
bookkeeping, launch order, KV-cache indexing, the S=1 / offset-table RoPE of the decode step. What it cannot pin: the
CUDA kernels (the `-m gpu` suite does)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
BF16 = torch.bfloat16


@pytest.fixture()
def cpu_engine(monkeypatch):
    import cpu_ops
    import llark_b200.engine as eng

    monkeypatch.setattr(eng, "ops", cpu_ops)

    def ensure_packed(self):  # the product refuses non-CUDA devices (no CPU path); the stand-in run lifts that check
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device)
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)
    return eng


def _build(cfg, params):
    from llark_b200.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    hf = WrappedLlamav2Config(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                              num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                              num_key_value_heads=cfg.num_heads, vocab_size=cfg.vocab_size,
                              max_position_embeddings=256, rms_norm_eps=cfg.rms_norm_eps)
    hf.mm_hidden_size = cfg.mm_hidden_size
    model = WrappedLlamav2ForCausalLM(hf)
    model.get_model().initialize_adapter_modules()
    model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = (
        cfg.audio_start_token, cfg.audio_end_token, cfg.audio_patch_token)
    return model.to(BF16)


def test_engine_forward_schedule_matches_oracle(cpu_engine):
    from oracle import llama_oracle as lo

    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    model = _build(cfg, params)
    with torch.no_grad():
        out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                    audio_encodings=batch["audio_encodings"])
    pb = {k: v.to(BF16) for k, v in params.items()}
    _, ref = lo.forward(pb, cfg, batch["input_ids"], batch["attention_mask"], None, batch["audio_encodings"].to(BF16))
    valid = batch["attention_mask"]
    a, b = out.logits.float()[valid], ref.float()[valid]
    assert ((a - b).norm() / b.norm()).item() < 2e-2
    # the ValueError of a missing end token travels through the flag word exactly as on the GPU
    bad = batch["input_ids"].clone()
    bad[0, 4 + 5 + 1] = 7
    with pytest.raises(ValueError):
        model(input_ids=bad, attention_mask=batch["attention_mask"], audio_encodings=batch["audio_encodings"])


def test_generate_cacheless_and_kv_cache_agree(cpu_engine, monkeypatch):
    from oracle import llama_oracle as lo
    from oracle.make_golden_generation import gen_inputs

    cfg, params, ids, enc = gen_inputs()
    model = _build(cfg, params)
    monkeypatch.setenv("LLARK_KV_CACHE", "0")
    ref = model.generate(input_ids=ids, audio_encodings=enc, max_new_tokens=6)
    assert ref.shape == (1, ids.shape[1] + 6)
    with torch.no_grad():
        full = model(input_ids=ref[:, :-1], audio_encodings=enc).logits.float()
    monkeypatch.setenv("LLARK_KV_CACHE", "1")
    out = model(input_ids=ids, audio_encodings=enc, use_cache=True)
    cache = out.past_key_values
    P = ids.shape[1]
    assert len(cache) == P and cache.k[0].shape == (1, 256, cfg.hidden_size)
    assert torch.allclose(out.logits.float(), full[:, :P], atol=1e-6)            # prefill == plain forward
    for t in range(P, P + 5):                                                     # teacher-forced decode steps
        step = model(input_ids=ref[:, : t + 1], past_key_values=cache, use_cache=True)
        assert len(cache) == t + 1 and step.logits.shape == (1, 1, cfg.vocab_size)
        a, b = step.logits.float()[0, 0], full[0, t]
        assert ((a - b).norm() / b.norm()).item() < 2e-2, t                      # same math, different reduction order
        assert a.argmax().item() == b.argmax().item() or (b.topk(2).values.diff().abs().item() < 0.05)
    got = model.generate(input_ids=ids, audio_encodings=enc, max_new_tokens=6)
    agree = (got == ref)[0, P:].float().mean().item()
    assert agree >= 0.8, (got, ref)   # identical up to near-ties of the random-weight fixture
    with pytest.raises(ValueError):
        model._engine.prefill(torch.zeros((1, 300), dtype=torch.int64), None, cache)


def test_engine_backward_and_optimizer_host_logic(cpu_engine, monkeypatch):
    """Manual backward into the flat gradient bucket (views, overwrite-after-step vs accumulate, frozen lm_head, detached
    text rows) and the fused optimizer's bookkeeping, against the oracle's autograd gradients."""
    import cpu_ops
    import llark_b200.optim as optim
    from oracle import llama_oracle as lo

    monkeypatch.setattr(optim, "ops", cpu_ops)
    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    model = _build(cfg, params)
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    opt = optim.LlarkAdamW(model, lr=1e-3)
    out = model(**batch)
    out.loss.backward()
    pb = {k: v.to(BF16) for k, v in params.items()}
    loss_o, _, grads_o = lo.forward_backward(pb, cfg, batch)
    assert abs(out.loss.item() - loss_o.item()) < 3e-3 * loss_o.item()
    named = dict(model.named_parameters())
    flat = model.engine.flat
    for name in ("model.mm_projector.weight", "model.layers.0.self_attn.q_proj.weight", "model.layers.1.mlp.down_proj.weight",
                 "model.norm.weight", "model.layers.0.input_layernorm.weight"):
        g = named[name].grad
        assert g.data_ptr() == flat.gview(name).data_ptr()           # .grad IS the bucket view
        ref = grads_o[name].float()
        assert ((g.float() - ref).norm() / ref.norm()).item() < 6e-2, name
    assert named["lm_head.weight"].grad is None                        # frozen (llamav2.py:412-415)
    ge = named["model.embed_tokens.weight"].grad.float()
    rows = ge.abs().sum(1).nonzero().flatten().tolist()
    assert sorted(rows) == sorted({cfg.audio_start_token, cfg.audio_end_token})  # only the marker rows (Appendix D.3)
    # accumulation: a second backward before step() adds into the bucket; after step() the next backward overwrites
    g1 = flat.grads.float().clone()
    model(**batch).loss.backward()
    assert torch.allclose(flat.grads.float(), 2 * g1, rtol=2e-2, atol=1e-3)
    opt.step()
    assert flat.consumed
    model(**batch).loss.backward()
    assert not torch.allclose(flat.grads.float(), 3 * g1, rtol=1e-2, atol=1e-3)
    l0 = out.loss.item()
    opt.step()
    with torch.no_grad():
        l1 = model(**batch).loss.item()
    assert l1 < l0


def test_activation_recompute_gives_the_same_gradients(cpu_engine):
    """--gradient_checkpointing True (m2t/models/llamav2.py:241,255-257): with recompute the forward keeps one tensor per
    layer and the backward regenerates the rest with the same kernels -- same loss, same gradient bits, more launches."""
    from oracle import llama_oracle as lo

    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=3)
    results = []
    for mode in ("0", "1", "auto"):
        model = _build(cfg, params)
        model.engine.recompute = mode
        if mode == "auto":
            model.gradient_checkpointing_enable()   # flag set, but nothing to save on this tiny CPU run -> no recompute
        out = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
        kept = [len(t) for t in model.engine._saved["layers"]]
        out.loss.backward()
        results.append((out.loss.item(), model.engine.flat.grads.clone(), model.engine.launches, kept,
                        model.engine.recomputed_last_step))
    (l0, g0, n0, k0, r0), (l1, g1, n1, k1, r1), (l2, g2, n2, k2, r2) = results
    assert (r0, r1, r2) == (False, True, False)
    assert set(k0) == {11} and set(k1) == {1} and k2 == k0
    assert l0 == l1 == l2
    assert torch.equal(g0, g1) and torch.equal(g0, g2)
    assert n1 == n0 + 8 * cfg.num_layers and n2 == n0


def test_trainer_checkpoint_resume_is_exact(cpu_engine, monkeypatch, tmp_path):
    """checkpoint-N (weights + optimizer.pt + trainer_state.json) -> a NEW trainer with resume_from_checkpoint=True must
    land on exactly the parameters of an uninterrupted run (same data order, cosine schedule by global_step, AdamW
    moments and step count restored)."""
    from types import SimpleNamespace

    import cpu_ops
    import llark_b200.optim as optim
    from llark_b200.trainer import WrappedTrainer
    from oracle import llama_oracle as lo

    monkeypatch.setattr(optim, "ops", cpu_ops)
    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    data = [lo.make_batch(cfg, B=1, S=32, F_frames=3, seed=s, pad_last=0) for s in range(4)]
    samples = [{"input_ids": b["input_ids"][0], "labels": b["labels"][0], "audio_encoding": b["audio_encodings"][0]}
               for b in data]

    class Stream(torch.utils.data.IterableDataset):
        def __init__(self, skip=0):
            self.skip = skip

        def __iter__(self):
            i = self.skip
            while True:
                yield samples[i % len(samples)]
                i += 1

    def collate(items):
        return {"input_ids": torch.stack([i["input_ids"] for i in items]),
                "labels": torch.stack([i["labels"] for i in items]),
                "attention_mask": torch.ones(len(items), 32, dtype=torch.bool),
                "audio_encodings": torch.stack([i["audio_encoding"] for i in items])}

    def make(out, max_steps, skip=0, save_steps=0):
        model = _build(cfg, params)
        for p in model.get_output_embeddings().parameters():
            p.requires_grad = False
        args = SimpleNamespace(max_steps=max_steps, per_device_train_batch_size=1, gradient_accumulation_steps=1,
                               learning_rate=2e-3, warmup_ratio=0.25, max_grad_norm=1.0, weight_decay=0.0,
                               logging_steps=1, save_steps=save_steps, output_dir=str(out), dataloader_num_workers=0,
                               tune_mm_mlp_adapter=True)
        return WrappedTrainer(model=model, tokenizer=None, args=args, train_dataset=Stream(skip), data_collator=collate)

    straight = make(tmp_path / "a", 4)
    straight.train()
    ref = straight.model.engine.flat.params.clone()

    # the interrupted run: same 4-step schedule, killed right after checkpoint-2 has been written
    first = make(tmp_path / "b", 4, save_steps=2)

    class Stop(Exception):
        pass

    real_save = first._save_optimizer

    def save_then_stop(ckpt):
        real_save(ckpt)
        raise Stop()

    first._save_optimizer = save_then_stop
    with pytest.raises(Stop):
        first.train()
    assert {"optimizer.pt", "pytorch_model.bin", "trainer_state.json"} <= set(os.listdir(tmp_path / "b" / "checkpoint-2"))
    assert os.path.exists(tmp_path / "b" / "mm_projector" / "checkpoint-2.bin")

    resumed = make(tmp_path / "b", 4, skip=2)
    resumed.train(resume_from_checkpoint=True)
    assert resumed.state["global_step"] == 4 and len(resumed.state["log_history"]) == 4
    assert resumed.optimizer.step_count == 4
    assert torch.equal(resumed.model.engine.flat.params, ref)


def test_backward_announces_final_layer_gradients_in_reverse_order(cpu_engine):
    """The overlap hook of the data-parallel step: slices handed to `grad_ready_hook` are contiguous, arrive in reverse
    layer order, cover exactly the decoder layers, and do not change after they were announced (also under accumulation,
    where the second backward adds into the bucket before announcing)."""
    from oracle import llama_oracle as lo

    cfg = lo.LlamaCfg(num_layers=5)
    model = _build(cfg, lo.init_params(cfg, seed=0))
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    eng = model.engine
    eng.comm_chunk_layers = 2
    calls = []
    eng.grad_ready_hook = lambda lo_, hi_: calls.append((lo_, hi_, eng.flat.grads[lo_:hi_].clone()))
    for rounds in (1, 2):                       # second pass accumulates into the live bucket
        calls.clear()
        model(**batch).loss.backward()
        flat = eng.flat
        spans = [(a, b) for a, b, _ in calls]
        assert spans == sorted(spans, reverse=True) and len(spans) == 4   # [4,5) [2,4) + the tail split [1,2) [0,1)
        assert spans[-1][0] == flat.offsets["model.layers.0.self_attn.q_proj.weight"][0]
        assert spans[0][1] == flat.offsets["model.norm.weight"][0]
        assert all(spans[i][0] == spans[i + 1][1] for i in range(len(spans) - 1))
        for a, b, snap in calls:
            assert torch.equal(flat.grads[a:b], snap)


def test_frozen_backbone_projector_only_training(cpu_engine, monkeypatch):
    """`--freeze_backbone True --tune_mm_mlp_adapter True` (m2t/train.py:79-80,143-152 + llamav2.py:408-415): only the
    projector and the embedding table are trainable. The backward still runs dgrad through every layer to reach the
    projector, writes no weight gradient for frozen tensors, and the fused optimizer steps a bucket that holds just the
    two trainable tensors."""
    import cpu_ops
    import llark_b200.optim as optim
    from oracle import llama_oracle as lo

    monkeypatch.setattr(optim, "ops", cpu_ops)
    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=1)
    model = _build(cfg, params)
    model.requires_grad_(False)
    for p in model.get_model().mm_projector.parameters():
        p.requires_grad = True
    for p in model.get_input_embeddings().parameters():
        p.requires_grad = True
    model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    opt = optim.LlarkAdamW(model, lr=1e-2)
    flat = model.engine.flat
    trainable = {"model.embed_tokens.weight", "model.mm_projector.weight", "model.mm_projector.bias"}
    assert flat.n_train <= sum(p.numel() for n, p in model.named_parameters() if n in trainable) + 3 * 128
    hooked = []
    model.engine.grad_ready_hook = lambda a, b: hooked.append((a, b))
    out = model(**batch)
    out.loss.backward()
    assert hooked == []                                           # no decoder layer owns a slice of the trainable bucket
    pb = {k: v.to(BF16) for k, v in params.items()}
    loss_o, _, grads_o = lo.forward_backward(pb, cfg, batch)
    assert abs(out.loss.item() - loss_o.item()) < 3e-3 * loss_o.item()
    named = dict(model.named_parameters())
    for n, p in named.items():
        if n in trainable:
            ref = grads_o[n].float()
            assert ((p.grad.float() - ref).norm() / (ref.norm() + 1e-30)).item() < 6e-2, n
        else:
            assert p.grad is None, n
    l0 = out.loss.item()
    before = named["model.layers.0.mlp.down_proj.weight"].detach().clone()
    for _ in range(3):
        opt.step()
        model(**batch).loss.backward()
    with torch.no_grad():
        assert model(**batch).loss.item() < l0
    assert torch.equal(named["model.layers.0.mlp.down_proj.weight"], before)          # frozen weights never move


def test_frozen_projector_backbone_training(cpu_engine, monkeypatch):
    """`--freeze_mm_mlp_adapter True` (m2t/train.py:154-158): the projector gets no gradient, everything else trains."""
    import cpu_ops
    import llark_b200.optim as optim
    from oracle import llama_oracle as lo

    monkeypatch.setattr(optim, "ops", cpu_ops)
    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=2)
    model = _build(cfg, params)
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    for p in model.get_model().mm_projector.parameters():
        p.requires_grad = False
    model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    opt = optim.LlarkAdamW(model, lr=1e-3)
    out = model(**batch)
    out.loss.backward()
    _, _, grads_o = lo.forward_backward({k: v.to(BF16) for k, v in params.items()}, cfg, batch)
    named = dict(model.named_parameters())
    assert named["model.mm_projector.weight"].grad is None and named["model.mm_projector.bias"].grad is None
    for n in ("model.layers.0.self_attn.q_proj.weight", "model.layers.1.mlp.gate_proj.weight", "model.norm.weight"):
        ref = grads_o[n].float()
        assert ((named[n].grad.float() - ref).norm() / ref.norm()).item() < 6e-2, n
    proj_before = named["model.mm_projector.weight"].detach().clone()
    opt.step()
    assert torch.equal(named["model.mm_projector.weight"], proj_before)


def test_stand_in_ops_track_the_real_signatures():
    """tests/cpu_ops.py is only a valid stand-in while every function it offers exists in llark_b200.ops with the same
    leading parameters (the real wrappers may add optional output buffers; gemm takes its options as keywords)."""
    import inspect

    import cpu_ops
    import llark_b200.ops as real

    for name, fn in vars(cpu_ops).items():
        if name.startswith("_") or not inspect.isfunction(fn) or fn.__module__ != "cpu_ops" or name == "lib":
            continue
        assert hasattr(real, name), f"stand-in op {name} does not exist in llark_b200.ops"
        want = list(inspect.signature(getattr(real, name)).parameters)
        have = [p for p in inspect.signature(fn).parameters if p != "kw"]
        assert have == want[: len(have)], (name, have, want)
    offered = {n for n, f in vars(cpu_ops).items() if inspect.isfunction(f) and f.__module__ == "cpu_ops"}
    import ast

    for path in ("llark_b200/engine.py", "llark_b200/engine_mpt.py", "llark_b200/optim.py", "llark_b200/jukebox/model.py"):
        src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), path)).read()
        used = {n.attr for n in ast.walk(ast.parse(src)) if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name)
                and n.value.id == "ops"}
        used -= {"GEMM_PROFILE"}
        assert used <= offered, (path, sorted(used - offered))

