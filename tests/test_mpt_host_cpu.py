"""Host-logic tests of the CLAP + MPT-1B variant: llark_b200.models.mpt + llark_b200.engine_mpt run
on CPU with tests/cpu_ops.py standing in for the CUDA ops (five of which do not exist as kernels yet) and are checked
against oracle/mpt_oracle.py, which tests/test_mpt_oracle_cpu.py pins to the reference's own WrappedMPTForCausalLM.
Pinned here: parameter / state-dict names, the manual backward of the MPT block (LayerNorms incl. the q / k ones, ALiBi
attention, exact GELU), the tied-table gradient (head wgrad + embedding scatter), bucket accumulation semantics."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
BF16 = torch.bfloat16


@pytest.fixture()
def cpu_mpt(monkeypatch):
    import cpu_ops
    import llark_b200.engine as eng
    import llark_b200.engine_mpt as engm
    import llark_b200.optim as optim

    for mod in (eng, engm, optim):
        monkeypatch.setattr(mod, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device, cast_to_bf16=not torch.is_grad_enabled())
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)
    return engm


def _build(cfg, params):
    from llark_b200.models.mpt import WrappedMPTConfig, WrappedMPTForCausalLM

    hf = WrappedMPTConfig(d_model=cfg.d_model, n_heads=cfg.n_heads, n_layers=cfg.n_layers,
                          expansion_ratio=cfg.expansion_ratio, max_seq_len=cfg.max_seq_len, vocab_size=cfg.vocab_size,
                          no_bias=True, mm_hidden_size=cfg.mm_hidden_size,
                          attn_config={"alibi": True, "qk_ln": True, "attn_impl": "torch",
                                       "alibi_bias_max": cfg.alibi_bias_max})
    model = WrappedMPTForCausalLM(hf)
    model.get_model().initialize_adapter_modules()
    missing, unexpected = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = (
        cfg.audio_start_token, cfg.audio_end_token, cfg.audio_patch_token)
    return model.to(BF16)


@pytest.mark.parametrize("F_frames,detach", [(1, False), (4, True)])
def test_mpt_engine_forward_backward_matches_oracle(cpu_mpt, F_frames, detach):
    from oracle import mpt_oracle as mo

    cfg = mo.MPTCfg(d_model=256, n_heads=2)  # head_dim 128 as MPT-1B (2048 / 16)
    params = mo.init_params(cfg, seed=0)
    batch = mo.make_batch(cfg, B=2, S=40, F_frames=F_frames, seed=F_frames)
    model = _build(cfg, params)
    assert sorted(model.state_dict().keys()) == sorted(params.keys())   # the reference's key names
    if detach:
        model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    out = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()})
    out.loss.backward()
    pb = {k: v.to(BF16) for k, v in params.items()}
    loss_o, logits_o, grads_o = mo.forward_backward(pb, cfg, batch, detach_text=detach, autocast_dtype=BF16)
    assert abs(out.loss.item() - loss_o.item()) < 3e-3 * loss_o.item()
    valid = batch["attention_mask"]
    a, b = out.logits.float()[valid], logits_o.float()[valid]
    assert ((a - b).norm() / b.norm()).item() < 2e-2
    named = dict(model.named_parameters())
    for name, g_ref in grads_o.items():
        g = named[name].grad
        assert g is not None, name
        rel = ((g.float() - g_ref.float()).norm() / (g_ref.float().norm() + 1e-30)).item()
        assert rel < 6e-2, (name, rel)
    flat = model.engine.flat
    assert named["transformer.wte.weight"].grad.data_ptr() == flat.gview("transformer.wte.weight").data_ptr()
    # second backward accumulates into the bucket; after an optimizer step the next one overwrites
    from llark_b200.optim import LlarkAdamW

    g1 = flat.grads.float().clone()
    model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}).loss.backward()
    assert torch.allclose(flat.grads.float(), 2 * g1, rtol=3e-2, atol=2e-3)
    opt = LlarkAdamW(model, lr=1e-3)
    opt.step()
    l0 = out.loss.item()
    with torch.no_grad():
        l1 = model(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}).loss.item()
    assert l1 < l0


def test_mpt_config_envelope_and_errors(cpu_mpt):
    from llark_b200.models.mpt import WrappedMPTConfig, WrappedMPTForCausalLM
    from oracle import mpt_oracle as mo

    cfg = mo.MPTCfg(d_model=256, n_heads=2)
    params = mo.init_params(cfg, seed=0)
    batch = mo.make_batch(cfg, B=2, S=40, F_frames=4, seed=0)
    model = _build(cfg, params)
    bad = batch["input_ids"].clone()
    bad[0, 3 + 4 + 1] = 7
    with pytest.raises(ValueError):
        model(input_ids=bad, attention_mask=batch["attention_mask"], audio_encodings=batch["audio_encodings"])
    with pytest.raises(NotImplementedError):
        model(input_ids=batch["input_ids"], use_cache=True)
    plain = WrappedMPTForCausalLM(WrappedMPTConfig(d_model=256, n_heads=2, n_layers=1, vocab_size=64, no_bias=True))
    with pytest.raises(NotImplementedError, match="alibi=True and qk_ln=True"):
        plain(input_ids=torch.zeros((1, 4), dtype=torch.int64))
    from llark_b200.engine_mpt import alibi_slopes

    assert torch.allclose(alibi_slopes(8), torch.tensor([2.0 ** -(i + 1) for i in range(8)]))
    assert torch.allclose(alibi_slopes(6), mo.gen_slopes(6))


def test_mpt_backward_announces_block_gradients_for_the_overlapped_allreduce(cpu_mpt):
    """Data-parallel overlap: during backward the engine hands contiguous, final slices of the flat gradient
    bucket to `grad_ready_hook` in reverse block order; together they cover exactly the blocks' parameters, and a slice's
    contents do not change afterwards (the optimizer all-reduces it on a side stream meanwhile)."""
    from oracle import mpt_oracle as mo

    cfg = mo.MPTCfg(d_model=256, n_heads=2, n_layers=5)
    model = _build(cfg, mo.init_params(cfg, seed=0))
    batch = mo.make_batch(cfg, B=2, S=40, F_frames=1, seed=0)
    eng = model.engine
    eng.comm_chunk_layers = 2
    calls = []

    def hook(lo, hi):
        calls.append((lo, hi, eng.flat.grads[lo:hi].clone()))

    eng.grad_ready_hook = hook
    model(**batch).loss.backward()
    flat = eng.flat
    spans = [(lo, hi) for lo, hi, _ in calls]
    assert spans == sorted(spans, reverse=True) and len(spans) == 3              # blocks {4}, {2,3}, {0,1}
    first = flat.offsets["transformer.blocks.0.norm_1.weight"][0]
    last = flat.offsets["transformer.norm_f.weight"][0]
    assert spans[-1][0] == first and spans[0][1] == last
    assert all(spans[i][0] == spans[i + 1][1] for i in range(len(spans) - 1))
    for lo, hi, snap in calls:
        assert torch.equal(flat.grads[lo:hi], snap)                              # final when announced
    wte = flat.offsets["transformer.wte.weight"][0]
    assert not (first <= wte < last)                                             # the tied table is left to step()

