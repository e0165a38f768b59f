"""ORACLE support (test infrastructure only): tests/golden/mpt_tiny_ref.pt = forward / backward of the REFERENCE's own
WrappedMPTForCausalLM (m2t/models/mpt.py + m2t/llava/model/mpt/*, imported in place) on the seeded tiny model and batch
of oracle/mpt_oracle.py, in fp32 and under bf16 autocast, with and without the detached-text splice.

transformers 5.5 note: PretrainedConfig no longer keeps two attributes MPTConfig relies on
(`use_cache`, `tie_word_embeddings`); they are set on the config OBJECT after construction (the reference's source is not
touched), everything else runs as is with attn_impl="torch"."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mpt_oracle as mo  # noqa: E402
from oracle import ref_shims  # noqa: E402


def build_reference_model(cfg, params):
    from m2t.llava.model.mpt.configuration_mpt import init_config_defaults
    from m2t.models.mpt import WrappedMPTConfig, WrappedMPTForCausalLM

    attn_config = {"attn_type": "multihead_attention", "attn_pdrop": 0.0, "attn_impl": "torch", "qk_ln": True,
                   "clip_qkv": None, "softmax_scale": None, "prefix_lm": False, "attn_uses_sequence_id": False,
                   "alibi": True, "alibi_bias_max": cfg.alibi_bias_max}
    want = dict(d_model=cfg.d_model, n_heads=cfg.n_heads, n_layers=cfg.n_layers, expansion_ratio=cfg.expansion_ratio,
                max_seq_len=cfg.max_seq_len, vocab_size=cfg.vocab_size, resid_pdrop=0.0, emb_pdrop=0.0,
                learned_pos_emb=True, attn_config=attn_config, init_device="cpu", logit_scale=None, no_bias=True,
                verbose=0, embedding_fraction=1.0, norm_type="low_precision_layernorm", use_cache=False,
                init_config=dict(init_config_defaults))
    hf = WrappedMPTConfig(**want)
    for k, v in want.items():
        if not hasattr(hf, k):
            setattr(hf, k, v)
    if not hasattr(hf, "tie_word_embeddings"):
        hf.tie_word_embeddings = True
    hf.mm_hidden_size = cfg.mm_hidden_size
    model = WrappedMPTForCausalLM(hf)
    model.get_model().initialize_adapter_modules()
    missing, unexpected = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = (
        cfg.audio_start_token, cfg.audio_end_token, cfg.audio_patch_token)
    return model


def run_reference(cfg, params, batch, bf16, detach_text):
    model = build_reference_model(cfg, params)
    if detach_text:
        model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    model.train()
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=bf16):
        out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"].clone(),
                    audio_encodings=batch["audio_encodings"])
    out.loss.backward()
    grads = {n: (p.grad.detach().float().clone() if p.grad is not None else None) for n, p in model.named_parameters()}
    return out.loss.detach().float(), out.logits.detach().float(), grads


def main():
    ref_shims.install()
    cfg = mo.MPTCfg()
    params = mo.init_params(cfg, seed=0)
    out = {"cfg": cfg.__dict__.copy(), "runs": {}}
    # four of the eight (frames, autocast, detached-text) combinations keep the fixture small while covering each value
    for F_frames, bf16, detach in ((1, False, False), (1, True, True), (4, False, True), (4, True, False)):
        batch = mo.make_batch(cfg, B=2, S=40, F_frames=F_frames, seed=F_frames)
        loss, logits, grads = run_reference(cfg, params, batch, bf16, detach)
        out["runs"][(F_frames, bf16, detach)] = {"loss": loss, "logits": logits, "grads": grads}
        print(F_frames, bf16, detach, float(loss))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "mpt_tiny_ref.pt"))


if __name__ == "__main__":
    main()
