"""ORACLE support (test infrastructure only): generate tests/golden/*.pt by running the REFERENCE's own classes
(imported in place from /root/reference) on seeded synthetic inputs.  Run in the build container:

    python oracle/make_golden.py

The fixtures pin oracle/llama_oracle.py (and, through it, the CUDA path) to the reference:
  llama_tiny_ref.pt   WrappedLlamav2ForCausalLM (m2t/models/llamav2.py) forward/backward, fp32 and bf16, both
                      splice paths (orig_embeds_params set / unset), frozen lm_head (SURVEY Appendix D 1,3)
  pool_ref.pt         jukebox/main.py windowed_average (:113-130) on a seeded [T,4800] array, frame 34
  data_ref.pt         m2t/data_modules.py tokenise/mask/collate index work on a synthetic tokenizer
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llama_oracle as lo  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_reference_model(cfg, params, dtype):
    from m2t.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    hf = WrappedLlamav2Config(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                              num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                              num_key_value_heads=cfg.num_heads, vocab_size=cfg.vocab_size,
                              max_position_embeddings=2048, rms_norm_eps=cfg.rms_norm_eps,
                              attn_implementation="eager")
    hf.mm_hidden_size = cfg.mm_hidden_size
    model = WrappedLlamav2ForCausalLM(hf)
    model.get_model().initialize_adapter_modules()
    missing, unexpected = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = (
        cfg.audio_start_token, cfg.audio_end_token, cfg.audio_patch_token)
    model.to(dtype)
    # trainable set of m2t/train.py + initialize_audio_tokenizer (llamav2.py:408-415): lm_head frozen
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    return model


def run_reference(cfg, params, batch, dtype, detach_text):
    model = build_reference_model(cfg, params, dtype)
    if detach_text:
        model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    model.train()
    # the reference trains under HF Trainer's bf16 autocast (scripts/training/train_llark.sh:34) on bf16 weights
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
        out = model(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"], labels=batch["labels"],
                    audio_encodings=batch["audio_encodings"])
    out.loss.backward()
    grads = {n: (p.grad.detach().float().clone() if p.grad is not None else None) for n, p in model.named_parameters()}
    return out.loss.detach().float(), out.logits.detach().float(), grads


def golden_llama():
    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    res = {"cfg": cfg.__dict__, "seed": 0}
    for dtype, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        for detach in (True, False):
            loss, logits, grads = run_reference(cfg, params, batch, dtype, detach)
            key = f"{tag}_{'detach' if detach else 'nodetach'}"
            small = {}
            for n, g in grads.items():
                if g is None:
                    small[n] = None
                    continue
                entry = {"norm": g.norm().item()}
                if g.numel() <= 5000:
                    entry["full"] = g.to(torch.bfloat16) if dtype == torch.bfloat16 else g
                else:
                    entry["head"] = g.flatten()[:2048].clone()
                    if "embed_tokens" in n:
                        entry["nonzero_rows"] = g.abs().sum(-1).nonzero().flatten().tolist()
                small[n] = entry
            res[key] = {"loss": loss, "logits": logits.to(torch.bfloat16) if dtype == torch.bfloat16 else logits,
                        "grads": small}
            print(key, "loss", loss.item())
    # the reference's two ValueError conditions (llamav2.py:146-152, :169-175)
    model = build_reference_model(cfg, params, torch.float32)
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    bad["input_ids"][0, 4 + 5 + 1] = 7  # remove the end token of sample 0
    errs = []
    try:
        model(**bad)
    except ValueError as e:
        errs.append(str(e))
    res["value_errors"] = errs
    torch.save(res, os.path.join(GOLD, "llama_tiny_ref.pt"))


def golden_pool():
    mod = ref_shims.load_reference_jukebox_main()
    g = torch.Generator().manual_seed(7)
    T_valid = 700
    acts = torch.randn(T_valid, 4800, generator=g)
    frame_len = int(mod.ACTS_SAMPLE_RATE // 10)
    pooled = mod.windowed_average(acts, frame_len).squeeze(0)
    torch.save({"seed": 7, "T_valid": T_valid, "frame_len": frame_len, "acts_sample_rate": mod.ACTS_SAMPLE_RATE,
                "pooled_head": pooled[:, :64].clone(), "pooled_sum": pooled.double().sum().item(),
                "shape": tuple(pooled.shape),
                "full_len_frames": int(mod.T // frame_len)}, os.path.join(GOLD, "pool_ref.pt"))
    print("pool", pooled.shape, frame_len)


if __name__ == "__main__":
    ref_shims.install()
    os.makedirs(GOLD, exist_ok=True)
    golden_llama()
    golden_pool()
    try:
        from oracle.make_golden_data import golden_data

        golden_data()
    except ImportError:
        pass
