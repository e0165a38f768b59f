"""Helpers shared by the GPU tests and __graft_entry__.smoke(): build the model mirror from a name->tensor
parameter dict (the oracle's `init_params` layout, which is the reference's state_dict layout)."""
import torch

BF16 = torch.bfloat16


def build_model(cfg, params, detach_text=True, freeze_lm_head=True, dtype=BF16):
    from llark_b200.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    hf = WrappedLlamav2Config(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                              num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                              num_key_value_heads=cfg.num_heads, vocab_size=cfg.vocab_size,
                              max_position_embeddings=2048, rms_norm_eps=cfg.rms_norm_eps)
    hf.mm_hidden_size = cfg.mm_hidden_size
    model = WrappedLlamav2ForCausalLM(hf)
    model.get_model().initialize_adapter_modules()
    missing, unexpected = model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = (
        cfg.audio_start_token, cfg.audio_end_token, cfg.audio_patch_token)
    model.to(device="cuda", dtype=dtype)
    if freeze_lm_head:
        for p in model.get_output_embeddings().parameters():
            p.requires_grad = False
    if detach_text:
        model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    return model


def to_cuda(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}




# ---- measured parity errors: every tolerance check of the GPU parity tests goes through check_close, which records the
# measured value next to its bound; tests/conftest.py writes the table to gpurun_out/parity_measured.txt at session end
PARITY_LOG = []


def check_close(name, measured, tol):
    measured = float(measured)
    PARITY_LOG.append((name, measured, float(tol)))
    assert measured < tol, f"{name}: measured {measured:.4e} >= bound {tol:.1e}"
