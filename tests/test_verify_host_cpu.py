"""Host logic of the fp32 verification mode (llark_b200/verify.py, `LLARK_VERIFY=fp32`) on CPU: the schedule -- 3xTF32
composition of every linear from split halves, fused QKV / gate|up weights, splice, RoPE tables, shifted CE -- run over
tests/cpu_ops.py's fp32 stand-ins must reproduce the REFERENCE's own fp32 logits and loss (tests/golden/llama_tiny_ref.pt).
The kernels themselves are checked on the GPU (tests/test_model_gpu.py::test_fp32_verification_mode_meets_the_1e3_logit_criterion)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
def test_fp32_forward_schedule_reproduces_the_reference_fp32_run(monkeypatch, mode):
    import cpu_ops
    import llark_b200.verify as verify
    from llark_b200.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM
    from oracle import llama_oracle as lo

    gold = torch.load(os.path.join(GOLD, "llama_tiny_ref.pt"), weights_only=False)
    cfg = lo.LlamaCfg(**gold["cfg"])
    params = lo.init_params(cfg, seed=gold["seed"])
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    hf = WrappedLlamav2Config(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                              num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                              num_key_value_heads=cfg.num_heads, vocab_size=cfg.vocab_size, max_position_embeddings=2048,
                              rms_norm_eps=cfg.rms_norm_eps)
    hf.mm_hidden_size = cfg.mm_hidden_size
    model = WrappedLlamav2ForCausalLM(hf)
    model.get_model().initialize_adapter_modules()
    model.load_state_dict({k: v.clone() for k, v in params.items()}, strict=False)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = (
        cfg.audio_start_token, cfg.audio_end_token, cfg.audio_patch_token)

    monkeypatch.setattr(verify, "ops", cpu_ops)
    monkeypatch.setenv("LLARK_VERIFY_GEMM", mode)
    # forward_fp32 refuses non-CUDA parameters (there is no CPU product path): let this test's CPU tensors through
    monkeypatch.setattr(verify, "_require_cuda", lambda dev: None)
    loss, logits = verify.forward_fp32(model, batch["input_ids"], batch["attention_mask"], batch["labels"],
                                       batch["audio_encodings"])
    ref = gold["fp32_detach"]
    valid = batch["attention_mask"]
    err = ((logits[valid].double() - ref["logits"][valid].double()).norm() / ref["logits"][valid].double().norm()).item()
    assert err < 1e-5, err          # fp32 CPU math of the same schedule: rounding-level agreement
    assert abs(loss.item() - ref["loss"].item()) < 1e-5 * ref["loss"].item()
