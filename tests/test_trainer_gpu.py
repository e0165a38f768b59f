"""WrappedTrainer end to end on the GPU with a tiny model: tar shards -> TarShardDataset -> collator -> fused
step (grad accumulation 2, cosine schedule, clip, AdamW) -> mm_projector side checkpoint (trainer.py:35-63)."""
import io
import json
import os
import pickle
import sys
import tarfile
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(__file__))


def _write_shards(tmp_path, n_shards=2, per_shard=4, frames=5, width=64):
    g = torch.Generator().manual_seed(0)
    paths = []
    for s in range(n_shards):
        p = tmp_path / f"train-{s:03d}.tar"
        with tarfile.open(p, "w") as tf:
            for i in range(per_shard):
                enc = torch.randn(frames, width, generator=g)
                js = {"response": [{"question": "what genre is this?", "answer": "this is jazz with drums."}]}
                for ext, payload in (("audio_encoding.pyd", pickle.dumps(enc)), ("json", json.dumps(js).encode())):
                    ti = tarfile.TarInfo(f"clip{s}_{i}.{ext}")
                    ti.size = len(payload)
                    tf.addfile(ti, io.BytesIO(payload))
        paths.append(str(p))
    return str(tmp_path / ("train-{000..%03d}.tar" % (n_shards - 1)))


def test_wrapped_trainer_reduces_loss_and_saves_projector(tmp_path):
    from synth_tokenizer import SynthTokenizer
    from oracle import llama_oracle as lo
    from llark_b200 import data_modules as dm
    from llark_b200.testing import build_model
    from llark_b200.trainer import WrappedTrainer

    tok = SynthTokenizer(vocab_size=296, model_max_length=128)
    tok.add_special_tokens({"pad_token": "[PAD]"})          # 296
    tok.add_tokens(["<audio_patch>"], special_tokens=True)   # 297
    tok.add_tokens(["<audio_start>", "<audio_end>"], special_tokens=True)  # 298, 299
    cfg = lo.LlamaCfg(vocab_size=304, audio_patch_token=297, audio_start_token=298, audio_end_token=299, pad_token=296)
    model = build_model(cfg, lo.init_params(cfg, seed=0))
    url = _write_shards(tmp_path)
    data_args = SimpleNamespace(is_multimodal=True, sep_audio_conv_front=False, audio_folder=None,
                                mm_use_audio_start_end=True, train_data_path=url, eval_data_path=None,
                                apply_task_sample_probs=False, task_sample_probs=None, eval_data_subsample=None)
    module = dm.make_data_module(tok, data_args)
    out_dir = tmp_path / "out"
    args = SimpleNamespace(per_device_train_batch_size=2, gradient_accumulation_steps=2, learning_rate=2e-3,
                           weight_decay=0.0, max_grad_norm=1.0, warmup_ratio=0.1, max_steps=12, logging_steps=1,
                           save_steps=12, output_dir=str(out_dir), tune_mm_mlp_adapter=True, dataloader_num_workers=0)
    trainer = WrappedTrainer(model=model, tokenizer=None, args=args, **module)
    state = trainer.train()
    hist = state["log_history"]
    assert len(hist) == 12
    assert hist[-1]["loss"] < 0.7 * hist[0]["loss"], (hist[0]["loss"], hist[-1]["loss"])
    assert all(h["grad_norm"] > 0 for h in hist)
    assert abs(hist[0]["learning_rate"]) < 1e-12 and hist[2]["learning_rate"] == pytest.approx(2e-3, rel=1e-6)
    trainer.save_state()
    side = out_dir / "mm_projector" / "checkpoint-12.bin"
    assert side.exists()
    sd = torch.load(side)
    assert set(sd) == {"model.embed_tokens.weight", "model.mm_projector.weight", "model.mm_projector.bias"}
    assert (out_dir / "checkpoint-12" / "pytorch_model.bin").exists()
    assert (out_dir / "trainer_state.json").exists()
    # resume: weights come back bit-exact, step counter restored
    model2 = build_model(cfg, lo.init_params(cfg, seed=1))
    t2 = WrappedTrainer(model=model2, tokenizer=None, args=args, **module)
    t2._load_from_checkpoint(True)
    assert t2.state["global_step"] == 12
    a = dict(model.named_parameters())["model.mm_projector.weight"]
    b = dict(model2.named_parameters())["model.mm_projector.weight"]
    assert torch.equal(a.detach().cpu(), b.detach().cpu())
