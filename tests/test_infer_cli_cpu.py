"""llark_b200.infer_cli (own counterpart of scripts/inference/infer_from_webdataset.py / infer_from_encodings.py) on CPU:
checkpoint folder -> load_pretrained_model -> eval shards or a folder of .npy embeddings -> greedy generation with the
"###" stop -> CSV with the reference's columns. tests/cpu_ops.py stands in for the CUDA ops and `.cuda()` is a no-op
(host logic only; the CUDA generate path is tested in tests/test_model_gpu.py)."""
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))


@pytest.fixture()
def run_dir(tmp_path, monkeypatch):
    from transformers import AutoTokenizer

    import cpu_ops
    import llark_b200.engine as eng
    from llark_b200.models.llamav2 import WrappedLlamav2ForCausalLM
    from test_reference_train_dropin_cpu import _shard, _tiny_checkpoint

    monkeypatch.setattr(eng, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device, cast_to_bf16=not torch.is_grad_enabled())
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    rng = random.Random(0)
    name = str(tmp_path / "meta-llama" / "Llama-2-tiny")      # the family is chosen by this substring, as upstream
    ck = os.path.join(name, "checkpoint-5")
    os.makedirs(ck)
    _tiny_checkpoint(ck, rng)
    tok = AutoTokenizer.from_pretrained(ck)
    tok.add_special_tokens({"pad_token": "[PAD]"})
    m = WrappedLlamav2ForCausalLM.from_pretrained(ck)
    m.config.mm_hidden_size = 12
    m.get_model().initialize_adapter_modules()
    m.config.tune_mm_mlp_adapter = False
    m.resize_token_embeddings(len(tok))
    m.initialize_audio_tokenizer(mm_use_audio_start_end=True, tokenizer=tok, device="cpu")
    m.save_pretrained(ck)
    os.remove(os.path.join(ck, "model.safetensors"))
    torch.save(m.state_dict(), os.path.join(ck, "pytorch_model.bin"))   # single-file checkpoint (WEIGHTS_NAME)
    tok.model_max_length = 512
    tok.save_pretrained(ck)
    shard = str(tmp_path / "eval-000000.tar")
    _shard(shard, rng, n=3)
    return name, shard


@pytest.mark.parametrize("fixed_prompt", [None, "Describe the audio."])
def test_infer_cli_over_eval_shards(run_dir, tmp_path, fixed_prompt):
    import pandas as pd

    from llark_b200 import infer_cli

    name, shard = run_dir
    out = str(tmp_path / "res" / "out.csv")
    argv = ["--model_name_or_path", name, "--ckpt-num", "5", "--eval_data_path", shard, "--bf16", "True", "--use_cpu",
            "True", "--mm_hidden_size", "12", "--lr_scheduler_type", "cosine", "--max-samples", "1", "--max-new-tokens",
            "5", "--outfile", out]
    if fixed_prompt:
        argv += ["--prompt", fixed_prompt]
    rows = infer_cli.main(argv)
    df = pd.read_csv(out)
    assert list(df.columns) == ["example_id", "prompt_text", "original_completion_text", "model_completion_text"]
    assert list(df["example_id"]) == ["clip000", "clip001"] and len(rows) == 2     # --max-samples 1 -> indices 0 and 1
    assert df["model_completion_text"].notna().all() and df["original_completion_text"].notna().all()
    if fixed_prompt:
        assert (df["prompt_text"] == fixed_prompt).all()
    else:
        assert df["prompt_text"].str.contains("<audio_start>").all()


def test_infer_cli_over_embedding_folder(run_dir, tmp_path):
    import numpy as np
    import pandas as pd

    from llark_b200 import infer_cli

    name, _ = run_dir
    enc_dir = tmp_path / "enc"
    enc_dir.mkdir()
    g = np.random.default_rng(0)
    for i in range(2):
        np.save(enc_dir / f"song{i}.npy", g.standard_normal((4, 12)).astype(np.float32))
    out = str(tmp_path / "folder.csv")
    base = ["--model_name_or_path", name, "--ckpt-num", "5", "--audio-encodings-dir", str(enc_dir), "--bf16", "True",
            "--use_cpu", "True", "--lr_scheduler_type", "cosine", "--max-new-tokens", "4", "--outfile", out]
    with pytest.raises(ValueError, match="--prompt"):
        infer_cli.main(base)
    infer_cli.main(base + ["--prompt", "What genre is this song?"])
    df = pd.read_csv(out)
    assert list(df.columns) == ["example_id", "prompt_text", "model_completion_text"]
    assert sorted(os.path.basename(e) for e in df["example_id"]) == ["song0", "song1"]
    assert (df["prompt_text"] == "What genre is this song?").all()


def test_trainer_save_then_load_pretrained_model_keeps_the_projector(run_dir, tmp_path):
    """WrappedTrainer._save writes a single pytorch_model.bin; load_pretrained_model must bring the
    trained mm_projector back (upstream only reads it from shards) instead of a freshly initialised nn.Linear."""
    from types import SimpleNamespace

    from transformers import AutoTokenizer

    from llark_b200.models.llamav2 import WrappedLlamav2ForCausalLM
    from llark_b200.models.utils import load_pretrained_model
    from llark_b200.trainer import WrappedTrainer

    name, _ = run_dir
    src = os.path.join(name, "checkpoint-5")
    tok = AutoTokenizer.from_pretrained(src)
    model, _ = load_pretrained_model(name, 5, torch_dtype=torch.float32, device="cpu")
    with torch.no_grad():   # "trained" projector and embedding rows
        g = torch.Generator().manual_seed(3)
        model.get_model().mm_projector.weight.copy_(torch.randn(model.get_model().mm_projector.weight.shape, generator=g))
        model.get_model().mm_projector.bias.copy_(torch.randn(model.get_model().mm_projector.bias.shape, generator=g))
        model.get_input_embeddings().weight[-2:].add_(1.0)
    trainer = WrappedTrainer.__new__(WrappedTrainer)
    trainer.model, trainer.tokenizer = model, tok
    trainer.args = SimpleNamespace(tune_mm_mlp_adapter=False)
    dst = os.path.join(name, "checkpoint-9")
    trainer._save(dst)
    assert os.path.exists(os.path.join(dst, "pytorch_model.bin"))
    back, _ = load_pretrained_model(name, 9, torch_dtype=torch.float32, device="cpu")
    assert isinstance(back, WrappedLlamav2ForCausalLM)
    for attr in ("weight", "bias"):
        assert torch.equal(getattr(back.get_model().mm_projector, attr).float(),
                           getattr(model.get_model().mm_projector, attr).float()), attr
    assert torch.equal(back.get_input_embeddings().weight.float(), model.get_input_embeddings().weight.float())
