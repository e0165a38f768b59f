"""The reference's own `m2t/train.py::train()` run UNCHANGED on the CUDA engine (the GPU twin of
tests/test_reference_train_dropin_cpu.py, which proves the same entry point over a CPU stand-in for the ops): imported
from the verbatim copy of the reference that `__graft_entry__.build()` stages under the git-ignored baseline/_ref/
(/root/reference does not exist on the GPU box), with the three module aliases of INTEGRATION.md section 2. Every kernel of
the training step runs through libllark_b200.so; nothing is patched but `sys.modules`. Skipped when the copy is absent."""
import json
import os
import random
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_COPY = os.path.join(ROOT, "baseline", "_ref")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF_COPY, "m2t")),
                                 reason="baseline/_ref (copy of the reference made by build()) is not present")]


def test_reference_train_py_runs_unchanged_on_the_cuda_engine(tmp_path, monkeypatch):
    import io
    import pickle
    import tarfile

    from oracle import ref_shims
    from test_reference_train_dropin_cpu import WORDS, _tiny_checkpoint

    import llark_b200.data_modules
    import llark_b200.models
    import llark_b200.models.llamav2
    import llark_b200.trainer

    monkeypatch.setattr(ref_shims, "REFERENCE_ROOT", REF_COPY)
    ref_shims.install()
    for name, mod in (("m2t.models.llamav2", llark_b200.models.llamav2), ("m2t.data_modules", llark_b200.data_modules),
                      ("m2t.models.trainer", llark_b200.trainer)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.delitem(sys.modules, "m2t.train", raising=False)
    import m2t.train as ref_train
    from m2t.arguments import DataArguments, ModelArguments

    assert os.path.realpath(ref_train.__file__).startswith(os.path.realpath(REF_COPY))
    rng = random.Random(0)
    ck = str(tmp_path / "tiny-llama-ckpt")  # "llama" in the path selects the Llama branches of train()
    base_vocab = _tiny_checkpoint(ck, rng)   # hidden 256, 2 heads of 128 (the tcgen05 attention's head size), 2 layers
    shard = str(tmp_path / "train-000000.tar")
    g = torch.Generator().manual_seed(0)
    with tarfile.open(shard, "w") as tf:
        for i in range(6):
            enc = torch.randn(1, 4, 64, generator=g)  # [1, F, A] as jukebox/main.py writes it; A = 64 keeps rows 16-B aligned
            js = {"response": [{"question": " ".join(rng.choices(WORDS, k=5)) + "?",
                                "answer": " ".join(rng.choices(WORDS, k=6)) + "."}]}
            for name, data in ((f"clip{i:03d}.audio_encoding.pyd", pickle.dumps(enc)),
                               (f"clip{i:03d}.json", json.dumps(js).encode())):
                ti = tarfile.TarInfo(name)
                ti.size = len(data)
                tf.addfile(ti, io.BytesIO(data))
    margs = ModelArguments(model_name_or_path=ck, version="v0", tune_mm_mlp_adapter=True, mm_use_audio_start_end=True,
                           mm_hidden_size=64)
    dargs = DataArguments(train_data_path=shard, is_multimodal=True)
    targs = types.SimpleNamespace(
        bits=16, bf16=True, fp16=False, lora_enable=False, cache_dir=None, model_max_length=512, fsdp=None,
        freeze_mm_mlp_adapter=False, device=torch.device("cuda", 0), output_dir=str(tmp_path / "out"), local_rank=-1,
        max_steps=4, per_device_train_batch_size=2, gradient_accumulation_steps=1, learning_rate=1e-3, weight_decay=0.0,
        warmup_ratio=0.0, max_grad_norm=1.0, should_save=True, save_steps=0, logging_steps=1, dataloader_num_workers=0,
        seed=0)
    out = ref_train.train(model_args=margs, data_args=dargs, training_args=targs)

    trainer, tok = out["trainer"], out["tokenizer"]
    model = trainer.model
    assert type(model).__module__ == "llark_b200.models.llamav2" and type(trainer).__module__ == "llark_b200.trainer"
    assert next(model.parameters()).is_cuda and model.engine.launches > 0        # the CUDA engine did the steps
    assert len(tok) == base_vocab + 4 and model.get_input_embeddings().weight.shape[0] == len(tok)
    assert not any(p.requires_grad for p in model.get_output_embeddings().parameters())
    files = sorted(os.listdir(targs.output_dir))
    assert "mm_projector.bin" in files and "pytorch_model.bin" in files and "trainer_state.json" in files
    state = json.load(open(os.path.join(targs.output_dir, "trainer_state.json")))
    hist = [h["loss"] for h in state.get("log_history", []) if "loss" in h]
    assert len(hist) == 4 and all(l == l and l < 20 for l in hist) and hist[-1] < hist[0]
    sd = torch.load(os.path.join(targs.output_dir, "mm_projector.bin"))
    assert {"model.mm_projector.bias", "model.mm_projector.weight"} <= set(sd)
