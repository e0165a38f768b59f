"""Drop-in proof at the reference's own entry point (BASELINE configs[0], the plumbing case): the REFERENCE's
`m2t/train.py::train()` is imported in place from /root/reference and run UNCHANGED with the three module aliases of
INTEGRATION.md §2 (m2t.models.llamav2, m2t.data_modules, m2t.models.trainer -> llark_b200), from
`WrappedLlamav2ForCausalLM.from_pretrained` through tokenizer surgery, adapter initialisation, `make_data_module` over a
webdataset-layout tar shard, `WrappedTrainer.train()` and `safe_save_model_for_hf_trainer`.

No GPU exists in the CPU suite, so tests/cpu_ops.py (test infrastructure) stands in for the CUDA ops: what is proven
here is the API surface and the host logic, not the kernels. Skipped where /root/reference is absent (the GPU box)."""
import io
import json
import os
import pickle
import random
import sys
import tarfile
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="the reference checkout is not present")

WORDS = ("rock jazz tempo key of C minor guitar drums fast slow bright dark what is the genre describe a chat between "
         "curious human and an artificial intelligence assistant gives helpful detailed polite answers to questions "
         "Human Assistant audio").split()


def _tiny_checkpoint(ck, rng):
    """A local `model_name_or_path`: byte-level BPE tokenizer that prepends <s> like Llama's + random tiny weights."""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors, trainers
    from transformers import PreTrainedTokenizerFast

    from llark_b200.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    corpus = [" ".join(rng.choices(WORDS, k=12)) + "." for _ in range(500)]
    tk = Tokenizer(models.BPE(unk_token="<unk>"))
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.train_from_iterator(corpus, trainers.BpeTrainer(vocab_size=300, special_tokens=["<unk>", "<s>", "</s>"],
                                                       initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    tk.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, bos_token="<s>", eos_token="</s>", unk_token="<unk>")
    fast.save_pretrained(ck)
    cfg = WrappedLlamav2Config(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                               num_key_value_heads=2, vocab_size=len(fast), max_position_embeddings=1024)
    torch.manual_seed(0)
    WrappedLlamav2ForCausalLM(cfg).save_pretrained(ck)
    return len(fast)


def _shard(path, rng, n=6):
    g = torch.Generator().manual_seed(0)
    with tarfile.open(path, "w") as tf:
        for i in range(n):
            enc = torch.randn(1, 4, 12, generator=g)  # [1, F, A] as jukebox/main.py writes it
            js = {"response": [{"question": " ".join(rng.choices(WORDS, k=5)) + "?",
                                "answer": " ".join(rng.choices(WORDS, k=6)) + "."}]}
            for name, data in ((f"clip{i:03d}.audio_encoding.pyd", pickle.dumps(enc)),
                               (f"clip{i:03d}.json", json.dumps(js).encode())):
                ti = tarfile.TarInfo(name)
                ti.size = len(data)
                tf.addfile(ti, io.BytesIO(data))


def test_reference_train_py_runs_unchanged_over_llark_b200(tmp_path, monkeypatch):
    import cpu_ops
    import llark_b200.data_modules
    import llark_b200.engine as eng
    import llark_b200.models
    import llark_b200.models.llamav2
    import llark_b200.optim as optim
    import llark_b200.trainer

    ref_shims.install()
    # --- the three aliases, before m2t.train is imported
    for name, mod in (("m2t.models.llamav2", llark_b200.models.llamav2), ("m2t.data_modules", llark_b200.data_modules),
                      ("m2t.models.trainer", llark_b200.trainer)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.delitem(sys.modules, "m2t.train", raising=False)
    import m2t.train as ref_train
    from m2t.arguments import DataArguments, ModelArguments

    # --- CPU stand-in for the CUDA ops (see module docstring)
    monkeypatch.setattr(eng, "ops", cpu_ops)
    monkeypatch.setattr(optim, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device)
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)

    rng = random.Random(0)
    ck = str(tmp_path / "tiny-llama-ckpt")  # "llama" in the path selects the Llama branches of train()
    base_vocab = _tiny_checkpoint(ck, rng)
    shard = str(tmp_path / "train-000000.tar")
    _shard(shard, rng)
    margs = ModelArguments(model_name_or_path=ck, version="v0", tune_mm_mlp_adapter=True, mm_use_audio_start_end=True,
                           mm_hidden_size=12)
    dargs = DataArguments(train_data_path=shard, is_multimodal=True)
    # transformers.TrainingArguments needs `accelerate` (absent here); train() only reads attributes
    targs = types.SimpleNamespace(
        bits=16, bf16=True, fp16=False, lora_enable=False, cache_dir=None, model_max_length=512, fsdp=None,
        freeze_mm_mlp_adapter=False, device=torch.device("cpu"), output_dir=str(tmp_path / "out"), local_rank=-1,
        max_steps=3, per_device_train_batch_size=2, gradient_accumulation_steps=1, learning_rate=1e-3, weight_decay=0.0,
        warmup_ratio=0.0, max_grad_norm=1.0, should_save=True, save_steps=0, logging_steps=1, dataloader_num_workers=0,
        seed=0)
    out = ref_train.train(model_args=margs, data_args=dargs, training_args=targs)

    assert sorted(out) == ["data_module", "tokenizer", "trainer"]
    trainer, tok = out["trainer"], out["tokenizer"]
    model = trainer.model
    assert type(model).__module__ == "llark_b200.models.llamav2" and type(trainer).__module__ == "llark_b200.trainer"
    # tokenizer surgery of train.py:108-116 + initialize_audio_tokenizer: [PAD], <audio_patch>, <audio_start>, <audio_end>
    assert len(tok) == base_vocab + 4 and model.get_input_embeddings().weight.shape[0] == len(tok)
    aec = model.get_model().audio_encoder_config
    assert [aec.audio_start_token, aec.audio_end_token] == tok.convert_tokens_to_ids(["<audio_start>", "<audio_end>"])
    assert model.config.use_cache is False and model.config.tune_mm_mlp_adapter is True
    assert next(model.get_model().mm_projector.parameters()).dtype == torch.bfloat16
    assert not any(p.requires_grad for p in model.get_output_embeddings().parameters())  # lm_head frozen (:412-415)
    files = sorted(os.listdir(targs.output_dir))
    assert "mm_projector.bin" in files and "pytorch_model.bin" in files and "trainer_state.json" in files
    state = json.load(open(os.path.join(targs.output_dir, "trainer_state.json")))
    hist = [h["loss"] for h in state.get("log_history", []) if "loss" in h]
    assert hist and len(hist) == 3 and hist[-1] < hist[0]
    sd = torch.load(os.path.join(targs.output_dir, "mm_projector.bin"))
    # the side file keeps the projector and the embedding table (keys_to_match of m2t/models/trainer.py:112-120)
    assert {"model.mm_projector.bias", "model.mm_projector.weight"} <= set(sd)
    assert all(("mm_projector" in k) or ("embed_tokens" in k) for k in sd)


def test_reference_train_py_takes_its_mpt_branch_over_llark_b200(tmp_path, monkeypatch):
    """The `"mpt" in model_name_or_path` branch of m2t/train.py (WrappedMPTForCausalLM, fast tokenizer, tied embeddings):
    same unchanged `train()`, with m2t.models.mpt aliased as well. HOST LOGIC ONLY -- five of the MPT ops have no CUDA
    kernel yet (llark_b200/engine_mpt.py); the stand-in implements them in PyTorch for this run."""
    import cpu_ops
    import llark_b200.data_modules
    import llark_b200.engine as eng
    import llark_b200.engine_mpt as engm
    import llark_b200.models.llamav2
    import llark_b200.models.mpt
    import llark_b200.optim as optim
    import llark_b200.trainer
    from llark_b200.models.mpt import WrappedMPTConfig, WrappedMPTForCausalLM
    from tokenizers import Tokenizer, models, pre_tokenizers, trainers
    from transformers import PreTrainedTokenizerFast

    ref_shims.install()
    for name, mod in (("m2t.models.llamav2", llark_b200.models.llamav2), ("m2t.models.mpt", llark_b200.models.mpt),
                      ("m2t.data_modules", llark_b200.data_modules), ("m2t.models.trainer", llark_b200.trainer)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.delitem(sys.modules, "m2t.train", raising=False)
    import m2t.train as ref_train
    from m2t.arguments import DataArguments, ModelArguments

    for mod in (eng, engm, optim):
        monkeypatch.setattr(mod, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device)
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)

    rng = random.Random(1)
    ck = str(tmp_path / "tiny-mpt-ckpt")
    assert "llama" not in ck
    corpus = [" ".join(rng.choices(WORDS, k=12)) + "." for _ in range(500)]
    tk = Tokenizer(models.BPE(unk_token="<unk>"))
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.train_from_iterator(corpus, trainers.BpeTrainer(vocab_size=300, special_tokens=["<unk>", "<|endoftext|>"],
                                                       initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    fast = PreTrainedTokenizerFast(tokenizer_object=tk, eos_token="<|endoftext|>", unk_token="<unk>")  # no BOS, as GPT-NeoX
    fast.save_pretrained(ck)
    cfg = WrappedMPTConfig(d_model=256, n_heads=2, n_layers=2, vocab_size=len(fast), max_seq_len=1024, no_bias=True,
                           attn_config={"alibi": True, "qk_ln": True, "attn_impl": "torch"})
    torch.manual_seed(0)
    WrappedMPTForCausalLM(cfg).save_pretrained(ck)
    shard = str(tmp_path / "train-000000.tar")
    _shard(shard, rng)
    margs = ModelArguments(model_name_or_path=ck, version="v0", tune_mm_mlp_adapter=True, mm_use_audio_start_end=True,
                           mm_hidden_size=12)
    dargs = DataArguments(train_data_path=shard, is_multimodal=True)
    targs = types.SimpleNamespace(
        bits=16, bf16=True, fp16=False, lora_enable=False, cache_dir=None, model_max_length=512, fsdp=None,
        freeze_mm_mlp_adapter=False, device=torch.device("cpu"), output_dir=str(tmp_path / "out"), local_rank=-1,
        max_steps=3, per_device_train_batch_size=2, gradient_accumulation_steps=1, learning_rate=1e-3, weight_decay=0.0,
        warmup_ratio=0.0, max_grad_norm=1.0, should_save=True, save_steps=0, logging_steps=1, dataloader_num_workers=0,
        seed=0)
    out = ref_train.train(model_args=margs, data_args=dargs, training_args=targs)
    model = out["trainer"].model
    assert type(model).__name__ == "WrappedMPTForCausalLM" and type(model).__module__ == "llark_b200.models.mpt"
    assert model.get_input_embeddings() is model.get_output_embeddings()                     # tied
    assert model.get_input_embeddings().weight.shape[0] == len(out["tokenizer"])
    state = json.load(open(os.path.join(targs.output_dir, "trainer_state.json")))
    hist = [h["loss"] for h in state["log_history"]]
    assert len(hist) == 3 and hist[-1] < hist[0]
    sd = torch.load(os.path.join(targs.output_dir, "pytorch_model.bin"))
    assert "transformer.blocks.1.attn.q_ln.weight" in sd and "transformer.mm_projector.weight" in sd

    # ... and what that run leaves behind loads through the MPT branch of load_pretrained_model
    # (m2t/models/utils.py:149-169: from_pretrained -> projector from pytorch_model.bin -> whole state dict -> audio token
    # ids from the tokenizer; no initialize_audio_tokenizer), as the inference scripts do with torch_dtype=float16
    import llark_b200.models.utils as mutils

    name = str(tmp_path / "mosaicml" / "mpt-tiny")
    ck3 = os.path.join(name, "checkpoint-3")
    os.makedirs(ck3)
    out["tokenizer"].save_pretrained(ck3)
    model.config.save_pretrained(ck3)
    torch.save(sd, os.path.join(ck3, "pytorch_model.bin"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)

    def ensure_packed_any_dtype(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device, cast_to_bf16=not torch.is_grad_enabled())
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed_any_dtype)
    loaded, tok = mutils.load_pretrained_model(name, 3, torch_dtype=torch.float16, device="cpu")
    assert type(loaded).__module__ == "llark_b200.models.mpt" and len(tok) == len(out["tokenizer"])
    lsd = loaded.state_dict()
    assert set(lsd) == set(sd)
    for k in ("transformer.mm_projector.weight", "transformer.mm_projector.bias", "transformer.wte.weight",
              "transformer.blocks.0.attn.Wqkv.weight"):
        # backbone tensors arrive through from_pretrained(torch_dtype=float16); the projector is created afterwards in
        # the default dtype and filled from the bf16 file -- either way within one fp16 rounding of what was saved
        assert torch.allclose(lsd[k].float(), sd[k].float(), rtol=1e-3, atol=1e-4), k
    aec = loaded.transformer.audio_encoder_config
    assert [aec.audio_start_token, aec.audio_end_token] == tok.convert_tokens_to_ids(["<audio_start>", "<audio_end>"])
    assert loaded.config.mm_use_audio_start_end is True
    ids = torch.tensor([[5, 6, aec.audio_start_token, tok.convert_tokens_to_ids("<audio_patch>"), aec.audio_end_token, 7, 8]])
    enc = torch.randn(1, 1, 12, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = loaded(input_ids=ids, audio_encodings=enc).logits.float()
        b = model(input_ids=ids, audio_encodings=enc).logits.float()
    assert ((a - b).norm() / b.norm()).item() < 3e-2          # fp16-loaded weights rounded to bf16 vs the bf16 originals
    # greedy, cache-less generate of the MPT wrapper == argmax of repeated full forwards (what m2t/infer.py:146-152 needs)
    got = loaded.generate(input_ids=ids, audio_encodings=enc, max_new_tokens=3, labels=None, example_id="x")
    want = ids.clone()
    with torch.no_grad():
        for _ in range(3):
            nxt = loaded(input_ids=want, audio_encodings=enc).logits[:, -1].float().argmax(-1)
            want = torch.cat([want, nxt[:, None]], dim=1)
            if nxt.item() == tok.eos_token_id:
                break
    assert torch.equal(got[:, : want.shape[1]], want) and got.shape[1] <= ids.shape[1] + 3
    with pytest.raises(NotImplementedError):
        loaded.generate(input_ids=ids, audio_encodings=enc, do_sample=True)
    other = tmp_path / "other" / "model" / "checkpoint-3"       # neither family in the name: `model` stays unbound upstream
    os.makedirs(other)
    out["tokenizer"].save_pretrained(str(other))
    torch.save({}, other / "pytorch_model.bin")
    with pytest.raises(UnboundLocalError):
        mutils.load_pretrained_model(str(tmp_path / "other" / "model"), 3)
