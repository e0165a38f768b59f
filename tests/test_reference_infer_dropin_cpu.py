"""Drop-in proof for the inference path at the reference's own entry point: the REFERENCE's
`scripts/inference/infer_from_webdataset.py::main()` is imported in place and run UNCHANGED over llark_b200's modules
aliased under their `m2t.*` names -- `load_pretrained_model` on a sharded fp16-loaded checkpoint (projector from the
shards), the tar-shard eval data module, `get_prompt_end_token_sequence`, `extract_prompt_tokens`,
`KeywordsStoppingCriteria`, `model.generate`, `infer_with_prompt` (fixed-prompt mode), `extract_response_tokens`, CSV out.

CPU run: tests/cpu_ops.py stands in for the CUDA ops and `.cuda()` is made a no-op (test infrastructure only), so this
proves the API surface and host logic; the CUDA generate path is tested in tests/test_model_gpu.py. Skipped where
/root/reference is absent."""
import functools
import importlib.util
import os
import random
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="the reference checkout is not present")


def _setup(tmp_path, monkeypatch):
    """Aliases + CPU harness + a trained-looking sharded checkpoint and an eval shard; returns (name, shard)."""
    from transformers import AutoTokenizer

    import cpu_ops
    import llark_b200.conversation_utils
    import llark_b200.data_modules
    import llark_b200.engine as eng
    import llark_b200.generate
    import llark_b200.infer
    import llark_b200.models.llamav2
    import llark_b200.models.utils
    import llark_b200.tokenizer
    import llark_b200.trainer
    import llark_b200.utils
    from llark_b200.models.llamav2 import WrappedLlamav2ForCausalLM
    from oracle.make_golden_formats import write_sharded_checkpoint
    from test_reference_train_dropin_cpu import _shard, _tiny_checkpoint

    ref_shims.install()
    aliases = {"m2t.models.llamav2": llark_b200.models.llamav2, "m2t.data_modules": llark_b200.data_modules,
               "m2t.models.trainer": llark_b200.trainer, "m2t.models.utils": llark_b200.models.utils,
               "m2t.generate": llark_b200.generate, "m2t.infer": llark_b200.infer,
               "m2t.conversation_utils": llark_b200.conversation_utils, "m2t.tokenizer": llark_b200.tokenizer,
               "m2t.utils": llark_b200.utils}
    for name, mod in aliases.items():
        monkeypatch.setitem(sys.modules, name, mod)
    # --- CPU harness: ops stand-in, no-op .cuda(), checkpoint loaded onto the CPU
    monkeypatch.setattr(eng, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device, cast_to_bf16=not torch.is_grad_enabled())
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(llark_b200.models.utils, "load_pretrained_model",
                        functools.partial(llark_b200.models.utils.load_pretrained_model, device="cpu"))

    # --- a trained checkpoint as the training run leaves it: `<name>/checkpoint-N`, sharded .bin + index, tokenizer with
    #     the four added tokens; "meta-llama/Llama-2" in the path selects the Llama branch and the end-sequence quirk
    rng = random.Random(0)
    name = str(tmp_path / "meta-llama" / "Llama-2-tiny")
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
    state = {k: v.clone() for k, v in m.state_dict().items()}
    write_sharded_checkpoint(ck, state)
    tok.model_max_length = 512
    tok.save_pretrained(ck)
    shard = str(tmp_path / "eval-000000.tar")
    _shard(shard, rng, n=3)
    return name, shard


def _load_script(filename):
    spec = importlib.util.spec_from_file_location(
        "ref_" + filename[:-3], os.path.join(ref_shims.REFERENCE_ROOT, "scripts", "inference", filename))
    script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(script)
    return script


@pytest.mark.parametrize("fixed_prompt", [None, "Describe the audio."])
def test_reference_inference_script_runs_unchanged_over_llark_b200(tmp_path, monkeypatch, fixed_prompt):
    import pandas as pd

    name, shard = _setup(tmp_path, monkeypatch)
    script = _load_script("infer_from_webdataset.py")
    from m2t.arguments import DataArguments, ModelArguments

    margs = ModelArguments(model_name_or_path=name, version="v0", mm_use_audio_start_end=True, mm_hidden_size=12)
    dargs = DataArguments(eval_data_path=shard, is_multimodal=True)
    targs = types.SimpleNamespace(bits=16, bf16=True, fp16=False)
    out = str(tmp_path / "res" / "out.csv")
    script.main(data_args=dargs, model_args=margs, training_args=targs, outfile=out, ckpt_num=5, max_samples=2,
                prompt=fixed_prompt, max_new_tokens=5)

    df = pd.read_csv(out)
    assert list(df.columns) == ["example_id", "prompt_text", "original_completion_text", "model_completion_text"]
    assert list(df["example_id"]) == ["clip000", "clip001", "clip002"]
    assert df["model_completion_text"].notna().all() and df["original_completion_text"].notna().all()
    if fixed_prompt is None:
        assert df["prompt_text"].str.contains("<audio_start>").all()     # the decoded prompt of each element
    else:
        assert (df["prompt_text"] == fixed_prompt).all()


def test_reference_infer_from_encodings_script(tmp_path, monkeypatch):
    """scripts/inference/infer_from_encodings.py::main(): a folder of `.npy` embeddings as the Jukebox CLI writes them
    ([frames, width] float32 numpy arrays go straight into infer_with_prompt) + one fixed prompt."""
    import numpy as np
    import pandas as pd

    name, _ = _setup(tmp_path, monkeypatch)
    script = _load_script("infer_from_encodings.py")
    from m2t.arguments import DataArguments, ModelArguments

    enc_dir = tmp_path / "encodings"
    enc_dir.mkdir()
    g = np.random.default_rng(0)
    for i in range(2):
        np.save(enc_dir / f"clip{i}.npy", g.standard_normal((4, 12)).astype(np.float32))
    margs = ModelArguments(model_name_or_path=name, version="v0", mm_use_audio_start_end=True, mm_hidden_size=12)
    dargs = DataArguments(is_multimodal=True)
    targs = types.SimpleNamespace(bits=16, bf16=True, fp16=False)
    out = str(tmp_path / "res2" / "out.csv")
    script.main(data_args=dargs, model_args=margs, training_args=targs, audio_encodings_dir=str(enc_dir), outfile=out,
                ckpt_num=5, prompt="What instruments do you hear?", max_new_tokens=4)
    df = pd.read_csv(out)
    assert list(df.columns) == ["example_id", "prompt_text", "model_completion_text"] and len(df) == 2
    assert (df["prompt_text"] == "What instruments do you hear?").all() and df["model_completion_text"].notna().all()
