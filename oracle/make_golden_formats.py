"""ORACLE support (test infrastructure only): golden vectors for the on-disk formats and checkpoint formats either side
of the hot path, produced by the REFERENCE's own functions imported in place:
m2t/data_modules.py (_to_conversation_hf, repeat_shards, maybe_add_gcs_prefix, preprocess_multimodal, preprocess_for_lm,
hf_preprocess_encodings, gen_from_msgpack_shards, read_jsonl_data), m2t/models/utils.py (load_sharded_checkpoint,
load_sharded_mm_projector_weights), m2t/utils.py, m2t/gcs_utils.py (write_npy, read_audio_encoding).
Output: tests/golden/formats_ref.pt. The input builders below are shared with tests/test_formats_cpu.py, which feeds
the mirror the same inputs and compares bit for bit."""
import copy
import importlib.machinery
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# ------------------------------------------------------------------------------------------ shared synthetic inputs
def hf_batch(seed=0, n=5):
    """A dict-of-lists batch in the jsonl / msgpack element format (flat float encodings + shape)."""
    rng = random.Random(seed)
    words = "rock jazz tempo key of C minor guitar drums fast slow bright dark what is the genre describe".split()
    src = {"__key__": [], "json": [], "audio_encoding": [], "audio_encoding_shape": []}
    for i in range(n):
        frames = [3, 5, 1, 4, 2][i % 5]
        shape = [1, frames, 8] if i % 2 else [frames, 8]
        flat = [round(rng.uniform(-1, 1), 4) for _ in range(frames * 8)]
        resp = [{"question": " ".join(rng.choices(words, k=rng.randint(3, 7))) + "?",
                 "answer": " ".join(rng.choices(words, k=rng.randint(2, 9))) + "."} for _ in range(1 + i % 3)]
        js = {"response": resp, "uri": f"u{i}", "genre": rng.choice(words)}
        if i == 3:
            js["response"] = None      # skipped: no response
        if i == 4:
            js["response"] = "oops"    # skipped: not a list
        src["__key__"].append(f"clip{i:03d}")
        src["json"].append(js)
        src["audio_encoding"].append(flat)
        src["audio_encoding_shape"].append(shape)
    return src


def batch_to_examples(src):
    return [{k: src[k][i] for k in src} for i in range(len(src["__key__"]))]


SHARD_URLS = ["/data/mtt-train-000.tar", "/data/mtt-train-001.tar", "/data/fma-train-000.tar", "/data/giant-000.tar"]
TASK_PROBS = {"mtt": 2.0, "fma": 1.0, "giant": 0.5}


def list_sources(seed=1):
    """Two conversations in the list-of-sources form of preprocess_multimodal / preprocess_for_lm."""
    rng = random.Random(seed)
    words = "describe the tempo and key of this clip please what instruments play".split()
    out = []
    for i in range(2):
        q = " ".join(rng.choices(words, k=5)) + "?"
        a = " ".join(rng.choices(words, k=6)) + "."
        h = f"<audio>\n{q}" if i == 0 else f"{q}\n<audio>"
        out.append([{"from": "human", "value": h}, {"from": "gpt", "value": a}])
    return out, [(4, 8), (2, 8)]


def tiny_llama_config(cls):
    cfg = cls(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=4,
              vocab_size=100, max_position_embeddings=64)
    cfg.mm_hidden_size = 12
    return cfg


def tiny_state(model, seed):
    """Deterministic values for every tensor of the model's state dict (incl. the projector)."""
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(v.shape, generator=g) * 0.1 for k, v in model.state_dict().items()}


def write_sharded_checkpoint(folder, state, extra=None):
    """HF sharded layout: 2 shards + index json; `extra` adds keys that are in the index but not in the model."""
    os.makedirs(folder, exist_ok=True)
    keys = sorted(state.keys())
    half = len(keys) // 2
    shards = {"pytorch_model-00001-of-00002.bin": keys[:half], "pytorch_model-00002-of-00002.bin": keys[half:]}
    weight_map = {}
    for fname, ks in shards.items():
        sd = {k: state[k] for k in ks}
        if extra and fname.endswith("00002-of-00002.bin"):
            sd.update(extra)
        torch.save(sd, os.path.join(folder, fname))
        for k in sd:
            weight_map[k] = fname
    with open(os.path.join(folder, "pytorch_model.bin.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": weight_map}, f)


def make_tokenizer():
    from synth_tokenizer import SynthTokenizer

    tok = SynthTokenizer(vocab_size=96, model_max_length=200)
    return tok


# ------------------------------------------------------------------------------------------ runs (reference or mirror)
def run_formats(dm, gcs, seed=0):
    """Everything in data_modules / gcs_utils form; `dm` is the reference's module or the mirror."""
    out = {}
    random.seed(seed)
    conv = dm._to_conversation_hf(hf_batch())
    out["to_conversation_hf"] = {k: list(v) for k, v in conv.items()}
    np.random.seed(seed)
    out["repeat_shards"] = dm.repeat_shards(list(SHARD_URLS))[:64]
    np.random.seed(seed)
    out["repeat_shards_probs"] = dm.repeat_shards(list(SHARD_URLS), task_sample_probs=dict(TASK_PROBS))[:64]
    try:
        dm.repeat_shards(["/data/unknown-0.tar"], task_sample_probs=dict(TASK_PROBS))
        out["repeat_shards_error"] = None
    except ValueError as e:
        out["repeat_shards_error"] = str(e)
    out["gcs_prefix"] = [dm.maybe_add_gcs_prefix("gs://b/x.tar"), dm.maybe_add_gcs_prefix("/local/x.tar")]
    tok = make_tokenizer()
    tok.add_special_tokens({"pad_token": "[PAD]"})
    tok.add_tokens(["<audio_patch>", "<audio_start>", "<audio_end>"], special_tokens=True)
    for front in (False, True):
        sources, shapes = list_sources()
        cfg = dict(is_multimodal=True, sep_audio_conv_front=front, use_audio_start_end=True)
        mm = dm.preprocess_multimodal(copy.deepcopy(sources), cfg, shapes)
        lm = dm.preprocess_for_lm(copy.deepcopy(mm), tok)
        out[f"list_forms_front{int(front)}"] = {"mm": mm, "input_ids": lm["input_ids"], "labels": lm["labels"]}
    ex = batch_to_examples(hf_batch())[1]
    out["hf_preprocess_encodings"] = dm.hf_preprocess_encodings(dict(ex))["audio_encoding"]
    with tempfile.TemporaryDirectory() as tmp:
        import msgspec

        good = os.path.join(tmp, "a.msgpack")
        with open(good, "wb") as f:
            f.write(msgspec.msgpack.encode(batch_to_examples(hf_batch(seed=2, n=3))))
        bad = os.path.join(tmp, "b.msgpack")
        with open(bad, "wb") as f:
            f.write(b"\xc1not msgpack")
        out["msgpack"] = list(dm.gen_from_msgpack_shards([good, os.path.join(tmp, "missing.msgpack"), bad]))
        jl = os.path.join(tmp, "c.jsonl")
        with open(jl, "w") as f:
            for e in batch_to_examples(hf_batch(seed=3, n=2)):
                f.write(json.dumps(e) + "\n")
        out["jsonl"] = dm.read_jsonl_data(jl)
        ary = np.arange(24, dtype=np.float32).reshape(2, 12)
        gcs.write_npy(os.path.join(tmp, "reps", "sub", "clip7.npy"), ary)
        out["npy_files"] = [os.path.basename(x) for x in gcs.list_files_with_extension(os.path.join(tmp, "reps", "sub"), ".npy")]
        out["npy_torch"] = gcs.read_audio_encoding("clip7", os.path.join(tmp, "reps", "sub"))
        out["npy_numpy_type"] = type(gcs.read_audio_encoding("clip7", os.path.join(tmp, "reps", "sub"), numpy_to_torch=False)).__name__
        out["npy_missing"] = gcs.read_audio_encoding(12345, os.path.join(tmp, "reps", "sub"))
    return out


def run_checkpoints(model_cls, cfg_cls, mu, ut):
    """Sharded-checkpoint loading + m2t/utils.py on a tiny model; `mu` / `ut` are the reference's modules or the mirror."""
    out = {}
    torch.manual_seed(0)
    cfg = tiny_llama_config(cfg_cls)
    with tempfile.TemporaryDirectory() as tmp:
        src = model_cls(cfg)
        src.get_model().initialize_adapter_modules()
        state = tiny_state(src, seed=5)
        write_sharded_checkpoint(os.path.join(tmp, "full"), state, extra={"model.not_a_key": torch.zeros(2)})
        m1 = model_cls(cfg)
        m1.get_model().initialize_adapter_modules()
        res = mu.load_sharded_checkpoint(m1, os.path.join(tmp, "full"))
        out["full_missing"], out["full_unexpected"] = list(res.missing_keys), list(res.unexpected_keys)
        out["full_state"] = {k: v.clone() for k, v in m1.state_dict().items()}
        # a checkpoint that lacks lm_head and one layer-norm: missing keys are reported, nothing raises (upstream quirk)
        partial = {k: v for k, v in state.items() if not k.startswith("lm_head") and "layers.1.input_layernorm" not in k}
        write_sharded_checkpoint(os.path.join(tmp, "partial"), partial)
        m2 = model_cls(cfg)
        m2.get_model().initialize_adapter_modules()
        before = {k: v.clone() for k, v in m2.state_dict().items()}
        res = mu.load_sharded_checkpoint(m2, os.path.join(tmp, "partial"), strict=True)
        out["partial_missing"] = sorted(res.missing_keys)
        after = m2.state_dict()
        out["partial_untouched"] = sorted(k for k in before if torch.equal(before[k], after[k]))
        m3 = model_cls(cfg)
        m3.get_model().initialize_adapter_modules()
        mu.load_sharded_mm_projector_weights(m3, os.path.join(tmp, "full"))
        out["proj_state"] = {k: v.clone() for k, v in m3.state_dict().items() if "mm_projector" in k}
        try:
            mu.load_sharded_checkpoint(m3, os.path.join(tmp, "nowhere"))
            out["no_index_error"] = None
        except ValueError as e:
            out["no_index_error"] = str(e).replace(tmp, "<tmp>")
    A = types.SimpleNamespace
    out["autocast"] = [str(ut.get_autocast_type(A(bits=16, bf16=True, fp16=False))),
                       str(ut.get_autocast_type(A(bits=16, bf16=False, fp16=True))),
                       str(ut.get_autocast_type(A(bits=16, bf16=False, fp16=False))),
                       str(ut.get_autocast_type(A(bits=8, bf16=True, fp16=False)))]
    out["compute_dtype"] = [str(ut.get_compute_dtype(A(fp16=True, bf16=True))), str(ut.get_compute_dtype(A(fp16=False, bf16=True))),
                            str(ut.get_compute_dtype(A(fp16=False, bf16=False)))]
    torch.manual_seed(1)
    m4 = model_cls(cfg)
    tok = make_tokenizer()
    tok.base_vocab = 100
    ut.smart_tokenizer_and_embedding_resize({"pad_token": "[PAD]"}, tok, m4)
    out["resize"] = {"len": len(tok), "emb": m4.get_input_embeddings().weight.data.clone(),
                     "head": m4.get_output_embeddings().weight.data.clone()}
    return out


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    return m


def import_reference_gcs_utils():
    """m2t/gcs_utils.py imports librosa, soundfile and google.cloud.storage at module level and refuses to import
    without two environment variables; none of that is touched by the local-file functions exercised here."""
    os.environ.setdefault("GOOGLE_CLOUD_PROJECT", "none")
    os.environ.setdefault("GCS_BUCKET_NAME", "none")
    added = []
    for name in ("librosa", "soundfile", "google", "google.cloud", "google.cloud.storage"):
        if name not in sys.modules:
            sys.modules[name] = _stub(name)
            added.append(name)
    sys.modules["google.cloud"].storage = sys.modules["google.cloud.storage"]
    try:
        import m2t.gcs_utils as g
    finally:
        for name in added:
            sys.modules.pop(name, None)
    return g


def golden_formats():
    from oracle import ref_shims

    ref_shims.install()
    import m2t.data_modules as ref_dm
    import m2t.models.utils as ref_mu
    import m2t.utils as ref_ut
    from m2t.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    out = {"formats": run_formats(ref_dm, import_reference_gcs_utils()),
           "checkpoints": run_checkpoints(WrappedLlamav2ForCausalLM, WrappedLlamav2Config, ref_mu, ref_ut)}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "formats_ref.pt"))
    print("formats golden:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    golden_formats()
