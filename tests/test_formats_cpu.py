"""CPU parity tests of the on-disk formats and checkpoint formats either side of the hot path:
llark_b200.data_modules / gcs_utils / models.utils / utils are fed the inputs of oracle/make_golden_formats.py and must
reproduce, bit for bit, what the REFERENCE's own functions produced for them (tests/golden/formats_ref.pt)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _same(a, b, path="root"):
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        assert isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor), path
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), path
    elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        assert np.array_equal(np.asarray(a), np.asarray(b)), path
    elif isinstance(a, dict):
        assert isinstance(b, dict) and sorted(a.keys()) == sorted(b.keys()), (path, sorted(a.keys()), sorted(b.keys()))
        for k in a:
            _same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), (path, len(a), len(b) if hasattr(b, "__len__") else b)
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, (path, a, b)


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "formats_ref.pt"), weights_only=False)


def test_on_disk_formats_bit_exact_against_reference(gold):
    from oracle.make_golden_formats import run_formats
    import llark_b200.data_modules as dm
    import llark_b200.gcs_utils as gcs

    got = run_formats(dm, gcs)
    _same(got, gold["formats"])
    # what the golden run pinned, spelled out: skipped elements, both prompt orders, weighted resampling, error text
    conv = got["to_conversation_hf"]
    assert set(conv["__key__"]) == {"clip000", "clip001", "clip002"} and "response" not in conv["json"][0]
    assert {c[0]["value"].startswith("<audio>\n") for c in conv["conversations"]} == {True, False}
    assert got["repeat_shards_error"].startswith("probability for shard /data/unknown-0.tar not defined")
    assert got["gcs_prefix"] == ["pipe:gsutil cat gs://b/x.tar", "/local/x.tar"]
    assert len(got["msgpack"]) == 3 and got["npy_missing"] is None and got["npy_numpy_type"] == "ndarray"


def test_checkpoint_formats_bit_exact_against_reference(gold):
    from oracle.make_golden_formats import run_checkpoints
    import llark_b200.models.utils as mu
    import llark_b200.utils as ut
    from llark_b200.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    got = run_checkpoints(WrappedLlamav2ForCausalLM, WrappedLlamav2Config, mu, ut)
    _same(got, gold["checkpoints"])
    assert got["full_unexpected"] == ["model.not_a_key"] and got["full_missing"] == []
    assert "lm_head.weight" in got["partial_missing"] and "lm_head.weight" in got["partial_untouched"]
    assert sorted(got["proj_state"]) == ["model.mm_projector.bias", "model.mm_projector.weight"]
    assert got["autocast"] == ["torch.bfloat16", "torch.float16", "None", "torch.float32"]


def test_hf_style_shard_dataset_pipeline(tmp_path):
    """read_hf_dataset's pipeline on msgpack and jsonl shards: split by node, batch-8 flat-map, shuffle, three maps."""
    import json

    import msgspec

    from oracle.make_golden_formats import batch_to_examples, hf_batch, make_tokenizer
    import llark_b200.data_modules as dm

    ex = batch_to_examples(hf_batch(seed=4, n=5)) + batch_to_examples(hf_batch(seed=5, n=5))
    for i, e in enumerate(ex):
        e["__key__"] = f"k{i:02d}"
    shard = tmp_path / "s0.msgpack"
    shard.write_bytes(msgspec.msgpack.encode(ex))
    jl = tmp_path / "s0.jsonl"
    jl.write_text("".join(json.dumps(e) + "\n" for e in ex))
    tok = make_tokenizer()
    tok.add_special_tokens({"pad_token": "[PAD]"})
    tok.add_tokens(["<audio_patch>", "<audio_start>", "<audio_end>"], special_tokens=True)
    cfg = dict(is_multimodal=True, sep_audio_conv_front=False, use_audio_start_end=True)
    seen = {}
    for path in (str(shard), str(jl)):
        keys = []
        for rank in (0, 1):
            ds = dm.HFStyleShardDataset([path], cfg, tok, is_train=False, rank=rank, world=2, seed=3, shuffle_buffer=4)
            out = list(ds)
            for o in out:
                assert o["audio_encoding"].dim() == 2 and o["input_ids"].dtype == torch.int64
                assert o["labels"].shape == o["input_ids"].shape and "audio_encoding_shape" in o
                n_patch = int((o["input_ids"] == tok.convert_tokens_to_ids("<audio_patch>")).sum())
                assert n_patch == o["audio_encoding"].shape[0]
            keys.append(sorted(o["__key__"] for o in out))
        assert not set(keys[0]) & set(keys[1])          # ranks see disjoint examples
        seen[path] = sorted(keys[0] + keys[1])
    assert seen[str(shard)] == seen[str(jl)] and len(seen[str(jl)]) > 0
    coll = dm.DataCollatorForSupervisedDataset(tokenizer=tok)
    ds = dm.HFStyleShardDataset([str(jl)], cfg, tok, is_train=False, shuffle_buffer=4)
    # collator consumes the elements as it does the tar-shard ones
    batch = coll([{"input_ids": o["input_ids"], "labels": o["labels"], "audio_encoding": o["audio_encoding"]}
                  for o in list(ds)[:3]])
    assert batch["input_ids"].shape[0] == 3 and batch["attention_mask"].dtype == torch.bool


def test_jukebox_cli_batching_and_output_names(tmp_path, monkeypatch):
    """jukebox/main.py:203-256: sorted inputs, --batch_size / --batch_idx slicing, `.wav` -> `.npy`, lazy model load,
    "Invalid batch index". The embedding itself is stubbed (it needs the GPU; tests/test_jukebox_gpu.py covers it)."""
    import llark_b200.jukebox.main as jm

    inp, out = tmp_path / "in", tmp_path / "out"
    inp.mkdir()
    for name in ("c.wav", "a.wav", "b.wav", "d.wav", "e.wav"):
        (inp / name).write_bytes(b"")
    loads, seen = [], []
    monkeypatch.setattr(jm, "load_model", lambda *a, **k: loads.append(1) or ("hps", "vq", "prior"))

    def fake_acts(path, hps, vqvae, top_prior, meanpool=True, pool_frames_per_second=None):
        seen.append((os.path.basename(str(path)), pool_frames_per_second, meanpool))
        return np.full((2, 3), len(seen), dtype=np.float32)

    monkeypatch.setattr(jm, "get_acts_from_file", fake_acts)
    jm.main(["--input_dir", str(inp), "--output_dir", str(out), "--batch_size", "2", "--batch_idx", "1",
             "--pool-frames-per-second", "5"])
    assert seen == [("c.wav", 5, True), ("d.wav", 5, True)] and len(loads) == 1
    assert sorted(os.listdir(out)) == ["c.npy", "d.npy"] and np.load(out / "d.npy")[0, 0] == 2
    with pytest.raises(ValueError, match="Invalid batch index"):
        jm.main(["--input_dir", str(inp), "--output_dir", str(out), "--batch_size", "2", "--batch_idx", "3"])
    seen.clear()
    jm.main(["--input_dir", str(inp), "--output_dir", str(out)])
    assert [s[0] for s in seen] == ["a.wav", "b.wav", "c.wav", "d.wav", "e.wav"] and seen[0][1] == 10
    empty = tmp_path / "empty"
    empty.mkdir()
    n_loads = len(loads)
    jm.main(["--input_dir", str(empty), "--output_dir", str(out)])
    assert len(loads) == n_loads  # nothing to do -> the model is never built
