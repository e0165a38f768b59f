"""CPU tests of the host-side index work (llark_b200/data_modules.py) against golden vectors produced by the
REFERENCE's own m2t/data_modules.py functions (oracle/make_golden_data.py): bit-exact input_ids / labels /
attention_mask, prompt construction, the masking quirks of SURVEY Appendix D.4-D.7."""
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "data_ref.pt"), weights_only=False)


def test_pipeline_bit_exact_against_reference(gold):
    from oracle.make_golden_data import make_tokenizer, run_pipeline
    import llark_b200.data_modules as dm

    assert dm.DEFAULT_CONVERSATION_HEADER == gold["header"]
    tok = make_tokenizer()
    convs, samples, batches = run_pipeline(dm, tok, seed=0)
    assert [c["conversations"][0]["value"][:80] for c in convs] == gold["prompts"]
    assert len(samples) == len(gold["samples"])
    for s, g in zip(samples, gold["samples"]):
        assert torch.equal(s["input_ids"], g["input_ids"])
        assert torch.equal(s["labels"], g["labels"])
        assert tuple(s["audio_encoding"].shape) == g["enc_shape"]
        assert s["example_id"] == g["example_id"]
    for b, g in zip(batches, gold["batches"]):
        assert torch.equal(b["input_ids"], g["input_ids"])
        assert torch.equal(b["labels"], g["labels"])
        assert b["attention_mask"].dtype == torch.bool and torch.equal(b["attention_mask"], g["attention_mask"])
        assert isinstance(b["audio_encodings"], list) == g["enc_is_list"]
    assert [dm._get_audio_token_len(list(s)) for s in ([1, 240, 4800], [240, 4800], [4800], [3, 1, 2])] == gold[
        "audio_token_len"]
    assert [dm.concat_audio_token_and_prompt("q", True), dm.concat_audio_token_and_prompt("q", False)] == gold["concat"]


def test_masking_quirks():
    """Header fully masked; first two tokens of the human turn unmasked; assistant turn is the supervised part."""
    from oracle.make_golden_data import make_tokenizer
    import llark_b200.data_modules as dm

    tok = make_tokenizer()
    e = {"id": "x", "audio_encoding": [0.0] * 32, "audio_encoding_shape": [2, 16],
         "conversations": [{"from": "human", "value": "<audio>\nwhat genre?"}, {"from": "gpt", "value": "jazz."}]}
    cfg = dict(is_multimodal=True, sep_audio_conv_front=False, audio_folder=None, use_audio_start_end=True,
               audio_processor=None)
    out = dm.preprocess_for_lm_mappable(dm.preprocess_multimodal_mappable(e, cfg), tok)
    ids, labels = out["input_ids"], out["labels"]
    n_patch = (ids == tok.convert_tokens_to_ids("<audio_patch>")).sum().item()
    assert n_patch == 2
    start = (ids == tok.convert_tokens_to_ids("<audio_start>")).nonzero().item()
    assert ids[start + n_patch + 1] == tok.convert_tokens_to_ids("<audio_end>")
    header_len = tok(dm.DEFAULT_CONVERSATION_HEADER, return_tensors="pt").input_ids.shape[1]
    assert (labels[:header_len] == -100).all()
    assert (labels[header_len: header_len + 2] != -100).all()  # the reference leaves two tokens unmasked
    assert (labels != -100).sum().item() > 2


def test_tar_shard_reader_and_make_data_module(tmp_path):
    import io
    import json
    import pickle
    import tarfile
    from types import SimpleNamespace

    from oracle.make_golden_data import make_tokenizer, synthetic_elements
    import llark_b200.data_modules as dm

    paths = []
    for shard in range(2):
        p = tmp_path / f"train-{shard:03d}.tar"
        with tarfile.open(p, "w") as tf:
            for e in synthetic_elements(n=3, seed=shard)[:3]:
                for ext, payload in (("audio_encoding.pyd", pickle.dumps(e["audio_encoding.pyd"])),
                                     ("json", json.dumps(e["json"]).encode())):
                    ti = tarfile.TarInfo(f"{e['__key__']}s{shard}.{ext}")
                    ti.size = len(payload)
                    tf.addfile(ti, io.BytesIO(payload))
        paths.append(str(p))
    url = str(tmp_path / "train-{000..001}.tar")
    assert dm.expand_url_to_file_list(url) == paths
    args = SimpleNamespace(is_multimodal=True, sep_audio_conv_front=False, audio_folder=None,
                           mm_use_audio_start_end=True, train_data_path=None, eval_data_path=url,
                           apply_task_sample_probs=False, task_sample_probs=None, eval_data_subsample=None)
    tok = make_tokenizer()
    module = dm.make_data_module(tok, args)
    assert module["train_dataset"] is None
    random.seed(0)
    items = list(module["eval_dataset"])
    assert len(items) == 2 * (1 + 2 + 3)
    batch = module["data_collator"](items[:4])
    assert batch["input_ids"].shape[0] == 4 and batch["attention_mask"].dtype == torch.bool
    # rank sharding: two ranks see disjoint shards
    r0 = dm.TarShardDataset(paths, dm.make_mm_config(args), tok, False, rank=0, world=2)
    r1 = dm.TarShardDataset(paths, dm.make_mm_config(args), tok, False, rank=1, world=2)
    k0 = {i["example_id"] for i in r0}
    k1 = {i["example_id"] for i in r1}
    assert k0 and k1 and not (k0 & k1)
    # a training rank never starves -- with ONE shard and two ranks both ranks draw from it (wds
    # resampled=True), the Q/A pairs of a clip are spread by the 100-element buffer, and shard weights follow
    # task_sample_probs; eval_data_subsample thins the evaluation stream
    import itertools

    random.seed(0)
    one = dm.TarShardDataset(paths[:1], dm.make_mm_config(args), tok, True, rank=1, world=2)
    got = list(itertools.islice(iter(one), 20))
    assert len(got) == 20
    w = dm.TarShardDataset(paths, dm.make_mm_config(args), tok, True, rank=0, world=1,
                           task_sample_probs={"train-000": 1.0, "train-001": 0.0})
    ids = {i["example_id"] for i in itertools.islice(iter(w), 30)}
    assert ids and all(str(i).endswith("s0") for i in ids), ids
    with pytest.raises(ValueError):
        dm.TarShardDataset(paths, dm.make_mm_config(args), tok, True, task_sample_probs={"nomatch": 1.0})
    random.seed(0)
    args.eval_data_subsample = 0.5
    thin = list(dm.make_data_module(tok, args)["eval_dataset"])
    assert 0 < len(thin) < len(items)
