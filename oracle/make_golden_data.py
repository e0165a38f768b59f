"""ORACLE support (test infrastructure only): golden vectors for the index / tokenizer work, produced by the
REFERENCE's own m2t/data_modules.py functions (imported in place through oracle/ref_shims.py) on the synthetic
tokenizer of tests/synth_tokenizer.py. Output: tests/golden/data_ref.pt (bit-exact targets)."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def synthetic_elements(n=5, seed=0):
    rng = random.Random(seed)
    words = "rock jazz tempo key of C minor guitar drums fast slow bright dark what is the genre describe".split()
    g = torch.Generator().manual_seed(seed)
    elems = []
    for i in range(n):
        frames = [240, 240, 3, 7, 1][i % 5]
        shape = (1, frames, 16) if i % 2 else (frames, 16)
        enc = torch.randn(shape, generator=g)
        resp = [{"question": " ".join(rng.choices(words, k=rng.randint(3, 9))) + "?",
                 "answer": " ".join(rng.choices(words, k=rng.randint(2, 12))) + "."} for _ in range(1 + i % 3)]
        elems.append({"__key__": f"clip{i:03d}", "json": {"response": resp}, "audio_encoding.pyd": enc})
    elems.append({"__key__": "bad0", "json": {"response": None}, "audio_encoding.pyd": enc})  # skipped (:305-311)
    return elems


def make_tokenizer():
    from synth_tokenizer import SynthTokenizer

    tok = SynthTokenizer(vocab_size=32000, model_max_length=300)
    tok.add_special_tokens({"pad_token": "[PAD]"})
    tok.add_tokens(["<audio_patch>"], special_tokens=True)
    tok.add_tokens(["<audio_start>", "<audio_end>"], special_tokens=True)
    return tok


def run_pipeline(dm, tok, seed=0):
    """elements -> conversation dicts -> multimodal expansion -> tokenise/mask -> collate (batches of 3)."""
    cfg = dict(is_multimodal=True, sep_audio_conv_front=False, audio_folder=None, use_audio_start_end=True,
               audio_processor=None)
    random.seed(seed)
    convs = list(dm.webdataset_element_to_conversation(iter(synthetic_elements())))
    samples = []
    for e in convs:
        e["audio_encoding"] = e["audio_encoding"].flatten().tolist()
        e = dm.preprocess_multimodal_mappable(e, cfg)
        samples.append(dm.preprocess_for_lm_mappable(e, tok))
    coll = dm.DataCollatorForSupervisedDataset(tokenizer=tok)
    batches = [coll(samples[i: i + 3]) for i in range(0, len(samples), 3)]
    return convs, samples, batches


def golden_data():
    from oracle import ref_shims

    ref_shims.install()
    import m2t.data_modules as ref_dm

    tok = make_tokenizer()
    convs, samples, batches = run_pipeline(ref_dm, tok)
    out = {
        "header": ref_dm.DEFAULT_CONVERSATION_HEADER,
        "prompts": [c["conversations"][0]["value"][:80] for c in convs],
        "samples": [{"input_ids": s["input_ids"], "labels": s["labels"], "enc_shape": tuple(s["audio_encoding"].shape),
                     "example_id": s["example_id"]} for s in samples],
        "batches": [{"input_ids": b["input_ids"], "labels": b["labels"], "attention_mask": b["attention_mask"],
                     "enc_is_list": isinstance(b["audio_encodings"], list)} for b in batches],
        "audio_token_len": [ref_dm._get_audio_token_len(list(s)) for s in ([1, 240, 4800], [240, 4800], [4800], [3, 1, 2])],
        "concat": [ref_dm.concat_audio_token_and_prompt("q", True), ref_dm.concat_audio_token_and_prompt("q", False)],
    }
    torch.save(out, os.path.join(ROOT, "tests", "golden", "data_ref.pt"))
    print("data golden:", len(samples), "samples", len(batches), "batches")


if __name__ == "__main__":
    golden_data()
