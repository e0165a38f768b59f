"""Host-side mirror of the index / tokenizer work of m2t/data_modules.py: same
function names, argument meaning and quirks (bit-exact parity is required and tested against golden vectors
produced by the reference's own functions, tests/golden/data_ref.pt).

What is mirrored (reference file:line):
    DEFAULT_CONVERSATION_HEADER               data_modules.py:46 + llava/conversation.py:237-271,390 (conv_v1_2)
    _tokenize_fn                              :57-78
    _mask_targets                             :81-89   (first two tokens of a human turn stay unmasked, offsets come
                                                        from separately tokenised segments -- Appendix D.4)
    sentences_to_formatted_conversation       :92-109  (mutates sentence["value"] in place, as the reference does)
    preprocess_encodings                      :180-186
    DataCollatorForSupervisedDataset          :189-222 (bool attention_mask; list of encodings when ragged)
    _get_audio_token_len                      :225-231
    preprocess_multimodal_mappable            :234-258
    preprocess_for_lm_mappable                :261-284
    concat_audio_token_and_prompt             :287-292
    webdataset_element_to_conversation        :295-340
    make_mm_config / make_data_module         :671-717
The webdataset / GCS reader (:466-520) is replaced by `TarShardDataset`, a plain tarfile reader of the same
on-disk format (`<key>.audio_encoding.pyd` pickle + `<key>.json`), sharded by rank like wds.split_by_node.

On-disk formats either side of the path, same names and behaviour:
    read_jsonl_data                           :49-55
    preprocess_multimodal / preprocess_for_lm :112-178 (the list-of-sources forms)
    assert_hf_src_format / _to_conversation_hf :343-427 (batched dict-of-lists -> one QA pair per element)
    maybe_add_gcs_prefix / repeat_shards      :430-463 (task_sample_probs, 1024 x len(urls) resampled shard list)
    gen_from_msgpack_shards                   :599-615 (a shard is ONE msgpack list of examples)
    hf_preprocess_encodings                   :618-622
    read_hf_dataset                           :625-668 -> `HFStyleShardDataset` (jsonl / msgpack shards without the
                                              `datasets` dependency: split by node, batch-8 flat-map, shuffle buffer
                                              1000, then the three per-element maps)
"""
import copy
import io
import json
import logging
import os
import pickle
import random
import tarfile
from collections import defaultdict
from dataclasses import dataclass
from functools import partial
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from llark_b200.special_tokens import (DEFAULT_AUDIO_END_TOKEN, DEFAULT_AUDIO_PATCH_TOKEN, DEFAULT_AUDIO_START_TOKEN,
                                       DEFAULT_AUDIO_TOKEN, IGNORE_INDEX)

# conv_v1_2 of the reference's m2t/llava/conversation.py: only .system, .roles and .sep are used on this path.
CONVERSATION_SYSTEM = ("A chat between a curious human and an artificial intelligence assistant. "
                       "The assistant gives helpful, detailed, and polite answers to the human's questions.")
CONVERSATION_ROLES = ("Human", "Assistant")
CONVERSATION_SEP = "###"
DEFAULT_CONVERSATION_HEADER = f"{CONVERSATION_SYSTEM}\n\n"


def _tokenize_fn(strings: Sequence[str], tokenizer) -> Dict:
    """Tokenize a list of strings one by one (each gets its own BOS); lengths exclude padding."""
    tokenized_list = [
        tokenizer(text, return_tensors="pt", padding="longest", max_length=tokenizer.model_max_length,
                  truncation=True)
        for text in strings
    ]
    input_ids = labels = [t.input_ids[0] for t in tokenized_list]
    input_ids_lens = labels_lens = [t.input_ids.ne(tokenizer.pad_token_id).sum().item() for t in tokenized_list]
    return dict(input_ids=input_ids, labels=labels, input_ids_lens=input_ids_lens, labels_lens=labels_lens)


def _mask_targets(target, tokenized_lens, speakers):
    """Mask (in place) the header and every human turn except its first two tokens. Turn k starts at the sum of
    the separately-tokenised lengths before it (each of those carries its own BOS: the reference's off-by-BOS
    quirk is preserved on purpose)."""
    from itertools import accumulate

    starts = list(accumulate(tokenized_lens))  # starts[k] = offset of turn k (after the header)
    target[: starts[0]] = IGNORE_INDEX
    for begin, length, who in zip(starts, tokenized_lens[1:], speakers):
        if who == "human":
            target[begin + 2: begin + length] = IGNORE_INDEX


_ROLE_OF = {"human": CONVERSATION_ROLES[0], "gpt": CONVERSATION_ROLES[1]}


def sentences_to_formatted_conversation(header, source, get_conversation=True) -> str:
    """header + "### <Role>: <text>\n" per turn + trailing "### ". Each turn's "value" is REWRITTEN in place with
    its formatted form, exactly as the reference does (later code tokenises those rewritten values)."""
    pieces = [header]
    for turn in source:
        role = _ROLE_OF.get(turn["from"].lower(), "unknown")
        turn["value"] = f"### {role}: {turn['value']}\n"
        if get_conversation:
            pieces.append(turn["value"])
    pieces.append("### ")
    return "".join(pieces)


def preprocess_encodings(audio_encoding, audio_encoding_shape: List[int]) -> torch.Tensor:
    encoding = torch.Tensor(audio_encoding)
    encoding = encoding.reshape(audio_encoding_shape)
    if encoding.ndim == 3:
        encoding = torch.squeeze(encoding, 0)
    return encoding


@dataclass
class DataCollatorForSupervisedDataset(object):
    """Right-pad ids with pad_token_id and labels with -100; bool attention_mask; stack encodings if same shape."""

    tokenizer: Any

    def __call__(self, instances: Sequence[Dict]) -> Dict[str, torch.Tensor]:
        pad_id = self.tokenizer.pad_token_id
        width = max(int(inst["input_ids"].shape[0]) for inst in instances)
        n = len(instances)
        input_ids = torch.full((n, width), pad_id, dtype=instances[0]["input_ids"].dtype)
        labels = torch.full((n, width), IGNORE_INDEX, dtype=instances[0]["labels"].dtype)
        for row, inst in enumerate(instances):
            k = inst["input_ids"].shape[0]
            input_ids[row, :k] = inst["input_ids"]
            labels[row, : inst["labels"].shape[0]] = inst["labels"]
        batch = dict(input_ids=input_ids, labels=labels, attention_mask=input_ids.ne(pad_id))
        if "audio_encoding" not in instances[0]:
            logging.warning("key `audio_encoding` not detected in data collator inputs.")
            return batch
        encodings = [inst["audio_encoding"] for inst in instances]
        same = all(x is not None and x.shape == encodings[0].shape for x in encodings)
        batch["audio_encodings"] = torch.stack(encodings) if same else encodings
        return batch


def _get_audio_token_len(enc_shape: List[int]) -> int:
    if len(enc_shape) == 3 and enc_shape[0] == 1:
        return enc_shape[1]
    return enc_shape[0]


def preprocess_multimodal_mappable(e: Dict[str, Any], multimodal_cfg: Dict[str, Any]):
    assert not multimodal_cfg["sep_audio_conv_front"], \
        "need to implement logic for this; see datasets.preprocess_multimodal()."
    audio_token_len = _get_audio_token_len(e["audio_encoding_shape"])
    replace_token = DEFAULT_AUDIO_PATCH_TOKEN * audio_token_len
    if multimodal_cfg["use_audio_start_end"]:
        replace_token = DEFAULT_AUDIO_START_TOKEN + replace_token + DEFAULT_AUDIO_END_TOKEN
    out = []
    for sentence in e["conversations"]:
        conv = copy.deepcopy(sentence)
        conv["value"] = sentence["value"].replace(DEFAULT_AUDIO_TOKEN, replace_token)
        out.append(conv)
    e["conversations"] = out
    return e


def preprocess_for_lm_mappable(e: Dict[str, Any], tokenizer, header: str = DEFAULT_CONVERSATION_HEADER):
    source = e["conversations"]
    conversation = sentences_to_formatted_conversation(header, source)
    input_ids = _tokenize_fn([conversation], tokenizer)["input_ids"][0]
    target = copy.deepcopy(input_ids)
    tokenized_lens = _tokenize_fn([header] + [s["value"] for s in source], tokenizer)["input_ids_lens"]
    speakers = [sentence["from"] for sentence in source]
    _mask_targets(target, tokenized_lens, speakers)
    audio_encoding = preprocess_encodings(e["audio_encoding"], e["audio_encoding_shape"])
    return dict(input_ids=input_ids, labels=target, audio_encoding=audio_encoding, example_id=e["id"])


def concat_audio_token_and_prompt(prompt: str, audio_first: bool) -> str:
    if audio_first:
        return "\n".join(("<audio>", prompt))
    return "\n".join((prompt, "<audio>"))


def webdataset_element_to_conversation(src):
    """One (json.response[], audio_encoding.pyd) element -> one dict per question/answer pair; the audio token is
    placed before or after the question by a coin flip of Python's `random` (seed it in tests, Appendix D.6)."""
    for elem in src:
        if "json" not in elem or not elem["json"].get("response"):
            logging.warning(f"no valid json response for {elem['__key__']}; skipping")
            continue
        if not isinstance(elem["json"]["response"], list):
            logging.warning(f"invalid response for {elem['__key__']} (could be empty); skipping")
            continue
        try:
            audio_encoding = elem["audio_encoding.pyd"]
            audio_encoding_shape = list(audio_encoding.shape)
        except AttributeError:
            logging.warning(f"error reading encoding for {elem['__key__']}, potentially the file is corrupted; "
                            "skipping")
            continue
        for response in elem["json"]["response"]:
            question, answer = (response["question"], response["answer"])
            audio_first = random.uniform(0.0, 1.0) > 0.5
            prompt_text = concat_audio_token_and_prompt(question, audio_first)
            yield {"audio_encoding": audio_encoding, "audio_encoding_shape": audio_encoding_shape,
                   "id": elem["__key__"],
                   "conversations": [{"from": "human", "value": prompt_text}, {"from": "gpt", "value": answer}]}


class TarShardDataset(torch.utils.data.IterableDataset):
    """Reader of the reference's webdataset shards with the standard library only: consecutive tar members
    `<key>.audio_encoding.pyd` (pickled array/tensor) and `<key>.json` form one element. Shards are split across
    ranks (wds.split_by_node) and dataloader workers; corrupt members are skipped with a warning
    (wds.warn_and_continue)."""

    def __init__(self, urls, multimodal_cfg, tokenizer, is_train, rank=0, world=1, seed=0, rsample_frac=None,
                 task_sample_probs=None, shuffle_buffer=100):
        self.urls = list(urls)
        self.cfg, self.tokenizer, self.is_train = multimodal_cfg, tokenizer, is_train
        self.rank, self.world, self.seed = rank, world, seed
        self.rsample_frac, self.shuffle_buffer = rsample_frac, shuffle_buffer
        # shard weights of the training stream (repeat_shards, m2t/data_modules.py:441-463): the first
        # task_sample_probs key found in a shard's name; a shard no key matches is a ValueError
        self.weights = None
        if is_train and task_sample_probs is not None:
            logging.warning(f"applying task probs {task_sample_probs}")
            self.weights = []
            for shard in self.urls:
                hit = next((p for key, p in task_sample_probs.items() if key in shard), None)
                if hit is None:
                    raise ValueError(f"probability for shard {shard} not defined in probs {task_sample_probs}")
                self.weights.append(hit)
        if not self.urls:
            raise ValueError("TarShardDataset needs at least one shard")

    def _elements(self, path):
        cur_key, cur = None, {}
        with tarfile.open(path) as tf:
            for m in tf:
                if not m.isfile():
                    continue
                key, _, ext = m.name.partition(".")
                if key != cur_key and cur:
                    yield cur
                    cur = {}
                cur_key = key
                cur["__key__"] = key
                try:
                    data = tf.extractfile(m).read()
                    if ext == "json":
                        cur["json"] = json.loads(data)
                    elif ext.endswith("pyd"):
                        cur[ext] = pickle.load(io.BytesIO(data))
                except Exception as e:  # noqa: BLE001
                    logging.warning(f"skipping corrupt member {m.name}: {e!r}")
            if cur:
                yield cur

    def _shard_stream(self, rng, wid, nw):
        """Training (wds `resampled=True` over the repeat_shards list, m2t/data_modules.py:484-498): EVERY rank and
        worker draws shards with replacement from the full list (weighted by task_sample_probs), each from its own
        seeded generator, so no rank starves when there are fewer shards than ranks x workers. Evaluation: one pass over
        the shards of this rank / worker (wds.split_by_node + split_by_worker)."""
        if self.is_train:
            while True:
                yield rng.choices(self.urls, weights=self.weights, k=1)[0]
        else:
            yield from self.urls[self.rank::self.world][wid::nw]

    def __iter__(self):
        info = torch.utils.data.get_worker_info()
        wid, nw = (info.id, info.num_workers) if info else (0, 1)
        rng = random.Random(self.seed + 1000 * self.rank + wid)
        mm = partial(preprocess_multimodal_mappable, multimodal_cfg=self.cfg)
        lm = partial(preprocess_for_lm_mappable, tokenizer=self.tokenizer)

        def conversations():
            for path in self._shard_stream(rng, wid, nw):
                yield from webdataset_element_to_conversation(self._elements(path))

        def shuffled(it):
            # wds `.shuffle(100)` (:510-511): a 100-element buffer between the per-clip Q/A expansion and tokenisation,
            # so the Q/A pairs of one clip do not arrive back to back; applied when training or sub-sampling
            if not (self.is_train or self.rsample_frac is not None) or self.shuffle_buffer <= 1:
                yield from it
                return
            buf = []
            for e in it:
                if len(buf) < self.shuffle_buffer:
                    buf.append(e)
                    continue
                k = rng.randrange(self.shuffle_buffer)
                yield buf[k]
                buf[k] = e
            rng.shuffle(buf)
            yield from buf

        for e in shuffled(conversations()):
            if self.rsample_frac and rng.random() >= self.rsample_frac:  # wds `.rsample(p)` (:512-513)
                continue
            yield lm(mm(e))


def expand_url_to_file_list(url: str) -> List[str]:
    """Comma-separated list with {a..b} brace ranges (the subset of braceexpand the reference's paths use)."""
    import re

    out = []
    for part in url.split(","):
        m = re.search(r"\{(\d+)\.\.(\d+)\}", part)
        if not m:
            out.append(part)
            continue
        a, b = m.group(1), m.group(2)
        width = len(a) if a.startswith("0") else 0
        for i in range(int(a), int(b) + 1):
            out.append(part[: m.start()] + str(i).zfill(width) + part[m.end():])
    return out


def read_jsonl_data(file) -> List[Dict[Any, Any]]:
    """Read and parse a jsonl file line by line (m2t/data_modules.py:49-55)."""
    logging.debug(f"reading {file}")
    with open(file, "r") as f:
        return [json.loads(line) for line in f]


def _audio_placeholder(n_frames: int, with_markers: bool) -> str:
    """The token string one <audio> tag expands to: n_frames patch tokens, optionally bracketed by start / end."""
    body = DEFAULT_AUDIO_PATCH_TOKEN * n_frames
    return DEFAULT_AUDIO_START_TOKEN + body + DEFAULT_AUDIO_END_TOKEN if with_markers else body


def preprocess_multimodal(sources: Sequence[Sequence[Dict[str, str]]], multimodal_cfg: dict,
                          encoded_shapes: Sequence[Sequence[int]]):
    """List-of-sources form of the <audio> expansion (m2t/data_modules.py:112-143): every sentence of conversation i has
    its <audio> tag replaced by the placeholder for encoded_shapes[i] = (frames, width); with sep_audio_conv_front the
    tag is first moved to the front of the opening turn as "<audio>###Human: ...". Mutates and returns `sources`."""
    if not multimodal_cfg["is_multimodal"]:
        return sources
    with_markers = multimodal_cfg["use_audio_start_end"]
    for (n_frames, _width), conversation in zip(encoded_shapes, sources):
        if multimodal_cfg["sep_audio_conv_front"]:
            opening = conversation[0]
            assert DEFAULT_AUDIO_TOKEN in opening["value"]
            text = opening["value"].replace(DEFAULT_AUDIO_TOKEN, "").strip()
            opening["value"] = f"{DEFAULT_AUDIO_TOKEN}{CONVERSATION_SEP}{CONVERSATION_ROLES[0]}: {text}"
        placeholder = _audio_placeholder(n_frames, with_markers)
        for sentence in conversation:
            sentence["value"] = sentence["value"].replace(DEFAULT_AUDIO_TOKEN, placeholder)
    return sources


def preprocess_for_lm(sources: Sequence[Sequence[Dict[str, str]]], tokenizer) -> Dict:
    """List-of-sources form of tokenise + mask (m2t/data_modules.py:146-177): all conversations are tokenised in one
    padded call; labels start as a copy and each row is masked with segment lengths measured by tokenising the header and
    every turn on their own (the off-by-BOS quirk lives in _mask_targets)."""
    header = DEFAULT_CONVERSATION_HEADER
    texts = [sentences_to_formatted_conversation(header, conversation) for conversation in sources]
    input_ids = _tokenize_fn(texts, tokenizer)["input_ids"]
    labels = copy.deepcopy(input_ids)
    for row, conversation in zip(labels, sources):
        pieces = [header] + [turn["value"] for turn in conversation]
        seg_lens = _tokenize_fn(pieces, tokenizer)["input_ids_lens"]
        _mask_targets(row, seg_lens, [turn["from"] for turn in conversation])
    return dict(input_ids=input_ids, labels=labels)


def assert_hf_src_format(src):
    """A batch in the HF layout is a dict of equal-length lists; element i of the batch is {k: src[k][i]}
    (m2t/data_modules.py:343-365). Raises AssertionError otherwise."""
    assert isinstance(src, dict)
    names = list(src.keys())
    kinds = [(k, type(src[k])) for k in names]
    assert all(kind is list for _, kind in kinds), f"expected dict of lists, got: {kinds}"
    sizes = [(k, len(src[k])) for k in names]
    assert all(n == sizes[0][1] for _, n in sizes), f"expected lists of same length, got {sizes}"


def _qa_pairs(elem_json, key):
    """The usable (question, answer) pairs of one element, or None (with the reference's warning) if it has none."""
    responses = elem_json.get("response")
    if not responses:
        logging.warning(f"no response for {key}; skipping")
        return None
    if not isinstance(responses, list):
        logging.warning(f"invalid response for {key} (could be empty); skipping")
        return None
    return [(r["question"], r["answer"]) for r in responses]


def _to_conversation_hf(src: Dict[str, List]) -> Dict[str, List]:
    """Batched flat-map of the HF pipeline (m2t/data_modules.py:368-427): every (question, answer) of every element
    becomes its own element carrying the flat audio encoding, its shape, the id and the element's json minus the bulky
    `response`; one random.uniform draw per pair decides whether <audio> precedes or follows the question."""
    assert_hf_src_format(src)
    out = defaultdict(list)
    for key, elem_json, encoding, shape in zip(src["__key__"], src["json"], src["audio_encoding"],
                                               src["audio_encoding_shape"]):
        pairs = _qa_pairs(elem_json, key)
        if pairs is None:
            continue
        meta = {k: v for k, v in elem_json.items() if k != "response"}
        for question, answer in pairs:
            audio_first = random.uniform(0.0, 1.0) > 0.5
            prompt = concat_audio_token_and_prompt(question, audio_first)
            fields = (("audio_encoding", encoding), ("audio_encoding_shape", shape), ("__key__", key), ("id", key),
                      ("json", meta),
                      ("conversations", [{"from": "human", "value": prompt}, {"from": "gpt", "value": answer}]))
            for name, value in fields:
                out[name].append(value)
    return out


def maybe_add_gcs_prefix(f: str) -> str:
    """`gs://` objects are read through a gsutil pipe by webdataset (m2t/data_modules.py:430-433); only the string
    transformation is mirrored, GCS I/O itself is out of scope."""
    return f"pipe:gsutil cat {f}" if f.startswith("gs://") else f


def repeat_shards(urls: List[str], task_sample_probs: Optional[Dict[str, float]] = None):
    """HF Trainer needs an explicitly repeated shard list for an IterableDataset: 1024 x len(urls) draws with
    replacement through np.random.choice, uniform or weighted by the first task_sample_probs key found in each shard's
    name (m2t/data_modules.py:441-463; a shard no key matches is a ValueError)."""
    weights = None
    if task_sample_probs is not None:
        logging.warning(f"applying task probs {task_sample_probs}")
        raw = []
        for shard in urls:
            hit = next((p for key, p in task_sample_probs.items() if key in shard), None)
            if hit is None:
                raise ValueError(f"probability for shard {shard} not defined in probs {task_sample_probs}")
            raw.append(hit)
        weights = np.array(raw) / sum(raw)
    return np.random.choice(urls, size=1024 * len(urls), replace=True, p=weights).tolist()


def _msgpack_decode(data: bytes):
    try:
        import msgspec

        return msgspec.msgpack.decode(data)
    except ImportError:
        import msgpack

        return msgpack.unpackb(data, raw=False)


def gen_from_msgpack_shards(shards):
    """Examples of msgpack shards, one shard = ONE msgpack list (m2t/data_modules.py:599-615). A local path that does
    not exist or bytes that do not decode are skipped with a warning; anything but a list is an AssertionError."""
    for shard in shards:
        if "gs://" not in shard and not os.path.exists(shard):
            logging.warning(f"skipping nonexistent shard {shard}")
            continue
        try:
            with open(shard, "rb") as fh:
                examples = _msgpack_decode(fh.read())
        except Exception as e:  # noqa: BLE001
            logging.warning(f"error decoding data from shard {shard}: {e}; skipping.")
            continue
        assert isinstance(examples, list), f"expected list after decoding {shard}; got {type(examples)}"
        yield from examples


def hf_preprocess_encodings(src: Dict[str, Any]) -> Dict[str, Any]:
    """Reshape the flat encoding in place (m2t/data_modules.py:618-622)."""
    src["audio_encoding"] = preprocess_encodings(src["audio_encoding"], src["audio_encoding_shape"])
    return src


_SHUFFLE_BUFFER_SIZE = 1000  # the HF IterableDataset default the reference relies on (:523)


class HFStyleShardDataset(torch.utils.data.IterableDataset):
    """read_hf_dataset (m2t/data_modules.py:625-668) without the `datasets` package: jsonl / json or msgpack shards ->
    examples split by node (example i goes to rank i % world, what split_dataset_by_node does for a non-sharded
    stream) -> `_to_conversation_hf` over batches of 8 (flat-map) -> shuffle buffer of 1000 -> the <audio> expansion,
    tokenise + mask, encoding reshape. Element dicts carry the keys the collator reads."""

    def __init__(self, urls, multimodal_cfg, tokenizer, is_train, rank=0, world=1, seed=0, batch_size=8,
                 shuffle_buffer=_SHUFFLE_BUFFER_SIZE):
        self.urls = repeat_shards(list(urls)) if is_train else list(urls)
        self.cfg, self.tokenizer = multimodal_cfg, tokenizer
        self.rank, self.world, self.seed = rank, world, seed
        self.batch_size, self.shuffle_buffer = batch_size, shuffle_buffer

    def _examples(self):
        if self.urls[0].endswith(".jsonl") or self.urls[0].endswith(".json"):
            for u in self.urls:
                yield from read_jsonl_data(u)
        elif self.urls[0].endswith(".msgpack"):
            yield from gen_from_msgpack_shards(self.urls)
        else:
            raise ValueError(f"unsupported shard type {self.urls[0]} (expected .jsonl, .json or .msgpack)")

    def _flat(self):
        batch = defaultdict(list)
        n = 0
        for i, ex in enumerate(self._examples()):
            if i % self.world != self.rank:
                continue
            for k, v in ex.items():
                batch[k].append(v)
            n += 1
            if n == self.batch_size:
                out = _to_conversation_hf(dict(batch))
                for j in range(len(out.get("__key__", []))):
                    yield {k: v[j] for k, v in out.items()}
                batch, n = defaultdict(list), 0
        if n:
            out = _to_conversation_hf(dict(batch))
            for j in range(len(out.get("__key__", []))):
                yield {k: v[j] for k, v in out.items()}

    def __iter__(self):
        rng = random.Random(self.seed)
        mm = partial(preprocess_multimodal_mappable, multimodal_cfg=self.cfg)
        lm = partial(preprocess_for_lm_mappable, tokenizer=self.tokenizer)

        def finish(e):
            # datasets.map(batched=False) MERGES the returned dict into the example, which is why the reference can
            # still find `audio_encoding_shape` after preprocess_for_lm_mappable dropped it
            e = {**e, **mm(e)}
            e = {**e, **lm(e)}
            return hf_preprocess_encodings(e)

        buf = []
        for e in self._flat():
            if len(buf) < self.shuffle_buffer:
                buf.append(e)
                continue
            k = rng.randrange(self.shuffle_buffer)
            yield finish(buf[k])
            buf[k] = e
        rng.shuffle(buf)
        for e in buf:
            yield finish(e)


def read_hf_dataset(url: str, multimodal_cfg: Dict[str, Any], tokenizer, is_train: bool, rsample_frac=None):
    """m2t/data_modules.py:625-668: RANK / WORLD_SIZE come from the environment as in the reference."""
    return HFStyleShardDataset(expand_url_to_file_list(url), multimodal_cfg, tokenizer, is_train,
                               rank=int(os.environ["RANK"]), world=int(os.environ["WORLD_SIZE"]))


def make_mm_config(data_args):
    return dict(is_multimodal=data_args.is_multimodal, sep_audio_conv_front=data_args.sep_audio_conv_front,
                audio_folder=data_args.audio_folder,
                use_audio_start_end=getattr(data_args, "mm_use_audio_start_end", False),
                audio_processor=getattr(data_args, "audio_encoding_processor", None))


def make_data_module(tokenizer, data_args, data_collator_cls=DataCollatorForSupervisedDataset) -> Dict:
    """Dataset + collator for supervised fine-tuning (m2t/data_modules.py:681-717)."""
    import os

    multimodal_cfg = make_mm_config(data_args)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    train_dataset = eval_dataset = None
    if data_args.train_data_path:
        probs = (getattr(data_args, "task_sample_probs", None)
                 if getattr(data_args, "apply_task_sample_probs", False) else None)
        train_dataset = TarShardDataset(expand_url_to_file_list(data_args.train_data_path), multimodal_cfg, tokenizer,
                                        True, rank, world, task_sample_probs=probs)
    if data_args.eval_data_path:
        eval_dataset = TarShardDataset(expand_url_to_file_list(data_args.eval_data_path), multimodal_cfg, tokenizer,
                                       False, rank, world,
                                       rsample_frac=getattr(data_args, "eval_data_subsample", None))
    return dict(train_dataset=train_dataset, eval_dataset=eval_dataset, data_collator=data_collator_cls(tokenizer=tokenizer))
