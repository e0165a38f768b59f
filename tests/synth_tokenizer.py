"""Deterministic stand-in for the slow Llama tokenizer (no SentencePiece model exists offline). It implements
the HF calling convention the reference's data pipeline uses -- tokenizer(text, return_tensors="pt",
padding="longest", max_length=, truncation=True).input_ids, .pad_token_id, .model_max_length, add_tokens,
convert_tokens_to_ids, len() -- prepends BOS to every call (which is what produces the reference's off-by-BOS
masking quirk) and maps special tokens to single ids. Used by BOTH the reference functions (when the golden
vectors are generated) and the mirror, so parity of the index work is independent of the tokenizer's vocabulary."""
import re
import zlib
from types import SimpleNamespace

import torch


class SynthTokenizer:
    def __init__(self, vocab_size=32000, model_max_length=2048):
        self.base_vocab = vocab_size
        self.model_max_length = model_max_length
        self.bos_token_id, self.eos_token_id, self.unk_token_id = 1, 2, 0
        self.special = {}
        self.pieces = {}
        self.pad_token_id = None
        self.padding_side = "right"

    def __len__(self):
        return self.base_vocab + len(self.special)

    def add_tokens(self, tokens, special_tokens=False):
        n = 0
        for t in tokens:
            if t not in self.special:
                self.special[t] = self.base_vocab + len(self.special)
                n += 1
        return n

    def add_special_tokens(self, d):
        n = 0
        for k, t in d.items():
            n += self.add_tokens([t], special_tokens=True)
            if k == "pad_token":
                self.pad_token_id = self.special[t]
        return n

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self.special.get(tokens, self.unk_token_id)
        return [self.special.get(t, self.unk_token_id) for t in tokens]

    def _encode(self, text):
        ids = [self.bos_token_id]
        if self.special:
            pat = "(" + "|".join(re.escape(t) for t in sorted(self.special, key=len, reverse=True)) + ")"
            parts = re.split(pat, text)
        else:
            parts = [text]
        for part in parts:
            if part in self.special:
                ids.append(self.special[part])
                continue
            for piece in re.findall(r"\w+|[^\w\s]|\s+", part):
                tid = 3 + zlib.crc32(piece.encode("utf-8")) % (self.base_vocab - 3)
                self.pieces.setdefault(tid, piece)  # reverse vocabulary for decode(): first piece seen per id
                ids.append(tid)
        return ids

    def decode(self, ids, skip_special_tokens=False):
        """Inverse of _encode for ids this instance has produced (unknown ids render as <id>); BOS/EOS/UNK and the
        added special tokens are dropped when skip_special_tokens is set, as the HF tokenizers do."""
        rev = {v: k for k, v in self.special.items()}
        out = []
        for t in (ids.tolist() if hasattr(ids, "tolist") else list(ids)):
            t = int(t)
            if t in (self.bos_token_id, self.eos_token_id, self.unk_token_id) or t in rev:
                if not skip_special_tokens:
                    out.append(rev.get(t, {self.bos_token_id: "<s>", self.eos_token_id: "</s>"}.get(t, "<unk>")))
                continue
            out.append(self.pieces.get(t, f"<{t}>"))
        return "".join(out)

    def batch_decode(self, batch, skip_special_tokens=False):
        return [self.decode(row, skip_special_tokens=skip_special_tokens) for row in batch]

    def __call__(self, text, return_tensors=None, padding=None, max_length=None, truncation=False,
                 add_special_tokens=True):
        single = isinstance(text, str)
        texts = [text] if single else list(text)
        rows = []
        for t in texts:
            ids = self._encode(t)
            if not add_special_tokens:
                ids = ids[1:]
            if truncation and max_length is not None:
                ids = ids[:max_length]
            rows.append(ids)
        width = max(len(r) for r in rows)
        pad = self.pad_token_id if self.pad_token_id is not None else 0
        if return_tensors == "pt":
            out = torch.full((len(rows), width), pad, dtype=torch.long)
            for i, r in enumerate(rows):
                out[i, : len(r)] = torch.tensor(r, dtype=torch.long)
            return SimpleNamespace(input_ids=out)
        return SimpleNamespace(input_ids=rows[0] if single else rows)
