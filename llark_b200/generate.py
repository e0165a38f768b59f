"""Generation plumbing of the inference path: mirror of m2t/generate.py plus the decoding loop
behind `WrappedLlamav2ForCausalLM.generate`.

    KeywordsStoppingCriteria    m2t/generate.py:18-44 (same constructor, same __call__ contract)
    greedy_generate             the greedy branch of transformers' GenerationMixin.generate as m2t/infer.py:99-152 uses it
                                (batch of one prompt, stopping criteria, eos), run WITHOUT a KV cache: every step is one
                                full forward of the fused engine over the sequence so far. The KV-cache decode kernels are
                                the remaining part of this row; results are identical, cost is O(n^2).
"""
from typing import Callable, Optional, Sequence

import torch
import transformers


class KeywordsStoppingCriteria(transformers.StoppingCriteria):
    """Stop as soon as one of `keywords` shows up in the text generated after the prompt (`input_ids` is the prompt;
    its length is latched on the first call). Keywords that the tokenizer maps to exactly one id are also recognised
    directly on the newest token -- with a BOS-prepending tokenizer such as Llama's no keyword qualifies."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        self.tokenizer = tokenizer
        self.input_ids = input_ids
        self.start_len = None
        encoded = (tokenizer(word).input_ids for word in keywords)
        self.keyword_ids = [ids[0] for ids in encoded if isinstance(ids, list) and len(ids) == 1]

    def _newest_token_is_keyword(self, output_ids) -> bool:
        newest = output_ids[0, -1]
        return any(newest == kid for kid in self.keyword_ids)

    def _generated_text(self, output_ids) -> str:
        return self.tokenizer.batch_decode(output_ids[:, self.start_len:], skip_special_tokens=True)[0]

    def __call__(self, output_ids: torch.LongTensor, scores: torch.FloatTensor, **kwargs) -> bool:
        if self.start_len is None:
            self.start_len = self.input_ids.shape[1]
        if self._newest_token_is_keyword(output_ids):
            return True
        text = self._generated_text(output_ids)
        return any(word in text for word in self.keywords)


@torch.no_grad()
def greedy_generate(step_logits: Callable[[torch.Tensor], torch.Tensor], input_ids: torch.Tensor,
                    max_new_tokens: int, eos_token_id: Optional[Sequence[int]] = None,
                    stopping_criteria: Sequence[Callable] = (), pad_token_id: Optional[int] = None) -> torch.Tensor:
    """Greedy decoding as GenerationMixin does it for a decoder-only model: `step_logits(ids)` returns the logits of
    the LAST position for every row of `ids` [B, t]; argmax is appended; a row is finished when it emitted an eos id;
    finished rows are padded with pad_token_id; the loop ends when every row is finished, a stopping criterion fires
    (evaluated on the extended sequence, scores = the step's logits) or max_new_tokens is reached. Returns the full
    sequences [B, prompt + new]."""
    if input_ids.dim() != 2:
        raise ValueError("input_ids must be [batch, seq]")
    eos = set(int(e) for e in (eos_token_id or ()))
    if eos and pad_token_id is None:
        pad_token_id = min(eos)  # transformers: "Setting pad_token_id to eos_token_id"
    unfinished = torch.ones(input_ids.shape[0], dtype=torch.bool, device=input_ids.device)
    for _ in range(max(int(max_new_tokens), 0)):
        logits = step_logits(input_ids)
        next_tokens = logits.float().argmax(dim=-1)
        if eos:
            next_tokens = torch.where(unfinished, next_tokens, torch.full_like(next_tokens, pad_token_id))
        input_ids = torch.cat([input_ids, next_tokens[:, None].to(input_ids.dtype)], dim=1)
        if eos:
            hit = torch.zeros_like(unfinished)
            for e in eos:
                hit |= next_tokens == e
            unfinished = unfinished & ~hit
        stop = any(bool(torch.as_tensor(c(input_ids, logits)).all()) for c in stopping_criteria)
        if stop or not bool(unfinished.any()):
            break
    return input_ids


@torch.no_grad()
def generate_with_model(model, step_logits: Callable[[torch.Tensor], torch.Tensor], inputs=None, input_ids=None,
                        stopping_criteria=None, max_new_tokens=None, max_length=None, do_sample=None, num_beams=None,
                        eos_token_id=None, pad_token_id=None) -> torch.Tensor:
    """The argument handling of `model.generate(...)` shared by the Llama and MPT wrappers: resolves the prompt, the
    token budget (max_new_tokens, else max_length / generation_config.max_length / 20 as transformers) and the eos /
    pad ids (argument, generation_config, config), switches the module to eval for the duration, and runs
    greedy_generate with the caller's `step_logits`. Sampling and beam search are refused."""
    if input_ids is None:
        input_ids = inputs
    if input_ids is None:
        raise ValueError("generate needs input_ids")
    if do_sample or (num_beams not in (None, 1)):
        raise NotImplementedError("llark_b200 generates greedily (do_sample=False, num_beams=1)")
    gc = getattr(model, "generation_config", None)
    if max_new_tokens is None:
        total = max_length if max_length is not None else (getattr(gc, "max_length", None) or 20)
        max_new_tokens = max(int(total) - input_ids.shape[1], 0)
    if eos_token_id is None:
        eos_token_id = getattr(gc, "eos_token_id", None)
        if eos_token_id is None:
            eos_token_id = getattr(model.config, "eos_token_id", None)
    if isinstance(eos_token_id, int):
        eos_token_id = [eos_token_id]
    if pad_token_id is None:
        pad_token_id = getattr(gc, "pad_token_id", None)
    if pad_token_id is None:
        pad_token_id = getattr(model.config, "pad_token_id", None)
    was_training = model.training
    model.eval()
    try:
        return greedy_generate(step_logits, input_ids, max_new_tokens, eos_token_id=eos_token_id,
                               stopping_criteria=list(stopping_criteria or []), pad_token_id=pad_token_id)
    finally:
        model.train(was_training)

