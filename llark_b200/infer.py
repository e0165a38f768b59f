"""Single-prompt inference with the name and behaviour of m2t/infer.py:99-152 (`infer_with_prompt`): the prompt goes
through exactly the training preprocessing with a placeholder answer, is cut right after the end-of-prompt marker, and is
handed to `model.generate` together with the audio encoding and the "###" stopping criterion."""
from typing import Any, Dict, Optional, Sequence

import torch

from llark_b200 import data_modules as dm
from llark_b200.conversation_utils import extract_prompt_tokens
from llark_b200.generate import KeywordsStoppingCriteria

PLACEHOLDER_ANSWER = "<empty>"
STOP_KEYWORDS = ["###"]


def _as_batch(t: torch.Tensor, ndim: int) -> torch.Tensor:
    return t.unsqueeze(0) if t.dim() < ndim else t


def build_inference_inputs(prompt_text: str, audio_encoding, end_seq: Sequence[int], multimodal_cfg: Dict[str, Any],
                           tokenizer, example_id: Optional[str] = None, audio_first: bool = False,
                           header: str = dm.DEFAULT_CONVERSATION_HEADER):
    """-> (leftover element fields, input_ids [1, t], audio_encodings [1, F, A], stopping criterion)."""
    turns = [{"from": "human", "value": dm.concat_audio_token_and_prompt(prompt_text, audio_first)},
             {"from": "gpt", "value": PLACEHOLDER_ANSWER}]
    elem = dict(audio_encoding=audio_encoding, audio_encoding_shape=list(audio_encoding.shape), example_id=example_id,
                id=example_id, conversations=turns)
    elem = dm.preprocess_for_lm_mappable(dm.preprocess_multimodal_mappable(elem, multimodal_cfg), tokenizer=tokenizer,
                                         header=header)
    encodings = _as_batch(elem.pop("audio_encoding"), 3)
    prompt_ids = _as_batch(extract_prompt_tokens(elem.pop("input_ids"), end_seq), 2)
    stop = KeywordsStoppingCriteria(keywords=list(STOP_KEYWORDS), tokenizer=tokenizer, input_ids=prompt_ids)
    return elem, prompt_ids, encodings, stop


def infer_with_prompt(prompt_text: str, model, audio_encoding, end_seq: Sequence[int], multimodal_cfg: Dict[str, Any],
                      tokenizer, example_id: Optional[str] = None, audio_first: bool = False,
                      header: str = dm.DEFAULT_CONVERSATION_HEADER, **generation_kwargs):
    rest, prompt_ids, encodings, stop = build_inference_inputs(prompt_text, audio_encoding, end_seq, multimodal_cfg,
                                                               tokenizer, example_id, audio_first, header)
    # the leftover fields (labels, example_id) ride along as keyword arguments, as upstream passes `**elem`
    return model.generate(**rest, **generation_kwargs, input_ids=prompt_ids.cuda(), audio_encodings=encodings.cuda(),
                          stopping_criteria=[stop])
