"""Mirror of m2t/infer.py:99-152: one prompt + one audio encoding -> generated token ids."""
from typing import Any, Dict, Optional, Sequence

import torch

from llark_b200.conversation_utils import extract_prompt_tokens
from llark_b200.data_modules import (DEFAULT_CONVERSATION_HEADER, concat_audio_token_and_prompt,
                                     preprocess_for_lm_mappable, preprocess_multimodal_mappable)
from llark_b200.generate import KeywordsStoppingCriteria


def build_inference_inputs(prompt_text: str, audio_encoding, end_seq: Sequence[int], multimodal_cfg: Dict[str, Any],
                           tokenizer, example_id: Optional[str] = None, audio_first: bool = False,
                           header: str = DEFAULT_CONVERSATION_HEADER):
    """Everything infer_with_prompt does before `model.generate` (same preprocessing as training, the "<empty>"
    assistant turn cut off at `end_seq`): returns (elem, input_ids [1, t], audio_encoding [1, F, A], stopping_criteria)."""
    prompt_text = concat_audio_token_and_prompt(prompt_text, audio_first)
    elem = {"audio_encoding": audio_encoding, "audio_encoding_shape": list(audio_encoding.shape),
            "example_id": example_id, "id": example_id,
            "conversations": [{"from": "human", "value": prompt_text}, {"from": "gpt", "value": "<empty>"}]}
    elem = preprocess_multimodal_mappable(elem, multimodal_cfg)
    elem = preprocess_for_lm_mappable(elem, tokenizer=tokenizer, header=header)
    audio_encoding = elem.pop("audio_encoding")
    if len(audio_encoding.shape) < 3:
        audio_encoding = torch.unsqueeze(audio_encoding, 0)
    input_ids = elem.pop("input_ids")
    input_ids = extract_prompt_tokens(input_ids, end_seq)
    if len(input_ids.shape) < 2:
        input_ids = torch.unsqueeze(input_ids, 0)
    stopping_criteria = KeywordsStoppingCriteria(keywords=["###"], tokenizer=tokenizer, input_ids=input_ids)
    return elem, input_ids, audio_encoding, stopping_criteria


def infer_with_prompt(prompt_text: str, model, audio_encoding, end_seq: Sequence[int], multimodal_cfg: Dict[str, Any],
                      tokenizer, example_id: Optional[str] = None, audio_first: bool = False,
                      header: str = DEFAULT_CONVERSATION_HEADER, **generation_kwargs):
    elem, input_ids, audio_encoding, stopping_criteria = build_inference_inputs(
        prompt_text, audio_encoding, end_seq, multimodal_cfg, tokenizer, example_id, audio_first, header)
    return model.generate(**elem, **generation_kwargs, input_ids=input_ids.cuda(),
                          audio_encodings=audio_encoding.cuda(), stopping_criteria=[stopping_criteria])
