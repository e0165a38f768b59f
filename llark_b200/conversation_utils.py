"""Mirror of m2t/conversation_utils.py:17-55 (example records and the prompt / response split of a token sequence)."""
from typing import Any, Dict, List

import numpy as np


def make_example(id: str, audio: str, audio_encoding: np.ndarray, audio_encoding_shape: List[int], prompt_question: str,
                 response: str) -> Dict[str, Any]:
    return {"id": id, "audio": audio, "audio_encoding": audio_encoding, "audio_encoding_shape": audio_encoding_shape,
            "response": [{"question": prompt_question, "answer": response}]}


def subsequence_pos(ary, subary):
    """(start, end) of the first occurrence of `subary` in `ary`; None when absent. As in the reference the last
    possible start position is NOT examined (range(len(ary) - s)), so a match flush with the end is missed."""
    assert isinstance(ary, list)
    assert isinstance(subary, list)
    s = len(subary)
    for start_idx in range(len(ary) - s):
        if ary[start_idx: start_idx + s] == subary:
            return start_idx, start_idx + s


def extract_prompt_tokens(input_ids, end_seq):
    """The input_ids of the prefix, up to and including `end_seq` (i.e. before the model's response)."""
    _, prompt_end_idx = subsequence_pos(input_ids.tolist(), end_seq)
    return input_ids[:prompt_end_idx]


def extract_response_tokens(input_ids, end_seq):
    """The input_ids of the model response (everything after `end_seq`)."""
    _, prompt_end_idx = subsequence_pos(input_ids.tolist(), end_seq)
    return input_ids[prompt_end_idx:]
