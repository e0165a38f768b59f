"""Example records and the prompt / response split of a token sequence, with the names and behaviour of
m2t/conversation_utils.py:17-55 (pinned by tests/test_generation_cpu.py against the reference's functions)."""
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


def make_example(id: str, audio: str, audio_encoding: np.ndarray, audio_encoding_shape: List[int], prompt_question: str,
                 response: str) -> Dict[str, Any]:
    """One dataset element in the `response: [{question, answer}]` layout the data pipeline consumes."""
    qa = {"question": prompt_question, "answer": response}
    return dict(id=id, audio=audio, audio_encoding=audio_encoding, audio_encoding_shape=audio_encoding_shape,
                response=[qa])


def subsequence_pos(ary, subary) -> Optional[Tuple[int, int]]:
    """(start, end) of the first window of `ary` equal to `subary`, None when there is none. Upstream never looks at the
    LAST window (it iterates range(len(ary) - len(subary))), so a match flush with the end is not found; kept."""
    assert isinstance(ary, list)
    assert isinstance(subary, list)
    width = len(subary)
    n_windows = len(ary) - width  # one fewer than exist
    start = 0
    while start < n_windows:
        if ary[start: start + width] == subary:
            return start, start + width
        start += 1
    return None


def _prompt_end(input_ids, end_seq) -> int:
    found = subsequence_pos(input_ids.tolist(), end_seq)
    # upstream unpacks the result directly, so a missing end sequence surfaces as the same TypeError
    _, stop = found
    return stop


def extract_prompt_tokens(input_ids, end_seq):
    """Everything up to and including `end_seq`: the prefix the model is conditioned on."""
    return input_ids[: _prompt_end(input_ids, end_seq)]


def extract_response_tokens(input_ids, end_seq):
    """Everything after `end_seq`: the model's answer."""
    return input_ids[_prompt_end(input_ids, end_seq):]
