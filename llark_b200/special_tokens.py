"""The token strings of the LLark prompt format and the label value the loss ignores; the names are part of the API
surface (`from m2t.special_tokens import ...`, m2t/special_tokens.py:17-25)."""
IGNORE_INDEX = -100  # torch.nn.CrossEntropyLoss ignore_index

# tokenizer-level specials of the Llama vocabulary (+ the pad token train.py adds)
DEFAULT_PAD_TOKEN, DEFAULT_BOS_TOKEN, DEFAULT_EOS_TOKEN, DEFAULT_UNK_TOKEN = "[PAD]", "<s>", "</s>", "<unk>"

# audio markup: the tag users write, and the three tokens it expands to (start, one patch per frame, end)
DEFAULT_AUDIO_TOKEN = "<audio>"
DEFAULT_AUDIO_START_TOKEN, DEFAULT_AUDIO_PATCH_TOKEN, DEFAULT_AUDIO_END_TOKEN = (
    "<audio_start>", "<audio_patch>", "<audio_end>")
