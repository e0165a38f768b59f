"""Model-side mirror of m2t/models/__init__.py:23-29 (AudioEncoderConfig)."""
from dataclasses import dataclass

from llark_b200.special_tokens import DEFAULT_AUDIO_END_TOKEN, DEFAULT_AUDIO_PATCH_TOKEN, DEFAULT_AUDIO_START_TOKEN


@dataclass
class AudioEncoderConfig:
    """Carrier of the audio special tokens. As in the reference the fields start out as the token STRINGS and are
    overwritten with token ids by initialize_audio_tokenizer (m2t/models/llamav2.py:386-419)."""

    use_audio_start_end: bool = True
    audio_start_token: object = DEFAULT_AUDIO_START_TOKEN
    audio_end_token: object = DEFAULT_AUDIO_END_TOKEN
    audio_patch_token: object = DEFAULT_AUDIO_PATCH_TOKEN
