"""Tokenizer construction and the end-of-prompt marker, with the names and behaviour of m2t/tokenizer.py:21-58."""
from typing import List

PROMPT_END = "\n### Assistant:"


def get_tokenizer(model_args, training_args):
    """AutoTokenizer for the backbone: right padding, the run's max length; the slow tokenizer for everything but MPT."""
    import transformers

    name = model_args.model_name_or_path
    options = {"cache_dir": training_args.cache_dir, "model_max_length": training_args.model_max_length,
               "padding_side": "right"}
    if "mpt" not in name:
        options["use_fast"] = False
    return transformers.AutoTokenizer.from_pretrained(name, **options)


def get_prompt_end_token_sequence(tokenizer, model_name: str, prompt_end_string=PROMPT_END) -> List[int]:
    """Token ids of the text that closes the prompt and opens the model's turn; sequences are cut into (prompt, response)
    at it. Llama-2's tokenizer emits a spurious prefix piece in front of the leading newline when the string is tokenised
    on its own (it does not appear inside a longer sequence), so that first id is dropped for Llama-2 models."""
    ids = tokenizer([prompt_end_string], add_special_tokens=False).input_ids[0]
    return ids[1:] if "meta-llama/Llama-2" in model_name else ids
