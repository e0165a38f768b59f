"""Mirror of m2t/tokenizer.py:21-58."""
from typing import List


def get_tokenizer(model_args, training_args):
    import transformers

    kw = dict(cache_dir=training_args.cache_dir, model_max_length=training_args.model_max_length, padding_side="right")
    if "mpt" not in model_args.model_name_or_path:
        kw["use_fast"] = False
    return transformers.AutoTokenizer.from_pretrained(model_args.model_name_or_path, **kw)


def get_prompt_end_token_sequence(tokenizer, model_name: str, prompt_end_string="\n### Assistant:") -> List[int]:
    """The token sequence that ends the prompt (and starts the model's generation); used to split sequences into
    (prompt, response). For Llama-2 the first token is dropped: the tokenizer puts a spurious prefix token in front of
    the leading newline which does not occur when the same text is tokenised inside a longer sequence."""
    end_seq = tokenizer([prompt_end_string], add_special_tokens=False).input_ids[0]
    if "meta-llama/Llama-2" in model_name:
        end_seq = end_seq[1:]
    return end_seq
