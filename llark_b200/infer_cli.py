"""Batch inference from the command line, for use without a reference checkout (the reference's own
scripts/inference/infer_from_webdataset.py and infer_from_encodings.py also run unchanged over llark_b200's modules,
INTEGRATION.md §2). One entry point, two sources:

    python -m llark_b200.infer_cli --model_name_or_path <run dir> --ckpt-num 100000 --bf16 True \\
        --eval_data_path 'eval-{000000..000021}.tar' [--prompt "Describe the audio."] --outfile results.csv
    python -m llark_b200.infer_cli --model_name_or_path <run dir> --ckpt-num 100000 --bf16 True \\
        --audio-encodings-dir embeddings/ --prompt "What genre is this song?" --outfile results.csv

Shards (`--eval_data_path`): each element's own prompt is completed, or `--prompt` replaces it; the CSV carries the
ground-truth completion as well (scripts/inference/infer_from_webdataset.py:52-151). Folder of `.npy` embeddings as the
Jukebox front end writes them (`--audio-encodings-dir`): `--prompt` is required (infer_from_encodings.py:48-121). Rows,
column names and the decoding recipe (greedy, stop at "###", completion = tokens after the end-of-prompt sequence) are the
reference's. So is its reading of `--max-samples N`: the loop stops AFTER element index N, i.e. N + 1 rows.
"""
import glob
import os
from typing import Iterator, Optional, Tuple

import torch

from llark_b200.arguments import DataArguments, ModelArguments, TrainingArguments
from llark_b200.conversation_utils import extract_prompt_tokens, extract_response_tokens
from llark_b200.data_modules import make_data_module, make_mm_config
from llark_b200.generate import KeywordsStoppingCriteria
from llark_b200.infer import infer_with_prompt
from llark_b200.models.utils import load_pretrained_model
from llark_b200.tokenizer import get_prompt_end_token_sequence
from llark_b200.utils import get_autocast_type


def _complete_own_prompt(model, tokenizer, example, end_seq, max_new_tokens) -> Tuple[str, torch.Tensor]:
    """The element carries a full conversation: keep its tokens up to the end-of-prompt sequence and let the model go on."""
    prompt_ids = extract_prompt_tokens(example["input_ids"], end_seq).unsqueeze(0).cuda()
    stop = KeywordsStoppingCriteria(keywords=["###"], tokenizer=tokenizer, input_ids=prompt_ids)
    out = model.generate(input_ids=prompt_ids, audio_encodings=example["audio_encoding"].unsqueeze(0).cuda(),
                         max_new_tokens=max_new_tokens, stopping_criteria=[stop])
    return tokenizer.decode(prompt_ids[0]), out


def _shard_examples(tokenizer, data_args) -> Iterator[dict]:
    return iter(make_data_module(tokenizer=tokenizer, data_args=data_args)["eval_dataset"])


def _folder_examples(folder) -> Iterator[dict]:
    import numpy as np

    for path in glob.glob(os.path.join(folder, "*.npy")):
        yield {"example_id": path.replace(".npy", ""), "audio_encoding": np.load(path)}


def run(model_args: ModelArguments, data_args: DataArguments, training_args: TrainingArguments, outfile: str,
        ckpt_num: int, prompt: Optional[str] = None, audio_encodings_dir: Optional[str] = None,
        max_samples: Optional[int] = None, max_new_tokens: Optional[int] = None, device: str = "cuda:0"):
    """Returns the list of row dicts it also writes to `outfile`."""
    import pandas as pd

    if not data_args.is_multimodal:
        raise AssertionError("llark inference is multimodal: --is_multimodal must stay True")
    from_folder = audio_encodings_dir is not None
    if from_folder and prompt is None:
        raise ValueError("--audio-encodings-dir has no prompts of its own: pass --prompt")
    if max_new_tokens is None:
        max_new_tokens = 512 if from_folder else 2048  # the defaults of the two reference scripts
    print("loading model and data...")
    model, tokenizer = load_pretrained_model(model_args.model_name_or_path, ckpt_num=ckpt_num, device=device)
    data_args.mm_use_audio_start_end = True
    end_seq = get_prompt_end_token_sequence(tokenizer, model_args.model_name_or_path)
    examples = _folder_examples(audio_encodings_dir) if from_folder else _shard_examples(tokenizer, data_args)
    mm_cfg = make_mm_config(data_args)
    os.makedirs(os.path.dirname(outfile) or ".", exist_ok=True)

    rows = []
    dev_type = torch.device(device).type
    with torch.autocast(device_type=dev_type, dtype=get_autocast_type(training_args)), torch.inference_mode():
        for i, ex in enumerate(examples):
            if prompt is None:
                prompt_text, generated = _complete_own_prompt(model, tokenizer, ex, end_seq, max_new_tokens)
            else:
                print(f"[DEBUG] inferring with fixed prompt: {prompt}")
                generated = infer_with_prompt(prompt, model=model, audio_encoding=ex["audio_encoding"],
                                              multimodal_cfg=mm_cfg, end_seq=end_seq, tokenizer=tokenizer,
                                              audio_first=True, max_new_tokens=max_new_tokens)
                prompt_text = prompt
            row = {"example_id": ex["example_id"], "prompt_text": prompt_text}
            if not from_folder:
                row["original_completion_text"] = tokenizer.decode(extract_response_tokens(ex["input_ids"], end_seq))
            row["model_completion_text"] = tokenizer.decode(extract_response_tokens(generated[0], end_seq))
            print(f"[PROMPT]\n{prompt_text}\n[MODEL COMPLETION]\n{row['model_completion_text']}\n" + "%" * 40)
            rows.append(row)
            if max_samples and i >= max_samples:
                break
    print(f"writing {len(rows)} results to {outfile}")
    pd.DataFrame(rows).to_csv(outfile, index=False)
    return rows


def main(argv=None):
    import transformers

    parser = transformers.HfArgumentParser((ModelArguments, DataArguments, TrainingArguments))
    parser.add_argument("--ckpt-num", type=int, required=True, help="step number of the trained checkpoint")
    parser.add_argument("--max-samples", type=int, default=None, help="stop after this element index")
    parser.add_argument("--audio-encodings-dir", default=None, help="folder of .npy embeddings (instead of shards)")
    parser.add_argument("--prompt", default=None, help="one prompt for every element, without conversation headers")
    parser.add_argument("--max-new-tokens", type=int, default=None)
    parser.add_argument("--outfile", default="infer_results.csv", help="csv file the results are written to")
    model_args, data_args, training_args, other = parser.parse_args_into_dataclasses(args=argv)
    device = str(training_args.device)
    return run(model_args, data_args, training_args, device=device, **vars(other))


if __name__ == "__main__":
    main()
