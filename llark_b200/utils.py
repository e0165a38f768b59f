"""Mirror of m2t/utils.py (dtype selection, state-dict dump, tokenizer / embedding resize): same names and behaviour.

    get_autocast_type                      m2t/utils.py:23-35
    get_compute_dtype                      :38-43
    safe_save_model_for_hf_trainer         :46-57
    smart_tokenizer_and_embedding_resize   :60-80
"""
import logging
from typing import Dict

import torch


def get_autocast_type(training_args):
    if training_args.bits == 16:
        if training_args.bf16:
            return torch.bfloat16
        elif training_args.fp16:
            return torch.float16
        # bits == 16 with neither flag: the reference falls off the end of the function and returns None
    else:
        logging.warning(f"Autocast logic for bits {training_args.bits} not implemented;"
                        "falling back to torch.float. If there are type mismatch errors"
                        "this could be the cause.")
        return torch.float


def get_compute_dtype(training_args):
    return torch.float16 if training_args.fp16 else (torch.bfloat16 if training_args.bf16 else torch.float32)


def safe_save_model_for_hf_trainer(trainer, output_dir: str):
    """Collects the state dict and dumps it to disk through trainer._save (which also writes the projector side file)."""
    if trainer.deepspeed:
        torch.cuda.synchronize()
        trainer.save_model(output_dir)
        return
    state_dict = trainer.model.state_dict()
    if trainer.args.should_save:
        cpu_state_dict = {key: value.cpu() for key, value in state_dict.items()}
        del state_dict
        trainer._save(output_dir, state_dict=cpu_state_dict)


def smart_tokenizer_and_embedding_resize(special_tokens_dict: Dict, tokenizer, model):
    """Add special tokens, resize the embeddings, initialise the new input / output rows with the mean of the old ones."""
    num_new_tokens = tokenizer.add_special_tokens(special_tokens_dict)
    model.resize_token_embeddings(len(tokenizer))
    if num_new_tokens > 0:
        input_embeddings = model.get_input_embeddings().weight.data
        output_embeddings = model.get_output_embeddings().weight.data
        input_embeddings_avg = input_embeddings[:-num_new_tokens].mean(dim=0, keepdim=True)
        output_embeddings_avg = output_embeddings[:-num_new_tokens].mean(dim=0, keepdim=True)
        input_embeddings[-num_new_tokens:] = input_embeddings_avg
        output_embeddings[-num_new_tokens:] = output_embeddings_avg
