"""Training-script helpers with the names and behaviour of m2t/utils.py (dtype selection :23-43, state-dict dump :46-57,
tokenizer / embedding growth :60-80); pinned against the reference's own functions by tests/test_formats_cpu.py."""
import logging
from typing import Dict

import torch

_HALF_BY_FLAG = (("bf16", torch.bfloat16), ("fp16", torch.float16))


def get_autocast_type(training_args):
    """16-bit runs autocast to whichever of bf16 / fp16 is switched on (bf16 wins; neither -> None, as upstream falls
    through); any other bit width is not handled by the reference and degrades to fp32 with a warning."""
    if training_args.bits != 16:
        logging.warning(f"Autocast logic for bits {training_args.bits} not implemented;"
                        "falling back to torch.float. If there are type mismatch errors"
                        "this could be the cause.")
        return torch.float
    for flag, dtype in _HALF_BY_FLAG:
        if getattr(training_args, flag):
            return dtype
    return None


def get_compute_dtype(training_args):
    """fp16 has priority over bf16 here (the opposite order of get_autocast_type -- upstream's choice), else fp32."""
    for flag, dtype in reversed(_HALF_BY_FLAG):
        if getattr(training_args, flag):
            return dtype
    return torch.float32


def safe_save_model_for_hf_trainer(trainer, output_dir: str):
    """Write the model through `trainer._save` from a CPU copy of the state dict (rank-0 only via args.should_save);
    under DeepSpeed the trainer's own save_model is used instead."""
    if trainer.deepspeed:
        torch.cuda.synchronize()
        trainer.save_model(output_dir)
        return
    on_device = trainer.model.state_dict()
    if not trainer.args.should_save:
        return
    host_copy = {name: tensor.cpu() for name, tensor in on_device.items()}
    del on_device
    trainer._save(output_dir, state_dict=host_copy)


def _mean_fill_tail(matrix: torch.Tensor, n_new: int) -> None:
    """Rows [-n_new:] <- mean of the rows before them (in place)."""
    matrix[-n_new:] = matrix[:-n_new].mean(dim=0, keepdim=True)


def smart_tokenizer_and_embedding_resize(special_tokens_dict: Dict, tokenizer, model):
    """Register the special tokens, grow both embedding matrices to the new vocabulary and start the added rows at
    the mean of the existing ones (the size may stop being a multiple of 64, as upstream notes)."""
    added = tokenizer.add_special_tokens(special_tokens_dict)
    model.resize_token_embeddings(len(tokenizer))
    if added > 0:
        _mean_fill_tail(model.get_input_embeddings().weight.data, added)
        _mean_fill_tail(model.get_output_embeddings().weight.data, added)
