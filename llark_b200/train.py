"""Training entry point: `python -m llark_b200.train <the flags of scripts/training/train_llark.sh>` (under torchrun for
more than one GPU), or `train(model_args, data_args, training_args)` from Python. It walks the same sequence as
m2t/train.py:53-278 for the regime that script uses -- 16-bit full fine-tuning, data parallel -- on llark_b200's own
classes (the reference's train() also runs unchanged over them; this module is for use without a
reference checkout):

    from_pretrained (MPT when "mpt" is in the name, else Llama-2)        m2t/train.py:62-76
    config.use_cache off, optional backbone freeze, cast to bf16 / fp16                 :77-106
    tokenizer: pad token added with mean-initialised rows; llama specials               :108-126
    projector created / loaded and cast, trainable flags, audio tokens registered       :128-182
    data module, one sample pushed through the collator (fails early on bad data)       :236-253
    WrappedTrainer.train (resuming when output_dir holds a checkpoint-*), save          :255-276
"""
import logging
import pathlib

import torch

from llark_b200.arguments import (DataArguments, ModelArguments, TrainingArguments, check_supported,
                                  get_bnb_model_args, parse_command_line)
from llark_b200.data_modules import make_data_module
from llark_b200.special_tokens import DEFAULT_BOS_TOKEN, DEFAULT_EOS_TOKEN, DEFAULT_PAD_TOKEN, DEFAULT_UNK_TOKEN
from llark_b200.tokenizer import get_tokenizer
from llark_b200.trainer import WrappedTrainer
from llark_b200.utils import get_compute_dtype, safe_save_model_for_hf_trainer, smart_tokenizer_and_embedding_resize

__all__ = ["train", "main", "ModelArguments", "DataArguments", "TrainingArguments"]


def _half_dtype(training_args):
    """The storage dtype --bf16 / --fp16 select at 16 bits (fp16 wins when both are set, as the two successive casts
    of m2t/train.py:101-106 leave it)."""
    if training_args.bits != 16:
        return None
    if training_args.fp16:
        return torch.float16
    return torch.bfloat16 if training_args.bf16 else None


def _load_backbone(model_args, training_args):
    if "mpt" in model_args.model_name_or_path:
        from llark_b200.models.mpt import WrappedMPTForCausalLM as model_cls
    else:
        from llark_b200.models.llamav2 import WrappedLlamav2ForCausalLM as model_cls
    extra = get_bnb_model_args(training_args, get_compute_dtype(training_args))
    model = model_cls.from_pretrained(model_args.model_name_or_path, cache_dir=training_args.cache_dir,
                                      mm_hidden_size=model_args.mm_hidden_size, **extra)
    model.config.use_cache = False
    if model_args.freeze_backbone:
        model.model.requires_grad_(False)
    dtype = _half_dtype(training_args)
    if dtype is not None:
        model.to(dtype)
    return model


def _prepare_tokenizer(model, model_args, training_args):
    tokenizer = get_tokenizer(model_args, training_args)
    if model_args.version != "v0":
        raise NotImplementedError(f"version {model_args.version} not implemented.")
    if tokenizer.pad_token is None:
        smart_tokenizer_and_embedding_resize(special_tokens_dict={"pad_token": DEFAULT_PAD_TOKEN}, tokenizer=tokenizer,
                                             model=model)
    if "llama" in model_args.model_name_or_path:
        tokenizer.add_special_tokens({"eos_token": DEFAULT_EOS_TOKEN, "bos_token": DEFAULT_BOS_TOKEN,
                                      "unk_token": DEFAULT_UNK_TOKEN})
    return tokenizer


def _attach_audio_path(model, tokenizer, model_args, data_args, training_args):
    """Projector, trainable set and audio tokens (m2t/train.py:128-182)."""
    inner = model.get_model()
    audio_config = inner.initialize_adapter_modules(pretrain_mm_mlp_adapter=model_args.pretrain_mm_mlp_adapter,
                                                    tune_mm_mlp_adapter=model_args.tune_mm_mlp_adapter,
                                                    fsdp=training_args.fsdp)["audio_config"]
    dtype = _half_dtype(training_args)
    if dtype is not None:
        inner.mm_projector.to(dtype)
    if not data_args.is_multimodal:
        raise AssertionError("llark training is multimodal: --is_multimodal must stay True")
    if model_args.tune_mm_mlp_adapter and training_args.freeze_mm_mlp_adapter:
        raise AssertionError("--tune_mm_mlp_adapter and --freeze_mm_mlp_adapter contradict each other")

    model.config.tune_mm_mlp_adapter = training_args.tune_mm_mlp_adapter = model_args.tune_mm_mlp_adapter
    model.config.freeze_mm_mlp_adapter = training_args.freeze_mm_mlp_adapter
    if model_args.freeze_backbone:
        print("[INFO] freezing backbone LM weights.")
        model.requires_grad_(False)
    else:
        print("[INFO] training mm backbone LLM weights.")
    if model_args.tune_mm_mlp_adapter:
        print("[INFO] training MM MLP adapter weights")
        inner.mm_projector.requires_grad_(True)
    if training_args.freeze_mm_mlp_adapter:
        print("[INFO] freezing mm projector weights.")
        inner.mm_projector.requires_grad_(False)

    use_markers = model_args.mm_use_audio_start_end
    model.config.mm_use_audio_start_end = data_args.mm_use_audio_start_end = use_markers
    audio_config.use_audio_start_end = training_args.use_audio_start_end = use_markers
    model.initialize_audio_tokenizer(mm_use_audio_start_end=use_markers, tokenizer=tokenizer,
                                     device=training_args.device, tune_mm_mlp_adapter=model_args.tune_mm_mlp_adapter,
                                     pretrain_mm_mlp_adapter=model_args.pretrain_mm_mlp_adapter)


def _describe(sample):
    return {k: f"shape: {getattr(v, 'shape', None)} dtype {getattr(v, 'dtype', None)}" for k, v in sample.items()}


def train(model_args: ModelArguments, data_args: DataArguments, training_args: TrainingArguments):
    check_supported(training_args)
    model = _load_backbone(model_args, training_args)
    tokenizer = _prepare_tokenizer(model, model_args, training_args)
    _attach_audio_path(model, tokenizer, model_args, data_args, training_args)
    data_module = make_data_module(tokenizer=tokenizer, data_args=data_args)
    logging.warning("data module initialized; fetching sample batch")
    first = next(iter(data_module["train_dataset"]))
    logging.warning(f"sample batch info: {_describe(first)}")
    logging.warning(f"sample batch collated info: {_describe(data_module['data_collator']([first]))}")

    trainer = WrappedTrainer(model=model, tokenizer=tokenizer, args=training_args, **data_module)
    resume = any(pathlib.Path(training_args.output_dir).glob("checkpoint-*"))
    if resume:
        trainer.train(resume_from_checkpoint=True)
    else:
        trainer.train()
    trainer.save_state()
    safe_save_model_for_hf_trainer(trainer=trainer, output_dir=training_args.output_dir)
    return {"trainer": trainer, "tokenizer": tokenizer, "data_module": data_module}


def main(argv=None):
    torch.multiprocessing.set_start_method("spawn", force=True)  # dataloader workers, as m2t/train.py:281
    model_args, data_args, training_args = parse_command_line(argv)
    return train(model_args=model_args, data_args=data_args, training_args=training_args)


if __name__ == "__main__":
    main()
