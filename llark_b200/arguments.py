"""Flag surface of the training / inference entry points: the three dataclasses `transformers.HfArgumentParser` fills
from the command line of scripts/training/train_llark.sh. Same field names and defaults as m2t/arguments.py:22-109, so a
launch script written for the reference parses unchanged; the fields that select code llark_b200 does not ship
(bitsandbytes 4/8-bit, LoRA, forced FSDP) are accepted by the parser and refused by `check_supported` /
`get_bnb_model_args` with a message instead of being silently ignored.
"""
import logging
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import transformers


def _opt(default=None, help=None):
    return field(default=default, metadata={"help": help} if help else None)


@dataclass
class ModelArguments:                                      # m2t/arguments.py:22-44
    model_name_or_path: Optional[str] = _opt("facebook/opt-125m", "checkpoint folder or hub name; 'mpt' in it selects MPT")
    version: Optional[str] = _opt("v0", "conversation / tokenizer recipe; only v0 exists")
    freeze_backbone: bool = _opt(False, "freeze every language-model parameter")
    tune_mm_mlp_adapter: bool = _opt(False, "train the audio projector")
    pretrain_mm_mlp_adapter: Optional[str] = _opt(None, "file with pretrained projector weights")
    mm_use_audio_start_end: bool = _opt(False, "wrap the audio patches in <audio_start> / <audio_end> (required)")
    mm_hidden_size: int = _opt(4800, "width of one audio frame: 4800 Jukebox, 512 CLAP")


def _default_task_probs():
    return {"captioning": 0.15, "reasoning": 0.55, "mir": 0.3}


@dataclass
class DataArguments:                                       # m2t/arguments.py:47-78
    train_data_path: str = _opt(None, "comma-separated shard specs of the training set")
    eval_data_path: str = _opt(None, "shard specs of the evaluation set (optional)")
    eval_data_subsample: float = _opt(None, "fraction of the evaluation set to keep")
    task_sample_probs: Dict[str, float] = field(default_factory=_default_task_probs)
    apply_task_sample_probs: bool = False
    is_multimodal: bool = True
    sep_audio_conv_front: bool = _opt(False, "put the audio tokens in a turn of their own (preprocess_multimodal)")
    audio_folder: Optional[str] = _opt(None)


@dataclass
class TrainingArguments:
    """m2t/arguments.py:81-109 extends transformers.TrainingArguments, whose constructor needs `accelerate` and builds
    the accelerate / DDP state WrappedTrainer does not use. This is a plain dataclass with the SAME flag names for
    everything scripts/training/train_llark.sh passes and everything llark_b200's trainer reads; other transformers flags
    on the command line are reported and ignored (parse_command_line)."""
    # --- what m2t/arguments.py adds
    cache_dir: Optional[str] = _opt(None)
    optim: str = _opt("adamw_torch", "only adamw_torch exists here: the fused bf16 AdamW kernel")
    remove_unused_columns: bool = _opt(False)
    freeze_mm_mlp_adapter: bool = _opt(False)
    force_fsdp: bool = _opt(False)
    model_max_length: int = _opt(512, "sequences are truncated to this length and right-padded in the collator")
    double_quant: bool = _opt(True, "bitsandbytes option (not shipped)")
    quant_type: str = _opt("nf4", "bitsandbytes option (not shipped)")
    bits: int = _opt(16, "16 = bf16 / fp16 weights; 4 and 8 are the bitsandbytes modes (not shipped)")
    lora_enable: bool = False
    lora_r: int = 64
    lora_alpha: int = 16
    lora_dropout: float = 0.05
    lora_weight_path: str = ""
    lora_bias: str = "none"
    # --- the transformers.TrainingArguments flags the launch script sets or the trainer reads (same defaults)
    output_dir: str = _opt("trainer_output")
    max_steps: int = _opt(-1, "required > 0: the training set is an endless stream")
    per_device_train_batch_size: int = _opt(8)
    gradient_accumulation_steps: int = _opt(1)
    gradient_checkpointing: bool = _opt(False, "recompute activations in the backward when they would not fit")
    learning_rate: float = _opt(5e-5)
    weight_decay: float = _opt(0.0)
    adam_beta1: float = _opt(0.9)
    adam_beta2: float = _opt(0.999)
    adam_epsilon: float = _opt(1e-8)
    max_grad_norm: float = _opt(1.0)
    lr_scheduler_type: str = _opt("linear", "only 'cosine' (what train_llark.sh uses) is implemented")
    warmup_ratio: float = _opt(0.0)
    warmup_steps: int = _opt(0)
    logging_steps: int = _opt(500)
    save_strategy: str = _opt("steps")
    save_steps: int = _opt(500)
    save_total_limit: Optional[int] = _opt(None)
    eval_strategy: str = _opt("no", "'steps' evaluates every --eval_steps optimizer steps when --eval_data_path is given")
    eval_steps: Optional[int] = _opt(None, "defaults to --logging_steps, as in transformers")
    per_device_eval_batch_size: int = _opt(8)
    bf16: bool = _opt(False)
    fp16: bool = _opt(False)
    tf32: Optional[bool] = _opt(None, "accepted; the Llama path computes in bf16 on tcgen05 either way")
    seed: int = _opt(42)
    dataloader_num_workers: int = _opt(0)
    ddp_find_unused_parameters: Optional[bool] = _opt(None)
    report_to: Optional[str] = _opt(None, "accepted and ignored: the loop logs JSON lines, no tracker integration")
    fsdp: Optional[str] = _opt(None)
    local_rank: int = _opt(-1)
    use_cpu: bool = _opt(False, "only for dry runs of the host logic in tests: there is no CPU compute path")

    def __post_init__(self):
        if self.local_rank == -1 and "LOCAL_RANK" in os.environ:
            self.local_rank = int(os.environ["LOCAL_RANK"])
        if self.lr_scheduler_type != "cosine":
            logging.warning(f"--lr_scheduler_type {self.lr_scheduler_type}: llark_b200's trainer always runs the cosine "
                            "schedule with warm-up of scripts/training/train_llark.sh")
        if self.save_strategy == "no":
            self.save_steps = 0

    @property
    def device(self):
        import torch

        if self.use_cpu:
            return torch.device("cpu")
        dev = torch.device("cuda", max(self.local_rank, 0))
        if torch.cuda.is_available():
            # one process per GPU: the C-ABI launches on the thread's current device, so make it this rank's
            # (transformers.TrainingArguments._setup_devices does the same)
            torch.cuda.set_device(dev)
        return dev

    @property
    def should_save(self) -> bool:
        return int(os.environ.get("RANK", "0")) == 0


def check_supported(training_args) -> None:
    """The regime of scripts/training/train_llark.sh is 16-bit full fine-tuning under DDP; the quantised / LoRA / FSDP
    branches of m2t/train.py:65-88,166-167,185-234 are outside this library and fail here, up front."""
    if getattr(training_args, "bits", 16) != 16:
        raise NotImplementedError(f"--bits {training_args.bits}: bitsandbytes quantisation is not part of llark_b200")
    if getattr(training_args, "lora_enable", False):
        raise NotImplementedError("--lora_enable: LoRA adapters are not part of llark_b200")
    fsdp = getattr(training_args, "fsdp", None)
    if fsdp is not None and len(fsdp) > 0:
        raise NotImplementedError("--fsdp: 7B parameters + bf16 AdamW fit one B200; llark_b200 is data-parallel only")


def get_bnb_model_args(training_args, compute_dtype) -> dict:
    """m2t/arguments.py:112-133 returns the from_pretrained kwargs of the 4/8-bit modes and {} at 16 bits."""
    del compute_dtype
    if getattr(training_args, "bits", 16) in (4, 8):
        raise NotImplementedError(f"--bits {training_args.bits}: bitsandbytes quantisation is not part of llark_b200")
    return {}


def write_args_to_file(args: List[str], dir: str) -> None:
    """`<dir>/args.txt` in the reference's layout (m2t/arguments.py:136-155): a flag and its value share a line that
    ends in a backslash. Ranks may race on the file; whoever writes last wins, and a concurrent removal is ignored."""
    os.makedirs(dir, exist_ok=True)
    path = os.path.join(dir, "args.txt")
    try:
        os.remove(path)
        logging.info(f"args file already exists at {path}; overwriting it.")
    except FileNotFoundError:
        pass
    logging.info(f"writing args to {path}")
    text = "".join(a + " " if a.startswith("-") else a + " \\\n" for a in args)
    with open(path, "w") as f:
        f.write(text)


# transformers renamed --evaluation_strategy (what train_llark.sh passes, transformers 4.29) to --eval_strategy
_RENAMED_FLAGS = {"--evaluation_strategy": "--eval_strategy"}


def parse_command_line(argv=None):
    """(ModelArguments, DataArguments, TrainingArguments) from argv, as m2t/train.py:282-283 does, plus the flag
    renames above so the reference's launch scripts keep working; transformers flags TrainingArguments does not carry are
    logged and ignored rather than fatal."""
    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    fixed = []
    for a in argv:
        key, eq, val = a.partition("=")
        fixed.append(_RENAMED_FLAGS.get(key, key) + eq + val)
    parser = transformers.HfArgumentParser((ModelArguments, DataArguments, TrainingArguments))
    *parsed, rest = parser.parse_args_into_dataclasses(args=fixed, return_remaining_strings=True)
    if rest:
        logging.warning(f"ignoring command-line arguments llark_b200 has no use for: {rest}")
    return tuple(parsed)
