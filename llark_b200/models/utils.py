"""Checkpoint formats either side of the hot path: mirror of m2t/models/utils.py with the same
function names, arguments and behaviour (pure host logic, tested on CPU against the reference's own functions).

    load_sharded_mm_projector_weights   m2t/models/utils.py:30-52
    load_sharded_checkpoint             :55-123  (transformers' loader + the mm_projector special case; like the reference
                                                  it computes missing / unexpected keys but never raises on them: the
                                                  `raise` is commented out upstream, :102)
    load_pretrained_model               :126-203 (both branches: "meta-llama/Llama-2" and "mosaicml/mpt" in the name)
The on-disk layout is HF's: `pytorch_model.bin` or `pytorch_model-XXXXX-of-YYYYY.bin` shards listed in
`pytorch_model.bin.index.json` (`weight_map`), projector tensors under `model.mm_projector.{weight,bias}`.
"""
import gc
import glob
import json
import os
from typing import Tuple

import torch

from llark_b200.special_tokens import DEFAULT_AUDIO_END_TOKEN, DEFAULT_AUDIO_START_TOKEN  # noqa: F401 (API parity)

WEIGHTS_NAME = "pytorch_model.bin"                    # transformers.trainer.WEIGHTS_NAME
WEIGHTS_INDEX_NAME = "pytorch_model.bin.index.json"   # transformers.trainer.WEIGHTS_INDEX_NAME


def _read_index(folder):
    index_file = os.path.join(folder, WEIGHTS_INDEX_NAME)
    if not os.path.isfile(index_file):
        raise ValueError(f"Can't find a checkpoint index ({WEIGHTS_INDEX_NAME}) in {folder}.")
    with open(index_file, "r", encoding="utf-8") as f:
        return json.load(f)


def _load_shard_into(model, path):
    """One shard: everything but the projector through load_state_dict(strict=False), the projector tensors through
    mm_projector.load_state_dict with the key prefix stripped (m2t/models/utils.py:42-52, :106-117)."""
    state_dict = torch.load(path, map_location="cpu")
    model.load_state_dict({k: v for k, v in state_dict.items() if "mm_projector" not in k}, strict=False)
    if any("mm_projector" in x for x in state_dict.keys()):
        print(f"loading mm_projector params from ckpt file {os.path.basename(path)}")
        model.get_model().mm_projector.load_state_dict(
            {k.split(".")[-1]: v for k, v in state_dict.items() if "mm_projector" in k})
    del state_dict


def _shard_paths(folder):
    return [os.path.join(folder, name) for name in set(_read_index(folder)["weight_map"].values())]


def load_sharded_mm_projector_weights(model, folder):
    """Walk the shards of `folder` and load the projector tensors (and, as upstream does, every other tensor that
    happens to live in those shards) into `model`."""
    for path in _shard_paths(folder):
        _load_shard_into(model, path)


def load_sharded_checkpoint(model, folder, strict=True):
    """transformers.modeling_utils.load_sharded_checkpoint with the mm_adapter special case: one shard in memory at a
    time. Returns torch's `_IncompatibleKeys(missing_keys, unexpected_keys)` computed from the index against the model's
    state dict. `strict` is accepted for signature parity; upstream builds an error message under it and then does NOT
    raise (the raise is commented out, m2t/models/utils.py:102), so neither does this."""
    in_checkpoint = _read_index(folder)["weight_map"].keys()
    in_model = model.state_dict().keys()
    report = torch.nn.modules.module._IncompatibleKeys([k for k in in_model if k not in in_checkpoint],
                                                      [k for k in in_checkpoint if k not in in_model])
    for path in _shard_paths(folder):
        _load_shard_into(model, path)
        gc.collect()  # the shard's tensors are freed before the next one is read
    return report


def _checkpoint_is_sharded(ckpt_dir) -> bool:
    if os.path.exists(os.path.join(ckpt_dir, WEIGHTS_NAME)):
        return False
    pattern = os.path.join(ckpt_dir, "pytorch_model-*of*.bin")
    found = glob.glob(pattern)
    assert len(found), f"no files found matching {pattern}"
    print(f"got checkpoint files {found}")
    return True


def load_pretrained_model(model_name: str, ckpt_num: int, torch_dtype=torch.float16, mm_use_audio_start_end=True,
                          device="cuda:0") -> Tuple[object, object]:
    """`<model_name>/checkpoint-<ckpt_num>` -> (model on `device`, tokenizer), m2t/models/utils.py:126-203: tokenizer
    from the checkpoint folder, backbone through from_pretrained, an EMPTY projector created by
    initialize_adapter_modules and then filled from the shards when the checkpoint is sharded, audio tokens registered
    last. The model family is chosen by substring of `model_name` exactly as upstream ("meta-llama/Llama-2" /
    "mosaicml/mpt"); anything else leaves upstream with an unbound `model`."""
    from transformers import AutoTokenizer

    from llark_b200.models.llamav2 import WrappedLlamav2ForCausalLM

    ckpt_dir = os.path.join(model_name, f"checkpoint-{ckpt_num}")
    tokenizer = AutoTokenizer.from_pretrained(ckpt_dir)
    sharded = _checkpoint_is_sharded(ckpt_dir)
    is_mpt = "mosaicml/mpt" in model_name
    if is_mpt:
        # CLAP + MPT-1B variant (:149-169): single-file checkpoints only -- upstream reads `ckpt_file`, which it binds
        # only when pytorch_model.bin exists, so a sharded MPT checkpoint dies there with a NameError too.
        from llark_b200.models.mpt import WrappedMPTForCausalLM

        if sharded:
            raise NameError("name 'ckpt_file' is not defined (the reference's MPT branch needs an unsharded checkpoint)")
        ckpt_file = os.path.join(ckpt_dir, WEIGHTS_NAME)
        model = WrappedMPTForCausalLM.from_pretrained(ckpt_dir, torch_dtype=torch_dtype)
        model.get_model().initialize_adapter_modules(tune_mm_mlp_adapter=False, pretrain_mm_mlp_adapter=ckpt_file,
                                                     fsdp=None)
        print(f"[DEBUG] loading {ckpt_file} weights manually")
        model.load_state_dict(torch.load(ckpt_file, map_location="cpu"))
        model.config.mm_use_audio_start_end = True
        aec = model.transformer.audio_encoder_config
        aec.audio_start_token, aec.audio_end_token = tokenizer.convert_tokens_to_ids(
            [DEFAULT_AUDIO_START_TOKEN, DEFAULT_AUDIO_END_TOKEN])
    elif "meta-llama/Llama-2" in model_name:
        model = WrappedLlamav2ForCausalLM.from_pretrained(ckpt_dir, torch_dtype=torch_dtype)
        model.get_model().initialize_adapter_modules(tune_mm_mlp_adapter=False, fsdp=None)
    else:
        raise UnboundLocalError(f"no model class for {model_name!r} (the reference leaves `model` unbound here)")
    if sharded:
        print("[DEBUG] loading mm projector parameters")
        load_sharded_mm_projector_weights(model, ckpt_dir)
    elif not is_mpt:
        # Deviation from upstream, which fills the projector only from shards (HF save_pretrained always shards a 7B
        # model): llark_b200's WrappedTrainer._save writes ONE pytorch_model.bin, from_pretrained drops its
        # `model.mm_projector.*` keys as unexpected (the module does not exist yet), so read them here -- otherwise
        # inference on such a checkpoint would run with the random projector initialize_adapter_modules just created.
        _load_shard_into(model, os.path.join(ckpt_dir, WEIGHTS_NAME))
    model.to(device)
    model.get_model().mm_projector.to(device)
    model.config.mm_use_audio_start_end = mm_use_audio_start_end
    if not is_mpt:  # :190-197 -- the MPT checkpoint already carries its audio tokens and their ids
        model.initialize_audio_tokenizer(mm_use_audio_start_end=mm_use_audio_start_end, tokenizer=tokenizer,
                                         device=device, tune_mm_mlp_adapter=model.config.tune_mm_mlp_adapter,
                                         pretrain_mm_mlp_adapter=None)
    return model, tokenizer
