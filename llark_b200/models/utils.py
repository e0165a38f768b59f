"""Checkpoint formats either side of the hot path: mirror of m2t/models/utils.py with the same
function names, arguments and behaviour (pure host logic, tested on CPU against the reference's own functions).

    load_sharded_mm_projector_weights   m2t/models/utils.py:30-52
    load_sharded_checkpoint             :55-123  (transformers' loader + the mm_projector special case; like the reference
                                                  it computes missing / unexpected keys but never raises on them: the
                                                  `raise` is commented out upstream, :102)
    load_pretrained_model               :126-203 (Llama-2 branch; the MPT branch belongs to the CLAP + MPT variant,
                                                  §8(f) row 1, which is not built: NotImplementedError)
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


def load_sharded_mm_projector_weights(model, folder):
    """Load sharded mm_projector weights (and, as the reference does, every other tensor of those shards) from folder."""
    index = _read_index(folder)
    for shard_file in list(set(index["weight_map"].values())):
        _load_shard_into(model, os.path.join(folder, shard_file))


def load_sharded_checkpoint(model, folder, strict=True):
    """transformers.modeling_utils.load_sharded_checkpoint plus the mm_adapter parameters. Shards are loaded one at a
    time and freed. Returns torch's `_IncompatibleKeys(missing_keys, unexpected_keys)`; `strict` only decides whether
    the (unused) error message is built, exactly as upstream."""
    index = _read_index(folder)
    shard_files = list(set(index["weight_map"].values()))
    loaded_keys = index["weight_map"].keys()
    model_keys = model.state_dict().keys()
    missing_keys = [key for key in model_keys if key not in loaded_keys]
    unexpected_keys = [key for key in loaded_keys if key not in model_keys]
    for shard_file in shard_files:
        _load_shard_into(model, os.path.join(folder, shard_file))
        gc.collect()
    return torch.nn.modules.module._IncompatibleKeys(missing_keys, unexpected_keys)


def load_pretrained_model(model_name: str, ckpt_num: int, torch_dtype=torch.float16, mm_use_audio_start_end=True,
                          device="cuda:0") -> Tuple[object, object]:
    """`<model_name>/checkpoint-<ckpt_num>` -> (model on `device`, tokenizer), m2t/models/utils.py:126-203."""
    from transformers import AutoTokenizer

    from llark_b200.models.llamav2 import WrappedLlamav2ForCausalLM

    ckpt_dir = os.path.join(model_name, f"checkpoint-{ckpt_num}")
    tokenizer = AutoTokenizer.from_pretrained(ckpt_dir)
    if os.path.exists(os.path.join(ckpt_dir, WEIGHTS_NAME)):
        sharded_ckpt = False
    else:
        ckpt_glob = os.path.join(ckpt_dir, "pytorch_model-*of*.bin")
        ckpt_files = glob.glob(ckpt_glob)
        assert len(ckpt_files), f"no files found matching {ckpt_glob}"
        print(f"got checkpoint files {ckpt_files}")
        sharded_ckpt = True
    if "mosaicml/mpt" in model_name:
        raise NotImplementedError("the CLAP + MPT-1B variant (SURVEY §8(f) row 1) is not part of llark_b200 yet")
    elif "meta-llama/Llama-2" in model_name:
        model = WrappedLlamav2ForCausalLM.from_pretrained(ckpt_dir, torch_dtype=torch_dtype)
        # does NOT load the adapter weights; it creates the module so that they can be loaded below
        model.get_model().initialize_adapter_modules(tune_mm_mlp_adapter=False, fsdp=None)
    else:
        raise UnboundLocalError(f"no model class for {model_name!r} (the reference leaves `model` unbound here)")
    if sharded_ckpt:
        print("[DEBUG] loading mm projector parameters")
        load_sharded_mm_projector_weights(model, ckpt_dir)
    model.to(device)
    model.get_model().mm_projector.to(device)
    model.config.mm_use_audio_start_end = mm_use_audio_start_end
    model.initialize_audio_tokenizer(mm_use_audio_start_end=model.config.mm_use_audio_start_end, tokenizer=tokenizer,
                                     device=device, tune_mm_mlp_adapter=model.config.tune_mm_mlp_adapter,
                                     pretrain_mm_mlp_adapter=None)
    return model, tokenizer
