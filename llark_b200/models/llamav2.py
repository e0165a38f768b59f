"""B200-native mirror of m2t/models/llamav2.py: same class names, methods, argument meaning and error behaviour,
with the arithmetic executed by hand-written sm_100a kernels (llark_b200.engine -> C-ABI) instead of
transformers/cuBLAS/ATen.

Drop-in contract: `WrappedLlamav2Config`, `WrappedLlamav2Model` (with `mm_projector`,
`audio_encoder_config`, `orig_embeds_params`, `initialize_adapter_modules`), `WrappedLlamav2ForCausalLM`
(`get_model`, `forward(..., audio_encodings=)`, `initialize_audio_tokenizer`, `prepare_inputs_for_generation`),
HF state_dict key names (`model.embed_tokens.weight`, `model.layers.{i}...`, `model.norm.weight`,
`model.mm_projector.{weight,bias}`, `lm_head.weight`) and the Auto* registration, so m2t/train.py runs unchanged
after `import llark_b200.models.llamav2 as llamav2` replaces `m2t.models.llamav2`.

The modules are ordinary HF `nn.Module`s (they only hold the parameters, re-pointed into the engine's flat bf16
bucket); `forward` never calls them.
"""
import logging
import os
from typing import List, Optional

import torch
import torch.nn as nn
from transformers import AutoConfig, AutoModelForCausalLM, LlamaConfig, LlamaForCausalLM, LlamaModel
from transformers.modeling_outputs import CausalLMOutputWithPast

from llark_b200.engine import LlamaEngine
from llark_b200.models import AudioEncoderConfig
from llark_b200.special_tokens import (
    DEFAULT_AUDIO_END_TOKEN,
    DEFAULT_AUDIO_PATCH_TOKEN,
    DEFAULT_AUDIO_START_TOKEN,
)


class WrappedLlamav2Config(LlamaConfig):
    """m2t/models/llamav2.py:39-43."""

    model_type = "wrapped_llamav2"
    mm_hidden_size: int = 4800  # width of the Jukebox embeddings after temporal averaging


class WrappedLlamav2Model(LlamaModel):
    """Parameter container + adapter initialisation (m2t/models/llamav2.py:46-93). The forward of the reference
    (:95-234, embed gather -> mm_projector -> splice -> LlamaModel.forward) is executed by LlamaEngine."""

    config_class = WrappedLlamav2Config

    def __init__(self, config: LlamaConfig):
        super().__init__(config)
        self.audio_encoder_config = AudioEncoderConfig()

    def initialize_adapter_modules(self, pretrain_mm_mlp_adapter=None, tune_mm_mlp_adapter=None, fsdp: bool = None):
        """m2t/models/llamav2.py:60-93: create mm_projector = Linear(mm_hidden_size, hidden_size), optionally load
        `*mm_projector*` keys from a checkpoint. `tune_mm_mlp_adapter` and `fsdp` are accepted and unused."""
        print("[INFO] ignoring parameter fsdp")
        del fsdp, tune_mm_mlp_adapter
        self.config.use_mm_proj = True
        if not hasattr(self, "mm_projector"):
            ref = self.embed_tokens.weight
            self.mm_projector = nn.Linear(self.config.mm_hidden_size, self.config.hidden_size, device=ref.device)
        if pretrain_mm_mlp_adapter is not None:
            weights = torch.load(pretrain_mm_mlp_adapter, map_location="cpu")
            self.mm_projector.load_state_dict(
                {k.split(".")[-1]: v for k, v in weights.items() if "mm_projector" in k})
        return dict(audio_config=AudioEncoderConfig())

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "WrappedLlamav2Model.forward is executed by the fused engine: call WrappedLlamav2ForCausalLM.forward")


class WrappedLlamav2ForCausalLM(LlamaForCausalLM):
    """m2t/models/llamav2.py:237-419."""

    config_class = WrappedLlamav2Config
    supports_gradient_checkpointing = True

    def __init__(self, config):
        super(LlamaForCausalLM, self).__init__(config)
        self.model = WrappedLlamav2Model(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()
        self._engine = LlamaEngine(self)
        # "sync": raise the splice ValueErrors in forward like the reference (one tiny D2H read per call);
        # "lazy": leave them in the device flag word until check_audio_token_errors() is called (bench / trainer).
        self.validate = "sync"

    def get_model(self):
        return self.model

    def _set_gradient_checkpointing(self, module, value=False):
        # The engine recomputes a layer's activations in the backward only when this flag is set AND the step's saved
        # activations would not fit in free HBM (LlamaEngine._want_recompute; a 7B step at S=512, B=16 keeps 35 GB).
        if isinstance(module, WrappedLlamav2Model):
            module.gradient_checkpointing = value

    def gradient_checkpointing_enable(self, *args, **kwargs):
        self.model.gradient_checkpointing = True

    @property
    def engine(self):
        return self._engine

    def check_audio_token_errors(self):
        self._engine.check_flags()

    def forward(
        self,
        input_ids: torch.LongTensor = None,
        attention_mask: Optional[torch.Tensor] = None,
        position_ids: Optional[torch.Tensor] = None,
        past_key_values: Optional[List[torch.FloatTensor]] = None,
        labels: Optional[torch.LongTensor] = None,
        use_cache: Optional[bool] = None,
        output_attentions: Optional[bool] = None,
        output_hidden_states: Optional[bool] = None,
        return_dict: Optional[bool] = None,
        audio_encodings=None,
        **kwargs,
    ):
        """Same signature and return value as m2t/models/llamav2.py:259-337. Training / scoring / cache-less
        generation. `past_key_values` / `use_cache=True` run the KV-cache path (prefill + one-token decode steps,
        SURVEY §8(f) row 2; verified on a B200 against the cache-less loop, tests/test_model_gpu.py)."""
        if os.environ.get("LLARK_VERIFY") == "fp32":
            # fp32 verification mode (north_star: fp32 logits within 1e-3 relative of the reference's fp32 run):
            # forward only, fp32 tensors end to end on the 3xTF32 tcgen05 path -- llark_b200/verify.py
            from llark_b200.verify import forward_fp32

            if torch.is_grad_enabled() and labels is not None and any(p.requires_grad for p in self.parameters()):
                raise RuntimeError("LLARK_VERIFY=fp32 is a forward-only checking mode: call it under torch.no_grad()")
            loss, logits = forward_fp32(self, input_ids, attention_mask, labels, audio_encodings)
            return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=None,
                                          attentions=None)
        if past_key_values is not None or use_cache:
            if os.environ.get("LLARK_KV_CACHE", "1") == "0":
                raise NotImplementedError("KV-cache decoding is switched off (LLARK_KV_CACHE=0): generate() then runs "
                                          "one full forward per token")
            return self._forward_cached(input_ids, attention_mask, past_key_values, labels, return_dict,
                                        audio_encodings)
        if position_ids is not None:
            S = input_ids.shape[1]
            if not torch.equal(position_ids.reshape(-1, S)[0].cpu(), torch.arange(S)):
                raise NotImplementedError("only default position_ids = arange(S) are supported")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("attention maps / hidden states are not materialised by the fused path")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        if audio_encodings is not None and getattr(self.config, "use_mm_proj", False):
            if not self.model.audio_encoder_config.use_audio_start_end:
                raise NotImplementedError("audio_encoder_config.use_audio_start_end=False is not implemented.")
        loss, logits = self._engine.run(input_ids, attention_mask, labels, audio_encodings)
        if self.validate == "sync" and audio_encodings is not None:
            self._engine.check_flags()
        if not return_dict:
            output = (logits,)
            return (loss,) + output if loss is not None else output
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=None,
                                      attentions=None)

    def _forward_cached(self, input_ids, attention_mask, past_key_values, labels, return_dict, audio_encodings):
        """use_cache=True: the first call (no cache) is the prefill over the whole prompt, later calls feed the newest
        token only (m2t/models/llamav2.py:346-347). Returns `past_key_values` = the engine's KVCache."""
        if labels is not None:
            raise NotImplementedError("labels with use_cache=True are not supported")
        if attention_mask is not None and not bool(attention_mask.to(torch.bool).all()):
            raise NotImplementedError("cached decoding needs unpadded prompts (the reference decodes one prompt at a time)")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        cache = past_key_values
        if cache is None or len(cache) == 0:
            s_max = min(int(getattr(self.config, "max_position_embeddings", 4096)), 4096)
            if cache is None:
                cache = self._engine.new_cache(input_ids.shape[0], s_max, input_ids.device)
            logits = self._engine.prefill(input_ids, audio_encodings, cache)
            if self.validate == "sync" and audio_encodings is not None:
                self._engine.check_flags()
        else:
            if input_ids.shape[1] != 1:
                input_ids = input_ids[:, -1:]
            logits = self._engine.decode_step(input_ids, cache)
        if not return_dict:
            return (logits, cache)
        return CausalLMOutputWithPast(loss=None, logits=logits, past_key_values=cache, hidden_states=None,
                                      attentions=None)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                      **kwargs):
        """m2t/models/llamav2.py:339-365 (generation plumbing; forwards audio_encodings on every step)."""
        if past_key_values:
            input_ids = input_ids[:, -1:]
        if inputs_embeds is not None and past_key_values is None:
            model_inputs = {"inputs_embeds": inputs_embeds}
        else:
            model_inputs = {"input_ids": input_ids}
        model_inputs.update({"attention_mask": attention_mask, "past_key_values": past_key_values,
                             "use_cache": kwargs.get("use_cache", True),
                             "audio_encodings": kwargs.get("audio_encodings", None)})
        return model_inputs

    @torch.no_grad()
    def generate(self, inputs=None, input_ids=None, audio_encodings=None, stopping_criteria=None, max_new_tokens=None,
                 max_length=None, do_sample=None, num_beams=None, eos_token_id=None, pad_token_id=None, **kwargs):
        """`model.generate(input_ids=, audio_encodings=, stopping_criteria=[...], max_new_tokens=...)` as called at
        m2t/infer.py:146-152 (and scripts/inference/*.py): greedy decoding of one prompt, returns the full sequences
        [B, prompt + new] like GenerationMixin. Sampling / beam search are not built. The prompt is prefilled once into a KV cache
        and every further token is one decode step (LLARK_KV_CACHE=0: a full forward per token instead); extra keyword arguments the reference passes
        along (`labels`, `example_id`, `id`) are ignored as transformers ignores / rejects them."""
        from llark_b200.generate import generate_with_model

        state = {"cache": None}

        def step_logits(ids):
            if os.environ.get("LLARK_KV_CACHE", "1") != "0":  # prefill once, then one token per step
                out = self.forward(input_ids=ids, audio_encodings=audio_encodings, past_key_values=state["cache"],
                                   use_cache=True, return_dict=True)
                state["cache"] = out.past_key_values
                return out.logits[:, -1, :]
            out = self.forward(input_ids=ids, audio_encodings=audio_encodings, use_cache=False, return_dict=True)
            return out.logits[:, -1, :]

        return generate_with_model(self, step_logits, inputs=inputs, input_ids=input_ids,
                                   stopping_criteria=stopping_criteria, max_new_tokens=max_new_tokens,
                                   max_length=max_length, do_sample=do_sample, num_beams=num_beams,
                                   eos_token_id=eos_token_id, pad_token_id=pad_token_id)

    def initialize_audio_tokenizer(self, mm_use_audio_start_end, tokenizer, device, tune_mm_mlp_adapter=False,
                                   pretrain_mm_mlp_adapter=None):
        """m2t/models/llamav2.py:367-419: add <audio_patch> (+ <audio_start>/<audio_end>), resize both embedding
        matrices, mean-initialise the rows of the two newest tokens, record ids, set the trainable flags."""
        del pretrain_mm_mlp_adapter
        aec = self.get_model().audio_encoder_config
        aec.use_audio_start_end = mm_use_audio_start_end
        tokenizer.add_tokens([DEFAULT_AUDIO_PATCH_TOKEN], special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        if mm_use_audio_start_end:
            num_new_tokens = tokenizer.add_tokens([DEFAULT_AUDIO_START_TOKEN, DEFAULT_AUDIO_END_TOKEN],
                                                  special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
            aec.audio_start_token, aec.audio_end_token = tokenizer.convert_tokens_to_ids(
                [DEFAULT_AUDIO_START_TOKEN, DEFAULT_AUDIO_END_TOKEN])
            if num_new_tokens > 0:
                inp = self.get_input_embeddings().weight.data
                out = self.get_output_embeddings().weight.data
                inp[-num_new_tokens:] = inp[:-num_new_tokens].mean(dim=0, keepdim=True)
                out[-num_new_tokens:] = out[:-num_new_tokens].mean(dim=0, keepdim=True)
            if tune_mm_mlp_adapter:
                self.get_model().orig_embeds_params = [
                    self.get_input_embeddings().weight.data.clone().to(device=device)]
                for p in self.get_input_embeddings().parameters():
                    p.requires_grad = True
                for p in self.get_output_embeddings().parameters():
                    p.requires_grad = False
        aec.audio_patch_token = tokenizer.convert_tokens_to_ids([DEFAULT_AUDIO_PATCH_TOKEN])[0]


def register():
    """Auto* registration (m2t/models/llamav2.py:422-423); tolerant of the reference having registered first."""
    try:
        AutoConfig.register("wrapped_llamav2", WrappedLlamav2Config)
        AutoModelForCausalLM.register(WrappedLlamav2Config, WrappedLlamav2ForCausalLM)
    except ValueError:
        logging.getLogger(__name__).info("wrapped_llamav2 already registered")


register()
