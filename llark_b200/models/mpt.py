"""Mirror of m2t/models/mpt.py (the CLAP + MPT-1B variant): same class names, methods, attribute
and state-dict key names (`transformer.wte.weight`, `transformer.blocks.{i}.{norm_1,attn.Wqkv,attn.q_ln,attn.k_ln,
attn.out_proj,norm_2,ffn.up_proj,ffn.down_proj}.weight`, `transformer.norm_f.weight`, `transformer.mm_projector.*`),
so `m2t/train.py` takes its `"mpt" in model_name_or_path` branch over it unchanged.

STATUS: runs on the B200 since round 2. The modules hold parameters; `forward` runs llark_b200.engine_mpt.MPTEngine, whose
schedule is verified against oracle/mpt_oracle.py on CPU (tests/test_mpt_host_cpu.py) and on the GPU kernels
(tests/test_mpt_gpu.py, part of `pytest -m gpu`). Supported configuration = what MPT-1B uses: multi-head attention with
ALiBi and q/k LayerNorm, low-precision LayerNorm, no biases, tied embeddings, no logit scale, attn_impl irrelevant."""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from transformers import AutoConfig, AutoModelForCausalLM, GenerationMixin, PretrainedConfig, PreTrainedModel
from transformers.modeling_outputs import CausalLMOutputWithPast

from llark_b200.engine_mpt import MPTEngine
from llark_b200.models import AudioEncoderConfig
from llark_b200.special_tokens import DEFAULT_AUDIO_END_TOKEN, DEFAULT_AUDIO_PATCH_TOKEN, DEFAULT_AUDIO_START_TOKEN

_ATTN_DEFAULTS = {"attn_type": "multihead_attention", "attn_pdrop": 0.0, "attn_impl": "triton", "qk_ln": False,
                  "clip_qkv": None, "softmax_scale": None, "prefix_lm": False, "attn_uses_sequence_id": False,
                  "alibi": False, "alibi_bias_max": 8}


class WrappedMPTConfig(PretrainedConfig):
    """m2t/models/mpt.py:34-36 over configuration_mpt.py:28-125 (the fields the training path reads)."""

    model_type = "wrapped_mpt"

    def __init__(self, d_model=2048, n_heads=16, n_layers=24, expansion_ratio=4, max_seq_len=2048, vocab_size=50368,
                 resid_pdrop=0.0, emb_pdrop=0.0, learned_pos_emb=True, attn_config=None, init_device="cpu",
                 logit_scale=None, no_bias=False, verbose=0, embedding_fraction=1.0,
                 norm_type="low_precision_layernorm", use_cache=False, mm_hidden_size=4800, **kwargs):
        self.d_model, self.n_heads, self.n_layers = d_model, n_heads, n_layers
        self.expansion_ratio, self.max_seq_len, self.vocab_size = expansion_ratio, max_seq_len, vocab_size
        self.resid_pdrop, self.emb_pdrop, self.learned_pos_emb = resid_pdrop, emb_pdrop, learned_pos_emb
        self.attn_config = {**_ATTN_DEFAULTS, **(attn_config or {})}
        self.init_device, self.logit_scale, self.no_bias, self.verbose = init_device, logit_scale, no_bias, verbose
        self.embedding_fraction, self.norm_type, self.use_cache = embedding_fraction, norm_type, use_cache
        self.mm_hidden_size = mm_hidden_size
        kwargs.pop("tie_word_embeddings", None)
        super().__init__(tie_word_embeddings=True, **kwargs)
        self.tie_word_embeddings = True
        if self.d_model % self.n_heads:
            raise ValueError("d_model must be divisible by n_heads")

    def check_supported(self):
        a = self.attn_config
        problems = []
        if a["attn_type"] != "multihead_attention" or a["prefix_lm"] or a["attn_uses_sequence_id"] or a["clip_qkv"]:
            problems.append("attn_config beyond multihead / causal / no clipping")
        if not (a["alibi"] and a["qk_ln"]):
            problems.append("alibi=True and qk_ln=True are required (what MPT-1B uses)")
        if not self.no_bias or self.logit_scale is not None or self.norm_type.lower() != "low_precision_layernorm":
            problems.append("no_bias=True, logit_scale=None, norm_type=low_precision_layernorm are required")
        if self.d_model // self.n_heads != 128:
            problems.append("head_dim must be 128 (the attention kernels' tile)")
        if problems:
            raise NotImplementedError("MPT configuration outside llark_b200's envelope: " + "; ".join(problems))


def _weight_only_norm(d):
    ln = nn.LayerNorm(d)
    ln.register_parameter("bias", None)  # no_bias=True strips every bias (modeling_mpt.py:77-82)
    return ln


class _Attention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.Wqkv = nn.Linear(d, 3 * d, bias=False)
        self.q_ln = _weight_only_norm(d)
        self.k_ln = _weight_only_norm(d)
        self.out_proj = nn.Linear(d, d, bias=False)


class _FFN(nn.Module):
    def __init__(self, d, ratio):
        super().__init__()
        self.up_proj = nn.Linear(d, ratio * d, bias=False)
        self.down_proj = nn.Linear(ratio * d, d, bias=False)


class _Block(nn.Module):
    """Parameter container with the names of blocks.py:23-66 (norm_1, attn, norm_2, ffn)."""

    def __init__(self, d, ratio):
        super().__init__()
        self.norm_1 = _weight_only_norm(d)
        self.attn = _Attention(d)
        self.norm_2 = _weight_only_norm(d)
        self.ffn = _FFN(d, ratio)


class WrappedMPTPreTrainedModel(PreTrainedModel):
    config_class = WrappedMPTConfig
    base_model_prefix = "model"
    _no_split_modules = ["_Block"]

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
        elif isinstance(module, nn.LayerNorm):
            module.weight.data.fill_(1.0)


class WrappedMPTModel(WrappedMPTPreTrainedModel):
    """m2t/models/mpt.py:39-234: parameters + adapter initialisation; the forward lives in MPTEngine."""

    def __init__(self, config):
        super().__init__(config)
        d = config.d_model
        self.wte = nn.Embedding(config.vocab_size, d)
        self.blocks = nn.ModuleList([_Block(d, config.expansion_ratio) for _ in range(config.n_layers)])
        self.norm_f = _weight_only_norm(d)
        self.audio_encoder_config = AudioEncoderConfig()
        self.gradient_checkpointing = False
        self.post_init()

    def get_input_embeddings(self):
        return self.wte

    def set_input_embeddings(self, value):
        self.wte = value

    def initialize_adapter_modules(self, pretrain_mm_mlp_adapter=None, tune_mm_mlp_adapter=False, fsdp: bool = None):
        """m2t/models/mpt.py:48-78."""
        print("[INFO] ignoring parameter fsdp")
        del fsdp, tune_mm_mlp_adapter
        self.config.use_mm_proj = True
        if not hasattr(self, "mm_projector"):
            self.mm_projector = nn.Linear(self.config.mm_hidden_size, self.config.d_model, device=self.wte.weight.device)
        if pretrain_mm_mlp_adapter is not None:
            weights = torch.load(pretrain_mm_mlp_adapter, map_location="cpu")
            self.mm_projector.load_state_dict({k.split(".")[-1]: v for k, v in weights.items() if "mm_projector" in k})
        return dict(audio_config=AudioEncoderConfig())

    def forward(self, *args, **kwargs):
        raise NotImplementedError("WrappedMPTModel.forward is executed by MPTEngine: call WrappedMPTForCausalLM.forward")


class WrappedMPTForCausalLM(WrappedMPTPreTrainedModel, GenerationMixin):
    """m2t/models/mpt.py:248-437."""

    supports_gradient_checkpointing = True

    def __init__(self, config):
        super().__init__(config)
        if not config.tie_word_embeddings:
            raise ValueError("MPTForCausalLM only supports tied word embeddings")
        self.transformer = WrappedMPTModel(config)
        self.logit_scale = None
        self.post_init()
        self._engine = MPTEngine(self)
        self.validate = "sync"

    @property
    def model(self):
        """Alias of self.transformer, matching the multimodal Llama model's interface (m2t/models/mpt.py:273-276)."""
        return self.transformer

    def get_model(self):
        return self.transformer

    @property
    def engine(self):
        return self._engine

    def check_audio_token_errors(self):
        self._engine.check_flags()

    def get_input_embeddings(self):
        return self.transformer.wte

    def set_input_embeddings(self, value):
        self.transformer.wte = value

    def get_output_embeddings(self):
        return self.transformer.wte

    def set_output_embeddings(self, new_embeddings):
        self.transformer.wte = new_embeddings

    def _set_gradient_checkpointing(self, module, value=False):
        if isinstance(module, WrappedMPTModel):
            module.gradient_checkpointing = value

    def gradient_checkpointing_enable(self, *args, **kwargs):
        self.transformer.gradient_checkpointing = True

    def forward(self, input_ids: torch.LongTensor, past_key_values: Optional[List[Tuple[torch.FloatTensor]]] = None,
                attention_mask: Optional[torch.ByteTensor] = None, prefix_mask: Optional[torch.ByteTensor] = None,
                sequence_id: Optional[torch.LongTensor] = None, labels: Optional[torch.LongTensor] = None,
                return_dict: Optional[bool] = None, output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None, use_cache: Optional[bool] = None, audio_encodings=None,
                **kwargs):
        """Signature and return value of m2t/models/mpt.py:285-335."""
        self.config.check_supported()
        if past_key_values is not None or use_cache:
            raise NotImplementedError("KV-cache decoding is not built for the MPT variant")
        if prefix_mask is not None or sequence_id is not None:
            raise NotImplementedError("prefix_lm / sequence_id attention is outside the supported MPT configuration")
        if output_attentions:
            raise NotImplementedError("output_attentions is not implemented yet for MPT")
        if output_hidden_states:
            raise NotImplementedError("hidden states are not materialised by the fused path")
        if attention_mask is not None and self.training and int(attention_mask[:, 0].sum()) != attention_mask.shape[0]:
            raise NotImplementedError("MPT does not support training with left padding.")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        if not return_dict:
            raise NotImplementedError("return_dict False is not implemented yet for MPT")
        loss, logits = self._engine.run(input_ids, attention_mask, labels, audio_encodings)
        if self.validate == "sync" and audio_encodings is not None:
            self._engine.check_flags()
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=None)

    @torch.no_grad()
    def generate(self, inputs=None, input_ids=None, audio_encodings=None, stopping_criteria=None, max_new_tokens=None,
                 max_length=None, do_sample=None, num_beams=None, eos_token_id=None, pad_token_id=None, **kwargs):
        """`model.generate(input_ids=, audio_encodings=, stopping_criteria=[...], max_new_tokens=...)` as m2t/infer.py:146-152
        calls it: greedy, one full forward per new token (no KV cache for this variant), full sequences returned."""
        from llark_b200.generate import generate_with_model

        del kwargs  # labels / example_id / id ride along from infer_with_prompt(**elem)

        def step_logits(ids):
            return self.forward(input_ids=ids, audio_encodings=audio_encodings, use_cache=False,
                                return_dict=True).logits[:, -1, :]

        return generate_with_model(self, step_logits, inputs=inputs, input_ids=input_ids,
                                   stopping_criteria=stopping_criteria, max_new_tokens=max_new_tokens,
                                   max_length=max_length, do_sample=do_sample, num_beams=num_beams,
                                   eos_token_id=eos_token_id, pad_token_id=pad_token_id)

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        """m2t/models/mpt.py:337-365."""
        if inputs_embeds is not None:
            raise NotImplementedError("inputs_embeds is not implemented for MPT yet")
        attention_mask = kwargs["attention_mask"].bool()
        if attention_mask[:, -1].sum() != attention_mask.shape[0]:
            raise NotImplementedError("MPT does not support generation with right padding.")
        if past_key_values is not None:
            input_ids = input_ids[:, -1].unsqueeze(-1)
        return {"input_ids": input_ids, "attention_mask": attention_mask, "prefix_mask": None, "sequence_id": None,
                "past_key_values": past_key_values, "use_cache": kwargs.get("use_cache", True),
                "audio_encodings": kwargs.get("audio_encodings", None)}

    def initialize_audio_tokenizer(self, mm_use_audio_start_end, tokenizer, device, tune_mm_mlp_adapter=False,
                                   pretrain_mm_mlp_adapter=None):
        """m2t/models/mpt.py:367-433: register the audio tokens, grow the (tied) table, start the two marker rows at the
        mean of the old rows, record the ids, set the trainable flags."""
        aec = self.get_model().audio_encoder_config
        aec.use_audio_start_end = mm_use_audio_start_end
        tokenizer.add_tokens([DEFAULT_AUDIO_PATCH_TOKEN], special_tokens=True)
        self.resize_token_embeddings(len(tokenizer))
        if mm_use_audio_start_end:
            n_new = tokenizer.add_tokens([DEFAULT_AUDIO_START_TOKEN, DEFAULT_AUDIO_END_TOKEN], special_tokens=True)
            self.resize_token_embeddings(len(tokenizer))
            aec.audio_start_token, aec.audio_end_token = tokenizer.convert_tokens_to_ids(
                [DEFAULT_AUDIO_START_TOKEN, DEFAULT_AUDIO_END_TOKEN])
            table = self.get_input_embeddings().weight.data
            if n_new > 0:
                # input and output embeddings are ONE tensor here; upstream assigns the (identical) mean twice
                table[-n_new:] = table[:-n_new].mean(dim=0, keepdim=True)
            if tune_mm_mlp_adapter:
                self.get_model().orig_embeds_params = [table.clone().to(device=device)]
                for p in self.get_input_embeddings().parameters():
                    p.requires_grad = True
                for p in self.get_output_embeddings().parameters():  # the same tensor: ends up frozen, as upstream
                    p.requires_grad = False
            if pretrain_mm_mlp_adapter:
                saved = torch.load(pretrain_mm_mlp_adapter, map_location="cpu")["transformer.wte.weight"]
                assert n_new == 2
                if table.shape == saved.shape:
                    table[-n_new:] = saved[-n_new:]
                elif saved.shape[0] == n_new:
                    table[-n_new:] = saved
                else:
                    raise ValueError(f"Unexpected embed_tokens_weight shape. Pretrained: {saved.shape}. Current: "
                                     f"{table.shape}. Numer of new tokens: {n_new}.")
        aec.audio_patch_token = tokenizer.convert_tokens_to_ids([DEFAULT_AUDIO_PATCH_TOKEN])[0]


def register():
    """m2t/models/mpt.py:436-437."""
    try:
        AutoConfig.register("wrapped_mpt", WrappedMPTConfig)
        AutoModelForCausalLM.register(WrappedMPTConfig, WrappedMPTForCausalLM)
    except ValueError:
        pass  # already registered (module imported twice under two names)


register()
