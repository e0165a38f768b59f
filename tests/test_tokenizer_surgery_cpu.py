"""SURVEY §8(a) row M3 / Appendix D.2 against the reference, side by side: the reference's own
`smart_tokenizer_and_embedding_resize` (m2t/utils.py:60-80) + `WrappedLlamav2ForCausalLM.initialize_audio_tokenizer`
(m2t/models/llamav2.py:367-419), imported in place, and llark_b200's mirrors run on copies of the same tiny checkpoint
and tokenizer with the same torch seed (transformers' resize draws the not-mean-initialised rows at random). Compared:
both embedding matrices element for element ([PAD] and the two marker rows = mean of the rows before them; the
<audio_patch> row is whatever the resize left), token ids, the recorded config ids, orig_embeds_params, trainable flags.
Skipped where /root/reference is absent."""
import copy
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="the reference checkout is not present")


@pytest.mark.parametrize("tune", [True, False])
def test_pad_and_audio_token_surgery_matches_reference(tmp_path, tune):
    from transformers import AutoTokenizer

    import llark_b200.models.llamav2 as ours_mod
    import llark_b200.utils as ours_utils
    from test_reference_train_dropin_cpu import _tiny_checkpoint

    ref_shims.install()
    import m2t.models.llamav2 as ref_mod
    import m2t.utils as ref_utils

    ck = str(tmp_path / "tiny-llama-ckpt")
    base_vocab = _tiny_checkpoint(ck, random.Random(0))
    sides = {}
    for name, mod, utils in (("ref", ref_mod, ref_utils), ("ours", ours_mod, ours_utils)):
        tok = AutoTokenizer.from_pretrained(ck)
        model = mod.WrappedLlamav2ForCausalLM.from_pretrained(ck, mm_hidden_size=12)
        model.get_model().initialize_adapter_modules(tune_mm_mlp_adapter=tune)
        torch.manual_seed(1234)
        utils.smart_tokenizer_and_embedding_resize(special_tokens_dict=dict(pad_token="[PAD]"), tokenizer=tok, model=model)
        after_pad = copy.deepcopy(model.get_input_embeddings().weight.data)
        model.initialize_audio_tokenizer(mm_use_audio_start_end=True, tokenizer=tok, device="cpu",
                                         tune_mm_mlp_adapter=tune, pretrain_mm_mlp_adapter=None)
        sides[name] = (tok, model, after_pad)
    (tr, mr, pr), (to, mo, po) = sides["ref"], sides["ours"]
    assert len(tr) == len(to) == base_vocab + 4
    specials = ["[PAD]", "<audio_patch>", "<audio_start>", "<audio_end>"]
    assert tr.convert_tokens_to_ids(specials) == to.convert_tokens_to_ids(specials) == list(range(base_vocab, base_vocab + 4))
    assert torch.equal(pr, po)
    # [PAD] row = mean of the original rows (m2t/utils.py:72-80)
    assert torch.allclose(po[-1], po[:-1].mean(0), atol=1e-6)
    for get in ("get_input_embeddings", "get_output_embeddings"):
        wr, wo = getattr(mr, get)().weight.data, getattr(mo, get)().weight.data
        assert wr.shape == wo.shape == (base_vocab + 4, mr.config.hidden_size)
        assert torch.equal(wr, wo), get
        # the two marker rows = mean of everything before them, <audio_patch> included (llamav2.py:396-406)
        assert torch.allclose(wo[-1], wo[:-2].mean(0), atol=1e-6) and torch.equal(wo[-1], wo[-2])
    ar, ao = mr.get_model().audio_encoder_config, mo.get_model().audio_encoder_config
    for f in ("audio_patch_token", "audio_start_token", "audio_end_token", "use_audio_start_end"):
        assert getattr(ar, f) == getattr(ao, f), f
    assert mr.config.vocab_size == mo.config.vocab_size
    for (nr, p_r), (no, p_o) in zip(mr.named_parameters(), mo.named_parameters()):
        assert nr == no and p_r.requires_grad == p_o.requires_grad, nr
    has_r = getattr(mr.get_model(), "orig_embeds_params", None)
    has_o = getattr(mo.get_model(), "orig_embeds_params", None)
    assert (has_r is None) == (has_o is None) == (not tune)
    if tune:
        assert torch.equal(has_r[0], has_o[0])
        assert not any(p.requires_grad for p in mo.get_output_embeddings().parameters())
        assert all(p.requires_grad for p in mo.get_input_embeddings().parameters())


def test_projector_initialisation_and_pretrained_load_match_reference(tmp_path):
    """M2 (m2t/models/llamav2.py:60-93): nn.Linear(mm_hidden_size, hidden_size) with bias, created once, filled from the
    `*mm_projector*` keys of a side file such as WrappedTrainer._save writes (other keys in the file are ignored)."""
    import llark_b200.models.llamav2 as ours_mod
    from test_reference_train_dropin_cpu import _tiny_checkpoint

    ref_shims.install()
    import m2t.models.llamav2 as ref_mod

    ck = str(tmp_path / "tiny-llama-ckpt")
    _tiny_checkpoint(ck, random.Random(0))
    g = torch.Generator().manual_seed(5)
    side = {"model.mm_projector.weight": torch.randn(256, 12, generator=g), "model.mm_projector.bias": torch.randn(256, generator=g),
            "model.embed_tokens.weight": torch.randn(8, 256, generator=g)}
    path = str(tmp_path / "mm_projector.bin")
    torch.save(side, path)
    out = {}
    for name, mod in (("ref", ref_mod), ("ours", ours_mod)):
        model = mod.WrappedLlamav2ForCausalLM.from_pretrained(ck, mm_hidden_size=12)
        assert not hasattr(model.get_model(), "mm_projector")
        res = model.get_model().initialize_adapter_modules(pretrain_mm_mlp_adapter=path, tune_mm_mlp_adapter=True, fsdp=None)
        first = model.get_model().mm_projector
        model.get_model().initialize_adapter_modules()                      # second call keeps the module
        assert model.get_model().mm_projector is first
        out[name] = (model, res)
    (mr, rr), (mo, ro) = out["ref"], out["ours"]
    assert sorted(rr) == sorted(ro) == ["audio_config"]
    assert type(rr["audio_config"]).__name__ == type(ro["audio_config"]).__name__ == "AudioEncoderConfig"
    # (upstream `audio_patch_token = str = ...` makes that one a class attribute; attribute access is what callers use)
    for f in ("use_audio_start_end", "audio_start_token", "audio_end_token", "audio_patch_token"):
        assert getattr(rr["audio_config"], f) == getattr(ro["audio_config"], f), f
    assert mr.config.use_mm_proj is True and mo.config.use_mm_proj is True
    for k in ("weight", "bias"):
        a, b = getattr(mr.get_model().mm_projector, k).data, getattr(mo.get_model().mm_projector, k).data
        assert torch.equal(a, b) and torch.equal(b, side["model.mm_projector." + k])
    assert set(mr.state_dict()) == set(mo.state_dict())                     # same key names incl. model.mm_projector.*
