"""ORACLE support (test infrastructure only): golden vectors for the generation plumbing from the
REFERENCE's own code imported in place: m2t/generate.py (KeywordsStoppingCriteria), m2t/conversation_utils.py,
m2t/tokenizer.py (get_prompt_end_token_sequence), m2t/infer.py (infer_with_prompt up to the model.generate call), and
greedy token sequences of the reference's WrappedLlamav2ForCausalLM.generate (transformers GenerationMixin) on the tiny
seeded model of oracle/llama_oracle.py. Output: tests/golden/generation_ref.pt.

transformers 5.5 note: the reference's prepare_inputs_for_generation (`if past_key_values:`) predates the Cache classes;
with use_cache=True the prefill is truncated to the last token and the first generated token is wrong (6 instead of
the forward pass's argmax 31 on the fixture). use_cache=False is the faithful run (identical to 4.29's cached result
by construction) and is what the fixture records."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_tokenizer(no_bos=False):
    from synth_tokenizer import SynthTokenizer

    class NoBos(SynthTokenizer):
        def __call__(self, text, **kw):
            kw.setdefault("add_special_tokens", False)
            return super().__call__(text, **kw)

    tok = (NoBos if no_bos else SynthTokenizer)(vocab_size=300, model_max_length=400)
    tok.add_special_tokens({"pad_token": "[PAD]"})
    tok.add_tokens(["<audio_patch>", "<audio_start>", "<audio_end>"], special_tokens=True)
    return tok


def run_host_logic(gen_mod, conv_mod, tok_mod, infer_mod):
    """The host-side pieces, on the reference's modules or the mirror's."""
    out = {}
    # --- KeywordsStoppingCriteria
    tok = make_tokenizer()
    prompt = tok("### Human: what key is this?\n### Assistant:", return_tensors="pt").input_ids
    crit = gen_mod.KeywordsStoppingCriteria(["###"], tok, prompt)
    cont_a = tok(" C minor, fairly slow", return_tensors="pt", add_special_tokens=False).input_ids
    cont_b = tok(" C minor.\n###", return_tensors="pt", add_special_tokens=False).input_ids
    out["stop_keyword_ids"] = list(crit.keyword_ids)
    out["stop_a"] = bool(crit(torch.cat([prompt, cont_a], 1), None))
    out["stop_b"] = bool(crit(torch.cat([prompt, cont_b], 1), None))
    out["stop_start_len"] = crit.start_len
    tok2 = make_tokenizer(no_bos=True)
    p2 = tok2("say stop now", return_tensors="pt").input_ids
    crit2 = gen_mod.KeywordsStoppingCriteria(["stop", "two words"], tok2, p2)
    out["stop2_keyword_ids"] = list(crit2.keyword_ids)
    out["stop2_fast"] = bool(crit2(torch.cat([p2, tok2(" please stop", return_tensors="pt").input_ids], 1), None))
    out["stop2_no"] = bool(crit2(torch.cat([p2, tok2(" go on", return_tensors="pt").input_ids], 1), None))
    # --- conversation_utils
    out["subseq"] = [conv_mod.subsequence_pos([1, 2, 3, 4, 5, 6], [3, 4]), conv_mod.subsequence_pos([1, 2, 3, 4], [3, 4]),
                     conv_mod.subsequence_pos([1, 2, 3], [9])]
    ids = torch.tensor([5, 6, 7, 8, 9, 10, 11])
    out["prompt_tokens"] = conv_mod.extract_prompt_tokens(ids, [7, 8])
    out["response_tokens"] = conv_mod.extract_response_tokens(ids, [7, 8])
    out["example"] = conv_mod.make_example("id0", "a.wav", None, [1, 2], "q?", "a.")
    # --- tokenizer.get_prompt_end_token_sequence
    out["end_seq_llama"] = tok_mod.get_prompt_end_token_sequence(tok, "meta-llama/Llama-2-7b-chat-hf")
    out["end_seq_other"] = tok_mod.get_prompt_end_token_sequence(tok, "some/other-model")
    # --- infer_with_prompt up to model.generate
    captured = {}

    class Capture:
        def generate(self, **kw):
            captured.update(kw)
            return "generated"

    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        enc = torch.arange(4 * 6, dtype=torch.float32).reshape(4, 6) / 10
        end_seq = tok_mod.get_prompt_end_token_sequence(tok, "meta-llama/Llama-2-7b-chat-hf")
        cfg = dict(is_multimodal=True, sep_audio_conv_front=False, use_audio_start_end=True)
        for audio_first in (False, True):
            captured.clear()
            res = infer_mod.infer_with_prompt("Describe the audio.", Capture(), enc, end_seq, cfg, tok, example_id="ex1",
                                              audio_first=audio_first, max_new_tokens=7)
            sc = captured.pop("stopping_criteria")
            out[f"infer_first{int(audio_first)}"] = {
                "result": res, "kwargs": {k: v for k, v in captured.items()}, "n_criteria": len(sc),
                "keywords": list(sc[0].keywords), "crit_input_ids": sc[0].input_ids}
    finally:
        torch.Tensor.cuda = old_cuda
    return out


GEN_CASES = [dict(name="plain", max_new_tokens=8), dict(name="eos", max_new_tokens=8, eos_from_step=2),
             dict(name="criteria", max_new_tokens=8, stop_on_step=3), dict(name="max_length", max_length=20)]


def gen_inputs():
    from oracle import llama_oracle as lo

    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=1, S=24, F_frames=5, seed=3, pad_last=0)
    return cfg, params, batch["input_ids"][:, :16].clone(), batch["audio_encodings"]


class StopOnToken:
    """A stopping criterion in the StoppingCriteria calling convention (ids, scores) -> bool."""

    def __init__(self, token):
        self.token = int(token)

    def __call__(self, input_ids, scores, **kw):
        return bool(input_ids[0, -1] == self.token)


def golden_generation():
    from oracle import ref_shims

    ref_shims.install()
    import m2t.conversation_utils as ref_conv
    import m2t.generate as ref_gen
    import m2t.infer as ref_infer
    import m2t.tokenizer as ref_tok
    from oracle.make_golden import build_reference_model
    import transformers

    out = {"host": run_host_logic(ref_gen, ref_conv, ref_tok, ref_infer)}
    cfg, params, ids, enc = gen_inputs()
    model = build_reference_model(cfg, params, torch.float32).eval()
    plain = model.generate(input_ids=ids, audio_encodings=enc, max_new_tokens=8, do_sample=False, use_cache=False)
    new = plain[0, ids.shape[1]:].tolist()
    seqs = {"plain": plain}
    seqs["eos"] = model.generate(input_ids=ids, audio_encodings=enc, max_new_tokens=8, do_sample=False, use_cache=False,
                                 eos_token_id=new[2])

    class HFStop(transformers.StoppingCriteria):
        def __call__(self, input_ids, scores, **kw):
            return bool(input_ids[0, -1] == new[3])

    seqs["criteria"] = model.generate(input_ids=ids, audio_encodings=enc, max_new_tokens=8, do_sample=False,
                                      use_cache=False, stopping_criteria=[HFStop()])
    seqs["max_length"] = model.generate(input_ids=ids, audio_encodings=enc, max_length=20, do_sample=False, use_cache=False)
    out["generate"] = {"input_ids": ids, "sequences": seqs, "new_tokens": new}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "generation_ref.pt"))
    print("generation golden:", {k: v.shape for k, v in seqs.items()}, new)


if __name__ == "__main__":
    golden_generation()
