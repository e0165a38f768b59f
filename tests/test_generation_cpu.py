"""CPU parity tests of the generation plumbing against golden vectors from the REFERENCE's own
m2t/generate.py, m2t/conversation_utils.py, m2t/tokenizer.py, m2t/infer.py and WrappedLlamav2ForCausalLM.generate
(tests/golden/generation_ref.pt, oracle/make_golden_generation.py). The decoding loop is driven here by the oracle's CPU
forward (test infrastructure); tests/test_model_gpu.py drives the same loop with the CUDA engine."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "generation_ref.pt"), weights_only=False)


def test_host_logic_bit_exact_against_reference(gold):
    from oracle.make_golden_generation import run_host_logic
    from test_formats_cpu import _same
    import llark_b200.conversation_utils as conv
    import llark_b200.generate as gen
    import llark_b200.infer as infer
    import llark_b200.tokenizer as tokm

    got = run_host_logic(gen, conv, tokm, infer)
    _same(got, gold["host"])
    assert got["stop_keyword_ids"] == [] and got["stop_a"] is False and got["stop_b"] is True
    assert len(got["stop2_keyword_ids"]) == 1 and got["stop2_fast"] is True and got["stop2_no"] is False
    assert got["subseq"] == [(2, 4), None, None]          # a match flush with the end is missed, as upstream
    assert len(got["end_seq_llama"]) == len(got["end_seq_other"]) - 1
    k = got["infer_first0"]["kwargs"]
    assert sorted(k) == ["audio_encodings", "example_id", "input_ids", "labels", "max_new_tokens"]
    assert k["input_ids"].dim() == 2 and k["audio_encodings"].shape == (1, 4, 6)


def test_greedy_loop_matches_reference_generate(gold):
    """greedy_generate + the oracle's CPU forward reproduce GenerationMixin.generate of the reference model:
    plain, eos (emitted then stop), stopping criterion, max_length."""
    from oracle import llama_oracle as lo
    from oracle.make_golden_generation import StopOnToken, gen_inputs
    from llark_b200.generate import greedy_generate

    cfg, params, ids, enc = gen_inputs()
    g = gold["generate"]
    assert torch.equal(ids, g["input_ids"])
    new = g["new_tokens"]

    def step(x):
        return lo.forward(params, cfg, x, None, None, enc)[1][:, -1, :]

    assert torch.equal(greedy_generate(step, ids, 8, eos_token_id=[2]), g["sequences"]["plain"])
    assert torch.equal(greedy_generate(step, ids, 8, eos_token_id=[new[2]]), g["sequences"]["eos"])
    assert torch.equal(greedy_generate(step, ids, 8, eos_token_id=[2], stopping_criteria=[StopOnToken(new[3])]),
                       g["sequences"]["criteria"])
    assert torch.equal(greedy_generate(step, ids, 20 - ids.shape[1], eos_token_id=[2]), g["sequences"]["max_length"])


def test_greedy_loop_batch_padding_and_limits():
    """Rows that emitted eos are padded while the others continue; zero budget returns the prompt."""
    from llark_b200.generate import greedy_generate

    V = 10
    script = {0: [3, 9, 4, 4], 1: [5, 6, 9, 7]}  # row -> next tokens; 9 = eos

    def step(ids):
        t = ids.shape[1] - 2
        out = torch.full((2, V), -1.0)
        for r in (0, 1):
            out[r, script[r][t]] = 1.0
        return out

    ids = torch.tensor([[1, 2], [1, 2]])
    out = greedy_generate(step, ids, 4, eos_token_id=[9], pad_token_id=0)
    assert out.tolist() == [[1, 2, 3, 9, 0], [1, 2, 5, 6, 9]]
    assert torch.equal(greedy_generate(step, ids, 0, eos_token_id=[9]), ids)
    with pytest.raises(ValueError):
        greedy_generate(step, ids[0], 2)
