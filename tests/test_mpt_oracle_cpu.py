"""Pins oracle/mpt_oracle.py (the CPU restatement of the CLAP + MPT-1B variant) to the REFERENCE's
own WrappedMPTForCausalLM: tests/golden/mpt_tiny_ref.pt holds its loss, logits and gradients on seeded inputs
(oracle/make_golden_mpt.py). fp32: 1e-5 on loss / logits, 1e-4 on gradients; bf16 autocast: 2e-2 / 6e-2 (rounding-order
noise of two bf16 pipelines). No product code is involved yet -- this is the oracle gate of that row."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "mpt_tiny_ref.pt"), weights_only=False)


@pytest.mark.parametrize("F_frames,bf16,detach", [(1, False, False), (1, True, True), (4, False, True), (4, True, False)])
def test_mpt_oracle_matches_reference(gold, F_frames, bf16, detach):
    from oracle import mpt_oracle as mo

    cfg = mo.MPTCfg(**gold["cfg"])
    params = mo.init_params(cfg, seed=0)
    batch = mo.make_batch(cfg, B=2, S=40, F_frames=F_frames, seed=F_frames)
    ref = gold["runs"][(F_frames, bf16, detach)]
    loss, logits, grads = mo.forward_backward(params, cfg, batch, detach_text=detach,
                                              autocast_dtype=torch.bfloat16 if bf16 else None)
    tol_l, tol_g = (2e-2, 6e-2) if bf16 else (1e-5, 1e-4)
    assert abs(loss.item() - ref["loss"].item()) <= tol_l * abs(ref["loss"].item())
    valid = batch["attention_mask"]
    assert _rel(logits.float()[valid], ref["logits"][valid]) < tol_l
    for name, g_ref in ref["grads"].items():
        g = grads[name]
        if g_ref is None:
            assert g is None or float(g.abs().max()) == 0.0, name
            continue
        assert _rel(g.float(), g_ref) < tol_g, name
    # tied table: even with the text rows detached, wte still receives the lm-head gradient on every row
    assert float(ref["grads"]["transformer.wte.weight"].abs().sum(1).min()) > 0


def test_mpt_oracle_value_errors_and_alibi_slopes():
    from oracle import mpt_oracle as mo

    cfg = mo.MPTCfg()
    params = mo.init_params(cfg, seed=0)
    batch = mo.make_batch(cfg, B=2, S=40, F_frames=4, seed=0)
    bad = batch["input_ids"].clone()
    bad[0, 3 + 4 + 1] = 7
    with pytest.raises(ValueError):
        mo.forward(params, cfg, bad, batch["attention_mask"], None, batch["audio_encodings"])
    assert torch.allclose(mo.gen_slopes(8), torch.tensor([2.0 ** -(i + 1) for i in range(8)]))
    assert mo.gen_slopes(6).shape == (6,)   # non-power-of-two head counts interleave (attention.py:462-464)
