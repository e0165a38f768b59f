"""CPU tests: the oracle restatement (oracle/llama_oracle.py) against the golden vectors produced by the
REFERENCE's own classes (oracle/make_golden.py imported m2t.models.llamav2 / jukebox.main from /root/reference).
The oracle is test infrastructure; these tests are what pins it before the CUDA path is compared with it."""
import os

import pytest
import torch

from oracle import llama_oracle as lo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "llama_tiny_ref.pt"), weights_only=False)


@pytest.mark.parametrize("detach", [True, False])
def test_oracle_fp32_matches_reference(gold, detach):
    cfg = lo.LlamaCfg(**gold["cfg"])
    params = lo.init_params(cfg, seed=gold["seed"])
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    loss, logits, grads = lo.forward_backward(params, cfg, batch, detach_text=detach)
    ref = gold[f"fp32_{'detach' if detach else 'nodetach'}"]
    assert abs(loss.item() - ref["loss"].item()) < 1e-5
    assert _rel(logits, ref["logits"]) < 1e-5
    for n, entry in ref["grads"].items():
        g = grads[n]
        if entry is None:
            assert g is None, n  # frozen lm_head
            continue
        assert abs(g.norm().item() - entry["norm"]) <= 1e-4 * max(entry["norm"], 1e-6), n
        if "full" in entry:
            assert _rel(g, entry["full"]) < 1e-4, n
        if "head" in entry:
            assert _rel(g.flatten()[:2048], entry["head"]) < 1e-4, n
        if "nonzero_rows" in entry:
            rows = g.abs().sum(-1).nonzero().flatten().tolist()
            assert rows == entry["nonzero_rows"], n
            if detach:  # only the two marker rows receive gradient
                assert rows == [cfg.audio_start_token, cfg.audio_end_token]


@pytest.mark.parametrize("detach", [True, False])
def test_oracle_bf16_matches_reference(gold, detach):
    """bf16 weights + bf16 activations: the regime the reference trains in. Two bf16 pipelines agree to a few
    bf16 ulps per op; tolerance 2e-2 relative Frobenius on logits/grads, 2e-3 on the loss."""
    cfg = lo.LlamaCfg(**gold["cfg"])
    params = lo.init_params(cfg, seed=gold["seed"], dtype=torch.bfloat16)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    loss, logits, grads = lo.forward_backward(params, cfg, batch, detach_text=detach)
    ref = gold[f"bf16_{'detach' if detach else 'nodetach'}"]
    assert abs(loss.item() - ref["loss"].item()) < 2e-3 * ref["loss"].item()
    assert _rel(logits, ref["logits"]) < 2e-2
    for n, entry in ref["grads"].items():
        if entry is None:
            assert grads[n] is None
            continue
        assert abs(grads[n].float().norm().item() - entry["norm"]) <= 5e-2 * max(entry["norm"], 1e-6), n


def test_reference_value_errors(gold):
    """llamav2.py:146-152: unbalanced start/end tokens raise ValueError; the oracle raises the same."""
    assert gold["value_errors"] and "start tokens and image end tokens" in gold["value_errors"][0]
    cfg = lo.LlamaCfg(**gold["cfg"])
    params = lo.init_params(cfg, seed=0)
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=0)
    batch["input_ids"][0, 4 + 5 + 1] = 7
    with pytest.raises(ValueError):
        lo.forward(params, cfg, batch["input_ids"], batch["attention_mask"], batch["labels"], batch["audio_encodings"])


def test_pool_oracle_matches_reference_windowed_average():
    """jukebox/main.py:113-130 (AvgPool1d(34, 34), floor) and the frame arithmetic of :10-22,157."""
    from oracle import jukebox_oracle as jo

    ref = torch.load(os.path.join(GOLD, "pool_ref.pt"), weights_only=False)
    assert ref["frame_len"] == jo.frame_len(10) == 34
    assert ref["full_len_frames"] == 240
    g = torch.Generator().manual_seed(ref["seed"])
    acts = torch.randn(ref["T_valid"], 4800, generator=g)
    pooled = jo.windowed_average(acts, ref["frame_len"])
    assert tuple(pooled.shape) == ref["shape"]
    assert torch.allclose(pooled[:, :64], ref["pooled_head"], atol=1e-6)
    assert abs(pooled.double().sum().item() - ref["pooled_sum"]) < 1e-3
