"""Host logic of SURVEY §8(a) rows J4 (get_cond: label rows, RangeEmbedding bins, conditioning tensors) and J7 (load_model
from upstream-format checkpoints: name map, "module." prefix, 36-of-72 non-strict load) on CPU, with tests/cpu_ops.py
standing in for the CUDA ops. The checker is oracle/jukebox_oracle.py's restatement of upstream's LabelConditioner
([UPSTREAM-RECALLED], parity unpinned as for the rest of the Jukebox network)."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))


@pytest.fixture()
def cpu_jukebox(monkeypatch):
    import cpu_ops
    import llark_b200.jukebox.main as jm
    import llark_b200.jukebox.model as jmodel

    monkeypatch.setattr(jmodel, "ops", cpu_ops)
    monkeypatch.setattr(jm, "ops", cpu_ops)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return jm, jmodel


def test_get_cond_matches_the_oracle_conditioner(cpu_jukebox):
    from oracle import jukebox_oracle as jo

    jm, jmodel = cpu_jukebox
    cfg = jmodel.JukeboxConfig(n_ctx=512, prior_width=64, prior_depth=1, heads=2, m_attn=0.5, blocks=16)
    ocfg = jo.JukeboxCfg(n_ctx=512, prior_width=64, prior_depth=1, heads=2, m_attn=0.5, blocks=16)
    state = jmodel.random_prior_state(cfg, seed=1, device="cpu")
    state.pop("x_cond"), state.pop("y_cond")
    tables = jo.init_label_params(ocfg, jo.LabelCfg(), seed=5)
    state.update(tables)
    top_prior = jmodel.JukeboxTopPrior(cfg, state, device="cpu", precision="tf32")
    hps = SimpleNamespace(sr=44100, n_samples=8)
    x_cond, y_cond = jm.get_cond(hps, top_prior)
    want_x, want_y = jo.get_cond(tables, ocfg, jo.LabelCfg(), n_samples=8)
    assert hps.sample_length == (62 * 44100 // 128) * 128                      # the side effect of jukebox/main.py:74-77
    assert tuple(x_cond.shape) == (1, 512, 64) and tuple(y_cond.shape) == (1, 1, 64)
    assert torch.equal(x_cond, want_x) and torch.equal(y_cond, want_y)
    # label rows: [total_length, offset, sample_length of this level, artist 0, genre 0]
    labels = top_prior.labeller.get_batch_labels([dict(artist="Unknown", genre="unknown", total_length=3_000_000,
                                                       offset=128 * 7, lyrics="")])
    y = top_prior.get_y(labels, 2)
    assert y.tolist() == [[3_000_000, 128 * 7 + 2 * 128, 512 * 128, 0, 0]]
    # bins are index work: bit-exact against the oracle for a second, non-default label row, mixed rows are not merged
    x2, y2, prime = top_prior.get_cond(None, torch.cat([y, top_prior.get_y(labels, 0)]))
    ox = jo.label_conditioner(tables, torch.cat([y, top_prior.get_y(labels, 0)]), ocfg, jo.LabelCfg())
    assert prime is None and torch.equal(x2, ox[1]) and torch.equal(y2, ox[0])
    with pytest.raises(AssertionError):                                        # RangeEmbedding's range check
        top_prior.get_cond(None, torch.tensor([[10, 0, 512 * 128, 0, 0]]))     # total length below min_duration


def test_load_model_from_upstream_format_checkpoints(cpu_jukebox, tmp_path):
    from oracle import jukebox_oracle as jo

    jm, jmodel = cpu_jukebox
    cfg = jmodel.JukeboxConfig(sample_length=128 * 64, l_bins=32, n_ctx=64, prior_width=32, prior_depth=2, heads=2,
                               m_attn=0.5, blocks=4)
    ocfg = jo.JukeboxCfg(sample_length=128 * 64, l_bins=32, n_ctx=64, prior_width=32, prior_depth=2, heads=2, m_attn=0.5,
                         blocks=4)
    g = torch.Generator().manual_seed(0)
    vq = jmodel.random_vqvae_state(cfg, seed=3)
    full = jmodel.JukeboxConfig(**{**cfg.__dict__, "prior_depth": 4})          # the file holds TWICE the blocks we build
    pr = jmodel.random_prior_state(full, seed=4, device="cpu")
    pr.pop("x_cond"), pr.pop("y_cond")
    pr.update(jo.init_label_params(ocfg, jo.LabelCfg(genre_bins=6, artist_bins=9, t_bins=8), seed=5))
    # write them under upstream's names (half precision transformer weights, "module." on the prior, unrelated extras)
    inv_v = {ours: theirs for theirs, ours in jo.upstream_vqvae_keys(ocfg).items()}
    up_v = {inv_v[k]: v for k, v in vq.items()}
    up_v["encoders.0.level_blocks.0.model.0.0.weight"] = torch.randn(64, 1, 4, generator=g)     # another level: ignored
    up_v["bottleneck.level_blocks.2.k_sum"] = torch.randn(32, 64, generator=g)
    ofull = jo.JukeboxCfg(**{**ocfg.__dict__, "prior_depth": 4})
    inv_p = {ours: theirs for theirs, ours in jo.upstream_prior_keys(ofull).items()}
    up_p = {"module." + inv_p[k]: (v.half() if ".w" in k and k.startswith("blk.") else v) for k, v in pr.items()}
    up_p["module.prior.x_out.weight"] = torch.randn(32, 32, generator=g)
    torch.save({"model": up_v, "step": 7}, tmp_path / "vqvae.pth.tar")
    torch.save({"model": up_p, "step": 9}, tmp_path / "prior_level_2.pth.tar")
    hps, vqvae, top_prior = jm.load_model("5b", cfg=cfg, vqvae_checkpoint=str(tmp_path / "vqvae.pth.tar"),
                                          prior_checkpoint=str(tmp_path / "prior_level_2.pth.tar"), device="cpu")
    assert hps.n_samples == 8 and hps.levels == 3 and vqvae.sample_length == cfg.sample_length
    for k, v in vq.items():
        assert torch.equal(vqvae.w[k], v), k
    w = top_prior.prior.w
    assert not any(k.startswith(("blk.2.", "blk.3.")) for k in w)              # blocks beyond prior_depth are dropped
    for k, v in pr.items():
        if k.startswith(("blk.2.", "blk.3.")):
            continue
        want = v.half().float() if (".w" in k and k.startswith("blk.")) else v
        got = w[k].t() if k.endswith((".c_attn.w", ".c_proj.w", ".fc.w", ".fc_proj.w")) else w[k]   # stored K-major
        assert torch.equal(got.float(), want), k
    assert top_prior.has_label_conditioner
    # a file that lacks tensors of the first 36 blocks is an error, not a silent random initialisation
    del up_p["module.prior.transformer._attn_mods.1.ln_1.bias"]
    torch.save({"model": up_p}, tmp_path / "broken.pth.tar")
    with pytest.raises(KeyError):
        jm.load_model("5b", cfg=cfg, vqvae_checkpoint=str(tmp_path / "vqvae.pth.tar"),
                      prior_checkpoint=str(tmp_path / "broken.pth.tar"), device="cpu")
    # the name maps of the product and of the oracle are the same table
    assert jmodel.vqvae_key_map(cfg) == jo.upstream_vqvae_keys(ocfg)
    assert jmodel.prior_key_map(cfg) == jo.upstream_prior_keys(ocfg)


def test_prior_3xtf32_schedule_matches_the_oracle(cpu_jukebox):
    """precision="3xtf32" (fp32-equivalent prior: every linear composed from TF32 halves, fp32 SIMT attention, separate
    quick-GELU): the host schedule over the CPU stand-ins reproduces the oracle's fp32 prior to rounding level."""
    from oracle import jukebox_oracle as jo

    jm, jmodel = cpu_jukebox
    ocfg = jo.tiny_cfg()
    cfg = jmodel.JukeboxConfig(sample_length=ocfg.sample_length, l_bins=ocfg.l_bins, n_ctx=ocfg.n_ctx,
                               prior_width=ocfg.prior_width, prior_depth=ocfg.prior_depth, heads=ocfg.heads,
                               m_attn=ocfg.m_attn, blocks=ocfg.blocks)
    pp = jo.init_prior_params(ocfg, seed=1)
    codes = torch.randint(0, ocfg.l_bins, (2, ocfg.n_ctx), generator=torch.Generator().manual_seed(9))
    ref = jo.prior_forward(codes, pp, ocfg)
    tp = jmodel.JukeboxTopPrior(cfg, pp, device="cpu", precision="3xtf32")
    tp.prior.only_encode = True
    out = tp.prior.forward(codes, x_cond=None, y_cond=None, encoder_kv=None, fp16=False)
    assert out.dtype == torch.float32 and ((out - ref).norm() / ref.norm()).item() < 1e-5
    with pytest.raises(ValueError):
        jmodel.PriorTransformer(cfg, pp, device="cpu", precision="fp64")
