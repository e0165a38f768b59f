"""Drop-in proof for the Jukebox front end at the reference's own functions: the REFERENCE's `jukebox/main.py`
(`get_z`, `get_cond`, `get_final_activations`, `get_acts_from_file`, imported in place) is run UNCHANGED over llark_b200's
model objects (`JukeboxVQVAE`, `JukeboxTopPrior`) -- i.e. the objects it normally gets from openai/jukebox's
make_vqvae / make_prior (jukebox/main.py:176-200) -- at the real geometry (1 048 576 samples, 8192 codes, width 4800,
block 64; one prior layer to keep the CPU run short) and compared with llark_b200's own `get_acts_from_file`.

CPU run: tests/cpu_ops.py stands in for the CUDA kernels and `torch.cuda.FloatTensor` for the CPU constructor (test
infrastructure only): this pins the object protocol (encode / get_cond / get_y / labeller / prior.forward / only_encode,
shapes and dtypes at every hand-over), not the kernels. Skipped where /root/reference is absent."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="the reference checkout is not present")


def test_reference_jukebox_main_runs_over_llark_objects(tmp_path, monkeypatch):
    import cpu_ops
    import llark_b200.jukebox.main as jm
    import llark_b200.jukebox.model as jmodel

    monkeypatch.setattr(jmodel, "ops", cpu_ops)
    monkeypatch.setattr(jm, "ops", cpu_ops)
    monkeypatch.setattr(torch.cuda, "FloatTensor", torch.FloatTensor, raising=False)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ref = ref_shims.load_reference_jukebox_main()

    cfg = jmodel.JukeboxConfig(prior_depth=1)
    vqvae = jmodel.JukeboxVQVAE(cfg, jmodel.random_vqvae_state(cfg, seed=0), device="cpu")
    # label tables instead of ready-made constants: the reference's get_cond (jukebox/main.py:71-98) then drives the
    # labeller -> get_y -> get_cond protocol for real
    state = jmodel.random_prior_state(cfg, seed=1, device="cpu")
    state.pop("x_cond"), state.pop("y_cond")
    state.update(jmodel.random_label_state(cfg, seed=2, device="cpu"))
    top_prior = jmodel.JukeboxTopPrior(cfg, state, device="cpu", precision="tf32")
    hps = SimpleNamespace(sr=44100, n_samples=8, name="samples", levels=3, hop_fraction=[0.5, 0.5, 0.125])

    rng = np.random.default_rng(0)
    n = 20 * 44100  # a 20 s clip: shorter than the 23.8 s window, so the padded tail must be dropped
    wave = rng.uniform(-1, 1, size=n).astype(np.float32)
    fpath = str(tmp_path / "clip.npy")
    np.save(fpath, wave)
    # the reference decodes with librosa (absent); hand it the same normalised mono waveform llark_b200 reads
    monkeypatch.setattr(ref, "load_audio_from_file", lambda p: jm.load_audio_from_file(p))

    with torch.no_grad():
        got_ref = ref.get_acts_from_file(fpath, hps, vqvae, top_prior, meanpool=True, pool_frames_per_second=10)
        got_own = jm.get_acts_from_file(fpath, hps, vqvae, top_prior, meanpool=True, pool_frames_per_second=10)
    latent = int(np.floor(8192 * n / 1048576))
    assert got_ref.shape == (latent // 34, 4800) and got_ref.dtype == np.float32
    assert got_own.shape == got_ref.shape
    assert np.allclose(got_own, got_ref, rtol=1e-5, atol=1e-6)
    assert hps.sample_length == (62 * 44100 // 128) * 128          # get_cond's side effect on hps (:74-77)
    assert top_prior.prior.only_encode is True
    # the individual hand-overs, through the reference's functions
    audio = ref.maybe_pad_audio_to_max_len(jm.load_audio_from_file(fpath))
    z = ref.get_z(audio, vqvae)
    assert tuple(z.shape) == (1, 8192) and z.dtype == torch.int64 and int(z.max()) < 2048
    x_cond, y_cond = ref.get_cond(hps, top_prior)
    assert tuple(x_cond.shape) == (1, 8192, 4800) and tuple(y_cond.shape) == (1, 1, 4800)
    acts = ref.get_final_activations(z, x_cond, y_cond, top_prior)
    assert tuple(acts.shape) == (1, 8192, 4800) and acts.dtype == torch.float32
    with pytest.raises(AssertionError):
        ref.get_z(audio[:1000], vqvae)                               # shorter than the expected window (:56-58)
