"""B200-native mirror of the reference's jukebox/main.py: same function names, argument meaning, error
behaviour and constants; the network arithmetic runs on llark_b200's sm_100a kernels (see model.py).

Reference lines are cited per function (paths relative to the reference root).
"""
import logging
import os
from math import floor

import numpy as np
import torch

from llark_b200 import ops
from llark_b200.jukebox.model import (JukeboxConfig, JukeboxTopPrior, JukeboxVQVAE, random_prior_state,
                                      random_vqvae_state)

JUKEBOX_SAMPLE_RATE = 44100
T = 8192  # time dimension of jukebox activations
JUKEBOX_EXPECTED_SAMPLES_LEN = 1048576
JUKEBOX_SAMPLE_SECONDS = JUKEBOX_EXPECTED_SAMPLES_LEN / JUKEBOX_SAMPLE_RATE
ACTS_SAMPLE_RATE = T / JUKEBOX_SAMPLE_SECONDS  # 344.53 (jukebox/main.py:10-22)


class EmptyFileError(ValueError):
    pass


def load_audio_from_file(fpath) -> np.ndarray:
    """jukebox/main.py:29-45: decode, resample to 44.1 kHz, channel mean, peak-normalise.
    The reference decodes with librosa (not in this image); this mirror reads .npy arrays and PCM .wav files
    (scipy) and resamples with a polyphase filter when the rate differs. An ndarray is accepted as-is."""
    if isinstance(fpath, np.ndarray):
        audio = fpath.astype(np.float32)
    else:
        path = os.fspath(fpath)
        try:
            if path.endswith(".npy"):
                audio = np.load(path).astype(np.float32)
            else:
                from scipy.io import wavfile

                sr, data = wavfile.read(path)
                if data.size == 0:
                    raise ValueError("empty audio")
                if np.issubdtype(data.dtype, np.integer):
                    data = data.astype(np.float32) / float(np.iinfo(data.dtype).max)
                audio = data.astype(np.float32).T if data.ndim == 2 else data.astype(np.float32)
                if sr != JUKEBOX_SAMPLE_RATE:
                    from math import gcd

                    from scipy.signal import resample_poly

                    g = gcd(int(sr), JUKEBOX_SAMPLE_RATE)
                    audio = resample_poly(audio, JUKEBOX_SAMPLE_RATE // g, int(sr) // g, axis=-1).astype(np.float32)
        except ValueError as ve:
            raise EmptyFileError(f"file {fpath} failed to read with exception {ve!r}; it is probably empty.")
    if audio.ndim == 1:
        audio = audio[np.newaxis]
    audio = audio.mean(axis=0)
    norm_factor = np.abs(audio).max() if audio.size else 0.0
    if norm_factor > 0:
        audio = audio / norm_factor
    return audio.flatten()


def maybe_pad_audio_to_max_len(audio: np.ndarray) -> np.ndarray:
    """jukebox/main.py:48-51."""
    if len(audio) < JUKEBOX_EXPECTED_SAMPLES_LEN:
        audio = np.pad(audio, (0, JUKEBOX_EXPECTED_SAMPLES_LEN - len(audio)))
    return audio


def get_z(audio, vqvae):
    """jukebox/main.py:54-68: VQ-VAE top-level codes of the first 1 048 576 samples; int64 [1, T]."""
    expected = getattr(vqvae, "sample_length", JUKEBOX_EXPECTED_SAMPLES_LEN)
    assert len(audio) >= expected, f"expected samples with shape {expected}; got shape {audio.shape}."
    audio = audio[:expected]
    x = torch.from_numpy(np.ascontiguousarray(audio[np.newaxis, :, np.newaxis], dtype=np.float32)).cuda()
    zs = vqvae.encode(x)
    z = zs[-1].flatten()[np.newaxis, :]
    if z.shape[-1] < min(T, vqvae.cfg.n_ctx):
        raise ValueError("Audio file is not long enough")
    return z


def get_cond(hps, top_prior):
    """jukebox/main.py:71-98: conditioning for the fixed metas (artist/genre "unknown", 62 s, offset 0), computed for
    hps.n_samples label rows of which row 0 is kept."""
    sample_length_in_seconds = 62
    hps.sample_length = (int(sample_length_in_seconds * hps.sr) // top_prior.raw_to_tokens) * top_prior.raw_to_tokens
    metas = [dict(artist="unknown", genre="unknown", total_length=hps.sample_length, offset=0,
                  lyrics="""lyrics go here!!!""")] * hps.n_samples
    labels = [None, None, top_prior.labeller.get_batch_labels(metas, "cuda")]
    x_cond, y_cond, prime = top_prior.get_cond(None, top_prior.get_y(labels[-1], 0))
    n_ctx = top_prior.cfg.n_ctx
    x_cond = x_cond[0, :n_ctx][np.newaxis, ...]
    y_cond = y_cond[0][np.newaxis, ...]
    return x_cond, y_cond


def get_final_activations(z, x_cond, y_cond, top_prior):
    """jukebox/main.py:101-110."""
    x = z[:, : top_prior.cfg.n_ctx]
    top_prior.prior.only_encode = True
    return top_prior.prior.forward(x, x_cond=x_cond, y_cond=y_cond, encoder_kv=None, fp16=False)


def windowed_average(acts, frame_len: int, ceil_mode=False):
    """jukebox/main.py:113-130: AvgPool1d(frame_len, stride=frame_len) over time; [T, W] -> [1, T//frame_len, W]."""
    assert acts.ndim == 2, "expected 2d inputs"
    if ceil_mode:
        raise NotImplementedError("the reference only ever calls windowed_average with ceil_mode=False")
    acts = acts.contiguous().float()
    return ops.jb_pool(acts, acts.shape[0], frame_len)[None]


def get_acts_from_file(fpath, hps, vqvae, top_prior, meanpool=True, pool_frames_per_second=None):
    """jukebox/main.py:133-173 -> np.ndarray [n_frames, W] (or [W] when pooling over all timesteps)."""
    audio = load_audio_from_file(fpath)
    n_ctx = top_prior.cfg.n_ctx
    expected = vqvae.sample_length
    input_audio_len = len(audio)
    latent_audio_len = floor(n_ctx * input_audio_len / expected)
    if len(audio) < expected:
        audio = np.pad(audio, (0, expected - len(audio)))
    z = get_z(audio, vqvae)
    x_cond, y_cond = get_cond(hps, top_prior)
    acts = get_final_activations(z, x_cond, y_cond, top_prior)
    acts = acts.squeeze(0).type(torch.float32)
    acts = acts[:latent_audio_len, :]
    if meanpool:
        logging.warning(f"mean pooling at f={pool_frames_per_second}")
        if not pool_frames_per_second:
            acts = ops.jb_pool(acts.contiguous(), acts.shape[0], acts.shape[0])[0]
        else:
            acts_rate = n_ctx / (expected / JUKEBOX_SAMPLE_RATE)
            frame_len = floor(acts_rate / pool_frames_per_second)
            acts = windowed_average(acts, frame_len)
            acts = torch.squeeze(acts, 0)
    acts = np.array(acts.cpu())
    logging.info(f"acts after pooling has shape {acts.shape}")
    return acts


def embed_batch(waveforms, vqvae, top_prior, pool_frames_per_second=10):
    """Batched form of get_acts_from_file for waveforms already in memory (BASELINE configs[1]: batch 8):
    fp32 [N, n_samples] (already mono) -> device tensor fp32 [N, n_frames, W]. Peak normalisation, padding /
    truncation, encoding, prior and pooling follow the per-file function above."""
    cfg = top_prior.cfg
    expected = vqvae.sample_length
    w = torch.as_tensor(waveforms, dtype=torch.float32)
    peak = w.abs().amax(dim=1, keepdim=True)
    w = torch.where(peak > 0, w / peak.clamp_min(1e-30), w)
    n = w.shape[1]
    latent_len = floor(cfg.n_ctx * n / expected)
    if n < expected:
        w = torch.nn.functional.pad(w, (0, expected - n))
    w = w[:, :expected].cuda()
    z = vqvae.encode(w[:, :, None])[-1]
    top_prior.prior.only_encode = True
    x_cond, y_cond = top_prior.default_cond()  # the fixed metas of get_cond, computed once and cached
    acts = top_prior.prior.forward(z[:, : cfg.n_ctx], x_cond=x_cond, y_cond=y_cond, encoder_kv=None, fp16=False)
    valid = min(latent_len, cfg.n_ctx)
    frame_len = floor((cfg.n_ctx / (expected / JUKEBOX_SAMPLE_RATE)) / pool_frames_per_second)
    return pool_batch(acts, valid, frame_len)


def pool_batch(acts, valid, frame_len):
    """windowed_average (jukebox/main.py:113-130) of the first `valid` positions of every clip of acts [N, T, W] in one
    launch (no per-clip copies): -> fp32 [N, valid // frame_len, W]."""
    return ops.jb_pool_batched(acts.float().contiguous(), valid, frame_len)


def load_model(model="5b", state=None, cfg=None, precision="tf32", seed=0, vqvae_checkpoint=None, prior_checkpoint=None,
               artist_ids=None, genre_ids=None, device="cuda"):
    """jukebox/main.py:176-200 -> (hps, vqvae, top_prior): the VQ-VAE built for sample_length 1048576 and the top-level
    prior with prior_depth forced to 36 (:197). Weights, in order of preference:
      * `vqvae_checkpoint` / `prior_checkpoint`: upstream's vqvae.pth.tar / prior_level_2.pth.tar, read with the name maps
        of model.py and loaded NON-STRICTLY like the reference's patched restore_model (make_models.py.patch:7-8): of the
        72 transformer blocks in the file the first 36 are used, every other extra tensor is ignored;
      * `state` = (vqvae_state, prior_state) dicts in this package's names;
      * otherwise seeded random weights of the upstream shapes (no Jukebox checkpoint is reachable offline).
    `precision`: "tf32" (fp32 storage, parity mode) or "bf16"."""
    from types import SimpleNamespace

    from llark_b200.jukebox.model import (load_upstream_state, prior_key_map, random_label_state, vqvae_key_map)

    if model != "5b":
        raise NotImplementedError("LLark uses the 5b prior")
    cfg = cfg or JukeboxConfig()
    hps = SimpleNamespace(sr=JUKEBOX_SAMPLE_RATE, n_samples=8, name="samples", levels=3,
                          hop_fraction=[0.5, 0.5, 0.125], sample_length=cfg.sample_length)
    if vqvae_checkpoint is not None or prior_checkpoint is not None:
        if vqvae_checkpoint is None or prior_checkpoint is None:
            raise ValueError("pass both vqvae_checkpoint and prior_checkpoint")
        vq_state, miss_v = load_upstream_state(vqvae_checkpoint, vqvae_key_map(cfg))
        prior_state, miss_p = load_upstream_state(prior_checkpoint, prior_key_map(cfg))
        if miss_v or miss_p:  # strict=False: report, do not raise (what torch's non-strict load returns as missing_keys)
            logging.warning(f"tensors missing from the checkpoints (left uninitialised would be wrong: raising): "
                            f"{(miss_v + miss_p)[:8]}{' ...' if len(miss_v + miss_p) > 8 else ''}")
            raise KeyError(f"{len(miss_v + miss_p)} tensors of the 36-block front end are not in the checkpoints")
    elif state is not None:
        vq_state, prior_state = state
    else:
        vq_state = random_vqvae_state(cfg, seed)
        prior_state = random_prior_state(cfg, seed + 1, device=device)
        prior_state.update(random_label_state(cfg, seed=seed + 2, device=device))
        prior_state.pop("x_cond", None)  # conditioning now comes from the label tables (get_cond)
        prior_state.pop("y_cond", None)
    vqvae = JukeboxVQVAE(cfg, vq_state, device=device)
    top_prior = JukeboxTopPrior(cfg, prior_state, device=device, precision=precision, artist_ids=artist_ids,
                                genre_ids=genre_ids)
    return hps, vqvae, top_prior


def _select_batch(paths, batch_size, batch_idx):
    """The `--batch_size / --batch_idx` slice of the sorted input list (both must be given to take effect)."""
    if batch_size is None or batch_idx is None:
        return paths
    starts = range(0, len(paths), batch_size)
    if batch_idx >= len(starts):
        raise ValueError("Invalid batch index")
    return paths[starts[batch_idx]: starts[batch_idx] + batch_size]


def main(argv=None):
    """Command line of jukebox/main.py:203-256: embed every file of --input_dir (optionally one batch of it) and save
    `<name>.npy` ([frames, 4800] float32, `.wav` in the name replaced) under --output_dir. The model is loaded lazily,
    only once a first input exists."""
    import os
    import pathlib
    from argparse import ArgumentParser

    ap = ArgumentParser()
    ap.add_argument("--batch_size", type=int, default=None)
    ap.add_argument("--batch_idx", type=int, default=None)
    ap.add_argument("--input_dir", default="/input", help="path to inputs")
    ap.add_argument("--output_dir", default="/output", help="path to outputs")
    ap.add_argument("--pool-frames-per-second", default=10, type=int,
                    help="Frames per second for pooling. Set to zero to pool over all timesteps.")
    args = ap.parse_args(argv)

    out_dir = pathlib.Path(args.output_dir)
    out_dir.mkdir(exist_ok=True)
    inputs = _select_batch(sorted(pathlib.Path(args.input_dir).iterdir()), args.batch_size, args.batch_idx)
    model = None
    for path in inputs:
        if model is None:
            model = load_model()
        hps, vqvae, top_prior = model
        with torch.no_grad():
            acts = get_acts_from_file(path, hps, vqvae, top_prior, meanpool=True,
                                      pool_frames_per_second=args.pool_frames_per_second)
        np.save(os.path.join(out_dir, os.path.basename(path).replace(".wav", ".npy")), acts)


if __name__ == "__main__":
    main()
