"""BASELINE configs[1]: Jukebox front end only -- 25 s @ 44.1 kHz synthetic waveforms -> [240, 4800] embeddings,
batch 8, 1 x B200, random weights of the upstream shapes (36-layer 5b prior). Prints one JSON line with clips/s
and per-stage timings / roofline numbers: conv encoder GB/s against the 1.42 GB/clip fused-res-block
traffic figure and GFLOP/s against 41.2 GFLOP/clip; prior TFLOP/s against 40.77 TFLOP/clip; pool GB/s.

    python tools/bench_jukebox.py [--precision tf32|bf16] [--batch 8] [--steps 3] [--depth 36]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--depth", type=int, default=36)
    a = ap.parse_args()
    from llark_b200 import ops
    from llark_b200.jukebox import main as jm
    from llark_b200.jukebox.model import JukeboxConfig

    cfg = JukeboxConfig(prior_depth=a.depth)
    hps, vqvae, top_prior = jm.load_model("5b", cfg=cfg, precision=a.precision)
    g = torch.Generator().manual_seed(0)
    wav = torch.rand(a.batch, 1102500, generator=g) * 2 - 1  # 25 s @ 44.1 kHz
    peak = wav.abs().amax(dim=1, keepdim=True)
    x = (wav / peak)[:, : cfg.sample_length].cuda().contiguous()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    stages = {"conv_encoder": 0.0, "vq": 0.0, "prior": 0.0, "pool": 0.0}
    total = 0.0
    for it in range(a.warmup + a.steps):
        e = [ev() for _ in range(5)]
        e[0].record()
        lat = vqvae.encode_top_latents(x[:, None, :].contiguous())
        e[1].record()
        z = ops.jb_vq_encode(lat, vqvae.w["codebook"])
        e[2].record()
        top_prior.prior.only_encode = True
        acts = top_prior.prior.forward(z, None, None, None, False)
        e[3].record()
        pooled = torch.stack([ops.jb_pool(acts[i].contiguous(), cfg.n_ctx, 34) for i in range(a.batch)])
        e[4].record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            for k, (s0, s1) in zip(stages, zip(e[:-1], e[1:])):
                stages[k] += s0.elapsed_time(s1)
            total += e[0].elapsed_time(e[4])
    n = a.steps * a.batch
    per_clip = {k: v / n for k, v in stages.items()}
    scale = a.depth / 36.0
    out = {"workload": "configs[1]: Jukebox front end, 25 s clips, batch %d" % a.batch, "precision": a.precision,
           "prior_depth": a.depth, "clips_per_s": n / (total / 1000.0), "ms_per_clip": total / n,
           "ms_per_clip_by_stage": per_clip,
           "conv_encoder": {"gflops_achieved": 41.2 / per_clip["conv_encoder"] * 1000.0 / 1000.0,
                            "tflops": 41.2e-3 / (per_clip["conv_encoder"] / 1000.0),
                            "gbs_vs_1p42GB_per_clip": 1.42 / (per_clip["conv_encoder"] / 1000.0)},
           "prior": {"tflops": 40.77 * scale / (per_clip["prior"] / 1000.0)},
           "pool": {"gbs": 0.1619 / (per_clip["pool"] / 1000.0)}, "pooled_shape": list(pooled.shape)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
