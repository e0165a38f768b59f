"""bench.py -- the hot path's headline metric on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      (the reference's CPU path on the box's host cores)

Workload (BASELINE.json configs[2], named in `config.workload`): LLark Llama-2-7B + mm_projector training step on
precomputed Jukebox embeddings -- random-init weights of the real architecture (H 4096, I 11008, 32 layers, 32
heads, V 32004, projector 4800->4096), per-GPU micro-batch B (default 16) samples of S = 512 tokens (header + <audio_start> +
240 audio frames + <audio_end> + 256 random text ids; labels on the last 128 tokens), bf16 weights / grads / Adam
states as the reference trains (m2t/train.py:101-103), lm_head frozen, text-embedding rows detached. One "step" =
forward + backward + gradient all-reduce (N > 1) + global-norm clip + AdamW on one micro-batch per GPU.
metric = audio-instruct samples/s (whole job). Synthetic data, no network.

`value`   : inputs already resident in HBM when the timed region starts.
`e2e`     : the same step through the public API from PINNED HOST buffers (H2D of ids/labels/mask/encodings every
            step and a D2H read of the loss inside the timed region).
`roofline`: the tcgen05 GEMM core (dominant kernel): algorithmic GEMM FLOPs of the step / summed GEMM launch time,
            each launch bracketed by CUDA events on the launching stream inside the timed region.
`cpu_baseline`: the oracle restatement of the reference's PyTorch path (oracle/llama_oracle.py) on the host cores,
            on a bounded sample (see `sample`).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LLAMA7B = dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
               num_key_value_heads=32, vocab_size=32004, max_position_embeddings=4096, rms_norm_eps=1e-5)
SEQ, FRAMES, MM_HIDDEN = 512, 240, 4800
# trainable parameters of the step (lm_head frozen, m2t/models/llamav2.py:412-415): 32 layers + embed + final norm + projector
TRAINABLE_PARAMS = (32 * (4 * 4096 * 4096 + 3 * 4096 * 11008 + 2 * 4096) + 32004 * 4096 + 4096
                    + MM_HIDDEN * 4096 + 4096)
ID_PAD, ID_PATCH, ID_START, ID_END = 32000, 32001, 32002, 32003
METRIC = "audio-instruct samples/sec (25s clip + 256-tok) Llama-2-7B"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="llark", choices=["llark", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU micro-batch (samples)")
    ap.add_argument("--layers", type=int, default=32, help="debug only: fewer layers INVALIDATES the number")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="llama", choices=["llama", "e2e"],
                    help="llama = BASELINE configs[2] (precomputed Jukebox embeds, the default headline); "
                         "e2e = configs[3]: 25 s waveform -> Jukebox front end -> projector -> Llama step in-process")
    ap.add_argument("--jukebox-precision", default="bf16", choices=["bf16", "tf32"])
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synth_batch(B, seed):
    """One micro-batch as DataCollatorForSupervisedDataset would emit it (m2t/data_modules.py:195-222), on host."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 32000, (B, SEQ), generator=g)
    pos = 14
    ids[:, pos] = ID_START
    ids[:, pos + 1: pos + 1 + FRAMES] = ID_PATCH
    ids[:, pos + 1 + FRAMES] = ID_END
    labels = torch.full((B, SEQ), -100, dtype=torch.int64)
    labels[:, -128:] = ids[:, -128:]
    mask = torch.ones(B, SEQ, dtype=torch.bool)
    enc = torch.randn(B, FRAMES, MM_HIDDEN, generator=g) * 0.5
    return dict(input_ids=ids, labels=labels, attention_mask=mask, audio_encodings=enc)


def step_flops(B, L):
    """Algorithmic GEMM FLOPs of one training step on B samples (no recompute; frozen lm_head has dgrad only)."""
    H, I, V = LLAMA7B["hidden_size"], LLAMA7B["intermediate_size"], LLAMA7B["vocab_size"]
    tok = B * SEQ
    layer = 2.0 * tok * (4 * H * H + 3 * H * I)
    head = 2.0 * tok * H * V
    proj = 2.0 * B * FRAMES * MM_HIDDEN * H
    attn_fwd = 4.0 * B * 32 * SEQ * SEQ * 128 / 2
    return dict(gemm=3 * L * layer + 2 * head + 2 * proj, total=3 * L * layer + 2 * head + 2 * proj + 3.5 * L * attn_fwd)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], None, set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_slice_seconds(n_layers, threads, with_optimizer=False):
    """Time ONE sample (S=512) forward+backward through an n_layers slice of Llama-2-7B dims with the oracle."""
    from oracle import llama_oracle as lo

    torch.set_num_threads(threads)
    cfg = lo.LlamaCfg(hidden_size=4096, intermediate_size=11008, num_layers=n_layers, num_heads=32, vocab_size=32004,
                      mm_hidden_size=4800, audio_start_token=ID_START, audio_end_token=ID_END,
                      audio_patch_token=ID_PATCH, pad_token=ID_PAD)
    g = torch.Generator().manual_seed(0)
    params = {}
    for n in lo.param_names(cfg):
        shp = lo.param_shape(cfg, n)
        params[n] = (torch.randn(shp, generator=g) * 0.02 + (1.0 if len(shp) == 1 else 0.0)).to(torch.bfloat16)
    b = synth_batch(1, 0)
    t0 = time.perf_counter()
    _, _, grads = lo.forward_backward(params, cfg, b, detach_text=True)
    t_fb = time.perf_counter() - t0
    if not with_optimizer:
        return t_fb
    # clip + torch.optim.AdamW over this slice's trainable tensors (bf16, no master weights: the reference's regime),
    # reported per parameter so the caller can scale it to the 6.63 G trainable parameters of the full model
    try:
        leaves = []
        for n, g in grads.items():
            if g is None:
                continue
            t = params[n].detach().clone().requires_grad_(True)
            t.grad = g
            leaves.append(t)
        opt = torch.optim.AdamW(leaves, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
        t0 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(leaves, 1.0)
        opt.step()
        per_param = (time.perf_counter() - t0) / max(sum(t.numel() for t in leaves), 1)
    except Exception:  # the baseline stays usable without the optimizer term
        per_param = None
    return t_fb, per_param


_CPU_WARM = False


def cpu_baseline_value(threads, batch=16):
    """Bounded sample: 1-layer and 3-layer slices; full-model time = t1 + 31 * (t3 - t1) / 2 (extrapolated)."""
    global _CPU_WARM
    if not _CPU_WARM:
        cpu_slice_seconds(1, threads)  # untimed: thread pool, allocator and oneDNN primitive caches
        _CPU_WARM = True
    t1, per_param = cpu_slice_seconds(1, threads, with_optimizer=True)
    t3 = cpu_slice_seconds(3, threads)
    per_layer = max((t3 - t1) / 2.0, 1e-3)
    full = t1 + 31 * per_layer
    sample = (f"oracle (PyTorch CPU bf16, {threads} threads) fwd+bwd of 1 sample S=512 through 1- and 3-layer slices "
              f"of Llama-2-7B dims incl. projector/lm_head/CE: t1={t1:.2f}s t3={t3:.2f}s -> 32 layers extrapolated "
              f"{full:.1f}s/sample")
    if per_param is not None:
        t_opt = per_param * TRAINABLE_PARAMS
        full += t_opt / batch  # the GPU arm pays one optimizer step per micro-batch of `batch` samples: same share here
        sample += (f" + clip/AdamW (torch, bf16) timed on the 1-layer slice, scaled to {TRAINABLE_PARAMS / 1e9:.2f} G "
                   f"trainable parameters ({t_opt:.1f}s per step) and charged 1/{batch} per sample")
    else:
        sample += "; optimizer step not included"
    return 1.0 / full, sample


def run_reference(args):
    """--impl reference: the reference's CPU path (its PyTorch classes restated in oracle/) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    total = args.warmup + args.steps
    budget_s = 240.0
    t_start = time.perf_counter()
    done = 0
    for i in range(total):
        if i >= 2 and (time.perf_counter() - t_start) > budget_s:
            break  # bounded: keep the whole run within a few minutes
        v, sample = cpu_baseline_value(threads, args.batch)
        done += 1
        if i >= args.warmup or total <= args.warmup:
            vals.append(v)
    if not vals:
        vals = [v]
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(vals), "warmup": min(args.warmup, max(done - len(vals), 0)),
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[2]: LLark Llama-2-7B + mm_projector training step, precomputed Jukebox "
                                   "embeds, seq 512, bf16", "global_batch": 1, "seq_len": SEQ,
                       "parallelism": "cpu"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def build_model(layers):
    from llark_b200.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    cfg = dict(LLAMA7B)
    cfg["num_hidden_layers"] = layers
    hf = WrappedLlamav2Config(**cfg)
    hf.mm_hidden_size = MM_HIDDEN
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            model = WrappedLlamav2ForCausalLM(hf)
            model.get_model().initialize_adapter_modules()
    finally:
        torch.set_default_dtype(old)
    model.to(torch.bfloat16)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = ID_START, ID_END, ID_PATCH
    # trainable set of scripts/training/train_llark.sh + initialize_audio_tokenizer (llamav2.py:408-415)
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    model.get_model().orig_embeds_params = [None]  # text-embedding rows detached (only the marker rows get grads)
    model.validate = "lazy"
    return model


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from llark_b200.comm import GradComm

        comm = GradComm.from_torch_distributed()
    from llark_b200 import ops
    from llark_b200.optim import LlarkAdamW

    B = args.batch
    torch.manual_seed(1234)  # identical weights on every rank (DDP broadcasts rank 0's; same seed is equivalent)
    model = build_model(args.layers)
    opt = LlarkAdamW(model, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, comm=comm)

    n_batches = 4
    host = [synth_batch(B, seed=1000 + rank + world * i) for i in range(n_batches)]
    front = None
    if args.workload == "e2e":
        # configs[3]: the frozen Jukebox front end runs in the same process on the same GPU (no grad), its
        # [B,240,4800] output replaces the precomputed encodings of the batch.
        from llark_b200.jukebox import main as jm

        _, vqvae, top_prior = jm.load_model("5b", precision=args.jukebox_precision)
        g = torch.Generator().manual_seed(77 + rank)
        for b in host:
            b["waveform"] = torch.rand(B, 1102500, generator=g) * 2 - 1  # 25 s @ 44.1 kHz, mono
            del b["audio_encodings"]

        def front(batch):
            with torch.no_grad():
                enc = jm.embed_batch(batch["waveform"], vqvae, top_prior, pool_frames_per_second=10)
            out = {k: v for k, v in batch.items() if k != "waveform"}
            out["audio_encodings"] = enc
            return out
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.cuda() for k, v in b.items()} for b in host]

    def step_resident(i):
        b = resident[i % n_batches]
        out = model(**(front(b) if front else b))
        out.loss.backward()
        opt.step()
        return out.loss

    def step_e2e(i):
        b = {k: v.cuda(non_blocking=True) for k, v in pinned[i % n_batches].items()}
        out = model(**(front(b) if front else b))
        out.loss.backward()
        opt.step()
        return out.loss.item()  # D2H read of the step's result

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            ops.GEMM_PROFILE = []
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        prof = ops.GEMM_PROFILE
        ops.GEMM_PROFILE = None
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, prof

    for i in range(args.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    model.check_audio_token_errors()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_res, prof = timed(step_resident, args.steps, profile=True)
    launches = (model.engine.launches + opt.launches) * args.steps
    gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    gemm_flops = sum(f for _, _, f, _ in prof)
    n_gemm = len(prof)
    by_kind = {}
    for a, b, f, shp in prof:
        kind = {(0, 0): "fwd(K,K)", (0, 1): "dgrad(K,MN)", (1, 1): "wgrad(MN,MN)", (1, 0): "(MN,K)"}[shp[3:]]
        key = f"{kind} M{shp[0]} N{shp[1]} K{shp[2]}"
        t = by_kind.setdefault(key, [0.0, 0.0, 0])
        t[0] += f
        t[1] += a.elapsed_time(b)
        t[2] += 1
    by_kind = {k: {"tflops": round(v[0] / (v[1] / 1000.0) / 1e12, 1), "ms_total": round(v[1], 2), "launches": v[2]}
               for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1][1])}
    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        return
    pk, pk_kind = peaks()
    total_samples = B * world * args.steps
    value = total_samples / (ms_res / 1000.0)
    e2e_value = total_samples / (ms_e2e / 1000.0)
    fl = step_flops(B, args.layers)
    achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("configs[2]: LLark Llama-2-7B + mm_projector training step (fwd+bwd+clip+AdamW), "
                                "precomputed Jukebox embeds [240,4800], seq 512, bf16") if front is None else
                               ("configs[3]: end-to-end 25 s waveform -> Jukebox-5b front end (36-layer prior, "
                                f"{args.jukebox_precision}) -> mm_projector -> Llama-2-7B training step, seq 512, bf16"),
                   "global_batch": B * world, "per_gpu_batch": B, "seq_len": SEQ, "layers": args.layers,
                   "parallelism": f"dp{world}", "l2": "working set (13.5 GB weights + activations) >> 126 MB L2; "
                                                     "4 rotating input batches",
                   "valid": args.layers == 32},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "llark::gemm_kernel (tcgen05/TMEM/TMA)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_kind": f"{pk_kind} bf16_tflops_sustained", "gemm_launches_timed": n_gemm,
                     "gemm_share_of_step": gemm_ms / ms_res if ms_res else None,
                     "step_tflops_algorithmic": fl["total"] * world * args.steps / (ms_res / 1000.0) / 1e12 / world,
                     "flops_per_sample": fl["total"] / B, "gemm_by_shape": by_kind},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            v, sample = cpu_baseline_value(os.cpu_count() or 1, B)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": sample}
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"failed: {e!r}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
