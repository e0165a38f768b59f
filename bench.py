"""bench.py -- the hot path's headline metric on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      (the reference's CPU path on the box's host cores)

Workload (BASELINE.json configs[2], named in `config.workload`): LLark Llama-2-7B + mm_projector training step on
precomputed Jukebox embeddings -- random-init weights of the real architecture (H 4096, I 11008, 32 layers, 32
heads, V 32004, projector 4800->4096), per-GPU micro-batch B (default 32) samples of S = 512 tokens (header + <audio_start> +
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
    ap.add_argument("--batch", type=int, default=32, help="per-GPU micro-batch (samples)")
    ap.add_argument("--layers", type=int, default=32, help="debug only: fewer layers INVALIDATES the number")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the configs[1] (Jukebox only) and configs[3] (end-to-end) measurements added to the line")
    ap.add_argument("--no-torch-baseline", action="store_true",
                    help="skip the stock-PyTorch run of the reference's model class on the same GPU (N=1 only)")
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


# ------------------------------------------------------------------------------------------------ reference arms
REF_COPY = os.path.join(ROOT, "baseline", "_ref")  # build() copies /root/reference/{m2t,jukebox} here (git-ignored)
REF_LAYERS = 4  # the CPU arm times a REAL 4-layer model of Llama-2-7B dims; the 32-layer figure is extrapolated


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _trainable_params(layers):
    return (layers * (4 * 4096 * 4096 + 3 * 4096 * 11008 + 2 * 4096) + 32004 * 4096 + 4096 + MM_HIDDEN * 4096 + 4096)


def reference_classes():
    """The reference's own classes from the git-ignored copy under baseline/_ref (made by __graft_entry__.build() from
    /root/reference, which does not exist on the GPU box). None when the copy is absent."""
    if not os.path.isdir(os.path.join(REF_COPY, "m2t")):
        return None
    os.environ["LLARK_REFERENCE_ROOT"] = REF_COPY
    from oracle import ref_shims

    ref_shims.REFERENCE_ROOT = REF_COPY
    ref_shims.install()
    from m2t.data_modules import DataCollatorForSupervisedDataset
    from m2t.models.llamav2 import WrappedLlamav2Config, WrappedLlamav2ForCausalLM

    return WrappedLlamav2Config, WrappedLlamav2ForCausalLM, DataCollatorForSupervisedDataset


def build_reference_model(classes, layers, device, attn="sdpa"):
    """The reference's WrappedLlamav2ForCausalLM (m2t/models/llamav2.py:237-419) at Llama-2-7B dims, set up as
    m2t/train.py does for scripts/training/train_llark.sh: bf16 weights, projector, lm_head frozen, text rows detached."""
    Config, Model, _ = classes
    cfg = dict(LLAMA7B)
    cfg["num_hidden_layers"] = layers
    hf = Config(**cfg, attn_implementation=attn)
    hf.mm_hidden_size = MM_HIDDEN
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = Model(hf)
            model.get_model().initialize_adapter_modules()
    finally:
        torch.set_default_dtype(old)
    model.to(torch.bfloat16)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_start_token, aec.audio_end_token, aec.audio_patch_token = ID_START, ID_END, ID_PATCH
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    model.get_model().orig_embeds_params = [None]
    model.config.use_cache = False
    model.train()
    return model


class _PadTok:
    pad_token_id = ID_PAD


def reference_step_fn(classes, model, B, device):
    """One training step the way HF Trainer(adamw_torch, max_grad_norm 1.0, bf16 autocast) runs the reference's model:
    collate -> forward under autocast -> backward -> clip -> AdamW. Returns (step(i) -> (t_fwd_bwd, t_opt), n_params)."""
    collate = classes[2](tokenizer=_PadTok())
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    batches = []
    for i in range(2):
        b = synth_batch(B, seed=1000 + i)
        batches.append([dict(input_ids=b["input_ids"][j], labels=b["labels"][j], audio_encoding=b["audio_encodings"][j])
                        for j in range(B)])
    dev_type = torch.device(device).type

    def sync():
        if dev_type == "cuda":
            torch.cuda.synchronize()

    def step(i):
        t0 = time.perf_counter()
        batch = collate(batches[i % len(batches)])
        batch = {k: v.to(device) for k, v in batch.items()}
        with torch.autocast(dev_type, dtype=torch.bfloat16):
            out = model(**batch)
        out.loss.backward()
        sync()
        t1 = time.perf_counter()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        loss = float(out.loss.detach())
        sync()
        return t1 - t0, time.perf_counter() - t1, loss

    return step, sum(p.numel() for p in params)


def port_step_fn(layers):
    """Fallback when baseline/_ref is absent: the oracle restatement of the same classes (oracle/llama_oracle.py)."""
    from oracle import llama_oracle as lo

    cfg = lo.LlamaCfg(hidden_size=4096, intermediate_size=11008, num_layers=layers, num_heads=32, vocab_size=32004,
                      mm_hidden_size=4800, audio_start_token=ID_START, audio_end_token=ID_END,
                      audio_patch_token=ID_PATCH, pad_token=ID_PAD)
    g = torch.Generator().manual_seed(0)
    params = {}
    for n in lo.param_names(cfg):
        shp = lo.param_shape(cfg, n)
        params[n] = (torch.randn(shp, generator=g) * 0.02 + (1.0 if len(shp) == 1 else 0.0)).to(torch.bfloat16)
    leaves = {}
    opt = [None]

    def step(i):
        t0 = time.perf_counter()
        loss, _, grads = lo.forward_backward(params, cfg, synth_batch(1, 1000 + i % 2), detach_text=True)
        t1 = time.perf_counter()
        if opt[0] is None:
            for n, g_ in grads.items():
                if g_ is not None:
                    leaves[n] = params[n].detach().clone().requires_grad_(True)
            opt[0] = torch.optim.AdamW(list(leaves.values()), lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
        for n, t in leaves.items():
            t.grad = grads[n]
        torch.nn.utils.clip_grad_norm_(list(leaves.values()), 1.0)
        opt[0].step()
        return t1 - t0, time.perf_counter() - t1, float(loss)

    return step, _trainable_params(layers)


def cpu_reference_measure(warmup, steps, batch_charge, budget_s=200.0, layers=REF_LAYERS):
    """The reference's CPU path on the host cores, bounded: a REAL `layers`-layer model of Llama-2-7B dims (embed,
    projector, splice, decoder layers, lm_head, CE, clip, AdamW -- nothing skipped), one S=512 sample per step.
    The 32-layer figure scales the measured forward+backward by the algorithmic FLOP ratio and the measured clip+AdamW
    by the trainable-parameter ratio; both factors are reported. Median over the timed steps."""
    threads = host_threads()
    torch.set_num_threads(threads)
    classes = reference_classes()
    t_build = time.perf_counter()
    if classes is not None:
        kind = "reference"
        model = build_reference_model(classes, layers, "cpu")
        step, n_train = reference_step_fn(classes, model, 1, "cpu")
    else:
        kind = "port"
        step, n_train = port_step_fn(layers)
    t_build = time.perf_counter() - t_build
    t_start = time.perf_counter()
    times = []
    n_warm = 0
    for i in range(warmup + steps):
        if len(times) >= 1 and i >= 1 and (time.perf_counter() - t_start) > budget_s:
            break
        t_fb, t_opt, loss = step(i)
        if i < min(warmup, 1) or (i < warmup and (time.perf_counter() - t_start) < 0.25 * budget_s):
            n_warm += 1
            continue
        times.append((t_fb, t_opt))
    med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
    t_fb, t_opt = med([t[0] for t in times]), med([t[1] for t in times])
    f_full, f_meas = step_flops(1, 32)["total"], step_flops(1, layers)["total"]
    flops_factor = f_full / f_meas
    params_factor = TRAINABLE_PARAMS / n_train
    full_s = t_fb * flops_factor + t_opt * params_factor / batch_charge
    name = ("unmodified reference classes (m2t.models.llamav2.WrappedLlamav2ForCausalLM + m2t.data_modules."
            "DataCollatorForSupervisedDataset from baseline/_ref) + torch.optim.AdamW/clip_grad_norm_ as HF Trainer "
            "drives them") if kind == "reference" else "oracle/llama_oracle.py port of the reference classes"
    sample = (f"{name}; PyTorch CPU bf16 autocast, {threads} threads; REAL {layers}-layer model of Llama-2-7B dims, "
              f"1 sample S=512 per step, {len(times)} timed steps after {n_warm} warm-up, median fwd+bwd {t_fb:.2f}s + "
              f"clip/AdamW {t_opt:.2f}s; 32-layer step = fwd+bwd x {flops_factor:.3f} (algorithmic FLOP ratio) + "
              f"clip/AdamW x {params_factor:.3f} (trainable-parameter ratio) charged 1/{batch_charge} per sample as the "
              f"GPU arm pays it -> {full_s:.1f}s/sample")
    return {"value": 1.0 / full_s, "kind": kind, "cores": threads, "sample": sample, "timed_steps": len(times),
            "warmup_steps": n_warm, "measured_step_s": t_fb + t_opt, "measured_layers": layers,
            "extrapolation": {"flops_factor": flops_factor, "params_factor": params_factor,
                              "optimizer_charge": 1.0 / batch_charge, "full_step_s": full_s},
            "build_s": t_build, "all_steps_s": [round(a + b, 3) for a, b in times]}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = cpu_reference_measure(args.warmup, args.steps, args.batch)
    value = m["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": m["timed_steps"], "warmup": m["warmup_steps"],
            # ms_per_step is the MEASURED wall time of one bounded step (4-layer model); `value` is the 32-layer figure
            "ms_per_step": 1000.0 * m["measured_step_s"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[2]: LLark Llama-2-7B + mm_projector training step, precomputed Jukebox "
                                   "embeds, seq 512, bf16", "global_batch": 1, "per_gpu_batch": 1, "seq_len": SEQ,
                       "parallelism": "cpu", "measured_layers": m["measured_layers"]},
            "extrapolation": m["extrapolation"],
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": m["cores"], "kind": m["kind"],
                             "sample": m["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def gpu_torch_baseline(B, steps, warmup):
    """The fair GPU comparator: the reference's own model class on the same B200 through
    stock PyTorch (cuBLAS GEMMs, SDPA attention, ATen elementwise, torch.optim.AdamW), bf16 autocast, same synthetic
    batch, same step (fwd + bwd + clip + AdamW on one micro-batch). Falls back to halving B on out-of-memory."""
    classes = reference_classes()
    if classes is None:
        return {"unavailable": "baseline/_ref (copy of the reference's m2t package made by build()) is absent"}
    import gc

    last_err = None
    while B >= 1:
        model = step = None
        try:
            torch.manual_seed(1234)
            model = build_reference_model(classes, 32, "cuda")
            step, _ = reference_step_fn(classes, model, B, "cuda")
            for i in range(warmup):
                step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            fl = step_flops(B, 32)["total"]
            res = {"value": B / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "per_gpu_batch": B, "steps": steps,
                   "warmup": warmup, "tflops_algorithmic": fl / (ms / 1000.0) / 1e12,
                   "impl": "m2t.models.llamav2.WrappedLlamav2ForCausalLM (baseline/_ref) .cuda(), torch "
                           f"{torch.__version__} bf16 autocast, sdpa attention, torch.optim.AdamW + clip_grad_norm_, "
                           "no gradient checkpointing",
                   "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
            return res
        except torch.OutOfMemoryError as e:
            last_err = repr(e)[:200]
            B //= 2
        finally:
            del model, step
            gc.collect()
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
    return {"unavailable": f"out of memory at every batch size: {last_err}"}


# ------------------------------------------------------------------------------------------------ Jukebox front end
def jukebox_measure(precision="bf16", batch=8, steps=3, warmup=1, depth=36, objs=None):
    """BASELINE configs[1]: Jukebox front end only -- 25 s @ 44.1 kHz synthetic waveforms -> [240, 4800] embeddings,
    batch 8, 1 x B200, random weights of the upstream shapes (36-layer 5b prior). Per-stage CUDA-event timings and the
    roofline numbers of SURVEY section 8(d): conv encoder GB/s against the 1.42 GB/clip fused-res-block traffic figure and
    FP32 TFLOP/s against 41.2 GFLOP/clip; prior TFLOP/s against 40.77 TFLOP/clip (per-shape GEMM rates from events
    around every launch); pool GB/s against 161.9 MB/clip."""
    from llark_b200 import ops
    from llark_b200.jukebox import main as jm
    from llark_b200.jukebox.model import JukeboxConfig

    if objs is None:
        cfg = JukeboxConfig(prior_depth=depth)
        _, vqvae, top_prior = jm.load_model("5b", cfg=cfg, precision=precision)
    else:
        vqvae, top_prior = objs
        cfg = top_prior.cfg
    g = torch.Generator().manual_seed(0)
    wav = torch.rand(batch, 1102500, generator=g) * 2 - 1  # 25 s @ 44.1 kHz
    peak = wav.abs().amax(dim=1, keepdim=True)
    x = (wav / peak)[:, : cfg.sample_length].cuda().contiguous()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    stages = {"conv_encoder": 0.0, "vq": 0.0, "prior": 0.0, "pool": 0.0}
    total = 0.0
    prof = None
    for it in range(warmup + steps):
        e = [ev() for _ in range(5)]
        if it == warmup + steps - 1:
            ops.GEMM_PROFILE = []
        e[0].record()
        lat = vqvae.encode_top_latents(x[:, None, :].contiguous())
        e[1].record()
        z = ops.jb_vq_encode(lat, vqvae.w["codebook"])
        e[2].record()
        top_prior.prior.only_encode = True
        acts = top_prior.prior.forward(z, *top_prior.default_cond(), None, False)
        e[3].record()
        pooled = jm.pool_batch(acts, cfg.n_ctx, 34)
        e[4].record()
        torch.cuda.synchronize()
        if ops.GEMM_PROFILE is not None:
            prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        if it >= warmup:
            for k, (s0, s1) in zip(stages, zip(e[:-1], e[1:])):
                stages[k] += s0.elapsed_time(s1)
            total += e[0].elapsed_time(e[4])
        del lat, z, acts
    n = steps * batch
    per_clip = {k: v / n for k, v in stages.items()}
    scale = cfg.prior_depth / 36.0
    by_shape = {}
    for a, b, f, shp in prof or []:
        t = by_shape.setdefault(f"M{shp[0]} N{shp[1]} K{shp[2]}", [0.0, 0.0, 0])
        t[0] += f
        t[1] += a.elapsed_time(b)
        t[2] += 1
    gemm_ms = sum(v[1] for v in by_shape.values())
    by_shape = {k: {"tflops": round(v[0] / (v[1] / 1000.0) / 1e12, 1), "ms_total": round(v[1], 2), "launches": v[2]}
                for k, v in by_shape.items()}
    pk, _ = peaks()
    conv_s = per_clip["conv_encoder"] / 1000.0
    return {"workload": "configs[1]: Jukebox front end only, 25 s @ 44.1 kHz synthetic clips -> [240,4800], batch %d, "
                        "1 GPU, levels 0/1 of the VQ-VAE skipped (dead work in the reference)" % batch,
            "precision": {"conv_encoder": ("fp32 SIMT" if os.environ.get("LLARK_JB_CONV_TC") == "0" else
                                           "res-blocks 3xTF32 mma.sync (fp32-equivalent), down / projection convs fp32 SIMT"),
                          "prior": precision}, "prior_depth": cfg.prior_depth,
            "clips_per_s": n / (total / 1000.0), "ms_per_clip": total / n, "ms_per_clip_by_stage": per_clip,
            "conv_encoder": {"fp32_tflops": 41.2e-3 / conv_s, "gbs_vs_1p42GB_per_clip": 1.42 / conv_s,
                             "frac_of_hbm_peak": 1.42 / conv_s / pk["hbm_gbs"]},
            "prior": {"tflops": 40.77 * scale / (per_clip["prior"] / 1000.0),
                      "frac_of_sustained_bf16_peak": 40.77 * scale / (per_clip["prior"] / 1000.0) /
                      pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                      "gemm_share_of_prior": gemm_ms / (stages["prior"] / steps) if prof else None,
                      "gemm_by_shape": by_shape},
            "pool": {"gbs": 0.1619 / (per_clip["pool"] / 1000.0)}, "pooled_shape": list(pooled.shape)}


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


def multi_gpu_checks(comm, opt, model, world, rank):
    """Correctness bits of the N > 1 path, run once after the warm-up steps (outside the timed region):
    (1) the library's bucket all-reduce (llark_grad_bucket_allreduce on the engine's communicator) against the exact
    expected sum -- integer-valued bf16 data regenerated from every rank's seed, so the sum is exact -- and against
    torch.distributed.all_reduce of a copy; (2) after the warm-up optimizer steps every rank holds bit-identical
    parameters (what DDP guarantees: identical initial weights + identical averaged gradients)."""
    import torch.distributed as dist

    n = 1 << 24
    def data(r):
        g = torch.Generator(device="cuda").manual_seed(4242 + r)
        return torch.randint(-8, 9, (n,), device="cuda", generator=g).to(torch.bfloat16)

    mine = data(rank)
    a, b = mine.clone(), mine.clone()
    comm.allreduce_(a)
    dist.all_reduce(b)
    want = sum(data(r).float() for r in range(world))
    torch.cuda.synchronize()
    ok_sum = bool(torch.equal(a.float(), want)) and bool(torch.equal(a, b))
    flat = model.engine.flat
    digest = flat.params[: flat.n_train].view(torch.int16).to(torch.int64).sum().reshape(1)
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(int(d.item()) == int(all_d[0].item()) for d in all_d)
    t = torch.tensor([int(ok_sum), int(same)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"bucket_allreduce_exact": bool(t[0].item()), "elements": n, "replicas_bit_identical_after_warmup":
            bool(t[1].item()), "world": world}


def measure_extras(args, model, opt, B, world, rank, timed):
    from llark_b200.jukebox import main as jm

    out = {}
    _, vqvae, top_prior = jm.load_model("5b", precision=args.jukebox_precision)
    if world == 1:
        out["configs1_jukebox"] = jukebox_measure(args.jukebox_precision, batch=8, steps=2, warmup=1,
                                                   objs=(vqvae, top_prior))
    g = torch.Generator().manual_seed(77 + rank)
    host = []
    for i in range(2):
        b = synth_batch(B, seed=2000 + rank + world * i)
        del b["audio_encodings"]
        b["waveform"] = torch.rand(B, 1102500, generator=g) * 2 - 1  # 25 s @ 44.1 kHz, mono
        host.append({k: v.pin_memory() for k, v in b.items()})
    chunk = 8  # clips per front-end pass (bounds the prior's fp32 activations at 1.26 GB per tensor)

    def step(i):
        b = {k: v.cuda(non_blocking=True) for k, v in host[i % 2].items()}
        with torch.no_grad():
            enc = torch.cat([jm.embed_batch(b["waveform"][j:j + chunk], vqvae, top_prior, pool_frames_per_second=10)
                             for j in range(0, B, chunk)])
        res = model(input_ids=b["input_ids"], labels=b["labels"], attention_mask=b["attention_mask"],
                    audio_encodings=enc)
        res.loss.backward()
        opt.step()
        return res.loss.item()

    step(0)
    steps = 2
    ms, _ = timed(step, steps)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    out["configs3_e2e"] = {
        "workload": "configs[3]: end-to-end 25 s waveform (pinned host) -> Jukebox-5b front end (VQ-VAE level 2, 36-layer "
                    f"prior {args.jukebox_precision}, pool) -> mm_projector -> Llama-2-7B training step, seq 512, bf16",
        "value": B * world * steps / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": 1,
        "per_gpu_batch": B, "global_batch": B * world, "n_gpus": world, "h2d_bytes_per_step": h2d,
        "d2h_bytes_per_step": 4}
    return out


_REAL_STDOUT = sys.stdout


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else the run prints (the reference's own
    `[INFO] ...` prints during model construction, warnings) is routed to stderr by main()."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    args = parse()
    sys.stdout = sys.stderr  # keep stdout clean for the JSON line (see emit)
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
    torch_base = None
    if world == 1 and not args.no_torch_baseline and args.layers == 32 and args.workload == "llama":
        try:
            torch_base = gpu_torch_baseline(B, steps=max(2, min(args.steps, 4)), warmup=2)
        except Exception as e:  # a comparator failure must not take the measurement down
            torch_base = {"unavailable": f"failed: {e!r}"[:300]}
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
    mg_checks = multi_gpu_checks(comm, opt, model, world, rank) if world > 1 else None
    sampler = ClockSampler(local) if rank == 0 else None
    ms_res, prof = timed(step_resident, args.steps, profile=True)
    ms_prof = ms_res  # the roofline is measured INSIDE the timed region (CUDA events around every GEMM launch)
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

    # ---- the other BASELINE configs, measured in the same process and carried as extra fields of the one JSON line:
    # configs[3] (waveform -> Jukebox front end -> projector -> Llama step, every N) and configs[1] (front end only, N=1)
    extras = {}
    if front is None and not args.no_extras and args.layers == 32:
        try:
            extras = measure_extras(args, model, opt, B, world, rank, timed)
        except Exception as e:  # never take the headline down
            extras = {"configs3_e2e": {"unavailable": f"failed: {e!r}"[:300]}}
    if rank != 0:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
        return
    pk, pk_kind = peaks()
    total_samples = B * world * args.steps
    value = total_samples / (ms_res / 1000.0)
    e2e_value = total_samples / (ms_e2e / 1000.0)
    fl = step_flops(B, args.layers)
    achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    # DRAM traffic of the GEMM kernel: dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu pass over the
    # bench's OWN shapes (tools/gpu/r2_evidence.sh -> profiles/gemm_traffic.json, keyed like gemm_by_shape); `traffic` is
    # the launch-weighted mean over this step's launches, `traffic_by_shape` carries each shape next to its algorithmic bytes
    traffic, traffic_by_shape = None, None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        shapes = tj.get("by_shape") or {}
        hit = {k: shapes[k] for k in by_kind if k in shapes}
        if hit:
            n_l = sum(by_kind[k]["launches"] for k in hit)
            traffic = sum(by_kind[k]["launches"] * hit[k]["dram_bytes_per_launch"] for k in hit) / n_l
            traffic_by_shape = {k: {"dram_bytes_per_launch": v["dram_bytes_per_launch"],
                                    "algorithmic_bytes_per_launch": v["algorithmic_bytes_per_launch"],
                                    "ratio": v["traffic_over_algorithmic"]} for k, v in hit.items()}
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
                     "gemm_share_of_step": gemm_ms / ms_prof if ms_prof else None,
                     "step_tflops_algorithmic": fl["total"] * world * args.steps / (ms_res / 1000.0) / 1e12 / world,
                     "flops_per_sample": fl["total"] / B, "gemm_by_shape": by_kind,
                     "traffic_by_shape": traffic_by_shape},
        "clocks": clocks,
    }
    if mg_checks is not None:
        line["multi_gpu_checks"] = mg_checks
    if extras:
        line.update(extras)
    if torch_base is not None:
        line["gpu_torch_baseline"] = torch_base
        if torch_base.get("value"):
            line["gpu_torch_baseline"]["llark_speedup"] = value / torch_base["value"]
    if not args.no_cpu_baseline and world == 1:
        try:
            m = cpu_reference_measure(1, 2, B, budget_s=100.0)
            line["cpu_baseline"] = {"value": m["value"], "unit": UNIT, "cores": m["cores"], "kind": m["kind"],
                                    "sample": m["sample"], "extrapolation": m["extrapolation"]}
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                    "sample": f"failed: {e!r}"}
    emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
