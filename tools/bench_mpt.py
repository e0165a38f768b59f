"""BASELINE configs[4]: CLAP + MPT-1B training step -- one [1, 512] CLAP embedding per sample
spliced into a 256-token sequence, MPT-1B dims (m2t/llava/model/mpt/configuration_mpt.py:30-36: d_model 2048, 16 heads,
24 layers, expansion 4, vocab 50368; ALiBi, q/k LayerNorm, no biases, tied head), bf16, fwd + bwd + clip + AdamW.
First run on a B200 in round 2 (profiles/r2a_bench_mpt.json). Under torchrun it is the data-parallel step on N GPUs
(configs[4] names 2 x B200): one rank per GPU, the flat-bucket NCCL all-reduce overlapped with backward as in bench.py.

    python tools/bench_mpt.py [--batch 32] [--steps 5] [--warmup 3] [--layers 24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tools/bench_mpt.py
    python tools/bench_mpt.py --cpu-standin --layers 1 --d-model 256 --heads 2 --vocab 512 --batch 2
        (no GPU: exercises this script's own logic with the test-only PyTorch stand-in for the ops; prints no rates)

One JSON line: samples/s resident and end to end (pinned host batch -> device, loss read back), step TFLOP/s against the
algorithmic FLOPs below, the GEMM share of the step.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

BF16 = torch.bfloat16
SEQ, MM_HIDDEN = 256, 512


def step_flops(B, S, d, L, V, exp):
    """Algorithmic FLOPs of one step: per token 2 x weights forward, 4 x backward (dgrad + wgrad) for the blocks and the
    tied head (its table is trainable: the head contributes dgrad AND wgrad), causal attention 4*S*d/2 per token per
    layer forward and 2.5 x that backward, projector negligible (one row per sample)."""
    T = B * S
    block_w = 4 * d * d + 2 * exp * d * d
    gemm_fwd = 2.0 * T * (L * block_w + V * d)
    attn_fwd = L * B * (4.0 * S * S * d / 2)
    return {"fwd": gemm_fwd + attn_fwd, "bwd": 2 * gemm_fwd + 2.5 * attn_fwd, "total": 3 * gemm_fwd + 3.5 * attn_fwd}


def build_model(a, device):
    from llark_b200.models.mpt import WrappedMPTConfig, WrappedMPTForCausalLM

    hf = WrappedMPTConfig(d_model=a.d_model, n_heads=a.heads, n_layers=a.layers, expansion_ratio=4, max_seq_len=2048,
                          vocab_size=a.vocab + 4, no_bias=True, mm_hidden_size=MM_HIDDEN,
                          attn_config={"alibi": True, "qk_ln": True, "attn_impl": "torch", "alibi_bias_max": 8})
    old = torch.get_default_dtype()
    torch.set_default_dtype(BF16)
    try:
        with torch.device(device):
            model = WrappedMPTForCausalLM(hf)
            model.get_model().initialize_adapter_modules()
    finally:
        torch.set_default_dtype(old)
    model.to(BF16)
    with torch.no_grad():  # MPT's own init is out of scope here: any non-degenerate weights time the same
        for n, p in model.named_parameters():
            if p.dim() == 1 and "bias" not in n:
                p.fill_(1.0)
            elif p.dim() == 1:
                p.zero_()
            else:
                p.normal_(0.0, 0.02)
    aec = model.get_model().audio_encoder_config
    aec.use_audio_start_end = True
    aec.audio_patch_token, aec.audio_start_token, aec.audio_end_token = a.vocab + 1, a.vocab + 2, a.vocab + 3
    return model


def synth_batch(a, seed):
    g = torch.Generator().manual_seed(seed)
    B = a.batch
    ids = torch.randint(3, a.vocab, (B, SEQ), generator=g)
    ids[:, 3], ids[:, 4], ids[:, 5] = a.vocab + 2, a.vocab + 1, a.vocab + 3  # <audio_start> <audio_patch> <audio_end>
    labels = torch.full((B, SEQ), -100, dtype=torch.int64)
    labels[:, SEQ // 2:] = ids[:, SEQ // 2:]
    return dict(input_ids=ids, attention_mask=torch.ones(B, SEQ, dtype=torch.bool), labels=labels,
                audio_encodings=torch.randn(B, 1, MM_HIDDEN, generator=g) * 0.5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--d-model", type=int, default=2048)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--vocab", type=int, default=50368)
    ap.add_argument("--cpu-standin", action="store_true")
    a = ap.parse_args()

    if a.cpu_standin:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cpu_ops
        import llark_b200.engine as eng
        import llark_b200.engine_mpt as engm
        import llark_b200.optim as optim

        for mod in (eng, engm, optim):
            mod.ops = cpu_ops

        def ensure_packed(self):  # the product refuses CPU tensors; the dry run packs them anyway (as tests/ do)
            named = self._named()
            if self.flat is None or not self.flat.is_current(named):
                self.flat = eng.FlatBuckets(named, named[0][1].device, cast_to_bf16=not torch.is_grad_enabled())
            return self.flat

        eng.LlamaEngine.ensure_packed = ensure_packed
        device = "cpu"
    else:
        if not torch.cuda.is_available():
            raise SystemExit("tools/bench_mpt.py needs a B200 (or --cpu-standin for a dry run of the script)")
        device = "cuda"
    from llark_b200 import ops
    from llark_b200.optim import LlarkAdamW

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    comm = None
    if world > 1 and device == "cuda":
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from llark_b200.comm import GradComm

        comm = GradComm.from_torch_distributed()
    torch.manual_seed(1234)
    model = build_model(a, device)
    opt = LlarkAdamW(model, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, comm=comm)
    host = [synth_batch(a, 1000 + rank + world * i) for i in range(4)]
    if device == "cuda":
        pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        resident = [{k: v.cuda() for k, v in b.items()} for b in host]
    else:
        pinned = resident = host

    def step_resident(i):
        out = model(**resident[i % 4])
        out.loss.backward()
        opt.step()
        return out.loss

    def step_e2e(i):
        b = {k: v.to(device, non_blocking=True) for k, v in pinned[i % 4].items()}
        out = model(**b)
        out.loss.backward()
        opt.step()
        return out.loss.item()

    if device == "cpu":
        t0 = time.time()
        l0 = float(step_resident(0))
        l1 = step_e2e(0)
        print(json.dumps({"dry_run": "cpu stand-in", "loss_step0": l0, "loss_step1": l1, "seconds": time.time() - t0,
                          "flops_per_sample": step_flops(a.batch, SEQ, a.d_model, a.layers, a.vocab + 4, 4)["total"]
                          / a.batch}))
        return

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
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        ms = e0.elapsed_time(e1)
        if world > 1:  # device time, max over ranks
            import torch.distributed as dist

            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, prof

    for i in range(a.warmup):
        step_resident(i)
    ms_res, prof = timed(step_resident, a.steps, profile=True)
    gemm_ms = sum(x.elapsed_time(y) for x, y, _, _ in prof)
    gemm_flops = sum(f for _, _, f, _ in prof)
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, a.steps)
    fl = step_flops(a.batch, SEQ, a.d_model, a.layers, a.vocab + 4, 4)
    n = a.batch * a.steps * world
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        if rank != 0:
            dist.destroy_process_group()
            return
    print(json.dumps({
        "workload": f"configs[4]: CLAP [1,512] + MPT-1B training step, seq {SEQ}, bf16, per-GPU batch {a.batch}, "
                    f"layers {a.layers}, dp{world}", "n_gpus": world,
        "samples_per_s": n / (ms_res / 1000.0), "ms_per_step": ms_res / a.steps,
        "e2e": {"samples_per_s": n / (ms_e2e / 1000.0), "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host[0].values()),
                "d2h_bytes_per_step": 4},
        "step_tflops_algorithmic_per_gpu": fl["total"] * a.steps / (ms_res / 1000.0) / 1e12,
        "flops_per_sample": fl["total"] / a.batch,
        "gemm": {"tflops": gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms else None,
                 "share_of_step": gemm_ms / ms_res if ms_res else None, "launches": len(prof)},
        "valid": a.layers == 24 and a.d_model == 2048 and a.heads == 16 and a.vocab == 50368,
    }))


if __name__ == "__main__":
    main()
