"""World-size-2 (gloo, CPU) test of the HOST logic of the data-parallel step: two ranks run
WrappedTrainer + LlarkAdamW over different micro-batches with gradient accumulation; the bucket slices the engine
announces as final during backward (`grad_ready_hook`, reverse layer order) are all-reduced as they appear, step()
reduces the remainder, scales by 1/world and applies clip + AdamW. Stand-ins (test infrastructure): tests/cpu_ops.py for
the CUDA ops, a gloo all-reduce for `GradComm.allreduce_`, inert objects for torch.cuda streams / events.

Checked: every element of the trainable bucket is reduced exactly once per optimizer step and only on the last
micro-batch of an accumulation window (DDP no_sync semantics); both ranks end bit-identical; the result equals a
single-process emulation of DDP (mean over ranks of the accumulated per-rank gradients) up to bf16 rounding."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.dirname(os.path.abspath(__file__))
STEPS, ACCUM, WORLD = 2, 2, 2


def _patch_cpu(eng, optim, cpu_ops):
    eng.ops = cpu_ops
    optim.ops = cpu_ops

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device)
        return self.flat

    eng.LlamaEngine.ensure_packed = ensure_packed


class _Inert:
    """torch.cuda.Stream / Event stand-in: ordering is trivially satisfied on one CPU thread."""

    def __init__(self, *a, **k):
        pass

    def record(self, *a, **k):
        pass

    def wait_event(self, *a, **k):
        pass


class GlooComm:
    """GradComm stand-in: sum all-reduce of a bucket slice through gloo (bf16 summed in fp32, rounded once, like NCCL's
    bf16 ring up to summation order at world 2)."""

    def __init__(self, world, rank, log):
        self.world, self.rank, self.log = world, rank, log

    def allreduce_(self, grads_flat, stream=None):
        self.log.append((grads_flat.storage_offset(), grads_flat.numel(), stream is not None))
        t = grads_flat.float()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        grads_flat.copy_(t.to(grads_flat.dtype))
        return grads_flat


def _model_and_data(rank_of_batch):
    from oracle import llama_oracle as lo
    from test_engine_host_cpu import _build

    cfg = lo.LlamaCfg(num_layers=6)       # comm_chunk_layers = 4 -> slices [4,6) and, tail split, [3,4) [2,3) [1,2) [0,1)
    model = _build(cfg, lo.init_params(cfg, seed=0))
    for p in model.get_output_embeddings().parameters():
        p.requires_grad = False
    model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    data = [lo.make_batch(cfg, B=2, S=40, F_frames=4, seed=200 + 10 * r + i) for r in rank_of_batch
            for i in range(STEPS * ACCUM)]
    return cfg, model, data


def _args(out):
    from types import SimpleNamespace

    return SimpleNamespace(max_steps=STEPS, per_device_train_batch_size=1, gradient_accumulation_steps=ACCUM,
                           learning_rate=1e-3, warmup_ratio=0.0, max_grad_norm=1.0, weight_decay=0.0, logging_steps=1,
                           save_steps=0, output_dir=out, dataloader_num_workers=0, tune_mm_mlp_adapter=True)


def _worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, TESTS)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)  # two workers share the host: avoid oversubscribing it
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_ops
    import llark_b200.engine as eng
    import llark_b200.optim as optim
    from llark_b200.trainer import WrappedTrainer

    _patch_cpu(eng, optim, cpu_ops)
    torch.cuda.Stream, torch.cuda.Event = _Inert, _Inert
    torch.cuda.current_stream = lambda *a, **k: _Inert()
    _, model, data = _model_and_data([rank])
    log = []
    trainer = WrappedTrainer(model=model, tokenizer=None, args=_args(os.path.join(tmp, f"r{rank}")),
                             train_dataset=data, data_collator=lambda items: items[0],
                             comm=GlooComm(world, rank, log))
    trainer.train()
    flat = model.engine.flat
    q.put((rank, flat.params.float().numpy().copy(), log, flat.n_train,      # numpy: pickled by value
           [h["grad_norm"] for h in trainer.state["log_history"]]))
    dist.barrier()


def test_world2_trainer_overlapped_bucket_allreduce(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, str(tmp_path), q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(WORLD)), key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, p0, log0, n_train, norms0), (_, p1, log1, _, norms1) = res
    p0, p1 = torch.from_numpy(p0), torch.from_numpy(p1)
    assert torch.equal(p0, p1) and norms0 == norms1                      # replicas stay bit-identical
    assert log0 == log1
    # one optimizer step = the trainable bucket covered exactly once (no gaps, no overlaps), hooks first
    per_step = len(log0) // STEPS
    assert per_step * STEPS == len(log0) and per_step >= 3
    for s in range(STEPS):
        calls = log0[s * per_step:(s + 1) * per_step]
        assert sum(1 for c in calls if c[2]) == 5                        # overlap slices on the side stream (tail split)
        spans = sorted((a, a + n) for a, n, _ in calls)
        assert spans[0][0] == 0 and spans[-1][1] == n_train
        assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))

    # single-process emulation of DDP: per rank accumulate ACCUM micro-batches of loss / ACCUM, average over ranks
    sys.path.insert(0, TESTS)
    import cpu_ops
    import llark_b200.engine as eng
    import llark_b200.optim as optim

    saved = (eng.ops, optim.ops, eng.LlamaEngine.ensure_packed)
    try:
        _patch_cpu(eng, optim, cpu_ops)
        from llark_b200.optim import LlarkAdamW
        from llark_b200.trainer import cosine_with_warmup

        _, model, data = _model_and_data([0, 1])
        per_rank = STEPS * ACCUM
        opt = LlarkAdamW(model, lr=1e-3, max_grad_norm=1.0)
        flat = model.engine.flat
        for step in range(STEPS):
            opt.param_groups[0]["lr"] = 1e-3 * cosine_with_warmup(step, STEPS, 0)
            rank_grads = []
            for r in range(WORLD):
                for m in range(ACCUM):
                    b = data[r * per_rank + step * ACCUM + m]
                    (model(**b).loss / ACCUM).backward()
                rank_grads.append(flat.grads[:flat.n_train].float().clone())
                flat.consumed = True                                      # next backward overwrites the bucket
            # the kernel receives the SUM and scales by 1 / world: present the emulation the same way at world 1
            flat.grads[:flat.n_train].copy_((torch.stack(rank_grads).sum(0) / WORLD).to(torch.bfloat16))
            flat.consumed = False
            opt.step()
        ref = flat.params.float().clone()
        _, fresh, _ = _model_and_data([0])
        init = fresh.engine.ensure_packed().params.float().clone()
    finally:
        eng.ops, optim.ops, eng.LlamaEngine.ensure_packed = saved
    # compare the UPDATE (parameters move by ~lr per step), relative to the size of the update itself
    upd_ref, upd = (ref - init)[:n_train], (p0 - init)[:n_train]
    cos = torch.nn.functional.cosine_similarity(upd, upd_ref, dim=0).item()
    assert cos > 0.98, cos
    assert abs(upd.norm().item() - upd_ref.norm().item()) < 0.03 * upd_ref.norm().item()
