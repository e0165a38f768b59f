"""Fused optimizer over the engine's flat buckets: gradient all-reduce (when data-parallel) -> global-norm clip ->
AdamW, as ONE NCCL call and two HBM-bound kernels (sum of squares, fused AdamW) over 6.6 G bf16 elements.

Reference: HF Trainer step = DDP all-reduce + clip_grad_norm_(max_grad_norm=1.0) + torch.optim.AdamW on bf16
parameters without master weights (m2t/arguments.py:84 optim="adamw_torch", m2t/train.py:101-103,
scripts/training/train_llark.sh:28-36: lr 5e-5, wd 0). Algorithmic traffic: 7 passes x 2 B x n.
"""
import torch

from . import ops

BF16 = torch.bfloat16


class LlarkAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0, comm=None):
        self.engine = model.engine
        flat = self.engine.ensure_packed()
        params = [p for _, p, o, _ in flat.entries if o < flat.n_train]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.comm = comm
        self._flat_id = id(flat)
        n = flat.n_train
        dev = flat.params.device
        self.exp_avg = torch.zeros(n, device=dev, dtype=BF16)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=BF16)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self._ws = torch.empty(ops.lib().llark_sumsq_workspace(), device=dev, dtype=torch.uint8)
        self.step_count = 0
        self.launches = 0
        # Overlap of the bucket all-reduce with backward: slices of the bucket are reduced on a side stream as soon
        # as the layers that own them have finished their wgrad (reverse layer order); step() reduces what is left
        # (embedding / final norm / projector) and joins the side stream. `sync_gradients` plays the role of DDP's
        # no_sync(): the trainer clears it on all but the last micro-batch of an accumulation window.
        self.sync_gradients = True
        self._reduced = []  # [(start, end)] slices already enqueued for this step
        self._comm_stream = None
        if comm is not None and comm.world > 1:
            self._comm_stream = torch.cuda.Stream(device=dev)
            self.engine.grad_ready_hook = self._on_grad_ready

    def grad_norm(self):
        """Global gradient norm of the last step (device scalar; reading it syncs)."""
        world = self.comm.world if self.comm is not None else 1
        return self.sumsq.sqrt() / world

    def _on_grad_ready(self, start, end):
        if not self.sync_gradients:
            return
        ev = torch.cuda.Event()
        ev.record()  # on the compute stream: the slice is final once this point is reached
        self._comm_stream.wait_event(ev)
        self.comm.allreduce_(self.engine.flat.grads[start:end], stream=self._comm_stream)
        self._reduced.append((start, end))

    @torch.no_grad()
    def step(self, closure=None):
        from .engine import _Nvtx

        with _Nvtx("llark.optimizer"):
            return self._step()

    def _step(self):
        flat = self.engine.flat
        if id(flat) != self._flat_id:
            raise RuntimeError("the model was re-packed after the optimizer was created; build the optimizer last")
        n = flat.n_train
        g = flat.grads[:n]
        world = 1
        self.launches = 0
        if self.comm is not None and self.comm.world > 1:
            world = self.comm.world
            # reduce whatever the overlap hooks have not covered, then join the side stream
            pos = 0
            for (a, b) in sorted(self._reduced) + [(n, n)]:
                if a > pos:
                    self.comm.allreduce_(g[pos:a])
                    self.launches += 1
                pos = max(pos, b)
            if self._reduced:
                done = torch.cuda.Event()
                done.record(self._comm_stream)
                torch.cuda.current_stream().wait_event(done)
                self.launches += len(self._reduced)
            self._reduced = []
        ops.grad_sumsq(g, self.sumsq, False, self._ws)
        grp = self.param_groups[0]
        self.step_count += 1
        ops.adamw_step(flat.params[:n], g, self.exp_avg, self.exp_avg_sq, self.sumsq, float(grp["lr"]),
                       grp["betas"][0], grp["betas"][1], grp["eps"], grp["weight_decay"], self.step_count,
                       grad_scale=1.0 / world, max_norm=self.max_grad_norm or 0.0, zero_grad=False)
        flat.consumed = True  # next backward overwrites the bucket (no 13 GB memset per step)
        self.launches += 3
        return None

    def zero_grad(self, set_to_none: bool = False):
        flat = self.engine.flat
        if set_to_none:
            super().zero_grad(set_to_none=True)
        elif flat is not None and not flat.consumed:
            flat.grads.zero_()
