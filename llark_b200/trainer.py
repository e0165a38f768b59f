"""`WrappedTrainer`: the training-loop object m2t/train.py:255-262 builds and drives
(`WrappedTrainer(model=, tokenizer=, args=, train_dataset=, eval_dataset=, data_collator=)`, `.train()`,
`.save_state()`, `._save()`), mirrored from m2t/models/trainer.py:34-129.

The reference subclasses transformers.Trainer (DDP + torch AdamW + autocast). On B200 the step is executed by the
fused engine, so this is a self-contained loop with the reference's semantics:
  * micro-batch loss divided by gradient_accumulation_steps; gradients accumulate in the flat bf16 bucket;
  * every `gradient_accumulation_steps` micro-batches: ONE NCCL all-reduce of the bucket (world > 1), global-norm
    clip at max_grad_norm, bf16 AdamW (no master weights), cosine schedule with warm-up ratio;
  * the loss is each rank's mean over its own non-ignored tokens, averaged across ranks (not a global token mean);
  * `_save` additionally writes {mm_projector, embed_tokens} tensors to mm_projector/checkpoint-N.bin when
    tune_mm_mlp_adapter is set (trainer.py:35-63); checkpoints keep the HF state_dict key names.
It does not depend on transformers.Trainer / accelerate (absent from this image).
"""
import json
import logging
import math
import os
import time
from typing import Optional

import torch

from llark_b200.optim import LlarkAdamW


def cosine_with_warmup(step, total, warmup):
    if step < warmup:
        return float(step) / float(max(1, warmup))
    progress = float(step - warmup) / float(max(1, total - warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))


def unwrap_model(model):
    """m2t/llava/train/llava_trainer.py:4-15."""
    return unwrap_model(model.module) if hasattr(model, "module") else model


class WrappedTrainer:
    def __init__(self, model=None, tokenizer=None, args=None, train_dataset=None, eval_dataset=None,
                 data_collator=None, comm=None, **_):
        self.model, self.tokenizer, self.args = model, tokenizer, args
        self.train_dataset, self.eval_dataset, self.data_collator = train_dataset, eval_dataset, data_collator
        self.deepspeed = None
        self.comm = comm
        self.state = {"global_step": 0, "log_history": []}
        self.optimizer = None
        model.validate = "lazy"
        seed = self._arg("seed", None)
        if seed is not None:  # transformers.Trainer.__init__ calls set_seed(args.seed): python, numpy and torch generators
            import random

            import numpy as np

            random.seed(seed)
            np.random.seed(seed)
            torch.manual_seed(seed)
        if self._arg("gradient_checkpointing", False):  # transformers.Trainer turns the flag into this call
            model.gradient_checkpointing_enable()
        # transformers.Trainer places the model on args.device in its constructor (m2t/train.py never does it itself)
        dev = self._arg("device", None)
        if dev is not None and next(model.parameters()).device != torch.device(dev):
            model.to(torch.device(dev))
        placed = next(model.parameters()).device
        if placed.type == "cuda" and torch.cuda.is_available():
            torch.cuda.set_device(placed)  # kernels are launched on the thread's current device

    def _arg(self, name, default):
        return getattr(self.args, name, default) if self.args is not None else default

    def _ensure_comm(self):
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.comm is None and world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                dist.init_process_group("nccl")
            from llark_b200.comm import GradComm

            self.comm = GradComm.from_torch_distributed()

    def create_optimizer(self):
        self._ensure_comm()
        self.optimizer = LlarkAdamW(
            self.model, lr=self._arg("learning_rate", 5e-5),
            betas=(self._arg("adam_beta1", 0.9), self._arg("adam_beta2", 0.999)), eps=self._arg("adam_epsilon", 1e-8),
            weight_decay=self._arg("weight_decay", 0.0), max_grad_norm=self._arg("max_grad_norm", 1.0), comm=self.comm)
        return self.optimizer

    def _batches(self):
        bs = self._arg("per_device_train_batch_size", 2)
        loader = torch.utils.data.DataLoader(self.train_dataset, batch_size=bs, collate_fn=self.data_collator,
                                             num_workers=self._arg("dataloader_num_workers", 0), pin_memory=True)
        while True:
            for batch in loader:
                yield batch

    @torch.no_grad()
    def evaluate(self, eval_dataset=None, max_batches=None):
        """Mean loss over the evaluation set, as transformers.Trainer.evaluate reports `eval_loss` (mean of the per-batch
        mean losses; the reference's make_data_module returns `eval_dataset`, m2t/data_modules.py:707-717). Every rank
        evaluates the shards it was dealt; with more than one rank the per-rank means are averaged."""
        dataset = eval_dataset if eval_dataset is not None else self.eval_dataset
        if dataset is None:
            raise ValueError("no evaluation dataset: pass --eval_data_path (or eval_dataset=)")
        bs = self._arg("per_device_eval_batch_size", None) or self._arg("per_device_train_batch_size", 2)
        loader = torch.utils.data.DataLoader(dataset, batch_size=bs, collate_fn=self.data_collator,
                                             num_workers=self._arg("dataloader_num_workers", 0), pin_memory=True)
        dev = next(self.model.parameters()).device
        was_training = self.model.training
        self.model.eval()
        total, count = torch.zeros((), device=dev, dtype=torch.float32), 0
        try:
            for i, batch in enumerate(loader):
                if max_batches is not None and i >= max_batches:
                    break
                batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else
                             [t.to(dev, non_blocking=True) for t in v]) for k, v in batch.items()}
                total += self.model(**batch).loss.float()
                count += 1
        finally:
            self.model.train(was_training)
        self.model.check_audio_token_errors()
        mean = (total / max(count, 1)).item()
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([mean], device=dev)
            dist.all_reduce(t)
            mean = (t / world).item()
        metrics = {"eval_loss": mean, "eval_batches": count, "step": self.state["global_step"]}
        self.state["log_history"].append(metrics)
        logging.warning(json.dumps(metrics))
        return metrics

    def train(self, resume_from_checkpoint=None):
        max_steps = self._arg("max_steps", -1)
        if max_steps is None or max_steps <= 0:
            raise ValueError("the dataset is an endless stream (with_epoch(1e9) in the reference): set --max_steps")
        accum = self._arg("gradient_accumulation_steps", 1)
        warmup = int(self._arg("warmup_steps", 0) or math.ceil(max_steps * self._arg("warmup_ratio", 0.0)))
        base_lr = self._arg("learning_rate", 5e-5)
        if self.optimizer is None:
            self.create_optimizer()
        if resume_from_checkpoint:
            self._load_from_checkpoint(resume_from_checkpoint)
        dev = next(self.model.parameters()).device
        if resume_from_checkpoint and self.state["global_step"] and hasattr(self.train_dataset, "seed"):
            # do not replay the samples the interrupted run already consumed: the resumed stream draws its shards and
            # shuffle buffer from a generator offset by the restored step count
            self.train_dataset.seed = int(self.train_dataset.seed) + 7919 * int(self.state["global_step"])
        it = self._batches()
        self.model.train()
        t0 = time.time()
        while self.state["global_step"] < max_steps:
            step = self.state["global_step"]
            self.optimizer.param_groups[0]["lr"] = base_lr * cosine_with_warmup(step, max_steps, warmup)
            losses = []
            for micro in range(accum):
                self.optimizer.sync_gradients = micro == accum - 1  # DDP no_sync() on all but the last micro-batch
                batch = next(it)
                batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else
                             [t.to(dev, non_blocking=True) for t in v]) for k, v in batch.items()}
                out = self.model(**batch)
                (out.loss / accum).backward()
                losses.append(out.loss.detach())
            # the reference raises its splice ValueErrors (llamav2.py:146-175) inside forward, i.e. BEFORE the update is
            # applied: read the device flag word (one 4-byte sync per optimizer step) before stepping
            self.model.check_audio_token_errors()
            self.optimizer.step()
            self.state["global_step"] += 1
            if self.state["global_step"] % max(1, self._arg("logging_steps", 1)) == 0:
                entry = {"step": self.state["global_step"], "loss": torch.stack(losses).mean().item(),
                         "learning_rate": self.optimizer.param_groups[0]["lr"],
                         "grad_norm": self.optimizer.grad_norm().item(), "elapsed_s": time.time() - t0}
                self.state["log_history"].append(entry)
                logging.warning(json.dumps(entry))
            eval_steps = self._arg("eval_steps", None) or self._arg("logging_steps", 1)
            if (self._arg("eval_strategy", self._arg("evaluation_strategy", "no")) == "steps" and self.eval_dataset is not None
                    and self.state["global_step"] % max(1, eval_steps) == 0):
                self.evaluate(max_batches=self._arg("eval_max_batches", None))
            save_steps = self._arg("save_steps", 0)
            if save_steps and self.state["global_step"] % save_steps == 0 and self._is_main():
                ckpt = os.path.join(self._arg("output_dir", "."), f"checkpoint-{self.state['global_step']}")
                self._save(ckpt)  # (the mm_projector side file is keyed by this folder name)
                self._save_optimizer(ckpt)
                with open(os.path.join(ckpt, "COMPLETE"), "w") as f:  # written last: marks a fully written folder
                    f.write(str(self.state["global_step"]))
                self._rotate_checkpoints()
        return self.state

    def _rotate_checkpoints(self):
        """--save_total_limit N (train_llark.sh passes 1): keep the N most recent checkpoint-* folders of output_dir, as
        transformers.Trainer._rotate_checkpoints does; the mm_projector/checkpoint-N.bin side files are not rotated."""
        import re
        import shutil

        limit = self._arg("save_total_limit", None)
        if limit is None or limit <= 0:
            return
        out = self._arg("output_dir", ".")
        found = []
        for name in os.listdir(out):
            m = re.fullmatch(r"checkpoint-(\d+)", name)
            if m and os.path.isdir(os.path.join(out, name)):
                found.append((int(m.group(1)), name))
        found.sort()
        if found and not os.path.exists(os.path.join(out, found[-1][1], "COMPLETE")):
            return  # never delete older checkpoints on the strength of a partially written newest one
        for _, name in found[:-limit]:
            logging.warning(f"deleting older checkpoint {name} (save_total_limit={limit})")
            shutil.rmtree(os.path.join(out, name), ignore_errors=True)

    # ------------------------------------------------------------------------------------------ optimizer state
    # transformers.Trainer writes optimizer.pt / scheduler.pt next to the weights of every checkpoint-N and restores
    # them under resume_from_checkpoint (which m2t/train.py:257-258 requests whenever a checkpoint-* folder exists).
    # Here the whole optimizer state is the two flat bf16 moment buffers and the step count; the cosine schedule is a
    # pure function of global_step, so no scheduler file is needed.
    def _save_optimizer(self, ckpt_dir):
        opt = self.optimizer
        torch.save({"exp_avg": opt.exp_avg.detach().cpu(), "exp_avg_sq": opt.exp_avg_sq.detach().cpu(),
                    "step_count": opt.step_count, "global_step": self.state["global_step"],
                    "n_train": int(opt.exp_avg.numel())}, os.path.join(ckpt_dir, "optimizer.pt"))
        with open(os.path.join(ckpt_dir, "trainer_state.json"), "w") as f:
            json.dump(self.state, f)

    def _load_optimizer(self, ckpt_dir):
        path = os.path.join(ckpt_dir, "optimizer.pt")
        if not os.path.exists(path):
            logging.warning(f"{path} not found: resuming with fresh AdamW moments")
            return
        st = torch.load(path, map_location="cpu")
        opt = self.optimizer
        if st["n_train"] != opt.exp_avg.numel():
            raise ValueError("optimizer.pt does not match the model's trainable parameter bucket")
        opt.exp_avg.copy_(st["exp_avg"])
        opt.exp_avg_sq.copy_(st["exp_avg_sq"])
        opt.step_count = int(st["step_count"])

    def _is_main(self):
        return int(os.environ.get("RANK", "0")) == 0

    def save_state(self):
        if not self._is_main():
            return
        out = self._arg("output_dir", ".")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "trainer_state.json"), "w") as f:
            json.dump(self.state, f)

    def save_model(self, output_dir=None, state_dict=None):
        if self._is_main():
            self._save(output_dir or self._arg("output_dir", "."), state_dict)

    def _save(self, output_dir: Optional[str] = None, state_dict=None):
        os.makedirs(output_dir, exist_ok=True)
        model = unwrap_model(self.model)
        sd = state_dict if state_dict is not None else model.state_dict()
        if self._arg("tune_mm_mlp_adapter", False):
            keep = {k: v.detach().cpu() for k, v in sd.items()
                    if any(s in k for s in ("mm_projector", "embed_tokens", "embed_in"))}
            folder = os.path.basename(os.path.normpath(output_dir))
            if folder.startswith("checkpoint-"):
                side = os.path.join(os.path.dirname(os.path.normpath(output_dir)), "mm_projector")
                os.makedirs(side, exist_ok=True)
                torch.save(keep, os.path.join(side, f"{folder}.bin"))
            else:
                torch.save(keep, os.path.join(output_dir, "mm_projector.bin"))
        torch.save({k: v.detach().cpu() for k, v in sd.items()}, os.path.join(output_dir, "pytorch_model.bin"))
        model.config.save_pretrained(output_dir)
        if self.tokenizer is not None and hasattr(self.tokenizer, "save_pretrained"):
            self.tokenizer.save_pretrained(output_dir)

    def _load_from_checkpoint(self, resume_from_checkpoint, model=None):
        """Load model weights (non-projector keys with strict=False, then projector keys) -- trainer.py:67-129."""
        model = model if model is not None else unwrap_model(self.model)
        path = resume_from_checkpoint
        if path is True:
            out = self._arg("output_dir", ".")
            cks = sorted((d for d in os.listdir(out) if d.startswith("checkpoint-")
                          and os.path.exists(os.path.join(out, d, "pytorch_model.bin"))),
                         key=lambda s: int(s.split("-")[1]))
            if not cks:
                return
            path = os.path.join(out, cks[-1])
        sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        proj = {k: v for k, v in sd.items() if "mm_projector" in k}
        rest = {k: v for k, v in sd.items() if "mm_projector" not in k}
        model.load_state_dict(rest, strict=False)
        if proj:
            model.get_model().mm_projector.load_state_dict({k.split(".")[-1]: v for k, v in proj.items()})
        step = os.path.basename(os.path.normpath(path)).split("-")[-1]
        if step.isdigit():
            self.state["global_step"] = int(step)
        state_file = os.path.join(path, "trainer_state.json")
        if os.path.exists(state_file):
            with open(state_file) as f:
                self.state["log_history"] = json.load(f).get("log_history", [])
        if self.optimizer is not None:
            self._load_optimizer(path)
