"""ORACLE support (test infrastructure only): tests/golden/trainstep_ref.pt -- a short optimisation TRAJECTORY of the
reference's own model class under the training regime of scripts/training/train_llark.sh, to pin SURVEY §8(a) rows
T1-T3 (pure-bf16 weights and AdamW moments, micro-batch losses divided by the accumulation count, clip at 1.0 on the
accumulated gradient, torch.optim.AdamW, cosine schedule with warm-up stepped once per optimizer step).

    python oracle/make_golden_trainstep.py

What runs: m2t/models/llamav2.py WrappedLlamav2ForCausalLM imported in place (oracle/ref_shims.py), cast to bf16 as
m2t/train.py:101-103 does, forward under torch.autocast(bf16) as `--bf16 True` makes transformers.Trainer do. The loop
around it restates transformers==4.29.2 `Trainer._inner_training_loop` (the version requirements.txt:4 pins; it is not
installed here -- 5.5's Trainer normalises the loss over the accumulation window differently, so it is NOT used):
    for each micro-batch:  loss = model(**batch).loss / gradient_accumulation_steps ; loss.backward()
    every `accum` micro-batches: clip_grad_norm_(params, max_grad_norm) ; optimizer.step() ; lr_scheduler.step() ;
                                 model.zero_grad()
with optimizer = torch.optim.AdamW (m2t/arguments.py:84 `optim="adamw_torch"`), betas (0.9, 0.999), eps 1e-8,
weight_decay 0, and lr_scheduler = transformers.get_cosine_schedule_with_warmup(ceil(max_steps * warmup_ratio), max_steps)
(unchanged between 4.29.2 and 5.5; imported from the installed transformers).
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llama_oracle as lo  # noqa: E402
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
HP = dict(max_steps=4, accum=2, lr=2e-3, warmup_ratio=0.25, max_grad_norm=1.0, B=2, S=48, F_frames=5)


def batches(cfg):
    return [lo.make_batch(cfg, B=HP["B"], S=HP["S"], F_frames=HP["F_frames"], seed=100 + i)
            for i in range(HP["max_steps"] * HP["accum"])]


def main():
    from transformers import get_cosine_schedule_with_warmup

    from oracle.make_golden import build_reference_model

    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    model = build_reference_model(cfg, params, torch.bfloat16)  # lm_head frozen, bf16 weights
    model.get_model().orig_embeds_params = [model.get_input_embeddings().weight.data.clone()]
    model.train()
    train_params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(train_params, lr=HP["lr"], betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    sched = get_cosine_schedule_with_warmup(opt, math.ceil(HP["max_steps"] * HP["warmup_ratio"]), HP["max_steps"])
    start = {n: p.detach().clone() for n, p in model.named_parameters()}
    data = batches(cfg)
    micro_losses, lrs, norms = [], [], []
    for step in range(HP["max_steps"]):
        lrs.append(sched.get_last_lr()[0])
        for m in range(HP["accum"]):
            b = data[step * HP["accum"] + m]
            with torch.autocast("cpu", dtype=torch.bfloat16):
                out = model(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                            audio_encodings=b["audio_encodings"])
            micro_losses.append(out.loss.detach().float().item())
            (out.loss / HP["accum"]).backward()
        norms.append(torch.nn.utils.clip_grad_norm_(train_params, HP["max_grad_norm"]).float().item())
        opt.step()
        sched.step()
        model.zero_grad()
    final = {}
    for n, p in model.named_parameters():
        d = (p.detach().float() - start[n].float())
        entry = {"delta_norm": d.norm().item(), "start_norm": start[n].float().norm().item()}
        if p.numel() <= 5000:
            entry["final"] = p.detach().clone()
        else:
            entry["final_head"] = p.detach().flatten()[:2048].clone()
            entry["delta_head"] = d.flatten()[:2048].clone()
        if "embed_tokens" in n:
            entry["changed_rows"] = d.abs().sum(-1).nonzero().flatten().tolist()
        final[n] = entry
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        b = data[0]
        end_loss = model(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                         audio_encodings=b["audio_encodings"]).loss.float().item()
    res = {"hp": HP, "cfg": cfg.__dict__, "seed": 0, "micro_losses": micro_losses, "lrs": lrs, "grad_norms": norms,
           "final": final, "loss_batch0_after": end_loss}
    torch.save(res, os.path.join(GOLD, "trainstep_ref.pt"))
    print("micro losses", [round(x, 4) for x in micro_losses])
    print("lrs", lrs, "grad norms", [round(x, 3) for x in norms], "batch0 loss after", round(end_loss, 4))
    print("lm_head delta", final["lm_head.weight"]["delta_norm"], "embed rows changed",
          final["model.embed_tokens.weight"]["changed_rows"])


if __name__ == "__main__":
    ref_shims.install()
    os.makedirs(GOLD, exist_ok=True)
    main()
