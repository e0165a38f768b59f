"""llark_b200's own training entry point (`python -m llark_b200.train`, the counterpart of m2t/train.py for use without
a reference checkout): the flags of scripts/training/train_llark.sh parse through llark_b200.arguments, and `main(argv)`
runs from_pretrained -> tokenizer surgery -> projector / audio tokens -> tar-shard data module -> WrappedTrainer ->
checkpoints on CPU (tests/cpu_ops.py stands in for the CUDA ops: host logic only), then resumes from the checkpoint it
left, as m2t/train.py:257-258 does when output_dir already holds one."""
import json
import os
import random
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))


def test_launch_script_flags_parse():
    from llark_b200.arguments import TrainingArguments, check_supported, parse_command_line, write_args_to_file

    argv = ("--model_name_or_path meta-llama/Llama-2-7b-chat-hf --output_dir /tmp/ck --gradient_checkpointing True "
            "--per_device_train_batch_size 2 --gradient_accumulation_steps 4 --learning_rate 5e-5 --freeze_backbone False "
            "--tune_mm_mlp_adapter True --save_total_limit 1 --mm_use_audio_start_end --report_to wandb --bf16 True "
            "--tf32 True --lr_scheduler_type cosine --warmup_ratio 0.03 --weight_decay 0. --max_steps 100000 "
            "--model_max_length 2048 --save_strategy steps --save_steps 5000 --logging_steps 1 "
            "--ddp_find_unused_parameters False --dataloader_num_workers 8 "
            "--train_data_path a-{000000..000999}.tar,b-{000000..000999}.tar --evaluation_strategy no").split()
    m, d, t = parse_command_line(argv)
    assert (m.model_name_or_path, m.tune_mm_mlp_adapter, m.mm_use_audio_start_end, m.freeze_backbone, m.mm_hidden_size) == (
        "meta-llama/Llama-2-7b-chat-hf", True, True, False, 4800)
    assert d.train_data_path == "a-{000000..000999}.tar,b-{000000..000999}.tar" and d.is_multimodal
    assert (t.max_steps, t.bf16, t.gradient_checkpointing, t.warmup_ratio, t.model_max_length, t.save_steps,
            t.gradient_accumulation_steps, t.per_device_train_batch_size, t.learning_rate, t.eval_strategy, t.bits,
            t.optim) == (100000, True, True, 0.03, 2048, 5000, 4, 2, 5e-5, "no", 16, "adamw_torch")
    check_supported(t)
    for bad in (dict(bits=4), dict(lora_enable=True), dict(fsdp="full_shard")):
        with pytest.raises(NotImplementedError):
            check_supported(TrainingArguments(lr_scheduler_type="cosine", **bad))
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        write_args_to_file(["--bf16", "True", "--max_steps", "3", "--mm_use_audio_start_end"], tmp)
        write_args_to_file(["--bf16", "True", "--max_steps", "3", "--mm_use_audio_start_end"], tmp)   # overwrite
        assert open(os.path.join(tmp, "args.txt")).read() == "--bf16 True \\\n--max_steps 3 \\\n--mm_use_audio_start_end "


def test_own_train_entry_point_runs_and_resumes(tmp_path, monkeypatch):
    import cpu_ops
    import llark_b200.engine as eng
    import llark_b200.optim as optim
    import llark_b200.train as entry
    from test_reference_train_dropin_cpu import _shard, _tiny_checkpoint

    monkeypatch.setattr(eng, "ops", cpu_ops)
    monkeypatch.setattr(optim, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device)
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)
    rng = random.Random(0)
    ck = str(tmp_path / "tiny-llama-ckpt")
    base_vocab = _tiny_checkpoint(ck, rng)
    shard = str(tmp_path / "train-000000.tar")
    _shard(shard, rng)
    out_dir = str(tmp_path / "out")
    argv = ["--model_name_or_path", ck, "--output_dir", out_dir, "--train_data_path", shard, "--mm_hidden_size", "12",
            "--tune_mm_mlp_adapter", "True", "--mm_use_audio_start_end", "--bf16", "True", "--use_cpu", "True",
            "--max_steps", "4", "--per_device_train_batch_size", "2", "--gradient_accumulation_steps", "1",
            "--learning_rate", "1e-3", "--lr_scheduler_type", "cosine", "--logging_steps", "1", "--save_steps", "2",
            "--gradient_checkpointing", "True", "--evaluation_strategy", "no", "--model_max_length", "512"]
    out = entry.main(argv)
    trainer, tok, model = out["trainer"], out["tokenizer"], out["trainer"].model
    assert len(tok) == base_vocab + 4 and model.get_input_embeddings().weight.shape[0] == len(tok)
    assert model.config.use_cache is False and model.get_model().gradient_checkpointing is True
    assert next(model.get_model().mm_projector.parameters()).dtype == torch.bfloat16
    assert not any(p.requires_grad for p in model.get_output_embeddings().parameters())
    files = set(os.listdir(out_dir))
    assert {"pytorch_model.bin", "mm_projector.bin", "trainer_state.json", "checkpoint-2", "checkpoint-4"} <= files
    hist = json.load(open(os.path.join(out_dir, "trainer_state.json")))["log_history"]
    assert [h["step"] for h in hist] == [1, 2, 3, 4] and hist[-1]["loss"] < hist[0]["loss"]

    # second launch with a larger budget: output_dir holds checkpoints -> resume from the latest (step 4), 2 more steps;
    # --save_total_limit 1 (as train_llark.sh) leaves only the newest checkpoint folder, the projector side files stay
    argv[argv.index("--max_steps") + 1] = "6"
    again = entry.main(argv + ["--save_total_limit", "1"])["trainer"]
    assert sorted(n for n in os.listdir(out_dir) if n.startswith("checkpoint-")) == ["checkpoint-6"]
    assert {"checkpoint-2.bin", "checkpoint-4.bin", "checkpoint-6.bin"} <= set(os.listdir(os.path.join(out_dir, "mm_projector")))
    steps = [h["step"] for h in again.state["log_history"]]
    assert again.state["global_step"] == 6 and steps[-2:] == [5, 6] and again.optimizer.step_count == 6


def test_periodic_evaluation(tmp_path, monkeypatch):
    """--evaluation_strategy steps --eval_steps 2 with an eval shard: `eval_loss` entries appear in the log at steps 2 and
    4, evaluation leaves parameters and the gradient bucket alone, and trainer.evaluate() can be called on its own."""
    import cpu_ops
    import llark_b200.engine as eng
    import llark_b200.optim as optim
    import llark_b200.train as entry
    from test_reference_train_dropin_cpu import _shard, _tiny_checkpoint

    monkeypatch.setattr(eng, "ops", cpu_ops)
    monkeypatch.setattr(optim, "ops", cpu_ops)

    def ensure_packed(self):
        named = self._named()
        if self.flat is None or not self.flat.is_current(named):
            self.flat = eng.FlatBuckets(named, named[0][1].device)
        return self.flat

    monkeypatch.setattr(eng.LlamaEngine, "ensure_packed", ensure_packed)
    rng = random.Random(0)
    ck = str(tmp_path / "tiny-llama-ckpt")
    _tiny_checkpoint(ck, rng)
    train_shard, eval_shard = str(tmp_path / "train-000000.tar"), str(tmp_path / "eval-000000.tar")
    _shard(train_shard, rng)
    _shard(eval_shard, rng, n=3)
    argv = ["--model_name_or_path", ck, "--output_dir", str(tmp_path / "out"), "--train_data_path", train_shard,
            "--eval_data_path", eval_shard, "--mm_hidden_size", "12", "--tune_mm_mlp_adapter", "True",
            "--mm_use_audio_start_end", "--bf16", "True", "--use_cpu", "True", "--max_steps", "4",
            "--per_device_train_batch_size", "2", "--per_device_eval_batch_size", "2", "--learning_rate", "1e-3",
            "--lr_scheduler_type", "cosine", "--logging_steps", "1", "--save_strategy", "no",
            "--evaluation_strategy", "steps", "--eval_steps", "2"]
    trainer = entry.main(argv)["trainer"]
    evals = [h for h in trainer.state["log_history"] if "eval_loss" in h]
    assert [h["step"] for h in evals] == [2, 4] and all(h["eval_batches"] == 2 for h in evals)   # 3 samples, batch 2
    assert all(0 < h["eval_loss"] < 20 for h in evals) and evals[1]["eval_loss"] < evals[0]["eval_loss"] + 0.5
    flat = trainer.model.engine.flat
    before, grads = flat.params.clone(), flat.grads.clone()
    again = trainer.evaluate()
    assert torch.equal(flat.params, before) and torch.equal(flat.grads, grads)
    # (the audio block goes before or after the question at random per pass, m2t/data_modules.py:328-329: not bit-equal)
    assert abs(again["eval_loss"] - evals[1]["eval_loss"]) < 0.1 and trainer.model.training

