"""Two-GPU test of the data-parallel exchange step (skipped on boxes with one GPU): each rank runs the CUDA
forward/backward on its own micro-batch, the flat gradient bucket is all-reduced by llark_grad_bucket_allreduce
(NCCL inside libllark_b200.so), the fused AdamW applies grad_scale = 1/world -- the result must equal a
single-process run that averages the two ranks' gradients (what DDP computes in the reference)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import llama_oracle as lo
    from llark_b200.comm import GradComm
    from llark_b200.optim import LlarkAdamW
    from llark_b200.testing import build_model, to_cuda

    cfg = lo.LlamaCfg()
    params = lo.init_params(cfg, seed=0)
    comm = GradComm.from_torch_distributed()
    batch = lo.make_batch(cfg, B=2, S=48, F_frames=5, seed=100 + rank)
    # phase A: no overlap (the hooks are muted like DDP's no_sync), step() reduces the whole bucket
    model = build_model(cfg, params)
    opt = LlarkAdamW(model, lr=1e-2, max_grad_norm=1.0, comm=comm)
    opt.sync_gradients = False
    out = model(**to_cuda(batch))
    out.loss.backward()
    torch.cuda.synchronize()
    local = model.engine.flat.grads.float().clone()
    opt.sync_gradients = True
    opt.step()
    new_params = model.engine.flat.params.float().cpu()
    # phase B: slices are all-reduced on the side stream while backward is still running
    model_b = build_model(cfg, params)
    opt_b = LlarkAdamW(model_b, lr=1e-2, max_grad_norm=1.0, comm=comm)
    model_b(**to_cuda(batch)).loss.backward()
    assert opt_b._reduced, "overlap hooks did not fire"
    opt_b.step()
    overlapped = model_b.engine.flat.params.float().cpu()
    assert torch.equal(overlapped, new_params), "overlapped and non-overlapped reductions must agree bit for bit"
    q.put((rank, local.cpu(), new_params, out.loss.item()))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bucket_allreduce_matches_gradient_average():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 100
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, g0, p0, l0), (_, g1, p1, l1) = res
    assert torch.equal(p0, p1), "ranks must hold identical parameters after the step"
    # single-process replay of the same update from the averaged gradient
    from oracle import llama_oracle as lo
    from llark_b200.optim import LlarkAdamW
    from llark_b200.testing import build_model

    cfg = lo.LlamaCfg()
    model = build_model(cfg, lo.init_params(cfg, seed=0))
    opt = LlarkAdamW(model, lr=1e-2, max_grad_norm=1.0)
    flat = model.engine.ensure_packed()
    avg = ((g0 + g1) / 2).to(torch.bfloat16)
    flat.grads.copy_(avg.cuda())
    opt.step()
    mine = flat.params.float().cpu()
    diff = (mine - p0).abs().max().item()
    assert diff <= 2e-2, diff  # bf16 rounding of sum-then-scale vs average-then-round
    assert (mine != p0).float().mean().item() < 0.2
