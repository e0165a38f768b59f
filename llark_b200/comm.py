"""Data-parallel exchange step: ONE NCCL all-reduce of the flat bf16 gradient bucket per optimizer step
(reference: HF Trainer + DDP, scripts/training/train_llark.sh:20-22,44).

The communicator lives in libllark_b200.so (llark_comm_init / llark_grad_bucket_allreduce); torch.distributed is
only the out-of-band channel that carries the 128-byte ncclUniqueId from rank 0 to the other ranks.
"""
import ctypes

import torch

from ._lib import check, lib


def shard_seeds(rank, world, n, base=0):
    """Synthetic-data partition: rank r draws seeds {base + r, base + r + world, ...}."""
    return [base + rank + i * world for i in range(n)]


def broadcast_unique_id(rank, group=None):
    """Rank 0 creates the ncclUniqueId; everyone receives it through torch.distributed (any backend)."""
    import torch.distributed as dist

    buf = (ctypes.c_ubyte * 128)()
    if rank == 0:
        check(lib().llark_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)), "llark_comm_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().tolist())


class GradComm:
    """Owns the NCCL communicator used for the gradient bucket."""

    def __init__(self, world, rank, unique_id: bytes):
        self.world, self.rank = world, rank
        self._comm = ctypes.c_void_p()
        idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(unique_id)
        check(lib().llark_comm_init(ctypes.byref(self._comm), ctypes.cast(idbuf, ctypes.c_void_p), world, rank),
              "llark_comm_init")

    @classmethod
    def from_torch_distributed(cls):
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        return cls(world, rank, broadcast_unique_id(rank))

    def allreduce_(self, grads_flat, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream()
        check(lib().llark_grad_bucket_allreduce(self._comm, ctypes.c_void_p(grads_flat.data_ptr()), grads_flat.numel(),
                                                ctypes.c_void_p(s.cuda_stream)), "llark_grad_bucket_allreduce")
        return grads_flat

    def close(self):
        if self._comm:
            lib().llark_comm_destroy(self._comm)
            self._comm = ctypes.c_void_p()
