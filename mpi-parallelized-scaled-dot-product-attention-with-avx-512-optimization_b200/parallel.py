"""torch.distributed plumbing for the one-process-per-GPU launch (torchrun).

torch.distributed is used only to bootstrap (broadcast the ncclUniqueId that the C
library's own NCCL communicator is built from) and to agree on timings; the data-path
collectives (allreduce MAX / SUM, reduce SUM -- attention-mpi.c:342,354,380) are issued
by the C library itself.  Works on the ``gloo`` backend on CPU for the host-logic tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, group=None) -> bytes:
    """Broadcast a fixed-size byte string from ``src`` to every rank (CPU tensor: gloo; or
    staged through the current CUDA device for an nccl-only group)."""
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if dist.get_rank(group) == src:
        if payload is None or len(payload) != nbytes:
            raise ValueError("source rank must supply exactly nbytes")
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def shard_rows(n: int, world: int, rank: int) -> tuple[int, int]:
    """(first row, row count) of ``rank``'s K/V shard -- owner_disp / owner_count through the C ABI."""
    from . import host
    return host.owner_disp(n, world, rank), host.owner_count(n, world, rank)


def bootstrap_context(precision="auto", q_batch: int = 0, kv_splits: int = 0, local_rank: int | None = None, group=None,
                      merge: str = "nccl2"):
    """Create the per-rank :class:`~host.Context` of a one-GPU-per-process job: rank 0 draws a
    ncclUniqueId, torch.distributed broadcasts it, every rank builds its communicator."""
    import os
    from . import host
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if local_rank is None:
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
    uid = None
    if world > 1:
        uid = broadcast_bytes(host.get_unique_id() if rank == 0 else None, 128, src=0, group=group)
    return host.Context(precision=precision, q_batch=q_batch, kv_splits=kv_splits, num_local=1, first_device=local_rank,
                        world_size=world, rank_base=rank, nccl_id=uid, merge=merge)


def max_over_ranks(value: float, group=None) -> float:
    """MAX-reduce a scalar over the ranks (the reference's MPI_Reduce(MAX) of the elapsed time, mpi.c:524)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
