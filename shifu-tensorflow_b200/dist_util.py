"""Host-side plumbing of the data-parallel path (one process per GPU).  torch.distributed is used for rendezvous only:
the NCCL communicator that carries the gradients lives inside libshifu_b200.so and is created from a unique id that
rank 0 makes and every rank receives here.  Works on any torch.distributed backend (gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List

import numpy as np


def broadcast_bytes(dist, payload_fn: Callable[[], bytes], n_bytes: int, rank: int, device="cpu") -> bytes:
    """rank 0 calls payload_fn() (e.g. capi.nccl_unique_id); everybody returns the same n_bytes."""
    import torch
    buf = torch.zeros(n_bytes, dtype=torch.uint8, device=device)
    if rank == 0:
        data = payload_fn()
        if len(data) != n_bytes:
            raise ValueError("payload must be %d bytes" % n_bytes)
        buf.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(dist, value: float, world: int, device="cpu") -> float:
    """timing rule: a multi-GPU number is the MAX over ranks"""
    if world == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_rows(n_rows: int, rank: int, world: int) -> np.ndarray:
    """row r of a global batch goes to rank r % world: every rank gets ceil/floor(n/world) rows"""
    return np.arange(rank, n_rows, world)


def shard_files(files: List[str], rank: int, world: int) -> List[str]:
    """round-robin file split across workers, like TrainingDataSet.java:65-82"""
    return [f for i, f in enumerate(files) if i % world == rank]


def all_gather_bytes(dist, payload: bytes, world: int, device="cpu") -> List[bytes]:
    """every rank contributes len(payload) bytes; returns the list in rank order (IPC handle exchange)"""
    import torch
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    outs = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def enable_peer_exchange(dist, trainer, world: int, device="cpu") -> bool:
    """switch a trainer from NCCL to the peer-memory all-reduce kernel (all ranks must call this)"""
    if world <= 1:
        return False
    trainer.set_peer_handles(all_gather_bytes(dist, trainer.ipc_handle(), world, device))
    dist.barrier()
    return True
