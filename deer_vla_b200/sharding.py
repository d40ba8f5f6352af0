"""Episode sharding across GPUs (robot_flamingo/eval/eval_utils.py:523-527, 565-568; eval_calvin.py:644).

The path shards by independent evaluation sequences: a full weight replica per rank, zero communication
during rollouts, and ONE collective at the end -- a gather of fixed-size action / exit-id tensors to rank 0
(the reference pickles python lists through ``gather_object``).  Works with NCCL (GPU tensors) and gloo.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_sequences(sequences: Sequence, rank: int, world: int) -> Sequence:
    """eval_sequences[rank*k:(rank+1)*k] with k = N / world; N must divide evenly (eval_utils.py:523-527)."""
    n = len(sequences)
    assert n % world == 0, "NUM_SEQUENCES must be divisible by the number of devices"
    k = n // world
    return sequences[rank * k:(rank + 1) * k]


def owner_of(sequence_index: int, n_sequences: int, world: int) -> int:
    return sequence_index // (n_sequences // world)


def gather_results(actions: torch.Tensor, exit_layers: torch.Tensor, dst: int = 0) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """actions [k, ..., 7] float, exit_layers [k, ...] int (same leading dims on every rank) -> on dst: the
    concatenation over ranks in rank order (== the reference's sequence order), elsewhere None."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return actions, exit_layers
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = torch.cat([actions.reshape(actions.shape[0], -1).float(),
                         exit_layers.reshape(exit_layers.shape[0], -1).float()], dim=1).contiguous()
    bufs: Optional[List[torch.Tensor]] = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, bufs, dst=dst)
    if rank != dst:
        return None
    allp = torch.cat(bufs, dim=0)
    na = actions[0].numel()
    acts = allp[:, :na].reshape(-1, *actions.shape[1:])
    exits = allp[:, na:].reshape(-1, *exit_layers.shape[1:]).to(exit_layers.dtype)
    return acts, exits


def all_gather_values(values: torch.Tensor) -> torch.Tensor:
    """Threshold calibration's collective (value_net.py:195-201): per-rank (n_exit, n_rank_samples) delta tables ->
    (n_exit, n_total) on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return values
    src = values.contiguous()
    if dist.get_backend() == "nccl" and not src.is_cuda:       # NCCL moves device tensors only
        src = src.cuda()
    bufs = [torch.zeros_like(src) for _ in range(dist.get_world_size())]
    dist.all_gather(bufs, src)
    return torch.cat(bufs, dim=1).to(values.device)
