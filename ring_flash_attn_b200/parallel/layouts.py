"""Sharding helpers: how a full sequence maps onto the ranks for each scheme.

The reference leaves data layout to the caller and only documents it in its tests
(/root/reference/test/test_zigzag_ring_flash_attn_func.py:9-14,
test_stripe_flash_attn_func.py:9-14, test_ring_flash_attn_varlen_func.py:9-15,
test_zigzag_ring_flash_attn_varlen_func.py:9-20).  Here the layouts are public utilities, with
inverses, so user code (and RoPE position ids) can be derived from one place.
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def shard_ring(x: torch.Tensor, rank: int, world: int, dim: int = 1) -> torch.Tensor:
    return x.chunk(world, dim=dim)[rank].contiguous()


def shard_zigzag(x: torch.Tensor, rank: int, world: int, dim: int = 1) -> torch.Tensor:
    chunks = x.chunk(2 * world, dim=dim)
    return torch.cat([chunks[rank], chunks[2 * world - 1 - rank]], dim=dim).contiguous()


def shard_stripe(x: torch.Tensor, rank: int, world: int, dim: int = 1) -> torch.Tensor:
    idx = torch.arange(rank, x.shape[dim], world, device=x.device)
    return x.index_select(dim, idx).contiguous()


def shard_ring_varlen(x: torch.Tensor, cu_seqlens: Sequence[int], rank: int, world: int) -> torch.Tensor:
    cu = [int(c) for c in cu_seqlens]
    return torch.cat([x[a:b].chunk(world, dim=0)[rank] for a, b in zip(cu[:-1], cu[1:])], dim=0).contiguous()


def shard_zigzag_varlen(x: torch.Tensor, cu_seqlens: Sequence[int], rank: int, world: int) -> torch.Tensor:
    cu = [int(c) for c in cu_seqlens]
    parts: List[torch.Tensor] = []
    for a, b in zip(cu[:-1], cu[1:]):
        ch = x[a:b].chunk(2 * world, dim=0)
        parts += [ch[rank], ch[2 * world - 1 - rank]]
    return torch.cat(parts, dim=0).contiguous()


def shard_llama3(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    return x.chunk(world, dim=0)[rank].contiguous()


def shard_zigzag_llama3(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Flat packed stream (T, ...): chunks ``rank`` and ``2W-1-rank`` of ``2W`` (zigzag_llama3_* entry points)."""
    return shard_zigzag(x, rank, world, dim=0)


def positions_zigzag_llama3(cu_seqlens: Sequence[int], rank: int, world: int, device=None) -> torch.Tensor:
    """Position inside its document of every local token of the zigzag-llama3 layout (what RoPE needs)."""
    cu = torch.as_tensor([int(c) for c in cu_seqlens], dtype=torch.long)
    total = int(cu[-1])
    flat = torch.arange(total)
    pos = flat - cu[torch.searchsorted(cu, flat, right=True) - 1]
    out = shard_zigzag_llama3(pos, rank, world)
    return out if device is None else out.to(device)


def positions(scheme: str, rank: int, world: int, seqlen_local: int, device=None) -> torch.Tensor:
    """Global position of every local token (what RoPE needs), batch layouts."""
    i = torch.arange(seqlen_local, device=device)
    if scheme == "ring":
        return rank * seqlen_local + i
    if scheme == "zigzag":
        c = seqlen_local // 2
        return torch.where(i < c, rank * c + i, (2 * world - 1 - rank) * c + (i - c))
    if scheme == "stripe":
        return i * world + rank
    raise ValueError(scheme)


def unshard(scheme: str, shards: Sequence[torch.Tensor], dim: int = 1) -> torch.Tensor:
    """Inverse of ``shard_<scheme>`` given every rank's shard (batch layouts)."""
    world = len(shards)
    if scheme == "ring":
        return torch.cat(list(shards), dim=dim)
    if scheme == "zigzag":
        halves = [s.chunk(2, dim=dim) for s in shards]
        front = [h[0] for h in halves]
        back = [h[1] for h in reversed(halves)]
        return torch.cat(front + back, dim=dim)
    if scheme == "stripe":
        stacked = torch.stack(list(shards), dim=dim + 1)  # (..., S_l, W, ...)
        shape = list(shards[0].shape)
        shape[dim] *= world
        return stacked.reshape(shape)
    raise ValueError(scheme)
