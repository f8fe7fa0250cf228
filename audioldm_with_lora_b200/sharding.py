"""Prompt/seed sharding of the sampling path across GPUs (one process per GPU, no data-path collective).

The reference samples in a single process (/root/reference/app.py:14, script/inference/generate_audio.py:47-52);
BASELINE.json configs 3 and 5 shard prompts over the 8 GPUs of a box.  Units (prompts) are independent, so the
only communication is the optional gather of the finished waveforms on the host.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of `n_items` owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_prompts(prompt_embeds: torch.Tensor, negative_prompt_embeds: Optional[torch.Tensor], rank: int, world: int):
    """This rank's slice of the prompt batch and the GLOBAL prompt indices it covers (the indices seed the initial
    latents, so a prompt's result does not depend on how many GPUs share the batch)."""
    s, e = shard_range(prompt_embeds.shape[0], rank, world)
    neg = None if negative_prompt_embeds is None else negative_prompt_embeds[s:e]
    return prompt_embeds[s:e], neg, list(range(s, e))


def gather_to_rank0(local: torch.Tensor, counts: List[int], group=None) -> Optional[torch.Tensor]:
    """Host-side gather of per-rank result rows (ragged by `counts`) to rank 0; other ranks get None."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    width = local.shape[1:]
    pad = max(counts)
    buf = local.new_zeros((pad,) + tuple(width))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
