"""Multi-GPU plumbing for the search: roots are independent units (trees never interact,
cnode.cpp:783,495; the network is row-independent in eval mode), so rank r owns a contiguous slice
of the global root batch and there is NO collective on the data path.  After a search the collector
may want the global batch: one all-gather of {visits int32[B,A], values f32[B]} (82 KB per rank at
B=1024, A=18) over NCCL/NVLink (gloo on CPU for tests)."""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_roots: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split [lo, hi) of the global root batch; remainders go to the lowest ranks."""
    base, rem = divmod(global_roots, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_search_results(visits: torch.Tensor, values: torch.Tensor, global_roots: int):
    """All-gathers per-rank (visits [b,A] int32, values [b] f32) into global tensors on every rank.
    Ranks may hold unequal slices (shard_range); slices are padded to the largest for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return visits, values
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(global_roots, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    A = visits.shape[1]
    pv = torch.full((bmax, A), -1, dtype=torch.int32, device=visits.device)
    pf = torch.zeros(bmax, dtype=torch.float32, device=values.device)
    pv[:visits.shape[0]] = visits
    pf[:values.shape[0]] = values
    gv = torch.empty(world * bmax, A, dtype=torch.int32, device=visits.device)
    gf = torch.empty(world * bmax, dtype=torch.float32, device=values.device)
    dist.all_gather_into_tensor(gv, pv)
    dist.all_gather_into_tensor(gf, pf)
    outv = torch.cat([gv[r * bmax:r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
    outf = torch.cat([gf[r * bmax:r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])
    return outv, outf
