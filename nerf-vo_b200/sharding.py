"""Ray-batch sharding for the data-parallel mapping step.

Rays are independent given the parameters, so a global batch of B rays is cut into contiguous shards, rank r taking rays
[r*B/W, (r+1)*B/W); every parameter (hash tables, MLPs, appearance embedding) is replicated.  The one exchange step is a
sum-all-reduce of the flat gradient buffer before the optimizer; every loss of the step is a MEAN over the local rays
(NS/model_components/losses.py:127,152,246; nn.MSELoss), so dividing the summed gradient by W reproduces the
single-process gradient at the same global batch — DistributedDataParallel's semantics (NS/pipelines/base_pipeline.py:281-283),
which the reference would use if NeRF-VO enabled it (it runs world_size=1, nerf_vo/mapping/nerfstudio.py:108).

Host-side only (no kernels): usable, and tested, on CPU tensors with the gloo backend."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Half-open ray range of `rank`.  B must divide evenly: the step's kernels are shape-specialised per rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n_global % world != 0:
        raise ValueError(f"global batch of {n_global} rays does not divide over {world} ranks")
    per = n_global // world
    return rank * per, (rank + 1) * per


def shard_batch(rays: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor], jitters: Optional[Sequence[torch.Tensor]], rank: int, world: int):
    """Slices every per-ray tensor of a global batch to this rank's shard (leading dimension = rays)."""
    n = next(iter(rays.values())).shape[0]
    a, b = shard_range(n, rank, world)
    cut = lambda d: {k: v[a:b] for k, v in d.items()}
    return cut(rays), cut(targets), None if jitters is None else [j[a:b] for j in jitters]


def allreduce_gradient_(flat_grad: torch.Tensor, group=None) -> float:
    """In-place SUM all-reduce of the flat gradient; returns the factor (1/W) the optimizer applies to obtain the mean.
    (The division is folded into the fused Adam kernel instead of a separate pass over the 70 MB buffer.)"""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def proposal_update_due(step: int, steps_since_update: int, warmup: int = 5000, update_every: int = 5) -> bool:
    """ProposalNetworkSampler's `updated` predicate (NS/model_components/ray_samplers.py:591 with the schedule of
    NS/models/nerfacto.py:202-207).  It depends on step counters only, hence is identical on every rank — required, because a
    rank that skipped the proposal backward would contribute zeros to an all-reduce the others fill."""
    import numpy as np

    sched = float(np.clip(np.interp(step, [0, warmup], [0, update_every]), 1, update_every))
    return steps_since_update > sched or step < 10


# ---- evaluation frames (config 5): image rows are independent units -----------------------------------------------------


def row_shard(height: int, rank: int, world: int) -> Tuple[int, int]:
    """Half-open row range of `rank` for an image of `height` rows: ceil(H/W) rows per rank, the last ranks may be short / empty
    (a frame height need not divide: 680 rows over 8 ranks = 85 each, 360 over 7 = 52 x 6 + 48)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    per = (height + world - 1) // world
    lo = min(rank * per, height)
    return lo, min(lo + per, height)


def gather_rows(local_rows: torch.Tensor, height: int, world: int, group=None) -> torch.Tensor:
    """All-gather of the per-rank row blocks [h_r, W, ...] into the full [H, W, ...] frame (every rank gets it).  Blocks are padded to
    ceil(H/world) rows so one fixed-size all_gather_into_tensor serves ragged shards."""
    per = (height + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    out = torch.empty((world * per,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:height]
