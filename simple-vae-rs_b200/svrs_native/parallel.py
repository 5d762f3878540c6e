"""Data-parallel rules of the fused step (SURVEY 8.4 row e), kept as small pure functions so they can be tested on CPU.

The reference's loss is NOT batch-size invariant: the NLL terms are SUMS over all elements while the KL terms are MEANS
over the batch (loss/cond_vae_loss.py:43-57).  For R ranks holding B_local samples each, the single-process
loss on the global batch is   sum_r [ mse_x,r + mse_y,r ] + (1/R) sum_r [ kld_u,r + kld_z,r ],   so each rank
back-propagates upstream gradients (1, 1/R, 1, 1/R) for (mse_x, kld_u, mse_y, kld_z) and the gradients are
all-reduced with SUM (not averaged, as stock DDP would do)."""
from __future__ import annotations

from typing import List, Tuple


def upstream_grad_scales(world: int) -> List[float]:
    """Upstream gradients of (mse_x, kld_u, mse_y, kld_z) on every rank before the SUM all-reduce."""
    return [1.0, 1.0 / world, 1.0, 1.0 / world]


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, as-even-as-possible split of n samples."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_offset(n: int, rank: int, world: int) -> int:
    """Global index of the rank's first sample: Philox eps is keyed by global sample index, so the noise a sample sees
    does not depend on how the batch is partitioned."""
    return shard_bounds(n, rank, world)[0]


def global_terms(local_terms, world: int):
    """Combine per-rank (mse_x, kld_u, mse_y, kld_z) SUMMED over ranks into the global-batch terms."""
    return [local_terms[0], local_terms[1] / world, local_terms[2], local_terms[3] / world]
