"""Host-side logic of the batch-sharded training loop, free of
any device dependency so that it is testable on CPU with world_size-2 gloo:
which samples a rank owns, how the gradient leaves of the TKProduct
(MnistCnnShaped2.hs:30-40) map into the one flat bucket that is all-reduced,
and what the exchange computes (sum over ranks, then 1/world: the mean of
per-shard mean losses/gradients equals the global mean for equal shards,
MnistCnnShaped2.hs:136).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def shard_range(rank: int, world: int, batch: int) -> tuple[int, int]:
    """Rank r owns samples [r*B/G, (r+1)*B/G); the batch must split evenly
    (the loss is a mean over the shard, so unequal shards would bias it)."""
    if batch % world != 0:
        raise ValueError(f"batch {batch} does not split evenly over {world} ranks")
    per = batch // world
    return rank * per, (rank + 1) * per


class BucketLayout:
    """Offsets of the gradient leaves (in parameter order) inside the flat
    bucket; one extra trailing slot carries the loss."""

    def __init__(self, shapes: Sequence[tuple]):
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(np.prod(s)) if len(s) else 1 for s in self.shapes]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.n_params = self.offsets[-1]
        self.n_total = self.n_params + 1

    def flatten(self, leaves: Sequence[np.ndarray], loss: float) -> np.ndarray:
        assert len(leaves) == len(self.shapes)
        out = np.empty(self.n_total, dtype=np.asarray(leaves[0]).dtype)
        for leaf, off, n, sh in zip(leaves, self.offsets, self.sizes, self.shapes):
            leaf = np.asarray(leaf)
            assert tuple(leaf.shape) == sh, (leaf.shape, sh)
            out[off:off + n] = leaf.reshape(-1)
        out[self.n_params] = loss
        return out

    def unflatten(self, bucket: np.ndarray):
        leaves = [bucket[o:o + n].reshape(sh) for o, n, sh in zip(self.offsets, self.sizes, self.shapes)]
        return leaves, float(bucket[self.n_params])


def exchange_scale(world: int) -> float:
    """Factor applied after the all-reduce SUM."""
    return 1.0 / world
