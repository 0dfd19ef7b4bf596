"""Multi-GPU driver: K-sharding of the weight application, one process per GPU.

Every output column depends only on the same input column and the (small,
replicated) weights, so the path shards embarrassingly along K = levels x times
(SURVEY.md section 8e).  Whole leading-axis slices (time steps) are dealt to the
ranks in contiguous blocks; each rank runs the same fused launch on its block.
There is NO collective in the data path.  ``gather`` is the optional epilogue for
callers who want the full result on every rank: an ``all_gather`` of the outputs
along the slice axis (NCCL over NVLink on the GPUs; ``gloo`` in the CPU tests).
"""

from __future__ import annotations


def shard_bounds(n_items, world_size, rank):
    """Contiguous balanced block ``[lo, hi)`` of ``n_items`` for ``rank``: the
    first ``n_items % world_size`` ranks get one extra item."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f'bad rank {rank} for world size {world_size}')
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_counts(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0]
            for r in range(world_size)]


class ShardedRemap:
    """Remap ``[T, ...]`` fields with the leading axis split over the ranks of a
    ``torch.distributed`` process group (weights replicated per rank).

    ``compute(local_field) -> local_out`` is the per-rank work; by default it is
    ``remapper.remap_array(local_field, remap_axes, threshold, return_torch=True)``.
    """

    def __init__(self, remapper=None, remap_axes=None, renormalization_threshold=None,
                 group=None, compute=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if compute is None:
            if remapper is None or remap_axes is None:
                raise ValueError('need a remapper and remap_axes, or a compute callable')
            if 0 in [int(a) for a in remap_axes]:
                raise ValueError('axis 0 is the sharded slice axis and cannot be remapped')

            def compute(local):
                return remapper.remap_array(local, remap_axes, renormalization_threshold,
                                            return_torch=True)
        self.compute = compute

    def local_slices(self, n_slices):
        return shard_bounds(n_slices, self.world, self.rank)

    def remap_local(self, field):
        """Remap this rank's block of ``field`` (the full ``[T, ...]`` array or a
        callable ``(lo, hi) -> block`` that loads only what this rank needs)."""
        if callable(field):
            raise TypeError('pass n_slices with a loader: use remap_local_from(loader, n_slices)')
        lo, hi = self.local_slices(field.shape[0])
        return self.compute(field[lo:hi])

    def remap_local_from(self, loader, n_slices):
        lo, hi = self.local_slices(n_slices)
        return self.compute(loader(lo, hi))

    def gather(self, local_out, n_slices):
        """All ranks receive the full ``[T, ...]`` output (optional epilogue)."""
        import torch
        if self.world == 1:
            return local_out
        counts = shard_counts(n_slices, self.world)
        most = max(counts)
        tail = tuple(local_out.shape[1:])
        pad = torch.zeros((most,) + tail, dtype=local_out.dtype, device=local_out.device)
        pad[:local_out.shape[0]] = local_out
        full = torch.empty((self.world * most,) + tail, dtype=local_out.dtype,
                           device=local_out.device)
        self.dist.all_gather_into_tensor(full, pad, group=self.group)
        parts = [full[r * most:r * most + counts[r]] for r in range(self.world)]
        return torch.cat(parts, dim=0)
