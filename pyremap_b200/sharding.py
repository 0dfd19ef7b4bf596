"""Multi-GPU driver: K-sharding of the weight application, one process per GPU.

Every output column depends only on the same input column and the (small,
replicated) weights, so the path shards embarrassingly along K = levels x times
(SURVEY.md section 8e).  Whole leading-axis slices (time steps) are dealt to the
ranks in contiguous blocks; each rank runs the same fused launch on its block.
There is NO collective in the data path.  Two small optional exchanges exist:

* the branch decision.  The reference chooses between its masked and its ``frac_b``
  branch once per *variable*: masked iff a threshold is given and the variable holds a
  NaN anywhere (``/root/reference/pyremap/remapper/remap_numpy.py:202-204,258-261``).
  A rank only sees its block, so every rank scans its block and the one-word flags are
  combined with an ``all_reduce(MAX)`` -- otherwise the result would depend on the world
  size whenever only some shards contain NaNs;
* ``gather``: for callers who want the full result on every rank, an ``all_gather`` of
  the outputs along the slice axis (NCCL over NVLink on the GPUs; ``gloo`` in the CPU
  tests).
"""

from __future__ import annotations

import inspect


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


def block_has_nan(block):
    """True iff this rank's block (numpy array, CPU or CUDA tensor) holds a NaN."""
    import numpy as np
    import torch
    if isinstance(block, torch.Tensor):
        if block.numel() == 0 or not block.is_floating_point():
            return False
        if block.is_cuda:
            from .engine import device_any_nan
            return device_any_nan(block)
        block = block.numpy()
    block = np.asarray(block)
    if block.size == 0 or block.dtype.kind != 'f':
        return False
    if block.dtype in (np.float32, np.float64) and block.flags.c_contiguous \
            and block.dtype.isnative:
        from . import _cabi
        try:
            return _cabi.host_any_nan(block)
        except _cabi.B200RemapError:      # library not loadable on an inspection-only host
            pass
    return bool(np.isnan(block).any())


class ShardedRemap:
    """Remap ``[T, ...]`` fields with the leading axis split over the ranks of a
    ``torch.distributed`` process group (weights replicated per rank).

    ``compute(local_field, masked) -> local_out`` is the per-rank work, where ``masked``
    is the variable-wide branch decision (``True``: masked renormalising branch,
    ``False``: ``frac_b`` branch).  By default it is
    ``remapper.remap_array(local_field, remap_axes, threshold, mode=..., return_torch=True)``.
    A one-argument ``compute(local_field)`` is accepted too (it then decides for itself).
    """

    def __init__(self, remapper=None, remap_axes=None, renormalization_threshold=None,
                 group=None, compute=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.threshold = renormalization_threshold
        self.remapper = remapper
        if compute is None:
            if remapper is None or remap_axes is None:
                raise ValueError('need a remapper and remap_axes, or a compute callable')
            if 0 in [int(a) for a in remap_axes]:
                raise ValueError('axis 0 is the sharded slice axis and cannot be remapped')

            def compute(local, masked):
                return remapper.remap_array(local, remap_axes, renormalization_threshold,
                                            return_torch=True,
                                            mode='masked' if masked else 'fracb')
            self._takes_branch = True
        else:
            try:
                params = [p for p in inspect.signature(compute).parameters.values()
                          if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
                self._takes_branch = len(params) >= 2
            except (TypeError, ValueError):
                self._takes_branch = False
        self.compute = compute

    def local_slices(self, n_slices):
        return shard_bounds(n_slices, self.world, self.rank)

    def _flag_device(self):
        import torch
        backend = self.dist.get_backend(self.group) if self.dist.is_initialized() else 'gloo'
        if 'nccl' in str(backend):
            return torch.device('cuda', torch.cuda.current_device())
        return torch.device('cpu')

    def decide_masked(self, local_block):
        """The reference's per-variable branch test over ALL ranks' blocks: masked iff a
        threshold is given and any block holds a NaN (one ``all_reduce(MAX)`` of a flag)."""
        if self.threshold is None:
            return False
        flag = block_has_nan(local_block)
        if self.world > 1:
            import torch
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self._flag_device())
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
            flag = bool(int(t.item()))
        return flag

    def _run(self, block):
        if not self._takes_branch:
            return self.compute(block)
        return self.compute(block, self.decide_masked(block))

    def remap_local(self, field):
        """Remap this rank's block of ``field`` (the full ``[T, ...]`` array or a
        callable ``(lo, hi) -> block`` that loads only what this rank needs)."""
        if callable(field):
            raise TypeError('pass n_slices with a loader: use remap_local_from(loader, n_slices)')
        lo, hi = self.local_slices(field.shape[0])
        return self._run(field[lo:hi])

    def remap_local_from(self, loader, n_slices):
        lo, hi = self.local_slices(n_slices)
        return self._run(loader(lo, hi))

    def sweep(self, loader, n_slices, chunk=8, sink=None, masked=None):
        """Remap this rank's block ``[lo, hi)`` of an ``n_slices`` variable chunk by chunk (a
        year of daily slices does not fit one GPU at once): ``loader(a, b)`` returns slices
        ``[a, b)``, ``sink(a, b, out)`` consumes each result (default: results are
        concatenated and returned).  The branch is decided once for the whole variable before
        the first chunk is remapped: pass ``masked`` when it is known, otherwise every rank
        scans its block through ``loader`` first and the flags are combined over the ranks."""
        import torch
        lo, hi = self.local_slices(n_slices)
        spans = [(a, min(hi, a + chunk)) for a in range(lo, hi, chunk)]
        if self._takes_branch and masked is None:
            flag = False
            if self.threshold is not None:
                for a, b in spans:
                    if block_has_nan(loader(a, b)):
                        flag = True
                        break
                if self.world > 1:
                    t = torch.tensor([1 if flag else 0], dtype=torch.int32,
                                     device=self._flag_device())
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
                    flag = bool(int(t.item()))
            masked = flag
        outs = []
        for a, b in spans:
            block = loader(a, b)
            out = self.compute(block, masked) if self._takes_branch else self.compute(block)
            if sink is not None:
                sink(a, b, out)
            else:
                outs.append(out)
        if sink is not None:
            return None
        return torch.cat(outs, dim=0) if outs else None

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
