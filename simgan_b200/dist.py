"""Minibatch data parallelism for the PPO and discriminator updates.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).
Parameters, Adam state and the rollout / expert buffers are replicated; every rank derives the SAME
index permutation from its identically-seeded CPU generator, processes the contiguous 1/G slice
``[row_begin,row_end)`` of every minibatch, and the flat gradient (+ loss sums) is summed with ONE
allreduce per optimizer step between the reduce phase and the clip+Adam epilogue, which every rank then
runs identically (no parameter broadcast).
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_rows, rank, world):
    """Contiguous near-equal split of a minibatch's rows; the union over ranks is exactly [0,n_rows)."""
    base, rem = divmod(n_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class DataParallel(object):
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.n_allreduce = 0
        self._cb = None

    def shard(self, n_rows):
        b, e = shard_bounds(n_rows, self.rank, self.world)
        if n_rows < self.world:          # some rank would get an empty shard: refuse on EVERY rank
            raise ValueError("minibatch of %d rows cannot be split over %d ranks" % (n_rows, self.world))
        return b, e

    def allreduce_(self, tensor):
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        self.n_allreduce += 1

    def make_callback(self, workspace):
        """C callback handed to sg_*_update (mode 1): wraps the gradient region of ``workspace`` (a torch
        uint8 CUDA tensor) as an fp32 view and enqueues the sum-allreduce on the current stream."""
        base = workspace.data_ptr()
        nbytes = workspace.numel()

        def _cb(ptr, n_floats, user):
            try:
                off = int(ptr) - base
                if off < 0 or off + 4 * n_floats > nbytes or off % 4:
                    return 2
                view = workspace[off:off + 4 * n_floats].view(torch.float32)
                self.allreduce_(view)
                return 0
            except Exception:      # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        self._cb = _lib.ALLREDUCE_FN(_cb)     # keep alive for the duration of the call
        return self._cb


def attach(ppo=None, disc=None, group=None):
    """Enable data-parallel updates on a PPO and/or Discriminator object when world_size > 1."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    dp = DataParallel(group)
    if ppo is not None:
        ppo.dp = dp
    if disc is not None:
        disc.dp = dp
    return dp
