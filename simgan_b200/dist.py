"""Minibatch data parallelism for the PPO and discriminator updates.

One process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).
Parameters, Adam state and the rollout / expert buffers are replicated; every rank derives the SAME
index permutation from its identically-seeded CPU generator, processes the contiguous 1/G slice
``[row_begin,row_end)`` of every minibatch, and the flat gradient (+ loss sums) is summed with ONE
allreduce per optimizer step between the reduce phase and the clip+Adam epilogue, which every rank then
runs identically (no parameter broadcast).

Two transports for that exchange:

* ``p2p`` (default on CUDA): the persistent step kernel itself pushes its locally reduced gradient slices into
  the peers' exchange buffers over NVLink (CUDA IPC peer pointers), publishes them with system-scope flags and
  adds the world's slices in rank order -- no collective call and no extra launch per step
  (csrc/sg_dp.cuh).  Needs shards of equal size on every rank.
* ``nccl``: one launch per phase and a ``torch.distributed.all_reduce`` of the flat gradient between the reduce
  and the clip+Adam phase (C callback ``sg_allreduce_fn``); also what the gloo CPU tests exercise.
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
    """``policy``: "always" shards every update; "auto" shards an update only when that removes work from its critical
    path -- a minibatch whose tiles all run concurrently on ONE GPU (tiles <= SMs: the 1024-row PPO minibatches and the
    128-triple discriminator batches of BASELINE configs[0-2]) gains nothing from being split and would only pay the
    per-step exchange, so under "auto" such an update is computed redundantly by every rank (identical seeds, deterministic
    kernels: the replicas stay bit-identical) and no gradient is exchanged."""

    def __init__(self, group=None, transport=None, policy="always"):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if transport is None:
            transport = "p2p" if (torch.cuda.is_available() and dist.get_backend(group) == "nccl") else "nccl"
        assert transport in ("p2p", "nccl")
        self.transport = transport
        assert policy in ("always", "auto")
        self.policy = policy
        self.n_allreduce = 0
        self._cb = None
        self._ctx = {}           # key -> sg_dp context (one per optimizer)

    def __getstate__(self):
        raise TypeError("simgan_b200.dist.DataParallel holds per-process handles (CUDA IPC mappings, a process group) and "
                        "cannot be pickled; PPO / Discriminator drop it from their own pickles")

    # ---- fused peer-memory exchange -----------------------------------------------------------------------
    def p2p_ok(self, n_rows):
        """The in-kernel exchange needs identical shard sizes (identical grids / slice tables) on every rank."""
        return self.transport == "p2p" and n_rows % self.world == 0

    def context(self, key, n_floats):
        """Create (once) the exchange context of one optimizer: allocate this rank's buffer, swap the CUDA IPC
        handles with the peers and map their buffers.  Collective: every rank must call it in the same order."""
        ctx = self._ctx.get(key)
        lib = _lib.lib()
        if ctx is not None and lib.sg_dp_capacity(ctx) >= n_floats + 4:
            return ctx
        if ctx is not None:
            lib.sg_dp_destroy(ctx)
        ctx = C.c_void_p()
        _lib.check(lib.sg_dp_create(self.rank, self.world, int(n_floats) + 4, C.byref(ctx)), "sg_dp_create")
        mine = C.create_string_buffer(64)
        _lib.check(lib.sg_dp_local_handle(ctx, mine), "sg_dp_local_handle")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=self.group)
        blob = b"".join(handles)
        _lib.check(lib.sg_dp_open_peers(ctx, blob), "sg_dp_open_peers")
        dist.barrier(group=self.group)
        self._ctx[key] = ctx
        return ctx

    def sum_trace_(self, trace, n_cols):
        """Per-step loss columns of a p2p-mode trace are this rank's partial sums: add them over the ranks."""
        part = trace[:, :n_cols].contiguous()
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
        trace[:, :n_cols] = part
        return trace

    def shards(self, n_rows, rows_per_tile):
        """Whether an update over minibatches of ``n_rows`` rows (tiles of ``rows_per_tile``) is sharded under the policy."""
        if self.policy == "always":
            return True
        sms = int(_lib.lib().sg_device_sm_count()) or 148
        return n_rows > rows_per_tile * sms

    def close(self):
        lib = _lib.lib()
        for ctx in self._ctx.values():
            lib.sg_dp_destroy(ctx)
        self._ctx = {}

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


def attach(ppo=None, disc=None, group=None, transport=None, policy="always"):
    """Enable data-parallel updates on a PPO and/or Discriminator object when world_size > 1."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    dp = DataParallel(group, transport, policy)
    if ppo is not None:
        ppo.dp = dp
    if disc is not None:
        disc.dp = dp
    return dp
