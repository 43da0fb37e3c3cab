"""torch.randperm streams of the CPU default generator, produced by ``sg_host_randperm_*`` (csrc/sg_host.cu).

``feed_forward_generator`` draws one ``torch.randperm(T*N)`` per PPO epoch on the CPU default generator
(third_party/a2c_ppo_acktr/storage.py:158-162).  For rollouts of millions of samples that serial walk is longer than the
epoch's kernel; ``PermutationStream`` hands the generator's mt19937 state to the C side, which produces the SAME
permutations on helper threads (engine on one, the Fisher-Yates walks of different epochs in parallel) while the caller
launches kernels, and stores the engine state back afterwards -- the generator ends exactly where the reference's draws
would have left it.

The byte layout of ``torch.get_rng_state()`` (legacy ``THGeneratorState``: seed u64, left i32, seeded i32, next u64,
state u64[624], normal cache) is checked by a self-test against ``torch.randperm`` on first use; if it ever differs the
stream falls back to calling ``torch.randperm`` (same results, just slower).
"""
import ctypes as C
import os
import struct

import numpy as np
import torch

from . import _lib

MIN_ELEMENTS = 1 << 17          # below this torch.randperm is a couple of ms and hides behind the running kernel
_STATE_BYTES = 5056
_KEY_OFF = 24
_usable = None


def _threads():
    env = os.environ.get("SIMGAN_HOST_THREADS")
    if env:
        return max(1, int(env))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    return max(1, min(4, (os.cpu_count() or 4) // max(1, local_world) - 1))


def _unpack(state):
    b = state.numpy().tobytes()
    _, left, seeded, nxt = struct.unpack_from("<QiiQ", b, 0)
    key = np.frombuffer(b, dtype=np.uint64, count=624, offset=_KEY_OFF).astype(np.uint32)
    pos = 624 if left == 1 else int(nxt)          # left counts down to the regeneration; left + next == 625 in between
    return key, pos, seeded


def _pack(state, key, pos):
    b = bytearray(state.numpy().tobytes())
    struct.pack_into("<i", b, 8, 625 - pos)
    struct.pack_into("<Q", b, 16, pos)
    b[_KEY_OFF:_KEY_OFF + 624 * 8] = key.astype(np.uint64).tobytes()
    return torch.frombuffer(b, dtype=torch.uint8).clone()


class PermutationStream(object):
    """``n_perms`` consecutive ``torch.randperm(n)`` results written into ``out`` (int32, (n_perms, n), host)."""

    def __init__(self, n, n_perms, out, owned=None):
        """``owned``: optional sequence of n_perms bools -- only those permutations are built into ``out`` (the engine still
        advances past the others: the ranks of one node split the walks and exchange the results); C stream only."""
        assert out.dtype == torch.int32 and out.is_contiguous() and tuple(out.shape) == (n_perms, n) and not out.is_cuda
        self.n, self.n_perms, self.out = n, n_perms, out
        mask = (1 << 64) - 1
        if owned is not None:
            assert len(owned) == n_perms <= 64 and n >= MIN_ELEMENTS and usable()
            mask = sum(1 << e for e, o in enumerate(owned) if o)
        self._h = None
        self._state = None
        self._done = 0
        if n_perms > 0 and n >= MIN_ELEMENTS and usable():
            self._state = torch.get_rng_state()
            key, pos, _ = _unpack(self._state)
            threads = _threads()
            if owned is not None:
                # a rank that owns two walks of the update runs them side by side: the second one is needed only a few
                # kernels after the first (the walks are memory-latency bound, so this pays even on a busy node)
                threads = max(threads, min(2, sum(1 for o in owned if o)))
            h = _lib.lib().sg_host_randperm_begin(key.ctypes.data, pos, n, n_perms, out.data_ptr(), threads, mask)
            if not h:
                _lib.check(1, "sg_host_randperm_begin")
            self._h = h

    def wait(self, e):
        """Block until permutation ``e`` is in ``out[e]`` (permutations must be taken in order)."""
        if self._h is not None:
            _lib.check(_lib.lib().sg_host_randperm_wait(self._h, e), "sg_host_randperm_wait")
        else:
            while self._done <= e:
                self.out[self._done].copy_(torch.randperm(self.n))
                self._done += 1
        return self.out[e]

    def finish(self):
        """Join the helper threads and leave the CPU default generator where the reference's draws would have."""
        if self._h is not None:
            key = np.empty(624, dtype=np.uint32)
            pos = C.c_int(0)
            h, self._h = self._h, None
            _lib.check(_lib.lib().sg_host_randperm_end(h, key.ctypes.data, C.byref(pos)), "sg_host_randperm_end")
            torch.set_rng_state(_pack(self._state, key, pos.value))
        else:
            self.wait(self.n_perms - 1) if self.n_perms else None

    def __del__(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.lib().sg_host_randperm_end(self._h, None, None)
            except Exception:
                pass


def randperm_i32(n):
    """``torch.randperm(n)`` (same values, same generator consumption) as an int32 tensor."""
    if n < MIN_ELEMENTS or not usable():
        return torch.randperm(n).to(torch.int32)
    out = torch.empty(1, n, dtype=torch.int32)
    ps = PermutationStream(n, 1, out)
    ps.wait(0)
    ps.finish()
    return out[0]


def randperm_prefix(n, m):
    """``torch.randperm(n)[:m]`` as int64 with the generator advanced by the whole draw: what ``zip(expert_loader,
    feed_forward_generator)`` consumes of the rollout sampler's permutation when the expert loader is the shorter one
    (third_party/a2c_ppo_acktr/algo/gail.py:159-166)."""
    if n < MIN_ELEMENTS or not usable():
        return torch.randperm(n)[:m]
    state = torch.get_rng_state()
    key, pos, _ = _unpack(state)
    out = torch.empty(max(m, 1), dtype=torch.int32)
    k2 = np.empty(624, dtype=np.uint32)
    p2 = C.c_int(0)
    _lib.check(_lib.lib().sg_host_randperm_prefix(key.ctypes.data, pos, n, m, out.data_ptr(), k2.ctypes.data, C.byref(p2)),
               "sg_host_randperm_prefix")
    torch.set_rng_state(_pack(state, k2, p2.value))
    return out[:m].long()


def usable():
    """One-time self-test: the C stream equals torch.randperm (values and generator state) on this torch build."""
    global _usable
    if _usable is None:
        _usable = False
        saved = torch.get_rng_state()
        try:
            if saved.numel() == _STATE_BYTES:
                ok = True
                for n in (1, 2, 1000, 4099):
                    torch.set_rng_state(saved)
                    torch.randperm(7)                    # move off the freshly seeded position
                    start = torch.get_rng_state()
                    want = [torch.randperm(n) for _ in range(3)]
                    after = torch.get_rng_state()
                    torch.set_rng_state(start)
                    key, pos, seeded = _unpack(start)
                    out = torch.empty(3, n, dtype=torch.int32)
                    h = _lib.lib().sg_host_randperm_begin(key.ctypes.data, pos, n, 3, out.data_ptr(), 2, (1 << 64) - 1)
                    if not h or not seeded:
                        ok = False
                        break
                    k2 = np.empty(624, dtype=np.uint32)
                    p2 = C.c_int(0)
                    _lib.lib().sg_host_randperm_end(h, k2.ctypes.data, C.byref(p2))
                    got_state = _pack(start, k2, p2.value)
                    ok = ok and all(torch.equal(out[i].long(), want[i]) for i in range(3))
                    if n > 1:
                        ok = ok and torch.equal(got_state, after)
                _usable = ok
        finally:
            torch.set_rng_state(saved)
    return _usable
