"""Rollout buffer with the reference's surface (third_party/a2c_ppo_acktr/storage.py:32-192).

Ten dense fp32 time-major tensors that callers index and assign directly
(main_gail_dyn_ppo.py:213-215, 276-292) -- the attributes ARE the interface.  Once ``.to(cuda)`` has
been called, ``insert`` / ``after_update`` are one fused block-copy launch, ``compute_returns`` is the
bit-exact return/GAE scan kernel and ``feed_forward_generator`` gathers all ten row-sets of a
minibatch in one launch, with indices drawn from the CPU default generator exactly as the
reference's BatchSampler(SubsetRandomSampler) does.
"""
import ctypes as C

import torch

from . import _lib

_T1_FIELDS = ("obs", "obs_feat", "recurrent_hidden_states", "masks", "bad_masks")
_ALL_FIELDS = ("obs", "obs_feat", "recurrent_hidden_states", "rewards", "value_preds", "returns",
               "action_log_probs", "actions", "masks", "bad_masks")


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.SgError("%s needs the rollout buffer on a CUDA device (call .to(device)); there is no CPU "
                           "fallback for the hot path" % what)


class RolloutStorage(object):
    def __init__(self, num_steps, num_processes, obs_shape, action_space, recurrent_hidden_state_size, feat_len=0):
        if action_space.__class__.__name__ == "Discrete":
            raise NotImplementedError("discrete action spaces are outside the hot path (all *CombinedEnv are Box)")
        T, N = num_steps, num_processes
        self.obs = torch.zeros(T + 1, N, *obs_shape)
        self.obs_feat = torch.zeros(T + 1, N, feat_len)
        self.recurrent_hidden_states = torch.zeros(T + 1, N, recurrent_hidden_state_size)
        self.rewards = torch.zeros(T, N, 1)
        self.value_preds = torch.zeros(T + 1, N, 1)
        self.returns = torch.zeros(T + 1, N, 1)
        self.action_log_probs = torch.zeros(T, N, 1)
        self.actions = torch.zeros(T, N, action_space.shape[0])
        self.masks = torch.ones(T + 1, N, 1)
        self.bad_masks = torch.ones(T + 1, N, 1)
        self.num_steps = num_steps
        self.step = 0

    def to(self, device):
        for k in _ALL_FIELDS:
            setattr(self, k, getattr(self, k).to(device))

    # ---- fused block copies ----------------------------------------------------------------------------
    @staticmethod
    def _copy_blocks(pairs):
        """pairs: list of (dst, src) same-numel fp32 tensors.  One launch on CUDA."""
        if not pairs:
            return
        if not pairs[0][0].is_cuda:
            for dst, src in pairs:          # host staging buffer before .to(device): plain data movement
                dst.copy_(src)
            return
        keep = []
        n = len(pairs)
        srcs, dsts, cnts = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int * n)()
        for i, (dst, src) in enumerate(pairs):
            s = src if torch.is_tensor(src) else torch.as_tensor(src)
            s = s.detach().to(device=dst.device, dtype=torch.float32)
            if s.numel() != dst.numel():
                s = s.expand_as(dst)
            s = s.contiguous()
            keep.append(s)
            assert dst.is_contiguous()
            srcs[i], dsts[i], cnts[i] = s.data_ptr(), dst.data_ptr(), dst.numel()
        rc = _lib.lib().sg_copy_blocks(srcs, dsts, cnts, n, _lib.current_stream())
        _lib.check(rc, "sg_copy_blocks")

    @_lib.on_device(lambda self, *a, **k: self.obs.device)
    def insert(self, obs, recurrent_hidden_states, actions, action_log_probs, value_preds, rewards, masks, bad_masks,
               obs_feat=None):
        """Slot step+1 for obs/feat/hxs/masks/bad_masks, slot step for the rest (storage.py:70-84)."""
        s = self.step
        pairs = [(self.obs[s + 1], obs)]
        if obs_feat is not None:
            pairs.append((self.obs_feat[s + 1], obs_feat))
        pairs += [(self.recurrent_hidden_states[s + 1], recurrent_hidden_states), (self.actions[s], actions),
                  (self.action_log_probs[s], action_log_probs), (self.value_preds[s], value_preds),
                  (self.rewards[s], rewards), (self.masks[s + 1], masks), (self.bad_masks[s + 1], bad_masks)]
        self._copy_blocks(pairs)
        self.step = (self.step + 1) % self.num_steps

    def mod_reward(self, offset, reverse_l):
        """Add ``offset`` to the last ``reverse_l`` reward slots (storage.py:86-94; no callers upstream)."""
        n = self.rewards.size(1)
        t = self.step
        for _ in range(reverse_l):
            t = (t - 1) % self.num_steps
            self.rewards[t] += offset.view(n, 1)

    @_lib.on_device(lambda self, *a, **k: self.obs.device)
    def after_update(self):
        """Slot T -> slot 0 for the five (T+1)-long tensors (storage.py:96-101)."""
        self._copy_blocks([(getattr(self, k)[0], getattr(self, k)[-1]) for k in _T1_FIELDS])

    # ---- returns ---------------------------------------------------------------------------------------
    @_lib.on_device(lambda self, *a, **k: self.obs.device)
    def compute_returns(self, next_value, use_gae, gamma, gae_lambda, use_proper_time_limits=True):
        """All four branches of storage.py:103-142 in the kernel sg_compute_returns (bit-exact)."""
        _require_cuda(self.rewards, "compute_returns")
        T, N = self.rewards.shape[:2]
        nv = next_value.detach().to(device=self.rewards.device, dtype=torch.float32).contiguous()
        tok = _lib.timer.start("compute_returns")
        rc = _lib.lib().sg_compute_returns(_lib.ptr(self.rewards), _lib.ptr(self.value_preds), _lib.ptr(self.masks),
                                           _lib.ptr(self.bad_masks), _lib.ptr(self.returns), _lib.ptr(nv), T, N,
                                           float(gamma), float(gae_lambda), int(bool(use_gae)),
                                           int(bool(use_proper_time_limits)), _lib.current_stream())
        _lib.timer.stop(tok)
        _lib.check(rc, "sg_compute_returns")

    # ---- minibatch sampler -----------------------------------------------------------------------------
    def flat_rows(self):
        """(S, D) row views addressed by flat sample id t*N+n (storage.py:169-181)."""
        def fl(x):
            return x.reshape(-1, x.shape[-1])
        return dict(obs=fl(self.obs[:-1]), obs_feat=fl(self.obs_feat[:-1]), next_obs_feat=fl(self.obs_feat[1:]),
                    hxs=fl(self.recurrent_hidden_states[:-1]), actions=fl(self.actions),
                    value_preds=fl(self.value_preds[:-1]), returns=fl(self.returns[:-1]), masks=fl(self.masks[:-1]),
                    action_log_probs=fl(self.action_log_probs))

    @staticmethod
    def sampler_permutation(batch_size):
        """The index stream of BatchSampler(SubsetRandomSampler(range(S)), mb, drop_last=True): one
        torch.randperm(S) on the CPU default generator per pass (storage.py:158-162)."""
        return torch.randperm(batch_size)

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None):
        _require_cuda(self.rewards, "feed_forward_generator")
        T, N = self.rewards.shape[:2]
        S = T * N
        if mini_batch_size is None:
            assert S >= num_mini_batch, (
                "PPO requires the number of processes ({}) * number of steps ({}) = {} to be greater than or equal "
                "to the number of PPO mini batches ({}).".format(N, T, S, num_mini_batch))
            mini_batch_size = S // num_mini_batch
        perm = self.sampler_permutation(S)
        fr = self.flat_rows()
        srcs = [fr["obs"], fr["hxs"], fr["actions"], fr["value_preds"], fr["returns"], fr["masks"],
                fr["action_log_probs"]]
        if advantages is not None:
            srcs.append(advantages.reshape(-1, 1))
        srcs += [fr["obs_feat"], fr["next_obs_feat"]]
        dev = self.rewards.device
        for i in range(S // mini_batch_size):
            idx = perm[i * mini_batch_size:(i + 1) * mini_batch_size].to(dev, non_blocking=False)
            outs = gather_rows(srcs, idx)
            if advantages is None:
                outs.insert(7, None)
            yield tuple(outs)

    def recurrent_generator(self, advantages, num_mini_batch):
        raise NotImplementedError("recurrent policies are outside the PPO+GAIL hot path (SURVEY.md section 2, row 1)")


@_lib.on_device(lambda srcs, idx: srcs[0].device)
def gather_rows(srcs, idx):
    """dst[i] = srcs[i][idx] for several (S, D_i) fp32 CUDA tensors in one launch (sg_gather_rows)."""
    n = len(srcs)
    rows = int(idx.numel())
    assert idx.dtype == torch.int64 and idx.is_cuda and idx.is_contiguous()
    outs, keep = [], []
    sp, dp, dims = (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_int * n)()
    for i, s in enumerate(srcs):
        s = s if s.is_contiguous() else s.contiguous()
        keep.append(s)
        d = s.shape[1]
        if d == 0:       # feat_len == 0: nothing to gather, keep the (mb, 0) shape
            outs.append(s.new_empty(rows, 0))
            sp[i], dp[i], dims[i] = s.data_ptr() or 1, 1, 1
            continue
        o = torch.empty(rows, d, device=s.device, dtype=torch.float32)
        outs.append(o)
        sp[i], dp[i], dims[i] = s.data_ptr(), o.data_ptr(), d
    live = [i for i, s in enumerate(srcs) if s.shape[1] > 0]
    if len(live) != n:
        sp2, dp2, dims2 = (C.c_void_p * len(live))(), (C.c_void_p * len(live))(), (C.c_int * len(live))()
        for j, i in enumerate(live):
            sp2[j], dp2[j], dims2[j] = sp[i], dp[i], dims[i]
        sp, dp, dims, n = sp2, dp2, dims2, len(live)
    if n:
        rc = _lib.lib().sg_gather_rows(sp, dp, dims, n, _lib.ptr(idx), rows, _lib.current_stream())
        _lib.check(rc, "sg_gather_rows")
    return outs
