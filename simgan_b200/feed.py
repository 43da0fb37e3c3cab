"""Rollout feed boundary: host vec-env  <->  device rollout buffer, one copy each way and ONE kernel per env step.

The reference's collection loop (third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py:209-236) issues per step a policy
forward (6 GEMMs + sampling), ``actions.cpu().numpy()`` (third_party/a2c_ppo_acktr/envs.py:199-205), three
``torch.from_numpy(...).to(device)`` / ``Tensor(...)`` uploads (envs.py:207-210, main_gail_dyn_ppo.py:230-236) and nine
``copy_`` calls (storage.py:70-84).  ``RolloutFeeder`` keeps the same data flow -- PyBullet envs stay on the host --
but packs a step's env outputs into one pinned block (one async H2D copy), runs ``sg_rollout_feed`` (insert + act in
one launch) and reads the next actions back with one async D2H copy.  ``Policy.act`` + ``RolloutStorage.insert`` remain
available and produce bit-identical buffers (tests/test_gpu_parity.py::test_rollout_feeder_matches_act_insert).
"""
import numpy as np
import torch

from . import _lib


class RolloutFeeder(object):
    def __init__(self, actor_critic, rollouts):
        self.ac, self.rs = actor_critic, rollouts
        kind = type(actor_critic).__name__
        if kind not in ("Policy", "SplitPolicy"):
            raise NotImplementedError("RolloutFeeder knows the Policy (MLPBase + DiagGaussian) and SplitPolicy parameter "
                                      "layouts; use act() + insert() with %s" % kind)
        # Policy: insert + act fused in sg_rollout_feed.  SplitPolicy (model_split.py:39-95, what the shipped scripts train):
        # the same staged block and the same two async copies, the insert as one sg_copy_blocks launch and the act as
        # sg_split_forward writing straight into the buffer slots.
        self.split = kind == "SplitPolicy"
        if not rollouts.obs.is_cuda:
            raise _lib.SgError("RolloutFeeder needs the rollout buffer on a CUDA device; there is no CPU fallback")
        self.dev = rollouts.obs.device
        self.T, self.N = rollouts.rewards.shape[:2]
        self.O, self.F, self.A = rollouts.obs.shape[-1], rollouts.obs_feat.shape[-1], rollouts.actions.shape[-1]
        n = int(_lib.lib().sg_rollout_stage_floats(self.O, self.F, self.N))
        self.stage_host = torch.empty(n, dtype=torch.float32).pin_memory()
        self.stage_dev = torch.empty(n, dtype=torch.float32, device=self.dev)
        N, O, F = self.N, self.O, self.F
        h = self.stage_host.numpy()
        o = 0
        self.h_obs = h[o:o + N * O].reshape(N, O); o += N * O
        self.h_feat = h[o:o + N * F].reshape(N, F); o += N * F
        self.h_reward = h[o:o + N]; o += N
        self.h_mask = h[o:o + N]; o += N
        self.h_bad = h[o:o + N]
        self.action_dev = torch.empty(N, self.A, dtype=torch.float32, device=self.dev)
        self.action_host = torch.empty(N, self.A, dtype=torch.float32).pin_memory()
        self._scratch_logp = torch.empty(N, 1, dtype=torch.float32, device=self.dev)
        self.done_event = torch.cuda.Event()

    def _launch_split(self, step, noise):
        import ctypes as C
        rs, ac, lib = self.rs, self.ac, _lib.lib()
        N, O, F, T = self.N, self.O, self.F, self.T
        slot = step + 1
        stream = _lib.current_stream()
        if step >= 0:
            st = self.stage_dev
            srcs = [st[:N * O], st[N * O:N * (O + F)], st[N * (O + F):N * (O + F) + N], st[N * (O + F) + N:N * (O + F) + 2 * N],
                    st[N * (O + F) + 2 * N:]]
            dsts = [rs.obs[slot], rs.obs_feat[slot], rs.rewards[step], rs.masks[slot], rs.bad_masks[slot]]
            keep = [(a, b) for a, b in zip(srcs, dsts) if a.numel() > 0]
            n = len(keep)
            src_p = (C.c_void_p * n)(*[_lib.ptr(a) for a, _ in keep])
            dst_p = (C.c_void_p * n)(*[_lib.ptr(b) for _, b in keep])
            cnt = (C.c_int * n)(*[a.numel() for a, _ in keep])
            _lib.check(lib.sg_copy_blocks(src_p, dst_p, cnt, n, stream), "sg_copy_blocks")
        last = slot >= T                                    # value of obs[T] doubles as next_value; no action slot there
        logp = self._scratch_logp if last else rs.action_log_probs[slot]
        rc = lib.sg_split_forward(_lib.ptr(ac.flat_params()), O, ac.hidden_size, ac.num_feet, _lib.ptr(rs.obs[slot]), N,
                                  _lib.ptr(noise), None, _lib.ptr(rs.value_preds[slot]), _lib.ptr(self.action_dev),
                                  _lib.ptr(logp), None, stream)
        _lib.check(rc, "sg_split_forward")
        if not last:
            rs.actions[slot].copy_(self.action_dev)

    @_lib.on_device(lambda self, *a, **k: self.dev)
    def _launch(self, step, deterministic):
        rs, ac = self.rs, self.ac
        flat = ac.flat_params()
        noise = None if deterministic else torch.randn(self.N, self.A, device=self.dev, dtype=torch.float32)
        if self.split:
            self._launch_split(step, noise)
            self.action_host.copy_(self.action_dev, non_blocking=True)
            self.done_event.record()
            self.done_event.synchronize()
            return self.action_host.numpy()
        rc = _lib.lib().sg_rollout_feed(
            _lib.ptr(flat), self.O, ac.hidden_size, self.A, self.F, self.N, self.T, step,
            _lib.ptr(self.stage_dev) if step >= 0 else None, _lib.ptr(noise), _lib.ptr(rs.obs), _lib.ptr(rs.obs_feat),
            _lib.ptr(rs.recurrent_hidden_states), _lib.ptr(rs.rewards), _lib.ptr(rs.value_preds),
            _lib.ptr(rs.action_log_probs), _lib.ptr(rs.actions), _lib.ptr(rs.masks), _lib.ptr(rs.bad_masks),
            _lib.ptr(self.action_dev), _lib.current_stream())
        _lib.check(rc, "sg_rollout_feed")
        self.action_host.copy_(self.action_dev, non_blocking=True)
        self.done_event.record()
        self.done_event.synchronize()           # the host envs need the actions: the one sync of the step
        return self.action_host.numpy()

    def begin(self, deterministic=False):
        """Start of a rollout: act on ``rollouts.obs[0]`` (set by ``envs.reset()`` / ``after_update``).
        Returns the (N, A) float32 action array for the first ``envs.step``."""
        self.rs.step = 0
        return self._launch(-1, deterministic)

    def step(self, obs, reward, done, bad_transition, sas_feat, deterministic=False):
        """One env step: ``obs (N,O)``, ``reward (N,)`` or ``(N,1)``, ``done (N,)`` bools, ``bad_transition (N,)`` bools
        (``'bad_transition' in info``), ``sas_feat (N,F)`` -- NumPy arrays straight from the vec-env.  Inserts them
        into slot ``rollouts.step`` and returns the actions for the next ``envs.step``."""
        s = self.rs.step
        np.copyto(self.h_obs, obs, casting="same_kind")
        np.copyto(self.h_feat, sas_feat, casting="same_kind")
        np.copyto(self.h_reward, np.asarray(reward).reshape(-1), casting="same_kind")
        np.subtract(1.0, np.asarray(done, dtype=np.float32), out=self.h_mask)          # masks = 0 where done
        np.subtract(1.0, np.asarray(bad_transition, dtype=np.float32), out=self.h_bad)
        self.stage_dev.copy_(self.stage_host, non_blocking=True)
        out = self._launch(s, deterministic)
        self.rs.step = (s + 1) % self.T
        return out


class ReturnNormalizer(object):
    """The reward scaling the reference's vec-env applies before the rollout buffer sees a reward
    (``VecNormalize(envs, gamma=gamma, ob=False)``, third_party/a2c_ppo_acktr/envs.py:120-125;
    third_party/a2c_ppo_acktr/baselines/common/vec_env/vec_normalize.py:50-58): a discounted running return per env,
    its running variance (``RunningMeanStd(shape=())``), reward / sqrt(var + eps) clipped to +-cliprew, return reset
    where an episode ended.  N numbers per env step on the host, NumPy float64 like the reference (the values are
    bit-identical to the reference class given the same NumPy); it sits where the env outputs are packed for the feed
    (``RolloutFeeder.step(..., reward=normalizer(reward, done))``) and is what the policy-refinement driver
    (third_party/a2c_ppo_acktr/main.py) trains on.  ``ret_rms`` is exposed under the reference's name."""

    def __init__(self, num_envs, gamma=0.99, cliprew=10.0, epsilon=1e-8):
        from .running_mean_std import RunningMeanStd
        self.ret_rms = RunningMeanStd(shape=())
        self.ret = np.zeros(num_envs)
        self.gamma, self.cliprew, self.epsilon = gamma, cliprew, epsilon

    def __call__(self, rews, news):
        self.ret = self.ret * self.gamma + rews
        self.ret_rms.update(self.ret)
        rews = np.clip(rews / np.sqrt(self.ret_rms.var + self.epsilon), -self.cliprew, self.cliprew)
        self.ret[np.asarray(news, dtype=bool)] = 0.
        return rews

    def reset(self):
        self.ret = np.zeros_like(self.ret)
