"""A PyBullet-free stand-in for ``make_vec_envs(...)`` of the reference (third_party/a2c_ppo_acktr/envs.py:89-137).

TEST INFRASTRUCTURE.  The reference's collection loop (main_gail_dyn_ppo.py:201-236) talks to a ``VecPyTorch`` wrapper:
``reset() -> obs`` (device tensor), ``step(action) -> (obs device tensor, reward CPU tensor (N,1), done bool array,
infos)`` (envs.py:199-210), where every ``infos[i]`` of a *CombinedEnv carries ``"sas_window"`` -- the ``2W+1`` lists
``[s_t..s_{t-W+1}, a_t..a_{t-W+1}, s_{t+1}]`` (my_pybullet_envs/hopper_env_combined_policy.py) -- plus ``"episode"``
at episode ends (bench.Monitor) and ``"bad_transition"`` on time-limit resets (envs.py:140-150).

This fake emits exactly that interface from counter-based NumPy streams: what it returns at step t depends only on
(seed, t), never on the actions, so the real ``main()`` (reference modules, CPU) and the twin driven by this package on a
GPU see identical environments.  No physics is simulated.
"""
import numpy as np
import torch


class _Space(object):
    def __init__(self, shape):
        self.shape = tuple(shape)


class Box(_Space):          # the reference dispatches on ``action_space.__class__.__name__ == "Box"`` (model.py:56)
    pass


class FakeVecEnv(object):
    def __init__(self, num_processes, device, seed=0, obs_dim=11, act_dim=3, s_dim=11, a_dim=3, window=10, ep_len=9,
                 reward_filter=None):
        # reward_filter(rews, news) -> rews: the VecNormalize return scaling of envs.py:120-125, which sits between the raw
        # envs and VecPyTorch in the reference's wrapper chain
        self.reward_filter = reward_filter
        self.N, self.device, self.seed = int(num_processes), device, int(seed)
        self.O, self.A, self.s_dim, self.a_dim, self.W, self.ep_len = obs_dim, act_dim, s_dim, a_dim, window, ep_len
        self.observation_space = _Space((obs_dim,))
        self.action_space = Box((act_dim,))
        self.t = 0
        self.actions_seen = []          # (device type, shape) of every action batch handed to step()

    def _rng(self, t, salt):
        return np.random.RandomState((self.seed * 1000003 + t * 7 + salt) % (2 ** 31 - 1))

    def _obs(self, t):
        return self._rng(t, 1).standard_normal((self.N, self.O)).astype(np.float32)

    def reset(self):
        self.t = 0
        return torch.from_numpy(self._obs(0)).float().to(self.device)

    def step(self, action):
        assert tuple(action.shape) == (self.N, self.A), action.shape
        self.actions_seen.append((action.device.type, tuple(action.shape)))
        action.cpu().numpy()            # envs.py:199-205: the action batch crosses to the host every step
        self.t += 1
        t = self.t
        obs = self._obs(t)
        r = self._rng(t, 2)
        reward = r.standard_normal(self.N).astype(np.float32)
        win = r.standard_normal((self.N, self.W, self.s_dim + self.a_dim)) * 0.5
        s_next = r.standard_normal((self.N, self.s_dim)) * 0.5
        done = np.array([(t + 3 * i) % self.ep_len == 0 for i in range(self.N)])
        if self.reward_filter is not None:
            reward = np.asarray(self.reward_filter(reward, done))
        infos = []
        for i in range(self.N):
            info = {"sas_window": [list(win[i, k, :self.s_dim]) for k in range(self.W)] +
                                  [list(win[i, k, self.s_dim:]) for k in range(self.W)] + [list(s_next[i])]}
            if done[i]:
                info["episode"] = {"r": float(10.0 + i + 0.01 * t), "l": self.ep_len}
                if (t + i) % 2 == 0:
                    info["bad_transition"] = True
            infos.append(info)
        return (torch.from_numpy(obs).float().to(self.device), torch.from_numpy(reward).unsqueeze(dim=1).float(), done,
                infos)

    def close(self):
        pass


class SamplingNoise(object):
    """The N(0,1) draws behind ``dist.sample()`` (model.py:96) as a counter-based stream, so that a CPU run of the
    reference and a CUDA run of this package sample the same actions (their native draws come from different generators:
    the CPU default generator there, the CUDA generator here)."""

    def __init__(self, seed):
        self.seed, self.calls = int(seed), 0

    def next(self, shape):
        rs = np.random.RandomState((self.seed * 7919 + self.calls) % (2 ** 31 - 1))
        self.calls += 1
        return torch.from_numpy(rs.standard_normal(tuple(shape)).astype(np.float32))
