"""Run the reference's REAL ``main()`` (third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py, unmodified, imported from
/root/reference) on the CPU against the fake vec-env of tests/fake_env.py.

TEST INFRASTRUCTURE (dev container only: needs /root/reference).  Used by ``oracle/make_golden_twin.py`` to write
tests/golden/twin_gail_dyn_ppo.npz and by tests/test_twin_vs_reference.py to keep tests/twin_main.py honest.

What is substituted, and why (nothing on the hot path):
  * ``sys.argv``                      -> the command line of the run
  * ``gym``                           -> stub module; ``gym.make(...)`` returns an object with ``getSourceCode()`` / ``close()``
                                         (main_gail_dyn_ppo.py:96-112 dumps the env's source next to the checkpoints)
  * ``third_party.a2c_ppo_acktr.envs``-> stub with ``make_vec_envs`` returning the fake vec-env and an empty ``VecNormalize``
                                         class (PyBullet and gym are not in the image)
  * ``numpy.infty``                   -> alias of ``numpy.inf`` for the duration of the run (removed in NumPy 2.0; :192 reads it)
  * ``gan_utils.load_sas_wpast_from_pickle`` -> a column-wise stacker: the reference's ragged ``np.array(sas)``
                                         (my_pybullet_envs/utils.py:193) raises on NumPy >= 1.24; same arrays otherwise
  * ``torch.load``                    -> default weights_only=False (torch >= 2.6 flipped the default; --warm-start loads a
                                         whole-object pickle, main.py:80-83)
  * ``my_pybullet_envs.laikago``      -> stub (main.py:40 imports mirror tables that only --dup-sym uses)
  * ``torch.normal``                  -> mean + std * eps with eps from the counter-based stream of tests/fake_env.SamplingNoise,
                                         so that a CUDA run (whose sampler draws from the CUDA generator) can replay the actions
The reference's own modules (model, storage, algo.ppo, algo.gail, distributions, utils, running_mean_std) run untouched.
"""
import importlib.util
import logging
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import ref_shim  # noqa: E402

_LOG_RE = re.compile(
    r"Updates (?P<j>\d+), num timesteps (?P<total_num_steps>\d+), FPS \d+ \n Last (?P<n_episodes>\d+) training episodes: "
    r"mean/median reward (?P<mean_reward>[-\d.]+)/(?P<median_reward>[-\d.]+), min/max reward (?P<min_reward>[-\d.]+)/"
    r"(?P<max_reward>[-\d.]+), dist en (?P<dist_entropy>\S+), l_pi (?P<value_loss>\S+), l_vf (?P<action_loss>\S+), "
    r"recent_gail_r (?P<recent_gail_r>\S+),loss_gail (?P<gail_loss>\S+), loss_gail_e (?P<gail_loss_e>\S+), "
    r"loss_gail_p (?P<gail_loss_p>\S+)\n")


def _safe_load_sas(pathname, downsample_freq=1, load_num_trajs=None):
    import pickle
    with open(pathname, "rb") as handle:
        saved_file = pickle.load(handle)
    n_trajs = len(saved_file)
    start_idx = torch.randint(0, downsample_freq, size=(n_trajs,)).long()
    sas = []
    for traj_idx, traj_tuples in saved_file.items():
        sas.extend(traj_tuples[start_idx[traj_idx]::downsample_freq])
        if load_num_trajs and traj_idx >= load_num_trajs - 1:
            break
    return [np.array([row[item] for row in sas]) for item in range(len(sas[0]))]


class _Bound(object):
    """Bind the reference's modules (and the stubs) under their canonical names for the duration of a block."""

    def __init__(self, extra):
        self.ref = ref_shim.load()
        self.extra = extra

    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items()
                      if k == "third_party" or k.startswith("third_party.") or k in ("pybullet", "gym", "my_pybullet_envs")
                      or k.startswith("my_pybullet_envs.")}
        for k in self.saved:
            del sys.modules[k]
        sys.modules.update(self.ref.modules)
        sys.modules.update(self.extra)
        return self.ref

    def __exit__(self, *exc):
        for k in list(sys.modules):
            if k == "third_party" or k.startswith("third_party.") or k in ("pybullet", "gym", "my_pybullet_envs") \
                    or k.startswith("my_pybullet_envs."):
                del sys.modules[k]
        sys.modules.update(self.saved)
        return False


def bound_reference(extra=None):
    return _Bound(extra or {})


_LOG_RE_MAIN = re.compile(
    r"Updates (?P<j>\d+), num timesteps (?P<total_num_steps>\d+), FPS \d+ \n Last (?P<n_episodes>\d+) training episodes: "
    r"mean/median reward (?P<mean_reward>[-\d.]+)/(?P<median_reward>[-\d.]+), min/max reward (?P<min_reward>[-\d.]+)/"
    r"(?P<max_reward>[-\d.]+), dist en (?P<dist_entropy>\S+), l_pi (?P<value_loss>\S+), l_vf (?P<action_loss>\S+) \n")


class _Permissive(types.ModuleType):
    """Stub package whose attributes are stub packages (``from gym import spaces`` etc. at import time of modules whose
    gym-dependent code never runs here)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Permissive(self.__name__ + "." + k)
        setattr(self, k, m)
        return m


def reference_vec_normalize(num_envs, gamma):
    """reward_filter for tests/fake_env.FakeVecEnv backed by the reference's REAL VecNormalize(venv, gamma=gamma, ob=False)
    (envs.py:120-125) wrapped around a minimal venv duck."""
    gym = _Permissive("gym")
    with bound_reference({"gym": gym, "gym.spaces": gym.spaces}):
        vec_env = importlib.import_module("third_party.a2c_ppo_acktr.baselines.common.vec_env")

        class _Venv(object):
            observation_space = None
            action_space = None

            def __init__(self):
                self.num_envs = num_envs
                self.out = None

            def step_wait(self):
                return self.out
        venv = _Venv()
        vn = vec_env.VecNormalize(venv, gamma=gamma, ob=False)

    def filt(rews, news):
        venv.out = (np.zeros((num_envs, 1)), rews, news, [{} for _ in range(num_envs)])
        return vn.step_wait()[1]
    return filt


def run(argv, make_env, noise, save_dir, main_file="main_gail_dyn_ppo.py"):
    """Returns (list of per-update log dicts parsed from the reference's own log line, save_path)."""
    gym_stub = types.ModuleType("gym")

    class _Dummy(object):
        def getSourceCode(self):
            return "# fake env: no source\n"

        def close(self):
            pass

        def reset(self):
            return None

    gym_stub.make = lambda *a, **k: _Dummy()
    laika_stub = types.ModuleType("my_pybullet_envs.laikago")          # main.py:40 imports its mirror tables (--dup-sym only)
    laika_stub.mirror_obs = laika_stub.mirror_action = None
    envs_stub = types.ModuleType("third_party.a2c_ppo_acktr.envs")
    envs_stub.VecNormalize = type("VecNormalize", (), {})
    envs_stub.make_vec_envs = lambda env_name, seed, num_processes, gamma, log_dir, device, allow_early_resets, **kw: \
        make_env(num_processes, device)

    records = []

    class _Grab(logging.Handler):
        def emit(self, record):
            records.append(record.getMessage())

    root_logger = logging.getLogger()
    old_handlers, old_level = list(root_logger.handlers), root_logger.level
    grab = _Grab()
    old_argv, old_normal, old_load, had_infty = sys.argv, torch.normal, torch.load, hasattr(np, "infty")
    with bound_reference({"gym": gym_stub, "third_party.a2c_ppo_acktr.envs": envs_stub,
                          "my_pybullet_envs.laikago": laika_stub}) as ref:
        try:
            sys.argv = ["main_gail_dyn_ppo.py"] + list(argv)
            if not had_infty:
                np.infty = np.inf
            torch.normal = lambda mean, std, **kw: mean + std * noise.next(mean.shape).to(mean.device)
            # torch >= 2.6 defaults torch.load to weights_only=True; the reference's checkpoints are whole-object pickles
            torch.load = lambda *a, **k: old_load(*a, **dict(dict(weights_only=False), **k))
            old_loader = ref.env_utils.load_sas_wpast_from_pickle
            ref.env_utils.load_sas_wpast_from_pickle = _safe_load_sas
            sys.modules["third_party.a2c_ppo_acktr"].envs = envs_stub
            spec = importlib.util.spec_from_file_location(
                "_simgan_ref_main", os.path.join(ref_shim.REF_ROOT, "third_party", "a2c_ppo_acktr", main_file))
            mod = importlib.util.module_from_spec(spec)
            root_logger.addHandler(grab)
            spec.loader.exec_module(mod)
            mod.main()
        finally:
            sys.argv, torch.normal, torch.load = old_argv, old_normal, old_load
            if not had_infty:
                del np.infty
            ref.env_utils.load_sas_wpast_from_pickle = old_loader
            for h in list(root_logger.handlers):
                if h not in old_handlers:
                    root_logger.removeHandler(h)
                    try:
                        h.close()
                    except Exception:
                        pass
            root_logger.setLevel(old_level)
    logs = []
    for msg in records:
        m = _LOG_RE.search(msg) or _LOG_RE_MAIN.search(msg)
        if m:
            d = {k: float(v) for k, v in m.groupdict().items()}
            # the reference's format call passes (dist_entropy, value_loss, action_loss) into "dist en {}, l_pi {}, l_vf {}"
            logs.append(d)
    return logs, os.path.join(save_dir, "ppo")
