"""tests/twin_main.py is a restatement of the reference's driver; this keeps it honest.  Where the reference tree is
mounted (dev container), the REAL ``main()`` of third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py and the twin -- both on
the reference's own modules, CPU, same fake vec-env, same sampling noise -- must produce identical log values and
bit-identical checkpoints, and the committed fixture tests/golden/twin_gail_dyn_ppo.npz must be what the real main()
produces today."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "twin_gail_dyn_ppo.npz")


def _reference_namespace(ref, safe_loader):
    return types.SimpleNamespace(Policy=ref.model.Policy, SplitPolicy=ref.model_split.SplitPolicy,
                                 algo=types.SimpleNamespace(PPO=ref.ppo.PPO), gail=ref.gail, utils=ref.utils,
                                 RolloutStorage=ref.storage.RolloutStorage, RunningMeanStd=ref.rms.RunningMeanStd,
                                 gan_utils=types.SimpleNamespace(load_sas_wpast_from_pickle=safe_loader,
                                                                 select_and_merge_sas=ref.env_utils.select_and_merge_sas))


def test_real_main_equals_twin_and_golden(tmp_path):
    from oracle import make_golden_twin as mg
    from oracle import run_reference_main as rrm
    import fake_env
    import twin_main

    cfg = mg.TWIN_CFG
    real_logs, real_path = mg.run_reference(str(tmp_path / "real"))
    assert len(real_logs) == 3

    # the twin on the same (reference) modules
    noise = fake_env.SamplingNoise(cfg["noise_seed"])
    old_normal = torch.normal
    twin_logs = []
    args = twin_main.default_args(gail=True, gail_dyn=True, gail_traj_path=mg.EXPERT, save_dir=str(tmp_path / "twin"),
                                  **{k: v for k, v in cfg.items() if k not in ("env_seed", "noise_seed")})
    with rrm.bound_reference() as ref:
        try:
            torch.normal = lambda mean, std, **kw: mean + std * noise.next(mean.shape).to(mean.device)
            envs = fake_env.FakeVecEnv(cfg["num_processes"], torch.device("cpu"), seed=cfg["env_seed"])
            twin_main.gail_dyn_ppo(args, envs, _reference_namespace(ref, rrm._safe_load_sas), twin_logs.append)
        finally:
            torch.normal = old_normal
    assert len(twin_logs) == len(real_logs)
    for a, b in zip(twin_logs, real_logs):
        for k in ("j", "total_num_steps", "n_episodes", "dist_entropy", "value_loss", "action_loss", "gail_loss", "gail_loss_e",
                  "gail_loss_p"):
            assert a[k] == b[k], (k, a[k], b[k])                  # '{}'-formatted Python floats round-trip exactly
        assert abs(a["recent_gail_r"] - b["recent_gail_r"]) <= 1e-6 * abs(b["recent_gail_r"])     # printed as float32
        for k in ("mean_reward", "median_reward", "min_reward", "max_reward"):
            assert round(a[k], 1) == b[k]                         # '{:.1f}'
    for j in range(3):
        rp, rd = mg.checkpoint_params(real_path, j, rrm.bound_reference())
        tp, td = mg.checkpoint_params(os.path.join(str(tmp_path / "twin"), "ppo"), j, rrm.bound_reference())
        for k in rp:
            assert torch.equal(rp[k], tp[k]), k
        for k in rd:
            assert torch.equal(rd[k], td[k]), k

    # the committed fixture is what the real main() produces
    z = np.load(GOLDEN)
    want = np.array([[d[k] for k in mg.LOG_KEYS] for d in real_logs])
    assert np.array_equal(z["logs"], want)
    rp, rd = mg.checkpoint_params(real_path, 2, rrm.bound_reference())
    for k, v in rp.items():
        assert np.array_equal(z["pol2_" + k], v.numpy()), k
    for k, v in rd.items():
        assert np.array_equal(z["disc2_" + k], v.numpy()), k


def test_real_main_py_equals_refinement_twin_and_golden(tmp_path):
    """Same for the PPO-only refinement driver (third_party/a2c_ppo_acktr/main.py, second shipped command): warm start from
    the checkpoint the GAIL run left, reset_critic + reset_variance, linear lr decay, VecNormalize'd rewards."""
    from oracle import make_golden_twin as mg
    from oracle import run_reference_main as rrm
    import fake_env
    import twin_main

    _, gail_path = mg.run_reference(str(tmp_path / "gail"))
    warm = os.path.join(gail_path, "FakeCombinedEnv-v1.pt")
    cfg = mg.REFINE_CFG
    real_logs, real_path = mg.run_reference_refinement(str(tmp_path / "real"), warm)

    noise = fake_env.SamplingNoise(cfg["noise_seed"])
    old_normal = torch.normal
    twin_logs = []
    args = twin_main.default_args(warm_start=warm, save_dir=str(tmp_path / "twin"),
                                  **{k: v for k, v in cfg.items() if k not in ("env_seed", "noise_seed")})
    envs = fake_env.FakeVecEnv(cfg["num_processes"], torch.device("cpu"), seed=cfg["env_seed"],
                               reward_filter=rrm.reference_vec_normalize(cfg["num_processes"], cfg["gamma"]))
    with rrm.bound_reference() as ref:
        try:
            torch.normal = lambda mean, std, **kw: mean + std * noise.next(mean.shape).to(mean.device)
            twin_main.policy_refinement(args, envs, _reference_namespace(ref, rrm._safe_load_sas), twin_logs.append)
        finally:
            torch.normal = old_normal
    assert len(twin_logs) == len(real_logs) == 3
    for a, b in zip(twin_logs, real_logs):
        for k in ("j", "n_episodes", "dist_entropy", "value_loss", "action_loss"):
            assert a[k] == b[k], (k, a[k], b[k])
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "twin_policy_refinement.npz"))
    assert np.array_equal(z["logs"], np.array([[d[k] for k in mg.REFINE_LOG_KEYS] for d in real_logs]))
    for j in range(3):
        rp = mg.policy_checkpoint(real_path, j, rrm.bound_reference())
        tp = mg.policy_checkpoint(os.path.join(str(tmp_path / "twin"), "ppo"), j, rrm.bound_reference())
        for k in rp:
            assert torch.equal(rp[k], tp[k]), k
            assert np.array_equal(z["pol%d_%s" % (j, k)], rp[k].numpy()), k


def test_return_normalizer_equals_reference_vec_normalize():
    """simgan_b200.feed.ReturnNormalizer against the reference's real VecNormalize(venv, gamma, ob=False) (envs.py:120-125,
    vec_normalize.py:50-58): bit-identical rewards and running statistics over a stream with episode ends."""
    from oracle import run_reference_main as rrm
    import simgan_b200 as sg
    ref_filter = rrm.reference_vec_normalize(5, 0.99)
    ours = sg.ReturnNormalizer(5, gamma=0.99)
    rs = np.random.RandomState(0)
    for t in range(200):
        rews = rs.standard_normal(5).astype(np.float32) * (1.0 + t % 7)
        news = rs.rand(5) < 0.1
        a = ref_filter(rews.copy(), news.copy())
        b = ours(rews.copy(), news.copy())
        assert a.dtype == b.dtype and np.array_equal(a, b), t


def test_split_policy_golden_is_what_the_real_main_produces(tmp_path):
    from oracle import make_golden_twin as mg
    logs, _ = mg.run_reference(str(tmp_path / "split"), mg.SPLIT_CFG, mg.SPLIT_ARGV, act_dim=7)
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "twin_gail_dyn_ppo_split.npz"))
    assert np.array_equal(z["logs"], np.array([[d[k] for k in mg.LOG_KEYS] for d in logs]))
