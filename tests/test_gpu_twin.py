"""The reference's training driver, end to end, on this package (drop-in boundary).

``twin_main.gail_dyn_ppo`` is the body of third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py:main() (bit-identical to the real
one on the reference's modules: tests/test_twin_vs_reference.py).  Here it runs on cuda:0 with every hot-path name resolved
the way the unmodified caller resolves it -- ``from third_party.a2c_ppo_acktr import algo, utils`` etc. after
``simgan_b200.compat.install()`` -- against the fake vec-env, and is held against what the REAL main() of the reference
logged and checkpointed on the CPU (tests/golden/twin_gail_dyn_ppo.npz, oracle/make_golden_twin.py).  Covers together what
the per-call parity tests cover apart: CPU ``Tensor(...)`` arguments into ``insert`` (:230-236), ``rollouts.rewards[step],
returns = predict_reward_combined(...)`` + host RunningMeanStd (:275-297), ``torch.save`` of the policy and of the
discriminator after every update (:307-320) and reloading them the way my_pybullet_envs/utils.py:24-56 does."""
import os
import types

import numpy as np
import pytest
import torch

import fake_env
import twin_main
from simgan_b200 import compat

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "twin_gail_dyn_ppo.npz")
EXPERT = os.path.join(HERE, "golden", "mini_expert.pkl")
# == oracle/make_golden_twin.py TWIN_CFG (the command line of the reference run behind the fixture)
CFG = dict(seed=1, num_processes=4, num_steps=32, num_mini_batch=4, ppo_epoch=3, gail_epoch=2, gail_batch_size=16,
           hidden_size=64, gail_traj_num=3, gail_downsample_frequency=1, gail_dis_hdim=100, num_env_steps=3 * 32 * 4,
           save_interval=1, log_interval=1)
ENV_SEED, NOISE_SEED = 5, 9
LOG_KEYS = ("j", "total_num_steps", "n_episodes", "mean_reward", "median_reward", "min_reward", "max_reward", "dist_entropy",
            "value_loss", "action_loss", "recent_gail_r", "gail_loss", "gail_loss_e", "gail_loss_p")


def caller_namespace():
    """The imports of main_gail_dyn_ppo.py:32-40, resolved through the compat aliases."""
    from third_party.a2c_ppo_acktr import algo, utils
    from third_party.a2c_ppo_acktr.algo import gail
    from third_party.a2c_ppo_acktr.model import Policy
    from third_party.a2c_ppo_acktr.model_split import SplitPolicy
    from third_party.a2c_ppo_acktr.storage import RolloutStorage
    from third_party.a2c_ppo_acktr.baselines.common.running_mean_std import RunningMeanStd
    from simgan_b200 import expert_data
    return types.SimpleNamespace(Policy=Policy, SplitPolicy=SplitPolicy, algo=algo, gail=gail, utils=utils,
                                 RolloutStorage=RolloutStorage, RunningMeanStd=RunningMeanStd, gan_utils=expert_data)


class replay_sampling_noise(object):
    """``act`` draws its N(0,1) noise with torch.randn(B, A, device=cuda) (simgan_b200/model.py:100); replay the stream the
    reference run used instead (tests/fake_env.SamplingNoise)."""

    def __init__(self, seed):
        self.noise = fake_env.SamplingNoise(seed)

    def __enter__(self):
        self.real = torch.randn
        noise, real = self.noise, self.real

        def randn(*size, **kw):
            dev = kw.get("device")
            if dev is not None and torch.device(dev).type == "cuda" and len(size) == 2:
                return noise.next(size).to(dev)
            return real(*size, **kw)
        torch.randn = randn
        return self.noise

    def __exit__(self, *exc):
        torch.randn = self.real
        return False


def _run_twin(tmp_path, act_dim=3, **over):
    compat.install()
    try:
        M = caller_namespace()
        cfg = dict(CFG)
        cfg.update(over)
        args = twin_main.default_args(cuda=True, gail=True, gail_dyn=True, gail_traj_path=EXPERT, save_dir=str(tmp_path), **cfg)
        envs = fake_env.FakeVecEnv(cfg["num_processes"], torch.device("cuda:0"), seed=ENV_SEED, act_dim=act_dim)
        logs = []
        with replay_sampling_noise(NOISE_SEED) as noise:
            out = twin_main.gail_dyn_ppo(args, envs, M, logs.append)
        return args, envs, logs, out, noise
    finally:
        compat.uninstall()


def test_gail_dyn_ppo_driver_matches_the_reference_run(tmp_path):
    z = np.load(GOLDEN)
    args, envs, logs, (actor_critic, discr, agent, rollouts), noise = _run_twin(tmp_path)
    want = z["logs"]
    assert len(logs) == want.shape[0] == 3
    assert noise.calls == 3 * CFG["num_steps"]
    assert all(d == "cuda" for d, _ in envs.actions_seen) and len(envs.actions_seen) == 3 * CFG["num_steps"]
    for row, d in zip(want, logs):
        w = dict(zip(LOG_KEYS, row))
        for k in ("j", "total_num_steps", "n_episodes"):
            assert d[k] == w[k], k
        for k in ("mean_reward", "median_reward", "min_reward", "max_reward"):
            assert round(d[k], 1) == w[k], k
        # north_star: losses within 1e-4 relative on the first outer iteration; the later iterations start from parameters
        # that already differ by fp32 reassociation, which Adam's sign-like first steps amplify
        tol = 1e-4 if w["j"] == 0 else 2e-3
        for k in ("dist_entropy", "value_loss", "gail_loss", "gail_loss_e", "gail_loss_p", "recent_gail_r"):
            assert abs(d[k] - w[k]) <= tol * max(abs(w[k]), 1e-3), (w["j"], k, d[k], w[k])
        assert abs(d["action_loss"] - w["action_loss"]) <= tol * max(abs(w["action_loss"]), 0.05), (w["j"], d["action_loss"])

    # ---- checkpoints: written after every update, whole-object pickles under the reference's class paths ------------------
    save_path = os.path.join(str(tmp_path), "ppo")
    names = sorted(os.listdir(save_path))
    for j in range(3):
        assert "FakeCombinedEnv-v1_%d.pt" % j in names and "FakeCombinedEnv-v1_%d_D.pt" % j in names
    raw = open(os.path.join(save_path, "FakeCombinedEnv-v1.pt"), "rb").read()
    assert b"third_party.a2c_ppo_acktr.model" in raw or b"simgan_b200.model" in raw
    # reload as my_pybullet_envs/utils.py:24-56 does on a CPU-only worker (torch.load(path, map_location="cpu"))
    compat.install()
    try:
        pol_cpu, ob_rms = torch.load(os.path.join(save_path, "FakeCombinedEnv-v1.pt"), map_location="cpu", weights_only=False)
        d_cpu = torch.load(os.path.join(save_path, "FakeCombinedEnv-v1_D.pt"), map_location="cpu", weights_only=False)
    finally:
        compat.uninstall()
    assert ob_rms is None and type(pol_cpu).__name__ == "Policy" and not next(pol_cpu.parameters()).is_cuda
    rhs = torch.zeros(1, pol_cpu.recurrent_hidden_state_size)
    masks = torch.zeros(1, 1)
    obs = torch.randn(1, 11)
    with torch.no_grad():
        v_cpu, a_cpu, _, _ = pol_cpu.act(obs, rhs, masks, deterministic=True)
        v_gpu, a_gpu, _, _ = actor_critic.act(obs.cuda(), rhs.cuda(), masks.cuda(), deterministic=True)
    assert torch.allclose(v_cpu, v_gpu.cpu(), atol=1e-5) and torch.allclose(a_cpu, a_gpu.cpu(), atol=1e-5)
    for (k, a), b in zip(pol_cpu.state_dict().items(), actor_critic.state_dict().values()):
        assert torch.equal(a, b.cpu()), k
    for a, b in zip(d_cpu.trunk.state_dict().values(), discr.trunk.state_dict().values()):
        assert torch.equal(a, b.cpu())
    # the reloaded discriminator scores on the CPU worker the way gail.py:212-216 does
    with torch.no_grad():
        p = d_cpu.predict_prob_single_step(torch.zeros(1, 14), torch.zeros(1, 11))
    assert p.shape == (1, 1) and 0.0 < float(p) < 1.0

    # ---- parameters after the last update vs the reference's checkpoint ---------------------------------------------------
    worst = 0.0
    for k, v in actor_critic.state_dict().items():
        w = torch.from_numpy(z["pol2_" + k])
        worst = max(worst, float((v.cpu() - w).abs().max()))
    for k, v in discr.trunk.state_dict().items():
        w = torch.from_numpy(z["disc2_" + k])
        worst = max(worst, float((v.cpu() - w).abs().max()))
    # 36 PPO + 12 D Adam steps of lr 3e-4: a parameter whose gradient sign is decided by rounding moves by up to 2*lr per step
    assert worst < 5e-3, worst


def test_gail_dyn_ppo_driver_with_split_policy(tmp_path):
    """The policy every shipped script trains (train_hopper_deform.sh:5: --use-split-pi --num-feet 1; hidden 100 here), held
    against the reference's REAL main() run with --use-split-pi (tests/golden/twin_gail_dyn_ppo_split.npz)."""
    z = np.load(os.path.join(HERE, "golden", "twin_gail_dyn_ppo_split.npz"))
    args, envs, logs, (actor_critic, discr, agent, rollouts), _ = _run_twin(tmp_path, act_dim=7, use_split_pi=True,
                                                                             num_feet=1, hidden_size=100)
    assert type(actor_critic).__name__ == "SplitPolicy" and len(logs) == 3
    for row, d in zip(z["logs"], logs):
        w = dict(zip(LOG_KEYS, row))
        assert d["n_episodes"] == w["n_episodes"]
        tol = 1e-4 if w["j"] == 0 else 2e-3
        for k in ("dist_entropy", "value_loss", "gail_loss", "gail_loss_e", "gail_loss_p", "recent_gail_r"):
            assert abs(d[k] - w[k]) <= tol * max(abs(w[k]), 1e-3), (w["j"], k, d[k], w[k])
        assert abs(d["action_loss"] - w["action_loss"]) <= tol * max(abs(w["action_loss"]), 0.05), (w["j"], d["action_loss"])
    worst = 0.0
    for k, v in actor_critic.state_dict().items():
        worst = max(worst, float((v.cpu() - torch.from_numpy(z["pol2_" + k])).abs().max()))
    assert worst < 5e-3, worst
    compat.install()
    try:
        pol_cpu, _ = torch.load(os.path.join(str(tmp_path), "ppo", "FakeCombinedEnv-v1.pt"), map_location="cpu",
                                weights_only=False)
    finally:
        compat.uninstall()
    assert type(pol_cpu).__name__ == "SplitPolicy"
    for a, b in zip(pol_cpu.state_dict().values(), actor_critic.state_dict().values()):
        assert torch.equal(a, b.cpu())


def test_policy_refinement_driver_matches_the_reference_run(tmp_path):
    """Second shipped command (train_hopper_deform.sh:7 -> third_party/a2c_ppo_acktr/main.py:69-88, 199-268): a warm-started
    policy gets reset_critic + reset_variance, then PPO with --use-linear-lr-decay --clip-param 0.1 --ppo-epoch 2
    --num-mini-batch 8 --lr 1.5e-4 --entropy-coef 0 on rewards scaled by the vec-env's return normaliser
    (VecNormalize(envs, gamma), here simgan_b200.ReturnNormalizer).  Held against the REAL main.py run of the reference
    (tests/golden/twin_policy_refinement.npz)."""
    import simgan_b200 as sg
    z = np.load(os.path.join(HERE, "golden", "twin_policy_refinement.npz"))
    cfg = dict(seed=3, num_processes=4, num_steps=32, num_mini_batch=8, ppo_epoch=2, lr=1.5e-4, entropy_coef=0.0,
               clip_param=0.1, hidden_size=64, num_env_steps=3 * 32 * 4, save_interval=1, log_interval=1,
               warm_start_logstd=-1.0, use_linear_lr_decay=True, gamma=0.99)
    # the warm-start checkpoint: the parameters the reference's GAIL run ended with, as a whole-object pickle of OUR class
    warm = sg.Policy((11,), fake_env.Box((3,)), base_kwargs={"recurrent": False, "hidden_size": 64})
    sd = warm.state_dict()
    for k in sd:
        sd[k] = torch.from_numpy(z["warm_" + k])
    warm.load_state_dict(sd)
    warm_path = str(tmp_path / "warm.pt")
    torch.save([warm, None], warm_path)

    compat.install()
    try:
        M = caller_namespace()
        args = twin_main.default_args(cuda=True, warm_start=warm_path, save_dir=str(tmp_path), **cfg)
        envs = fake_env.FakeVecEnv(4, torch.device("cuda:0"), seed=6, reward_filter=sg.ReturnNormalizer(4, gamma=0.99))
        logs = []
        with replay_sampling_noise(10):
            actor_critic, agent, rollouts = twin_main.policy_refinement(args, envs, M, logs.append)
    finally:
        compat.uninstall()
    keys = ("j", "total_num_steps", "n_episodes", "mean_reward", "median_reward", "min_reward", "max_reward", "dist_entropy",
            "value_loss", "action_loss")
    assert len(logs) == 3
    assert [round(d["lr"], 12) for d in logs] == [round(1.5e-4 * (1 - j / 3.0), 12) for j in range(3)]      # utils.py:68-72
    for row, d in zip(z["logs"], logs):
        w = dict(zip(keys, row))
        tol = 1e-4 if w["j"] == 0 else 2e-3
        assert d["n_episodes"] == w["n_episodes"]
        for k in ("dist_entropy", "value_loss"):
            assert abs(d[k] - w[k]) <= tol * abs(w[k]), (w["j"], k, d[k], w[k])
        assert abs(d["action_loss"] - w["action_loss"]) <= tol * max(abs(w["action_loss"]), 0.05), (w["j"], d["action_loss"])
    worst = 0.0
    for k, v in actor_critic.state_dict().items():
        worst = max(worst, float((v.cpu() - torch.from_numpy(z["pol2_" + k])).abs().max()))
    assert worst < 3e-3, worst
    # the critic really was re-initialised (main.py:84) and the log-std reset (main.py:85-86) before training moved them
    assert not np.allclose(z["warm_base.critic.0.weight"], z["pol0_base.critic.0.weight"], atol=1e-2)
    assert abs(float(actor_critic.dist.logstd._bias.mean()) + 1.0) < 0.05
