"""Write tests/golden/twin_gail_dyn_ppo.npz: what the reference's REAL main() (oracle/run_reference_main.py) logs and
checkpoints over three outer iterations on the fake vec-env -- the fixture tests/test_gpu_twin.py holds this package's
CUDA run of the same loop against.  TEST INFRASTRUCTURE; dev container only (needs /root/reference).

    python oracle/make_golden_twin.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import run_reference_main as rrm  # noqa: E402
import fake_env  # noqa: E402

EXPERT = os.path.join(ROOT, "tests", "golden", "mini_expert.pkl")
# one place for the run's command line: the twin tests build their args namespace from the same dict
TWIN_CFG = dict(seed=1, num_processes=4, num_steps=32, num_mini_batch=4, ppo_epoch=3, gail_epoch=2, gail_batch_size=16,
                hidden_size=64, gail_traj_num=3, gail_downsample_frequency=1, gail_dis_hdim=100, num_env_steps=3 * 32 * 4,
                save_interval=1, log_interval=1, env_seed=5, noise_seed=9)
LOG_KEYS = ("j", "total_num_steps", "n_episodes", "mean_reward", "median_reward", "min_reward", "max_reward", "dist_entropy",
            "value_loss", "action_loss", "recent_gail_r", "gail_loss", "gail_loss_e", "gail_loss_p")


def argv(save_dir, cfg=TWIN_CFG):
    return ["--env-name", "FakeCombinedEnv-v1", "--algo", "ppo", "--no-cuda", "--gail", "--gail-dyn",
            "--seed", str(cfg["seed"]), "--num-processes", str(cfg["num_processes"]), "--num-steps", str(cfg["num_steps"]),
            "--num-mini-batch", str(cfg["num_mini_batch"]), "--ppo-epoch", str(cfg["ppo_epoch"]),
            "--gail-epoch", str(cfg["gail_epoch"]), "--gail-batch-size", str(cfg["gail_batch_size"]),
            "--hidden-size", str(cfg["hidden_size"]), "--gail-traj-path", EXPERT, "--gail-traj-num", str(cfg["gail_traj_num"]),
            "--gail-downsample-frequency", str(cfg["gail_downsample_frequency"]), "--gail-dis-hdim", str(cfg["gail_dis_hdim"]),
            "--num-env-steps", str(cfg["num_env_steps"]), "--save-interval", str(cfg["save_interval"]),
            "--log-interval", str(cfg["log_interval"]), "--save-dir", save_dir, "--log-dir", os.path.join(save_dir, "log")]


def run_reference(save_dir, cfg=TWIN_CFG, extra_argv=(), act_dim=3):
    noise = fake_env.SamplingNoise(cfg["noise_seed"])
    logs, save_path = rrm.run(argv(save_dir, cfg) + list(extra_argv),
                              lambda n, dev: fake_env.FakeVecEnv(n, dev, seed=cfg["env_seed"], act_dim=act_dim), noise, save_dir)
    return logs, save_path


# what train_hopper_deform.sh:5 trains: --use-split-pi (model_split.py), one foot (7 action dims), hidden 100 here
SPLIT_ARGV = ("--use-split-pi", "--num-feet", "1")
SPLIT_CFG = dict(TWIN_CFG, hidden_size=100)


def checkpoint_params(save_path, j, bound):
    """state_dicts of the policy and the discriminator the run saved after update j (whole-object pickles of the
    reference's classes: loaded with its modules bound)."""
    with bound:
        pol, ob_rms = torch.load(os.path.join(save_path, "FakeCombinedEnv-v1_%d.pt" % j), weights_only=False)
        disc = torch.load(os.path.join(save_path, "FakeCombinedEnv-v1_%d_D.pt" % j), weights_only=False)
        assert ob_rms is None
        return ({k: v.clone() for k, v in pol.state_dict().items()}, {k: v.clone() for k, v in disc.trunk.state_dict().items()})


# second shipped command (train_hopper_deform.sh:7 -> third_party/a2c_ppo_acktr/main.py): PPO-only refinement of a warm-started
# policy with a fresh critic and log-std, linear lr decay, clip 0.1, 2 epochs x 8 minibatches, VecNormalize'd env rewards
REFINE_CFG = dict(seed=3, num_processes=4, num_steps=32, num_mini_batch=8, ppo_epoch=2, lr=1.5e-4, entropy_coef=0.0,
                  clip_param=0.1, hidden_size=64, num_env_steps=3 * 32 * 4, save_interval=1, log_interval=1,
                  warm_start_logstd=-1.0, use_linear_lr_decay=True, env_seed=6, noise_seed=10, gamma=0.99)
REFINE_LOG_KEYS = ("j", "total_num_steps", "n_episodes", "mean_reward", "median_reward", "min_reward", "max_reward",
                   "dist_entropy", "value_loss", "action_loss")


def refine_argv(save_dir, warm_start, cfg=REFINE_CFG):
    return ["--env-name", "FakeCombinedEnv-v1", "--algo", "ppo", "--no-cuda", "--seed", str(cfg["seed"]),
            "--num-processes", str(cfg["num_processes"]), "--num-steps", str(cfg["num_steps"]), "--lr", str(cfg["lr"]),
            "--entropy-coef", str(cfg["entropy_coef"]), "--ppo-epoch", str(cfg["ppo_epoch"]),
            "--num-mini-batch", str(cfg["num_mini_batch"]), "--num-env-steps", str(cfg["num_env_steps"]),
            "--use-linear-lr-decay", "--clip-param", str(cfg["clip_param"]), "--hidden-size", str(cfg["hidden_size"]),
            "--warm-start", warm_start, "--warm-start-logstd", str(cfg["warm_start_logstd"]),
            "--save-interval", str(cfg["save_interval"]), "--log-interval", str(cfg["log_interval"]), "--save-dir", save_dir,
            "--log-dir", os.path.join(save_dir, "log")]


def run_reference_refinement(save_dir, warm_start, cfg=REFINE_CFG):
    noise = fake_env.SamplingNoise(cfg["noise_seed"])

    def make_env(n, dev):
        return fake_env.FakeVecEnv(n, dev, seed=cfg["env_seed"], reward_filter=rrm.reference_vec_normalize(n, cfg["gamma"]))
    return rrm.run(refine_argv(save_dir, warm_start, cfg), make_env, noise, save_dir, main_file="main.py")


def policy_checkpoint(save_path, j, bound):
    with bound:
        pol, _ = torch.load(os.path.join(save_path, "FakeCombinedEnv-v1_%d.pt" % j), weights_only=False)
        return {k: v.clone() for k, v in pol.state_dict().items()}


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        logs, save_path = run_reference(tmp)
        # ---- refinement run, warm-started from the checkpoint the first run left --------------------------------------------
        rtmp = os.path.join(tmp, "refine")
        rlogs, rpath = run_reference_refinement(rtmp, os.path.join(save_path, "FakeCombinedEnv-v1.pt"))
        n_r = REFINE_CFG["num_env_steps"] // REFINE_CFG["num_steps"] // REFINE_CFG["num_processes"]
        assert len(rlogs) == n_r, (len(rlogs), n_r)
        rout = {"logs": np.array([[d[k] for k in REFINE_LOG_KEYS] for d in rlogs], dtype=np.float64)}
        for j in range(n_r):
            for k, v in policy_checkpoint(rpath, j, rrm.bound_reference()).items():
                rout["pol%d_%s" % (j, k)] = v.numpy()
        for k, v in checkpoint_params(save_path, 2, rrm.bound_reference())[0].items():
            rout["warm_" + k] = v.numpy()
        # ---- the GAIL driver with SplitPolicy ------------------------------------------------------------------------------
        stmp = os.path.join(tmp, "split")
        slogs, spath = run_reference(stmp, SPLIT_CFG, SPLIT_ARGV, act_dim=7)
        assert len(slogs) == 3
        sout = {"logs": np.array([[d[k] for k in LOG_KEYS] for d in slogs], dtype=np.float64)}
        spol, sdisc = checkpoint_params(spath, 2, rrm.bound_reference())
        for k, v in spol.items():
            sout["pol2_" + k] = v.numpy()
        for k, v in sdisc.items():
            sout["disc2_" + k] = v.numpy()
        spath_npz = os.path.join(ROOT, "tests", "golden", "twin_gail_dyn_ppo_split.npz")
        np.savez_compressed(spath_npz, **sout)
        print("wrote", spath_npz, "logs:\n", sout["logs"])
        rpath_npz = os.path.join(ROOT, "tests", "golden", "twin_policy_refinement.npz")
        np.savez_compressed(rpath_npz, **rout)
        print("wrote", rpath_npz, "logs:\n", rout["logs"])
        n_upd = TWIN_CFG["num_env_steps"] // TWIN_CFG["num_steps"] // TWIN_CFG["num_processes"]
        assert len(logs) == n_upd, (len(logs), n_upd)
        out["logs"] = np.array([[d[k] for k in LOG_KEYS] for d in logs], dtype=np.float64)
        for j in range(n_upd):
            pol, disc = checkpoint_params(save_path, j, rrm.bound_reference())
            for k, v in pol.items():
                out["pol%d_%s" % (j, k)] = v.numpy()
            for k, v in disc.items():
                out["disc%d_%s" % (j, k)] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "twin_gail_dyn_ppo.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "logs:\n", out["logs"])


if __name__ == "__main__":
    main()
