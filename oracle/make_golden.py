"""Generate tests/golden/* from the REAL reference (run in the dev container only).

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  Imports the unmodified reference through oracle/ref_shim.py, runs one outer
iteration's update phase (D-update x E_d -> reward relabel -> GAE -> PPO.update) exactly as
third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py:255-304 sequences it, on a seeded synthetic rollout,
and stores inputs, RNG states, intermediate and final outputs.  Also writes:
  * hopper_expert_sas_f32.npy   merged (17555,25) expert matrix of hopper_new11_deform_n200_3.pkl
                                (--gail-traj-num 200 --gail_downsample_frequency 1), fp32 as the
                                reference narrows it with Tensor(...) (main_gail_dyn_ppo.py:165)
  * laika_expert_sas_f32.npy    merged (15678,86) expert matrix of laika_70_deform_n200_0.pkl (--laika)
  * mini_expert.pkl / mini_expert_merged.npy   3-trajectory slice of that pkl in the on-disk format
                                (collect_tarsim_traj.py:261-265) + the reference's merged matrix
"""
import os
import pickle
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ppo_gail_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _policy_params(pol):
    b, d = pol.base, pol.dist
    t = dict(aw1=b.actor[0].weight, ab1=b.actor[0].bias, aw2=b.actor[2].weight, ab2=b.actor[2].bias,
             cw1=b.critic[0].weight, cb1=b.critic[0].bias, cw2=b.critic[2].weight, cb2=b.critic[2].bias,
             vw=b.critic_linear.weight, vb=b.critic_linear.bias, mw=d.fc_mean.weight, mb=d.fc_mean.bias,
             logstd=d.logstd._bias)
    return {k: v.detach().clone() for k, v in t.items()}


def run_case(ref, name, seed, T, N, O, A, H, F_, HD, expert, gail_epoch, gail_batch, ppo_epoch, nmb, ep_len):
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    pol = ref.model.Policy((O,), ref_shim.BoxSpace(A), base_kwargs={"recurrent": False, "hidden_size": H})
    agent = ref.ppo.PPO(pol, 0.2, ppo_epoch, nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    disc = ref.gail.Discriminator(F_, HD, torch.device("cpu"))
    g = {"meta_dims": np.array([T, N, O, A, H, F_, HD, gail_epoch, gail_batch, ppo_epoch, nmb, seed], dtype=np.int64)}
    p0 = _policy_params(pol)
    for k, v in p0.items():
        g["pol0_" + k] = v.numpy()
    for k, v in zip(orc.DISC_KEYS, disc.trunk.parameters()):
        g["disc0_" + k] = v.detach().numpy().copy()
    buf = orc.synth_rollout(T, N, O, A, F_, p0, seed=seed, ep_len=ep_len, feat_bank=expert)
    rs = ref.storage.RolloutStorage(T, N, (O,), ref_shim.BoxSpace(A), 1, F_)
    for k, v in buf.items():
        getattr(rs, k).copy_(v)
        g["buf_" + k] = v.numpy().copy()
    g["expert"] = expert.numpy()
    with torch.no_grad():
        next_value = pol.get_value(rs.obs[-1], rs.recurrent_hidden_states[-1], rs.masks[-1]).detach()
    g["next_value"] = next_value.numpy()

    loader = DataLoader(TensorDataset(expert), batch_size=gail_batch, shuffle=True,
                        drop_last=len(expert) > gail_batch)
    g["rng_before_disc"] = torch.get_rng_state().numpy()
    # explicit index streams, re-derived from the same generator state with the oracle's emulation
    st = torch.get_rng_state()
    e_idx, p_idx, alphas = [], [], []
    S = T * N
    for _ in range(gail_epoch):
        ei = orc.expert_loader_indices(len(expert), gail_batch, len(expert) > gail_batch)
        pi = orc.sampler_chunks(S, gail_batch)
        n = min(len(ei), len(pi))
        al = [torch.rand(gail_batch, 1) for _ in range(n)]
        e_idx.append(torch.stack(ei[:n])), p_idx.append(torch.stack(pi[:n])), alphas.append(torch.stack(al))
    g["disc_expert_idx"] = torch.stack(e_idx).numpy()
    g["disc_policy_idx"] = torch.stack(p_idx).numpy()
    g["disc_alpha"] = torch.stack(alphas).numpy()
    torch.set_rng_state(st)
    d_losses = [disc.update_gail_dyn(loader, rs) for _ in range(gail_epoch)]
    g["disc_losses"] = np.array(d_losses, dtype=np.float64)
    for k, v in zip(orc.DISC_KEYS, disc.trunk.parameters()):
        g["disc1_" + k] = v.detach().numpy().copy()

    gail_tar_length = len(expert) * 1.0 / 200 * 1
    n_done = (1.0 - rs.masks).sum().cpu().numpy() + N / 2
    d_sa = 1 - n_done / (n_done + (T * N) / gail_tar_length)
    r_sa = np.log(d_sa) - np.log(1 - d_sa)
    g["r_sa"] = np.array(r_sa, dtype=np.float64)
    g["gail_tar_length"] = np.array(gail_tar_length, dtype=np.float64)
    rms = ref.rms.RunningMeanStd(shape=())
    means, raw = [], []
    for step in range(T):
        rs.rewards[step], returns = disc.predict_reward_combined(rs.obs_feat[step + 1], 0.99, rs.masks[step],
                                                                 offset=-r_sa)
        raw.append(rs.rewards[step].clone())
        rms.update(returns.view(-1).cpu().numpy())
        rews = rs.rewards[step].view(-1).cpu().numpy()
        rews = np.clip(rews / np.sqrt(rms.var + 1e-7), -10.0, 10.0)
        rs.rewards[step] = torch.FloatTensor(rews).view(-1, 1)
        means.append(float(torch.mean(returns)))
    g["relabel_raw_reward"] = torch.stack(raw).numpy()
    g["relabel_rewards"] = rs.rewards.numpy().copy()
    g["relabel_disc_returns"] = disc.returns.numpy().copy()
    g["relabel_mean_returns"] = np.array(means, dtype=np.float64)
    g["relabel_rms"] = np.array([float(rms.mean), float(rms.var), float(rms.count)], dtype=np.float64)

    rs.compute_returns(next_value, True, 0.99, 0.95, True)
    g["gae_returns"] = rs.returns.numpy().copy()
    adv = rs.returns[:-1] - rs.value_preds[:-1]
    g["adv_mean_std"] = np.array([float(adv.mean()), float(adv.std())], dtype=np.float64)

    g["rng_before_ppo"] = torch.get_rng_state().numpy()
    st = torch.get_rng_state()
    g["ppo_perm"] = torch.stack([torch.randperm(S) for _ in range(ppo_epoch)]).numpy()
    torch.set_rng_state(st)
    g["ppo_losses"] = np.array(agent.update(rs), dtype=np.float64)
    for k, v in _policy_params(pol).items():
        g["pol1_" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name), **g)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in g.items() if k in ("ppo_losses", "disc_losses")},
          g["ppo_losses"], g["disc_losses"][-1])


def run_split_case(ref, name, seed, T, N, O, H, num_feet, ppo_epoch, nmb, entropy_coef, ep_len):
    """PPO.update of the REAL reference on its SplitPolicy (third_party/a2c_ppo_acktr/model_split.py), the policy
    class the shipped train_*.sh scripts use (--use-split-pi): inputs, index stream, losses and parameters after."""
    torch.set_num_threads(1)
    torch.manual_seed(seed)
    A = 7 * num_feet
    pol = ref.model_split.SplitPolicy((O,), ref_shim.BoxSpace(A), base_kwargs={"hidden_size": H, "num_feet": num_feet})
    agent = ref.ppo.PPO(pol, 0.2, ppo_epoch, nmb, 0.5, entropy_coef, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    g = {"meta_dims": np.array([T, N, O, A, H, num_feet, ppo_epoch, nmb, seed], dtype=np.int64),
         "entropy_coef": np.array(entropy_coef, dtype=np.float64)}
    p0 = {k: v.detach().clone() for k, v in zip(orc.SPLIT_KEYS, pol.parameters())}
    for k, v in p0.items():
        g["sp0_" + k] = v.numpy()
    # rollout contents: obs ~ N(0,1), masks ~ Bernoulli, actions/values/log-probs from the policy's own act()
    gen = torch.Generator().manual_seed(1000 + seed)
    buf = orc.new_buffer(T, N, O, A, 3)
    buf["obs"].copy_(torch.randn(T + 1, N, O, generator=gen))
    buf["masks"].copy_((torch.rand(T + 1, N, 1, generator=gen) >= 1.0 / ep_len).float())
    noise = torch.randn(T * N, A, generator=gen)
    v, a, lp = orc.split_act(p0, buf["obs"][:-1].reshape(T * N, O), noise=noise)
    buf["value_preds"][:-1] = v.view(T, N, 1)
    buf["actions"].copy_(a.view(T, N, A))
    buf["action_log_probs"].copy_(lp.view(T, N, 1))
    buf["rewards"].copy_(torch.randn(T, N, 1, generator=gen).clamp(-3, 3))
    rs = ref.storage.RolloutStorage(T, N, (O,), ref_shim.BoxSpace(A), 1, 3)
    for k, val in buf.items():
        getattr(rs, k).copy_(val)
    with torch.no_grad():
        next_value = pol.get_value(rs.obs[-1], rs.recurrent_hidden_states[-1], rs.masks[-1]).detach()
        ve, lpe, ente, _ = pol.evaluate_actions(rs.obs[3], None, None, rs.actions[3])
    g["next_value"] = next_value.numpy()
    g["eval_value"], g["eval_logp"], g["eval_entropy"] = ve.numpy(), lpe.numpy(), np.array(float(ente))
    rs.compute_returns(next_value, True, 0.99, 0.95, True)
    for k in buf:
        g["buf_" + k] = getattr(rs, k).numpy().copy()
    S = T * N
    g["rng_before_ppo"] = torch.get_rng_state().numpy()
    st = torch.get_rng_state()
    g["ppo_perm"] = torch.stack([torch.randperm(S) for _ in range(ppo_epoch)]).numpy()
    torch.set_rng_state(st)
    g["ppo_losses"] = np.array(agent.update(rs), dtype=np.float64)
    for k, val in zip(orc.SPLIT_KEYS, pol.parameters()):
        g["sp1_" + k] = val.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name), **g)
    print(name, g["ppo_losses"])


def main():
    assert ref_shim.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    torch.manual_seed(0)
    path = os.path.join(ref_shim.REF_ROOT, "hopper_new11_deform_n200_3.pkl")
    cols = orc.load_sas_wpast(path, downsample_freq=1, load_num_trajs=200)
    merged = ref.env_utils.select_and_merge_sas(cols, s_idx=np.array([0]), a_idx=np.array([0]))
    expert_full = torch.Tensor(merged)
    np.save(os.path.join(OUT, "hopper_expert_sas_f32.npy"), expert_full.numpy())
    print("expert", expert_full.shape)

    with open(path, "rb") as fh:
        trajs = pickle.load(fh)
    mini = {k: [[list(map(float, c)) for c in row] for row in trajs[k][:12]] for k in range(3)}
    with open(os.path.join(OUT, "mini_expert.pkl"), "wb") as fh:
        pickle.dump(mini, fh)
    torch.manual_seed(0)
    mcols = orc.load_sas_wpast(os.path.join(OUT, "mini_expert.pkl"), downsample_freq=2, load_num_trajs=3)
    np.save(os.path.join(OUT, "mini_expert_merged.npy"),
            ref.env_utils.select_and_merge_sas(mcols, s_idx=np.array([0]), a_idx=np.array([0])))

    # cfg-1 sizes (BASELINE.json configs[0]) with short epochs; Hopper dims, real expert rows
    run_case(ref, "hopper_cfg1_seed0.npz", 0, 128, 4, 14, 7, 64, 25, 100, expert_full[:1024].clone(),
             gail_epoch=2, gail_batch=128, ppo_epoch=2, nmb=32, ep_len=88.0)
    # ragged sizes: N, B, dims not multiples of anything; hidden 100 like the shipped scripts
    torch.manual_seed(1)
    run_case(ref, "ragged_seed1.npz", 1, 37, 3, 11, 5, 100, 19, 48, torch.randn(157, 19) * 0.7 + 0.3,
             gail_epoch=2, gail_batch=24, ppo_epoch=2, nmb=5, ep_len=7.0)
    # Laikago dims (O=64, A=28, H=256, F=86), tiny rollout
    torch.manual_seed(2)
    run_case(ref, "laika_dims_seed2.npz", 2, 24, 4, 64, 28, 256, 86, 100, torch.randn(96, 86),
             gail_epoch=1, gail_batch=32, ppo_epoch=1, nmb=3, ep_len=20.0)
    if "--all" in sys.argv or "--split" in sys.argv:
        split_cases(ref)


def laika_expert(ref):
    """Merged (15678, 86) expert matrix of laika_70_deform_n200_0.pkl (BASELINE configs[2] / [3]) through the
    reference's own select_and_merge_sas (my_pybullet_envs/utils.py:233-263), --gail-traj-num 200, downsample 1."""
    torch.manual_seed(0)
    path = os.path.join(ref_shim.REF_ROOT, "laika_70_deform_n200_0.pkl")
    cols = orc.load_sas_wpast(path, downsample_freq=1, load_num_trajs=200)
    merged = ref.env_utils.select_and_merge_sas(cols, s_idx=np.array([0]), a_idx=np.array([0]))
    expert = torch.Tensor(merged)
    np.save(os.path.join(OUT, "laika_expert_sas_f32.npy"), expert.numpy())
    print("laika expert", expert.shape)


def split_cases(ref):
    # SplitPolicy, shipped-script dims (Hopper: O=14, A=7, hidden 100, entropy_coef 0; train_hopper_deform.sh:5)
    run_split_case(ref, "split_hopper_seed3.npz", 3, 40, 4, 14, 100, 1, ppo_epoch=2, nmb=4, entropy_coef=0.0, ep_len=9.0)
    # Laikago-like: 4 feet (A=28), hidden 64, with an entropy bonus so that the state-dependent log-std heads get
    # the entropy gradient as well
    run_split_case(ref, "split_feet4_seed4.npz", 4, 24, 3, 64, 64, 4, ppo_epoch=1, nmb=3, entropy_coef=0.01, ep_len=9.0)


if __name__ == "__main__":
    if "--laika" in sys.argv:
        laika_expert(ref_shim.load())
    elif "--split" in sys.argv and "--all" not in sys.argv:
        split_cases(ref_shim.load())
    else:
        main()
