"""Parity at the REAL sizes of BASELINE.json configs[2], [3] and a 1/8 slice of configs[4] (cfg 2 lives in
test_gpu_parity.py::test_full_size_*): a prefix of one discriminator epoch step by step, the whole-rollout reward
relabel + GAE, and one full PPO epoch, against the CPU oracle on the same index streams.

  cfg3  Laikago 16 x 2048, O=64 A=28 H=256 F=86: 1024-row minibatches, real laika_70_deform_n200_0.pkl expert rows
  cfg4  Laikago 128 x 2048: 8192-row minibatches (several tiles per CTA and step: the accumulate code paths, 16-row tiles
        or the tensor-core tiles)
  cfg5s synthetic 512 x 1024 (1/8 of the env columns of configs[4]), O=111 (not a multiple of 4: scalar path), A=12, H=64:
        16384-row minibatches

Tolerances are the contract's: per-step losses within 1e-4 relative (scale = the column's largest magnitude), gradient
norms likewise, relabelled rewards / returns within 2e-4 of their largest magnitude, RunningMeanStd count exact."""
import os

import numpy as np
import pytest
import torch

from oracle import ppo_gail_oracle as orc

import gpu_util as gu
import simgan_b200 as sg
from golden_util import GOLDEN_DIR

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-4

SIZES = {
    "cfg3": dict(T=2048, N=16, O=64, A=28, H=256, F=86, expert="laika", ep_len=78.0),
    "cfg4": dict(T=2048, N=128, O=64, A=28, H=256, F=86, expert="laika", ep_len=78.0),
    "cfg5s": dict(T=1024, N=512, O=111, A=12, H=64, F=86, expert="synth", ep_len=78.0),
}


def _expert(kind, F):
    if kind == "laika":
        return torch.from_numpy(np.load(os.path.join(GOLDEN_DIR, "laika_expert_sas_f32.npy")))
    return torch.randn(16384, F, generator=torch.Generator().manual_seed(77))


def _workload(name, seed):
    c = SIZES[name]
    torch.manual_seed(seed)
    p = orc.init_policy(c["O"], c["H"], c["A"])
    d = orc.init_disc(c["F"], 100)
    expert = _expert(c["expert"], c["F"])
    buf = orc.synth_rollout(c["T"], c["N"], c["O"], c["A"], c["F"], p, seed=seed, ep_len=c["ep_len"], feat_bank=expert)
    return c, p, d, expert, buf


@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg5s"])
def test_full_size_disc_relabel_gae(name):
    from torch.utils.data import DataLoader, TensorDataset
    c, p, dpar, expert, buf = _workload(name, seed=11)
    T, N, O, A, F = c["T"], c["N"], c["O"], c["A"], c["F"]
    S, B, n = T * N, 128, 6
    g = torch.Generator().manual_seed(5)
    e_idx = torch.randperm(expert.shape[0], generator=g)[:n * B].view(n, B)
    p_idx = torch.randperm(S, generator=g)[:n * B].view(n, B)
    alpha = torch.rand(n, B, generator=g)
    ora = orc.DiscOracle(dpar)
    trace = []
    ora.update_epoch(expert, buf, batch_size=B, replay=(list(e_idx), list(p_idx), [a.view(B, 1) for a in alpha]), trace=trace)
    d = gu.make_disc(dpar, F, 100)
    rs = gu.make_storage(buf, O, A, F)
    loader = DataLoader(TensorDataset(expert.to(gu.DEV)), batch_size=B, shuffle=True, drop_last=True)
    d.update_gail_dyn(loader, rs, replay=(e_idx, p_idx, alpha))
    tr, tr_o = d.last_trace.double().numpy(), np.array(trace)
    assert tr.shape == tr_o.shape == (n, 3)
    assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * np.abs(tr_o)), np.abs(tr - tr_o) / np.abs(tr_o)
    # whole-rollout relabel + GAE with the updated discriminator
    o_rms = orc.RunningMeanStd(shape=())
    r_sa = orc.alive_bonus_offset(buf["masks"], T, N, expert.shape[0] / 200.0)
    orc.relabel_rewards(ora, o_rms, buf, 0.99, -r_sa)
    nv = orc.policy_forward(p, buf["obs"][-1])[0]
    orc.compute_returns(buf, nv, True, 0.99, 0.95, True)
    rms = sg.RunningMeanStd(shape=())
    d.relabel_rollout(rs, 0.99, -r_sa, rms)
    rs.compute_returns(nv.to(gu.DEV), True, 0.99, 0.95, True)
    ref = buf["rewards"]
    assert float((rs.rewards.cpu() - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    assert abs(float(rms.var) - float(o_rms.var)) <= 1e-4 * float(o_rms.var) and float(rms.count) == float(o_rms.count)
    ret_ref = buf["returns"][:-1]
    assert float((rs.returns.cpu()[:-1] - ret_ref).abs().max()) <= 2e-4 * float(ret_ref.abs().max())
    # size-independent property: GAE is exactly reproducible from the kernel's own rewards (bit-exact recurrence)
    buf2 = {k: getattr(rs, k).cpu().clone() for k in buf}
    orc.compute_returns(buf2, nv, True, 0.99, 0.95, True)
    assert torch.equal(buf2["returns"][:-1], rs.returns.cpu()[:-1])


@pytest.mark.parametrize("name,kernel_mode", [("cfg3", 0), ("cfg4", 0), ("cfg5s", 0), ("cfg4", 4), ("cfg3", 4), ("cfg5s", 4)])
def test_full_size_ppo_epoch(name, kernel_mode):
    """One PPO epoch = 32 minibatches at the config's real minibatch size, per-step losses and gradient norms vs the
    oracle.  kernel_mode 0 = automatic choice, 4 = the tcgen05 3xTF32 tensor-core tiles forced."""
    if kernel_mode == 4 and not hasattr(sg.PPO, "MMA_MODE"):
        pytest.skip("tensor-core tile path not built")
    c, p, dpar, expert, buf = _workload(name, seed=12)
    T, N, O, A, F, H = c["T"], c["N"], c["O"], c["A"], c["F"], c["H"]
    S = T * N
    buf["rewards"].copy_(torch.randn(T, N, 1, generator=torch.Generator().manual_seed(1)).clamp(-3, 3))
    nv = orc.policy_forward(p, buf["obs"][-1])[0]
    orc.compute_returns(buf, nv, True, 0.99, 0.95, True)
    hyper = orc.PPOHyper(ppo_epoch=1, num_mini_batch=32)
    ora = orc.PPOOracle(p, hyper)
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(9))
    mbs = S // 32
    trace = []
    ora.update(buf, index_chunks=[[perm[i * mbs:(i + 1) * mbs] for i in range(32)]], trace=trace)
    pol = gu.make_policy(p, O, H, A)
    agent = sg.PPO(pol, 0.2, 1, 32, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    agent.kernel_mode = kernel_mode
    rs = gu.make_storage(buf, O, A, F)
    agent.update(rs, permutations=perm.view(1, -1))
    tr, tr_o = agent.last_trace.double().numpy(), np.array(trace)
    scale = np.abs(tr_o).max(axis=0)
    scale[1] = max(scale[1], 0.5)
    assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * scale + 1e-6), np.abs(tr - tr_o).max(axis=0) / scale
    pm, po = gu.policy_params(pol), ora.params()
    for k in orc.POLICY_KEYS:
        assert torch.allclose(pm[k], po[k].reshape(-1), rtol=1e-3, atol=2e-5), k
