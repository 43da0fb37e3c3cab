"""CUDA hot path vs the CPU oracle and the golden vectors of the real reference.  All calls go through
the C ABI (ctypes) underneath the reference-shaped Python classes."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from golden_util import CASES, Golden
from oracle import ppo_gail_oracle as orc

pytestmark = pytest.mark.gpu

gu = None
sg = None


@pytest.fixture(autouse=True, scope="module")
def _gpu():
    global gu, sg
    assert torch.cuda.is_available(), "the -m gpu suite needs a CUDA device"
    import gpu_util
    import simgan_b200
    gu, sg = gpu_util, simgan_b200
    yield


# fp32 tolerances (north_star: bit-exact indexing/sampling, <=1e-4 rel on losses/returns)
LOSS_RTOL = 1e-4


# ------------------------------------------------------------------------------------------- returns / GAE
@pytest.mark.parametrize("T,N", [(37, 5), (128, 4), (2048, 16), (64, 300)])
@pytest.mark.parametrize("use_gae,proper", [(True, True), (True, False), (False, True), (False, False)])
def test_compute_returns_bitexact(T, N, use_gae, proper):
    torch.manual_seed(0)
    p = orc.init_policy(14, 64, 7)
    buf = orc.synth_rollout(T, N, 14, 7, 25, p, seed=T + N, ep_len=9.0)
    buf["bad_masks"][min(3, T), 0] = 0.0
    buf["rewards"].copy_(torch.randn(T, N, 1))
    nv = torch.randn(N, 1)
    rs = gu.make_storage(buf, 14, 7, 25)
    rs.compute_returns(nv.to(gu.DEV), use_gae, 0.99, 0.95, proper)
    orc.compute_returns(buf, nv, use_gae, 0.99, 0.95, proper)
    assert torch.equal(rs.returns.cpu(), buf["returns"])
    assert torch.equal(rs.value_preds.cpu(), buf["value_preds"])


@pytest.mark.parametrize("case", CASES)
def test_gae_golden_bitexact(case):
    g = Golden(case)
    buf = g.buffer()
    buf["rewards"].copy_(g.t("relabel_rewards"))
    rs = gu.make_storage(buf, g.O, g.A, g.F)
    rs.compute_returns(g.t("next_value").to(gu.DEV), True, 0.99, 0.95, True)
    assert torch.equal(rs.returns.cpu(), g.t("gae_returns"))


def test_adv_stats():
    from simgan_b200 import _lib
    g = Golden(CASES[0])
    ret = g.t("gae_returns").to(gu.DEV)
    vp = g.buffer()["value_preds"]
    vp[-1] = g.t("next_value")
    vp = vp.to(gu.DEV)
    lib = _lib.lib()
    out = torch.empty(2, device=gu.DEV)
    ws = torch.empty(int(lib.sg_adv_stats_workspace_bytes(g.S)), dtype=torch.uint8, device=gu.DEV)
    _lib.check(lib.sg_adv_stats(_lib.ptr(ret), _lib.ptr(vp), g.S, _lib.ptr(out), _lib.ptr(ws), _lib.current_stream()))
    mean, std = out.cpu().tolist()
    assert abs(mean - g.z["adv_mean_std"][0]) <= 1e-6 * max(1.0, abs(g.z["adv_mean_std"][0])) + 1e-7
    assert abs(std - g.z["adv_mean_std"][1]) <= 1e-6 * g.z["adv_mean_std"][1]


# ------------------------------------------------------------------------------------------------- sampler
def test_feed_forward_generator_bitexact():
    g = Golden("ragged_seed1.npz")
    buf = g.buffer()
    rs = gu.make_storage(buf, g.O, g.A, g.F)
    adv = torch.randn(g.T, g.N, 1)
    torch.manual_seed(11)
    mine = list(rs.feed_forward_generator(adv.to(gu.DEV), num_mini_batch=5))
    torch.manual_seed(11)
    ref = list(orc.feed_forward_batches(buf, adv, num_mini_batch=5))
    assert len(mine) == len(ref) == 5
    for mb, rb in zip(mine, ref):
        assert len(mb) == len(rb) == 10
        for x, y in zip(mb, rb):
            assert torch.equal(x.cpu(), y)
    torch.manual_seed(12)
    mine = list(rs.feed_forward_generator(None, mini_batch_size=24))
    torch.manual_seed(12)
    ref = list(orc.feed_forward_batches(buf, None, mini_batch_size=24))
    assert len(mine) == len(ref) and mine[0][7] is None
    assert torch.equal(mine[-1][-1].cpu(), ref[-1][-1])
    # RNG stream left where the reference leaves it
    after = torch.rand(3)
    torch.manual_seed(12)
    torch.randperm(g.S)
    assert torch.equal(torch.rand(3), after)


def test_insert_and_after_update():
    T, N, O, A, F = 3, 2, 14, 7, 25
    from oracle.ref_shim import BoxSpace
    rs = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F)
    rs.to(gu.DEV)
    buf = orc.new_buffer(T, N, O, A, F)
    step = 0
    gen = torch.Generator().manual_seed(0)
    for _ in range(4):
        args = [torch.randn(N, O, generator=gen), torch.zeros(N, 1), torch.randn(N, A, generator=gen),
                torch.randn(N, 1, generator=gen), torch.randn(N, 1, generator=gen), torch.randn(N, 1, generator=gen),
                (torch.rand(N, 1, generator=gen) > 0.3).float(), torch.ones(N, 1), torch.randn(N, F, generator=gen)]
        rs.insert(*[a.to(gu.DEV) for a in args])
        step = orc.buffer_insert(buf, step, *args)
        assert rs.step == step
    rs.after_update()
    orc.buffer_after_update(buf)
    for k, v in buf.items():
        assert torch.equal(getattr(rs, k).cpu(), v), k


# --------------------------------------------------------------------------------------------- actor-critic
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("B", [1, 16, 77])
def test_policy_forward(case, B):
    g = Golden(case)
    p = g.policy()
    pol = gu.make_policy(p, g.O, g.H, g.A)
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, g.O, generator=gen)
    act = torch.randn(B, g.A, generator=gen)
    v_o, lp_o, ent_o = orc.policy_evaluate(p, x, act)
    with torch.no_grad():
        v, lp, ent, _ = pol.evaluate_actions(x.to(gu.DEV), None, None, act.to(gu.DEV))
    assert gu.rel_err(v.cpu(), v_o, floor=0.5) < 2e-6      # value = 64-term dot product of O(1) terms
    assert gu.rel_err(lp.cpu(), lp_o) < 2e-6
    assert abs(float(ent) - float(ent_o)) < 2e-6 * abs(float(ent_o))
    assert gu.rel_err(pol.get_value(x.to(gu.DEV), None, None).cpu(), v_o, floor=0.5) < 2e-6
    # deterministic act = mean
    v_d, a_d, lp_d = orc.policy_act(p, x, deterministic=True)
    v2, a2, lp2, _ = pol.act(x.to(gu.DEV), None, None, deterministic=True)
    assert gu.rel_err(a2.cpu(), a_d, floor=0.05) < 2e-6 and gu.rel_err(lp2.cpu(), lp_d) < 2e-6
    # sampled act: same CUDA-generator stream as torch.normal(mean, std) in the reference's Normal.sample()
    torch.cuda.manual_seed(5)
    v3, a3, lp3, _ = pol.act(x.to(gu.DEV), None, None)
    torch.cuda.manual_seed(5)
    mean = a2
    std = pol.dist.logstd._bias.detach().t().exp().expand_as(mean)
    a_ref = torch.normal(mean, std)
    assert gu.rel_err(a3, a_ref) < 2e-6
    _, mean_o, logstd_o = orc.policy_forward(p, x)
    lp_chk = orc.gaussian_logp_entropy(mean_o, logstd_o, a3.cpu())[0]
    assert gu.rel_err(lp3.cpu(), lp_chk) < 5e-6


# ------------------------------------------------------------------------------------------------------ PPO
def _ppo_objects(g, mode):
    pol = gu.make_policy(g.policy(), g.O, g.H, g.A)
    agent = sg.PPO(pol, 0.2, g.ppo_epoch, g.nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    agent.kernel_mode = mode
    buf = g.buffer()
    buf["rewards"].copy_(g.t("relabel_rewards"))
    buf["returns"].copy_(g.t("gae_returns"))
    buf["value_preds"][-1] = g.t("next_value")
    rs = gu.make_storage(buf, g.O, g.A, g.F)
    return pol, agent, rs, buf


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", [1, 2, 0, 4])
def test_ppo_per_step_trace_vs_oracle(case, mode):
    """Every optimizer step's (value_loss, action_loss, entropy, grad_norm) against the oracle replaying the
    same recorded index stream; parameters after the whole update."""
    g = Golden(case)
    if mode == 4 and g.H not in (64, 128, 256):
        pytest.skip("the tensor-core tiles need hidden in {64,128,256}")
    pol, agent, rs, buf = _ppo_objects(g, mode)
    ora = orc.PPOOracle(g.policy(), g.hyper())
    trace = []
    out_o = ora.update(buf, index_chunks=g.ppo_chunks(), trace=trace)
    out = agent.update(rs, permutations=g.t("ppo_perm"))
    tr = agent.last_trace.double().numpy()
    tr_o = np.array(trace)
    scale = np.abs(tr_o).max(axis=0)
    scale[1] = max(scale[1], 0.5)      # action loss = mean of cancelling ratio*adv terms of magnitude E|adv| ~ 0.8
    # first step: identical state on both sides -> tight
    assert np.all(np.abs(tr[0] - tr_o[0]) <= 5e-6 * scale + 1e-7), (tr[0], tr_o[0])
    # all steps (fp32 reassociation compounds through Adam)
    assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * scale + 1e-6), np.abs(tr - tr_o).max(axis=0) / scale
    for i in range(3):
        assert abs(out[i] - out_o[i]) <= LOSS_RTOL * max(abs(out_o[i]), scale[i])
    po = ora.params()
    pm = gu.policy_params(pol)
    for k in orc.POLICY_KEYS:
        assert torch.allclose(pm[k], po[k].reshape(-1), rtol=1e-3, atol=2e-5), k


@pytest.mark.parametrize("case", CASES)
def test_ppo_update_golden_rng(case):
    """Drop-in call: PPO.update draws its own permutations from the CPU generator; compare with the losses
    the REAL reference produced from the same generator state."""
    g = Golden(case)
    pol, agent, rs, _ = _ppo_objects(g, 0)
    torch.set_rng_state(g.t("rng_before_ppo"))
    out = agent.update(rs)
    ref = g.z["ppo_losses"]
    assert abs(out[0] - ref[0]) <= LOSS_RTOL * abs(ref[0])
    assert abs(out[2] - ref[2]) <= LOSS_RTOL * abs(ref[2])
    assert abs(out[1] - ref[1]) <= LOSS_RTOL * max(abs(ref[1]), 0.05)   # action loss hovers around zero
    pm = gu.policy_params(pol)
    for k, v in g.policy("pol1").items():
        assert torch.allclose(pm[k], v.reshape(-1), rtol=1e-3, atol=2e-5), k
    # the generator was consumed exactly as the reference consumes it
    torch.set_rng_state(g.t("rng_before_ppo"))
    for _ in range(g.ppo_epoch):
        torch.randperm(g.S)
    expect = torch.rand(4)
    pol2, agent2, rs2, _ = _ppo_objects(g, 0)      # construction itself draws from the CPU generator
    torch.set_rng_state(g.t("rng_before_ppo"))
    agent2.update(rs2)
    assert torch.equal(torch.rand(4), expect)


def test_ppo_modes_agree():
    """phased (1) and persistent (2) kernels run the same arithmetic in the same order: bit-identical.  The
    resident kernel (0 = auto at these sizes) uses the column-owner tile, i.e. another fp32 summation order."""
    g = Golden(CASES[0])
    outs = []
    for mode in (1, 2, 0):
        pol, agent, rs, _ = _ppo_objects(g, mode)
        agent.update(rs, permutations=g.t("ppo_perm"))
        outs.append((agent.last_trace.clone(), pol.flat_params().cpu().clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    scale = outs[0][0].abs().max(dim=0).values.clamp_min(0.5)
    assert bool(((outs[2][0] - outs[0][0]).abs() <= 2e-5 * scale).all())
    assert torch.allclose(outs[2][1], outs[0][1], rtol=1e-4, atol=2e-6)


def test_ppo_lr_schedule_and_step_count():
    from simgan_b200.utils import update_linear_schedule
    g = Golden(CASES[0])
    pol, agent, rs, buf = _ppo_objects(g, 0)
    ora = orc.PPOOracle(g.policy(), g.hyper())
    for j in range(2):
        update_linear_schedule(agent.optimizer, j, 4, 3e-4)
        update_linear_schedule(ora.optimizer, j, 4, 3e-4)
        out = agent.update(rs, permutations=g.t("ppo_perm"))
        out_o = ora.update(buf, index_chunks=g.ppo_chunks())
        assert abs(out[0] - out_o[0]) <= 2e-4 * abs(out_o[0])
    assert agent.optimizer.step_count == 2 * g.ppo_epoch * g.nmb


# ------------------------------------------------------------------------------------------- discriminator
def _disc_objects(g, mode):
    from torch.utils.data import DataLoader, TensorDataset
    d = gu.make_disc(g.disc(), g.F, g.HD)
    d.kernel_mode = mode
    buf = g.buffer()
    rs = gu.make_storage(buf, g.O, g.A, g.F)
    expert = g.t("expert")
    loader = DataLoader(TensorDataset(expert.to(gu.DEV)), batch_size=g.gail_batch, shuffle=True,
                        drop_last=len(expert) > g.gail_batch)
    return d, rs, buf, expert, loader


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", [1, 2, 0])
def test_disc_update_vs_oracle(case, mode):
    g = Golden(case)
    d, rs, buf, expert, loader = _disc_objects(g, mode)
    ora = orc.DiscOracle(g.disc())
    for e in range(g.gail_epoch):
        trace = []
        out_o = ora.update_epoch(expert, buf, batch_size=g.gail_batch, replay=g.disc_replay(e), trace=trace)
        out = d.update_gail_dyn(loader, rs, replay=g.disc_replay(e))
        tr, tr_o = d.last_trace.double().numpy(), np.array(trace)
        if e == 0:
            assert np.all(np.abs(tr[0] - tr_o[0]) <= 1e-5 * np.abs(tr_o[0])), (tr[0], tr_o[0])
        assert np.all(np.abs(tr - tr_o) <= 2e-4 * np.abs(tr_o) + 1e-6), (np.abs(tr - tr_o) / np.abs(tr_o)).max(axis=0)
        for i in range(3):
            assert abs(out[i] - out_o[i]) <= LOSS_RTOL * abs(out_o[i])
    pm, po = gu.disc_params(d), ora.params()
    for k in orc.DISC_KEYS:
        assert torch.allclose(pm[k], po[k].reshape(-1), rtol=2e-3, atol=5e-5), k


@pytest.mark.parametrize("case", CASES)
def test_disc_update_golden_rng(case):
    """Drop-in call with the caller's DataLoader: index streams and alphas come from the CPU generator."""
    g = Golden(case)
    d, rs, buf, expert, loader = _disc_objects(g, 0)
    torch.set_rng_state(g.t("rng_before_disc"))
    outs = [d.update_gail_dyn(loader, rs) for _ in range(g.gail_epoch)]
    ref = g.z["disc_losses"]
    assert np.all(np.abs(np.array(outs) - ref) <= LOSS_RTOL * np.abs(ref)), (outs, ref)
    after = torch.rand(4)
    # reference consumption of the generator: replay with the real DataLoader + oracle sampler
    torch.set_rng_state(g.t("rng_before_disc"))
    for _ in range(g.gail_epoch):
        n = 0
        pol_chunks = None
        for eb in loader:
            if pol_chunks is None:
                pol_chunks = orc.sampler_chunks(g.S, g.gail_batch)
            if n >= len(pol_chunks):
                break
            torch.rand(g.gail_batch, 1)
            n += 1
    assert torch.equal(torch.rand(4), after)
    pm = gu.disc_params(d)
    for k, v in g.disc("disc1").items():
        assert torch.allclose(pm[k], v.reshape(-1), rtol=2e-3, atol=5e-5), k


def test_disc_modes_agree():
    """phased (1) and persistent (2) are bit-identical; the register-resident kernel (0 = auto for the
    instantiated hidden widths) sums in another order."""
    g = Golden(CASES[0])
    outs = []
    for mode in (1, 2, 0):
        d, rs, buf, expert, loader = _disc_objects(g, mode)
        d.update_gail_dyn(loader, rs, replay=g.disc_replay(0))
        outs.append((d.last_trace.clone(), d.flat_params().cpu().clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.allclose(outs[2][0], outs[0][0], rtol=2e-5, atol=1e-6)
    assert torch.allclose(outs[2][1], outs[0][1], rtol=1e-4, atol=2e-6)


# ------------------------------------------------------------------------------------------ reward relabel
@pytest.mark.parametrize("case", CASES)
def test_relabel_normalize_bitexact(case):
    """Everything after the discriminator forward is integer-like bookkeeping in fixed arithmetic: feed the
    reference's own raw rewards and require bit-identical rewards, returns and float64 RMS state."""
    from simgan_b200 import _lib
    g = Golden(case)
    T, N = g.T, g.N
    lib = _lib.lib()
    raw = g.t("relabel_raw_reward").reshape(T, N).contiguous().to(gu.DEV)
    masks = g.buffer()["masks"].to(gu.DEV)
    rewards = torch.empty(T, N, 1, device=gu.DEV)
    dret = torch.zeros(N, 1, device=gu.DEV)
    rms = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device=gu.DEV)
    mret = torch.empty(T, device=gu.DEV)
    ws = torch.empty(int(lib.sg_relabel_workspace_bytes(T, N)), dtype=torch.uint8, device=gu.DEV)
    _lib.check(lib.sg_relabel_normalize(_lib.ptr(raw), _lib.ptr(masks), _lib.ptr(rewards), T, N, 0.99, _lib.ptr(dret), 0,
                                        _lib.ptr(rms), _lib.ptr(mret), _lib.ptr(ws), _lib.current_stream()))
    assert torch.equal(rewards.cpu(), g.t("relabel_rewards"))
    assert torch.equal(dret.cpu(), g.t("relabel_disc_returns"))
    assert np.array_equal(rms.cpu().numpy(), g.z["relabel_rms"])
    assert np.allclose(mret.cpu().double().numpy(), g.z["relabel_mean_returns"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", CASES)
def test_relabel_rollout_vs_golden(case):
    """Full relabel incl. the discriminator forward (fp32 transcendental differences allowed)."""
    g = Golden(case)
    d = gu.make_disc(g.disc("disc1"), g.F, g.HD)
    rs = gu.make_storage(g.buffer(), g.O, g.A, g.F)
    rms = sg.RunningMeanStd(shape=())
    r_sa = float(g.z["r_sa"])
    from simgan_b200.algo.gail import alive_bonus_offset
    assert alive_bonus_offset(rs.masks, g.T, g.N, float(g.z["gail_tar_length"])) == r_sa
    mret = d.relabel_rollout(rs, 0.99, -r_sa, rms)
    ref = g.t("relabel_rewards")
    assert float((rs.rewards.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    assert gu.rel_err(d.returns.cpu(), g.t("relabel_disc_returns")) <= 1e-4
    assert abs(float(rms.var) - g.z["relabel_rms"][1]) <= 1e-4 * g.z["relabel_rms"][1]
    assert float(rms.count) == g.z["relabel_rms"][2]
    # per-step API gives the same stream (slow path used by an unmodified main loop)
    d2 = gu.make_disc(g.disc("disc1"), g.F, g.HD)
    rs2 = gu.make_storage(g.buffer(), g.O, g.A, g.F)
    for t in range(min(g.T, 6)):
        rew, ret = d2.predict_reward_combined(rs2.obs_feat[t + 1], 0.99, rs2.masks[t], offset=-r_sa)
        assert gu.rel_err(rew.cpu(), g.t("relabel_raw_reward")[t]) <= 1e-4
    # second pass keeps state (returns / rms persist across iterations)
    d.relabel_rollout(rs, 0.99, -r_sa, rms)
    assert float(rms.count) == g.z["relabel_rms"][2] + g.S


# ---------------------------------------------------------------------------------------- whole update phase
@pytest.mark.parametrize("case", CASES)
def test_full_update_phase_dropin(case):
    """D-update x E -> relabel -> GAE -> PPO.update through the reference-shaped classes, consuming the CPU
    generator from the recorded state, against the REAL reference's outputs (golden)."""
    from torch.utils.data import DataLoader, TensorDataset
    g = Golden(case)
    pol = gu.make_policy(g.policy(), g.O, g.H, g.A)
    agent = sg.PPO(pol, 0.2, g.ppo_epoch, g.nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    d = gu.make_disc(g.disc(), g.F, g.HD)
    rs = gu.make_storage(g.buffer(), g.O, g.A, g.F)
    expert = g.t("expert").to(gu.DEV)
    loader = DataLoader(TensorDataset(expert), batch_size=g.gail_batch, shuffle=True, drop_last=len(expert) > g.gail_batch)
    with torch.no_grad():
        next_value = pol.get_value(rs.obs[-1], rs.recurrent_hidden_states[-1], rs.masks[-1]).detach()
    assert gu.rel_err(next_value.cpu(), g.t("next_value")) < 2e-6
    torch.set_rng_state(g.t("rng_before_disc"))
    for _ in range(g.gail_epoch):
        dl = d.update_gail_dyn(loader, rs)
    assert np.all(np.abs(np.array(dl) - g.z["disc_losses"][-1]) <= LOSS_RTOL * np.abs(g.z["disc_losses"][-1]))
    rms = sg.RunningMeanStd(shape=())
    d.relabel_rollout(rs, 0.99, -float(g.z["r_sa"]), rms)
    rs.compute_returns(next_value, True, 0.99, 0.95, True)
    ref_ret = g.t("gae_returns")
    assert float((rs.returns.cpu() - ref_ret)[:-1].abs().max()) <= 2e-4 * float(ref_ret.abs().max())
    out = agent.update(rs)
    ref = g.z["ppo_losses"]
    assert abs(out[0] - ref[0]) <= 5e-4 * abs(ref[0])
    assert abs(out[2] - ref[2]) <= 1e-4 * abs(ref[2])
    rs.after_update()
    assert torch.equal(rs.obs[0], rs.obs[-1])


# ------------------------------------------------------------------------------ full BASELINE sizes (cfg 2)
def _cfg2_workload(seed=3):
    """T=2048, N=16 Hopper sizes: synthetic rollout from the oracle's generator + the real expert rows."""
    import os
    from golden_util import GOLDEN_DIR
    torch.manual_seed(seed)
    p = orc.init_policy(14, 64, 7)
    d = orc.init_disc(25, 100)
    expert = torch.from_numpy(np.load(os.path.join(GOLDEN_DIR, "hopper_expert_sas_f32.npy")))
    buf = orc.synth_rollout(2048, 16, 14, 7, 25, p, seed=seed, feat_bank=expert)
    return p, d, expert, buf


def test_full_size_disc_steps_and_relabel_vs_oracle():
    """cfg 2 sizes: the first 6 discriminator minibatches of an epoch step by step, then the whole-rollout relabel
    (32 768 rows) and GAE, against the oracle run on the same streams."""
    from torch.utils.data import DataLoader, TensorDataset
    p, dpar, expert, buf = _cfg2_workload()
    S, B, n = 2048 * 16, 128, 6
    g = torch.Generator().manual_seed(5)
    e_idx = torch.randperm(expert.shape[0], generator=g)[:n * B].view(n, B)
    p_idx = torch.randperm(S, generator=g)[:n * B].view(n, B)
    alpha = torch.rand(n, B, generator=g)
    ora = orc.DiscOracle(dpar)
    trace = []
    ora.update_epoch(expert, buf, batch_size=B, replay=(list(e_idx), list(p_idx), [a.view(B, 1) for a in alpha]), trace=trace)
    d = gu.make_disc(dpar, 25, 100)
    rs = gu.make_storage(buf, 14, 7, 25)
    loader = DataLoader(TensorDataset(expert.to(gu.DEV)), batch_size=B, shuffle=True, drop_last=True)
    d.update_gail_dyn(loader, rs, replay=(e_idx, p_idx, alpha))
    tr, tr_o = d.last_trace.double().numpy(), np.array(trace)
    assert tr.shape == tr_o.shape == (n, 3)
    assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * np.abs(tr_o)), np.abs(tr - tr_o) / np.abs(tr_o)
    # relabel + GAE over the full buffer with the updated discriminator
    o_rms = orc.RunningMeanStd(shape=())
    r_sa = orc.alive_bonus_offset(buf["masks"], 2048, 16, 87.8)
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


def test_full_size_ppo_epoch_vs_oracle():
    """cfg 2 sizes: one PPO epoch = 32 minibatches of 1024 rows, per-step losses and gradient norms vs the oracle."""
    p, dpar, expert, buf = _cfg2_workload(seed=4)
    buf["rewards"].copy_(torch.randn(2048, 16, 1, generator=torch.Generator().manual_seed(1)).clamp(-3, 3))
    nv = orc.policy_forward(p, buf["obs"][-1])[0]
    orc.compute_returns(buf, nv, True, 0.99, 0.95, True)
    hyper = orc.PPOHyper(ppo_epoch=1, num_mini_batch=32)
    ora = orc.PPOOracle(p, hyper)
    perm = torch.randperm(2048 * 16, generator=torch.Generator().manual_seed(9))
    trace = []
    ora.update(buf, index_chunks=[[perm[i * 1024:(i + 1) * 1024] for i in range(32)]], trace=trace)
    pol = gu.make_policy(p, 14, 64, 7)
    agent = sg.PPO(pol, 0.2, 1, 32, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    rs = gu.make_storage(buf, 14, 7, 25)
    agent.update(rs, permutations=perm.view(1, -1))
    tr, tr_o = agent.last_trace.double().numpy(), np.array(trace)
    scale = np.abs(tr_o).max(axis=0)
    scale[1] = max(scale[1], 0.5)
    assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * scale + 1e-6), np.abs(tr - tr_o).max(axis=0) / scale
    pm, po = gu.policy_params(pol), ora.params()
    for k in orc.POLICY_KEYS:
        assert torch.allclose(pm[k], po[k].reshape(-1), rtol=1e-3, atol=2e-5), k


# --------------------------------------------------------------------------------------------- edge cases
def test_ragged_minibatch_tail_and_single_env():
    """S not divisible by num_mini_batch (tail rows dropped like BatchSampler(drop_last=True)), rows-per-minibatch
    not a multiple of the 8-row tile, a single env column."""
    torch.manual_seed(0)
    T, N, O, A, H = 37, 1, 5, 2, 16
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, 3, p, seed=2, ep_len=7.0)
    nv = orc.policy_forward(p, buf["obs"][-1])[0]
    orc.compute_returns(buf, nv, True, 0.99, 0.95, True)
    hyper = orc.PPOHyper(ppo_epoch=2, num_mini_batch=5)              # 37 // 5 = 7 rows, 2 dropped per epoch
    ora = orc.PPOOracle(p, hyper)
    pol = gu.make_policy(p, O, H, A)
    agent = sg.PPO(pol, 0.2, 2, 5, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    rs = gu.make_storage(buf, O, A, 3)
    torch.manual_seed(42)
    out_o = ora.update(buf)
    torch.manual_seed(42)
    out = agent.update(rs)
    assert abs(out[0] - out_o[0]) <= 2e-4 * abs(out_o[0]) and abs(out[2] - out_o[2]) <= 1e-4 * abs(out_o[2])
    assert agent.last_trace.shape == (10, 4)


def test_num_mini_batch_larger_than_samples_raises():
    pol = gu.make_policy(orc.init_policy(5, 16, 2), 5, 16, 2)
    agent = sg.PPO(pol, 0.2, 1, 64, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    buf = orc.synth_rollout(4, 2, 5, 2, 3, orc.init_policy(5, 16, 2), seed=0)
    rs = gu.make_storage(buf, 5, 2, 3)
    with pytest.raises(AssertionError, match="PPO requires"):
        agent.update(rs)


def test_disc_epoch_bounded_by_shorter_stream_and_small_rollout():
    """zip(expert_loader, rollout sampler) stops at the shorter side (gail.py:162-163); a rollout smaller than one
    discriminator batch yields no step and the reference's division by n=0."""
    from torch.utils.data import DataLoader, TensorDataset
    torch.manual_seed(0)
    dpar = orc.init_disc(9, 48)
    expert = torch.randn(70, 9)
    buf = orc.synth_rollout(25, 2, 4, 2, 9, orc.init_policy(4, 16, 2), seed=1)       # S = 50 -> 3 batches of 16
    d = gu.make_disc(dpar, 9, 48)
    rs = gu.make_storage(buf, 4, 2, 9)
    loader = DataLoader(TensorDataset(expert.to(gu.DEV)), batch_size=16, shuffle=True, drop_last=True)   # 4 expert batches
    ora = orc.DiscOracle(dpar)
    torch.manual_seed(7)
    out_o = ora.update_epoch(expert, buf, batch_size=16)
    torch.manual_seed(7)
    out = d.update_gail_dyn(loader, rs)
    assert d.last_trace.shape[0] == 3
    assert all(abs(a - b) <= LOSS_RTOL * abs(b) for a, b in zip(out, out_o))
    loader_big = DataLoader(TensorDataset(expert.to(gu.DEV)), batch_size=64, shuffle=True, drop_last=True)
    with pytest.raises(ZeroDivisionError):
        d.update_gail_dyn(loader_big, rs)


# ----------------------------------------------------------------------------------------- rollout feed step
def test_rollout_feeder_matches_act_insert():
    """One fused launch per env step (RolloutFeeder) == Policy.act + RolloutStorage.insert of the reference-shaped
    loop (main_gail_dyn_ppo.py:209-236), bit for bit, including wrap-around and the value of the last observation."""
    from oracle.ref_shim import BoxSpace
    T, N, O, A, F, H = 6, 5, 14, 7, 25, 64
    torch.manual_seed(0)
    p = orc.init_policy(O, H, A)
    pol = gu.make_policy(p, O, H, A)
    rs_a = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F); rs_a.to(gu.DEV)
    rs_b = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F); rs_b.to(gu.DEV)
    rng = np.random.RandomState(0)
    obs0 = rng.randn(N, O).astype(np.float32)
    rs_a.obs[0].copy_(torch.from_numpy(obs0)); rs_b.obs[0].copy_(torch.from_numpy(obs0))
    env = [(rng.randn(N, O).astype(np.float32), rng.randn(N).astype(np.float32), rng.rand(N) < 0.3, rng.rand(N) < 0.1,
            rng.randn(N, F)) for _ in range(T)]
    # reference-shaped loop
    torch.cuda.manual_seed(11)
    acts_a = []
    for step in range(T):
        with torch.no_grad():
            value, action, logp, hxs = pol.act(rs_a.obs[step], rs_a.recurrent_hidden_states[step], rs_a.masks[step])
        acts_a.append(action.cpu().numpy())
        obs, rew, done, bad, feat = env[step]
        masks = torch.tensor([[0.0] if d else [1.0] for d in done])
        bad_masks = torch.tensor([[0.0] if b else [1.0] for b in bad])
        rs_a.insert(torch.from_numpy(obs).to(gu.DEV), hxs, action, logp, value, torch.from_numpy(rew).unsqueeze(1),
                    masks, bad_masks, torch.Tensor(feat))
    # fused feed
    torch.cuda.manual_seed(11)
    feeder = sg.RolloutFeeder(pol, rs_b)
    acts_b = [feeder.begin().copy()]
    for step in range(T):
        obs, rew, done, bad, feat = env[step]
        acts_b.append(feeder.step(obs, rew, done, bad, feat).copy())
    for a, b in zip(acts_a, acts_b):
        assert np.array_equal(a, b)
    for k in ("obs", "obs_feat", "recurrent_hidden_states", "rewards", "actions", "action_log_probs", "masks", "bad_masks"):
        assert torch.equal(getattr(rs_a, k), getattr(rs_b, k)), k
    assert torch.equal(rs_a.value_preds[:-1], rs_b.value_preds[:-1])
    with torch.no_grad():
        nv = pol.get_value(rs_a.obs[-1], None, None)
    assert torch.equal(rs_b.value_preds[-1], nv)             # slot T already holds next_value
    assert rs_b.step == rs_a.step == 0


def test_rollout_feeder_vs_oracle_insert_and_act():
    """The fused feed step against the ORACLE's buffer_insert + policy_act (A2C/storage.py:70-84, A2C/model.py:89-101) on
    the same env outputs and the same CUDA-generator noise stream: everything the step copies is bit-exact, what the
    policy computes (actions, log-probs, values) agrees to fp32 forward tolerance (2e-6 of the largest magnitude)."""
    from oracle.ref_shim import BoxSpace
    T, N, O, A, F, H = 7, 6, 14, 7, 25, 64
    torch.manual_seed(1)
    p = orc.init_policy(O, H, A)
    pol = gu.make_policy(p, O, H, A)
    rs = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F); rs.to(gu.DEV)
    buf = orc.new_buffer(T, N, O, A, F)
    rng = np.random.RandomState(3)
    obs0 = rng.randn(N, O).astype(np.float32)
    rs.obs[0].copy_(torch.from_numpy(obs0)); buf["obs"][0].copy_(torch.from_numpy(obs0))
    env = [(rng.randn(N, O).astype(np.float32), rng.randn(N).astype(np.float32), rng.rand(N) < 0.3, rng.rand(N) < 0.1,
            rng.randn(N, F)) for _ in range(T)]
    # the noise the feeder will draw: one randn(N, A) on the CUDA generator per act
    torch.cuda.manual_seed(21)
    noise = [torch.randn(N, A, device=gu.DEV).cpu() for _ in range(T + 1)]
    # oracle loop (main_gail_dyn_ppo.py:209-236)
    step = 0
    acts_o = []
    for t in range(T):
        value, action, logp = orc.policy_act(p, buf["obs"][step], noise=noise[t])
        acts_o.append(action)
        obs, rew, done, bad, feat = env[t]
        masks = torch.tensor([[0.0] if d else [1.0] for d in done])
        bad_masks = torch.tensor([[0.0] if b else [1.0] for b in bad])
        step = orc.buffer_insert(buf, step, torch.from_numpy(obs), torch.zeros(N, 1), action, logp, value,
                                 torch.from_numpy(rew).unsqueeze(1), masks, bad_masks, torch.Tensor(feat))
    nv_o = orc.policy_forward(p, buf["obs"][-1])[0]
    # fused feed
    torch.cuda.manual_seed(21)
    feeder = sg.RolloutFeeder(pol, rs)
    acts = [torch.from_numpy(feeder.begin().copy())]
    for t in range(T):
        obs, rew, done, bad, feat = env[t]
        acts.append(torch.from_numpy(feeder.step(obs, rew, done, bad, feat).copy()))
    assert rs.step == step == 0
    for k in ("obs", "obs_feat", "recurrent_hidden_states", "rewards", "masks", "bad_masks"):
        assert torch.equal(getattr(rs, k).cpu(), buf[k]), k
    for a, b in zip(acts, acts_o):
        assert gu.rel_err(a, b) <= 2e-6
    for k in ("actions", "action_log_probs"):
        assert gu.rel_err(getattr(rs, k).cpu(), buf[k]) <= 2e-6, k
    assert gu.rel_err(rs.value_preds[:-1].cpu(), buf["value_preds"][:-1]) <= 2e-6
    assert gu.rel_err(rs.value_preds[-1].cpu(), nv_o) <= 2e-6        # slot T already holds next_value


# ------------------------------------------------------------------------------------ remaining API surface
def test_legacy_discriminator_update_and_mod_reward():
    """Discriminator.update (the (state, action)-split form, gail.py:91-152) maps onto the same kernel over
    concatenated rows; RolloutStorage.mod_reward (storage.py:86-94) adds an offset to the last slots."""
    from torch.utils.data import DataLoader, TensorDataset
    from oracle.ref_shim import BoxSpace
    torch.manual_seed(0)
    T, N, O, A = 20, 4, 6, 3
    p = orc.init_policy(O, 16, A)
    buf = orc.synth_rollout(T, N, O, A, 2, p, seed=5)
    rs = gu.make_storage(buf, O, A, 2)
    dpar = orc.init_disc(O + A, 48)
    d = gu.make_disc(dpar, O + A, 48)
    es, ea = torch.randn(64, O), torch.randn(64, A)
    loader = DataLoader(TensorDataset(es.to(gu.DEV), ea.to(gu.DEV)), batch_size=16, shuffle=True, drop_last=True)
    # oracle: the same update on the concatenated matrices with the same streams
    expert = torch.cat([es, ea], 1)
    fl = orc.flat_views(buf)
    pol_rows = torch.cat([fl["obs"], fl["actions"]], 1)
    torch.manual_seed(3)
    e_idx, p_idx, alpha = sg.Discriminator.draw_epoch_indices(64, 16, True, T * N)
    ora = orc.DiscOracle(dpar)
    tot = [0.0, 0.0, 0.0]
    for i in range(e_idx.shape[0]):
        out = ora.step(expert[e_idx[i]], pol_rows[p_idx[i]], alpha[i].view(-1, 1))
        tot = [a + b for a, b in zip(tot, out)]
    torch.manual_seed(3)
    got = d.update(loader, rs)
    assert all(abs(a - b / e_idx.shape[0]) <= 1e-4 * abs(b / e_idx.shape[0]) for a, b in zip(got, tot))
    before = rs.rewards.clone()
    rs.step = 3
    rs.mod_reward(torch.full((N,), 0.5, device=gu.DEV), 2)
    assert torch.equal(rs.rewards[2], before[2] + 0.5) and torch.equal(rs.rewards[1], before[1] + 0.5)
    assert torch.equal(rs.rewards[0], before[0]) and torch.equal(rs.rewards[3], before[3])


def test_disc_hidden_width_without_register_kernel():
    """Hidden widths other than 48/64/100/128 take the shared-memory-resident generic tile (disc_persistent_kernel<true>);
    odd feature / hidden sizes exercise the scalar (non-float4) micro-kernel paths."""
    from torch.utils.data import DataLoader, TensorDataset
    for F_, HD in ((9, 40), (7, 30)):
        torch.manual_seed(F_)
        dpar = orc.init_disc(F_, HD)
        expert = torch.randn(80, F_)
        buf = orc.synth_rollout(30, 2, 4, 2, F_, orc.init_policy(4, 16, 2), seed=F_)
        loader = DataLoader(TensorDataset(expert.to(gu.DEV)), batch_size=16, shuffle=True, drop_last=True)
        outs = []
        for mode in (0, 1):
            d = gu.make_disc(dpar, F_, HD)
            d.kernel_mode = mode
            rs = gu.make_storage(buf, 4, 2, F_)
            torch.manual_seed(21)
            outs.append((d.update_gail_dyn(loader, rs), d.last_trace.clone(), d.flat_params().cpu().clone()))
        ora = orc.DiscOracle(dpar)
        torch.manual_seed(21)
        out_o = ora.update_epoch(expert, buf, batch_size=16)
        for got in outs:
            assert all(abs(a - b) <= LOSS_RTOL * abs(b) for a, b in zip(got[0], out_o)), (got[0], out_o)
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])     # same tile code: bit-identical


@pytest.mark.parametrize("perturb", [False, True])
def test_early_draws_do_not_change_results(perturb):
    """Three outer iterations (D x 3 -> relabel -> GAE -> PPO) with the early index draws of simgan_b200/_spec.py on
    and off: identical parameters, per-step traces and CPU generator state; the early path must actually have been
    taken (staged block used) in the steady state, and a caller that consumes the generator between calls voids it."""
    from torch.utils.data import DataLoader, TensorDataset
    from simgan_b200 import _spec
    g = Golden(CASES[0])

    def run(enabled):
        _spec.reset()
        _spec.enabled = enabled
        pol = gu.make_policy(g.policy(), g.O, g.H, g.A)
        agent = sg.PPO(pol, 0.2, g.ppo_epoch, g.nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
        d = gu.make_disc(g.disc(), g.F, g.HD)
        rs = gu.make_storage(g.buffer(), g.O, g.A, g.F)
        expert = g.t("expert").to(gu.DEV)
        loader = DataLoader(TensorDataset(expert), batch_size=g.gail_batch, shuffle=True,
                            drop_last=len(expert) > g.gail_batch)
        rms = sg.RunningMeanStd(shape=())
        torch.set_rng_state(g.t("rng_before_disc"))
        traces, taken = [], 0
        for it in range(3):
            with torch.no_grad():
                nv = pol.get_value(rs.obs[-1], rs.recurrent_hidden_states[-1], rs.masks[-1]).detach()
            for e in range(3):
                pre = d.__dict__.get("_predraw")
                taken += int(_spec.still_valid(pre, d.__dict__.get("_last_key")))
                d.update_gail_dyn(loader, rs)
                traces.append(d.last_trace.clone())
                if perturb and it == 1 and e == 1:
                    torch.rand(2)
            d.relabel_rollout(rs, 0.99, 0.1, rms)
            rs.compute_returns(nv, True, 0.99, 0.95, True)
            taken += int(_spec.still_valid(agent._predraw, agent._last_S))
            agent.update(rs)
            traces.append(agent.last_trace.clone())
            rs.after_update()
        return traces, pol.flat_params().cpu().clone(), d.flat_params().cpu().clone(), torch.get_rng_state(), taken

    try:
        t0, p0, d0, s0, k0 = run(False)
        t1, p1, d1, s1, k1 = run(True)
    finally:
        _spec.enabled = True
        _spec.reset()
    assert k0 == 0 and k1 >= (5 if perturb else 6), (k0, k1)
    assert torch.equal(s0, s1) and torch.equal(p0, p1) and torch.equal(d0, d1)
    assert len(t0) == len(t1) and all(torch.equal(a, b) for a, b in zip(t0, t1))


def test_division_through_reciprocal_matches_ieee_division():
    """The RunningMeanStd chain of the relabel divides through a precomputed correctly-rounded reciprocal plus two
    fused residual corrections; 2.7e9 pseudo-random operand pairs over the divisor ranges the chain can see
    (count + N from 1e-4 up to 1e12) and far beyond must give the IEEE quotient bit for bit."""
    from simgan_b200 import _lib
    lib = _lib.lib()
    bad = torch.zeros(1, dtype=torch.int64, device=gu.DEV)
    for seed, (lo, hi) in enumerate([(1e-4, 1e3), (1.0, 1e7), (1e3, 1e12), (1e-200, 1e200)]):
        _lib.check(lib.sg_selftest_division(1234 + seed, 1184, 2200, lo, hi, _lib.ptr(bad), _lib.current_stream()))
    assert int(bad.cpu()) == 0


@pytest.mark.parametrize("T,N,scale,start", [(2500, 5, 1.0, None), (1500, 16, 1e-3, (0.3, 2.5, 1e9)),
                                             (40, 3, 1e6, None), (1100, 2, 1.0, (0.0, 1.0, 9007199254740971.0))])
def test_relabel_running_stats_bitexact_vs_numpy(T, N, scale, start):
    """Relabel bookkeeping against the reference loop's arithmetic (NumPy float32 batch moments, float64 Chan merge,
    clip(reward / sqrt(var + 1e-7))) over more than one chunk of the chain, from a fresh and from a warm RMS state;
    the last case starts at a count whose sums have an all-ones significand, which takes the plain-division redo path."""
    from simgan_b200 import _lib
    lib = _lib.lib()
    gen = torch.Generator().manual_seed(T + N)
    raw = (torch.randn(T, N, generator=gen) * scale).contiguous()
    masks = (torch.rand(T + 1, N, 1, generator=gen) > 0.05).float()
    rms0 = (0.0, 1.0, 1e-4) if start is None else start
    # reference arithmetic on the host
    ref = orc.RunningMeanStd()
    ref.mean, ref.var, ref.count = np.float64(rms0[0]), np.float64(rms0[1]), rms0[2]
    ret = None
    ref_rewards = np.empty((T, N), dtype=np.float32)
    for t in range(T):
        ret = raw[t].clone() if ret is None else ret * 0.99 * masks[t, :, 0] + raw[t]
        ref.update(ret.numpy())
        ref_rewards[t] = np.clip(raw[t].numpy() / np.sqrt(ref.var + 1e-7), -10.0, 10.0)
    rewards = torch.empty(T, N, 1, device=gu.DEV)
    dret = torch.zeros(N, 1, device=gu.DEV)
    rms = torch.tensor(list(rms0), dtype=torch.float64, device=gu.DEV)
    mret = torch.empty(T, device=gu.DEV)
    ws = torch.empty(int(lib.sg_relabel_workspace_bytes(T, N)), dtype=torch.uint8, device=gu.DEV)
    raw_d, masks_d = raw.to(gu.DEV), masks.to(gu.DEV)
    _lib.check(lib.sg_relabel_normalize(_lib.ptr(raw_d), _lib.ptr(masks_d), _lib.ptr(rewards), T, N, 0.99,
                                        _lib.ptr(dret), 0, _lib.ptr(rms), _lib.ptr(mret), _lib.ptr(ws), _lib.current_stream()))
    got = rms.cpu().numpy()
    assert got[0] == float(ref.mean) and got[1] == float(ref.var) and got[2] == float(ref.count), (got, ref.mean, ref.var, ref.count)
    assert np.array_equal(rewards.cpu().numpy()[:, :, 0], ref_rewards)
    assert torch.equal(dret.cpu()[:, 0], ret)


@pytest.mark.parametrize("O,H,A", [(11, 16, 3), (13, 64, 7), (20, 256, 5)])
def test_large_minibatch_uses_16_row_tiles(O, H, A):
    """Minibatches of >= 4*8*#SM rows run 16-row tiles and several tiles per CTA (per-CTA partial gradients
    accumulated across tiles); a ragged last tile included.  kernel_mode 0 = column-owner kernel (or, at H=256 where
    the weights do not fit shared memory, the generic tile through L2), kernel_mode 2 = generic tile through L2.
    Against the oracle and against each other."""
    torch.manual_seed(1)
    T, N = 1251, 8                                             # 10008 samples -> 2 minibatches of 5004 rows (312.75 tiles)
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, 3, p, seed=4, ep_len=50.0)
    nv = orc.policy_forward(p, buf["obs"][-1])[0]
    orc.compute_returns(buf, nv, True, 0.99, 0.95, True)
    hyper = orc.PPOHyper(ppo_epoch=2, num_mini_batch=2)
    ora = orc.PPOOracle(p, hyper)
    torch.manual_seed(7)
    trace = []
    ora.update(buf, trace=trace)
    tr_o = np.array(trace)
    outs = []
    for mode in (0, 2):
        pol = gu.make_policy(p, O, H, A)
        agent = sg.PPO(pol, 0.2, 2, 2, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
        agent.kernel_mode = mode
        torch.manual_seed(7)
        agent.update(gu.make_storage(buf, O, A, 3))
        outs.append((agent.last_trace.double().numpy(), pol.flat_params().cpu().clone()))
    scale = np.abs(tr_o).max(axis=0)
    scale[1] = max(scale[1], 0.5)
    for tr, _ in outs:
        assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * scale + 1e-6), np.abs(tr - tr_o).max(axis=0) / scale
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-3, atol=2e-5)
