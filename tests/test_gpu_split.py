"""SplitPolicy on the GPU (sg_split_forward / sg_split_ppo_update) vs the CPU oracle and the golden vectors of the
real reference (third_party/a2c_ppo_acktr/model_split.py + algo/ppo.py)."""
import numpy as np
import pytest
import torch

from oracle import ppo_gail_oracle as orc
from oracle.ref_shim import BoxSpace
from test_split_oracle import SPLIT_CASES, SplitGolden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _make(g, which="sp0"):
    import simgan_b200 as sg
    sp = sg.SplitPolicy((g.O,), BoxSpace(g.A), base_kwargs={"hidden_size": g.H, "num_feet": g.feet})
    for q, k in zip(sp.parameters(), orc.SPLIT_KEYS):
        q.data.copy_(g.t("%s_%s" % (which, k)).reshape(q.shape))
    sp.to(DEV)
    return sp


def _storage(g):
    import gpu_util as gu
    return gu.make_storage(g.buffer(), g.O, g.A, 3)


def _rel(a, b, floor=1e-30):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))


@pytest.mark.parametrize("case", SPLIT_CASES)
@pytest.mark.parametrize("B", [1, 13, 64])
def test_split_forward(case, B):
    g = SplitGolden(case)
    p = g.params()
    sp = _make(g)
    gen = torch.Generator().manual_seed(B)
    x, act = torch.randn(B, g.O, generator=gen), torch.randn(B, g.A, generator=gen)
    v_o, lp_o, ent_o = orc.split_evaluate(p, x, act)
    with torch.no_grad():
        v, lp, ent, _ = sp.evaluate_actions(x.to(DEV), None, None, act.to(DEV))
        assert _rel(sp.get_value(x.to(DEV), None, None).cpu(), v_o, 0.5) < 3e-6
    assert _rel(v.cpu(), v_o, 0.5) < 3e-6 and _rel(lp.cpu(), lp_o) < 3e-6
    assert abs(float(ent) - float(ent_o)) < 3e-6 * abs(float(ent_o))
    v_d, a_d, lp_d = orc.split_act(p, x, deterministic=True)
    v2, a2, lp2, _ = sp.act(x.to(DEV), None, None, deterministic=True)
    assert _rel(a2.cpu(), a_d, 0.01) < 3e-6 and _rel(lp2.cpu(), lp_d) < 3e-6
    # sampled: action = mean + exp(logstd) * N(0,1) drawn from the CUDA generator, like Normal.sample()
    torch.cuda.manual_seed(3)
    _, a3, lp3, _ = sp.act(x.to(DEV), None, None)
    torch.cuda.manual_seed(3)
    noise = torch.randn(B, g.A, device=DEV)
    _, a_n, lp_n = orc.split_act(p, x, noise=noise.cpu())
    assert _rel(a3.cpu(), a_n, 0.05) < 1e-5 and _rel(lp3.cpu(), lp_n) < 1e-5


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_split_ppo_update_vs_oracle_and_golden(case):
    import simgan_b200 as sg
    g = SplitGolden(case)
    sp = _make(g)
    agent = sg.PPO(sp, 0.2, g.ppo_epoch, g.nmb, 0.5, g.entropy_coef, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    rs = _storage(g)
    ora = orc.PPOOracle(g.params(), g.hyper(), keys=orc.SPLIT_KEYS, evaluate=orc.split_evaluate)
    trace = []
    out_o = ora.update(g.buffer(), index_chunks=g.chunks(), trace=trace)
    out = agent.update(rs, permutations=g.t("ppo_perm"))
    tr, tr_o = agent.last_trace.double().numpy(), np.array(trace)
    scale = np.abs(tr_o).max(axis=0)
    scale[1] = max(scale[1], 0.5)
    assert np.all(np.abs(tr[0] - tr_o[0]) <= 1e-5 * scale + 1e-7), (tr[0], tr_o[0])
    assert np.all(np.abs(tr - tr_o) <= 1e-4 * scale + 1e-6), np.abs(tr - tr_o).max(axis=0) / scale
    ref = g.z["ppo_losses"]                                         # the REAL reference's outputs
    assert abs(out[0] - ref[0]) <= 1e-4 * abs(ref[0]) and abs(out[2] - ref[2]) <= 1e-4 * abs(ref[2])
    assert abs(out[1] - ref[1]) <= 1e-4 * max(abs(ref[1]), 0.05)
    for q, k in zip(sp.parameters(), orc.SPLIT_KEYS):
        assert torch.allclose(q.detach().cpu().reshape(-1), g.t("sp1_" + k).reshape(-1), rtol=1e-3, atol=2e-5), k
    # drop-in call: draws its own permutations from the CPU generator, like feed_forward_generator
    sp2 = _make(g)
    agent2 = sg.PPO(sp2, 0.2, g.ppo_epoch, g.nmb, 0.5, g.entropy_coef, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    torch.set_rng_state(g.t("rng_before_ppo"))
    out2 = agent2.update(_storage(g))
    assert abs(out2[0] - ref[0]) <= 1e-4 * abs(ref[0])


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_split_ppo_shared_memory_w2_copy_is_bit_identical(case):
    """kernel_mode 0 keeps the three H x H matrices in a shared-memory copy refreshed by TMA after every Adam step,
    kernel_mode 2 reads them through L2: same micro-kernels, same summation order -> identical parameters and traces."""
    import simgan_b200 as sg
    g = SplitGolden(case)
    outs = []
    for mode in (0, 2):
        sp = _make(g)
        agent = sg.PPO(sp, 0.2, g.ppo_epoch, g.nmb, 0.5, g.entropy_coef, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
        agent.kernel_mode = mode
        agent.update(_storage(g), permutations=g.t("ppo_perm"))
        outs.append((agent.last_trace.clone(), sp.flat_params().cpu().clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_rollout_feeder_with_split_policy(case):
    """RolloutFeeder driving the policy the shipped scripts train (model_split.py:39-95): against the drop-in calls
    (SplitPolicy.act + RolloutStorage.insert, the same CUDA-generator stream) every buffer is bit-identical, and what the
    step copies is bit-exact against the ORACLE's buffer_insert (A2C/storage.py:70-84)."""
    import simgan_b200 as sg
    g = SplitGolden(case)
    sp = _make(g)
    T, N, O, A, F = 6, 5, g.O, g.A, 9
    rng = np.random.RandomState(4)
    obs0 = rng.randn(N, O).astype(np.float32)
    env = [(rng.randn(N, O).astype(np.float32), rng.randn(N).astype(np.float32), rng.rand(N) < 0.3, rng.rand(N) < 0.1,
            rng.randn(N, F)) for _ in range(T)]

    def fresh():
        rs = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F)
        rs.to(DEV)
        rs.obs[0].copy_(torch.from_numpy(obs0))
        return rs
    # drop-in calls (main_gail_dyn_ppo.py:209-236) + the oracle's insert on the host
    rs_a = fresh()
    buf = orc.new_buffer(T, N, O, A, F)
    buf["obs"][0].copy_(torch.from_numpy(obs0))
    torch.cuda.manual_seed(8)
    step_o = 0
    acts_a = []
    for t in range(T):
        with torch.no_grad():
            value, action, logp, hxs = sp.act(rs_a.obs[rs_a.step], rs_a.recurrent_hidden_states[rs_a.step], rs_a.masks[rs_a.step])
        acts_a.append(action.cpu())
        obs, rew, done, bad, feat = env[t]
        masks = torch.tensor([[0.0] if d else [1.0] for d in done])
        bad_masks = torch.tensor([[0.0] if b else [1.0] for b in bad])
        rs_a.insert(torch.from_numpy(obs).to(DEV), hxs, action, logp, value, torch.from_numpy(rew).unsqueeze(1), masks, bad_masks,
                    torch.Tensor(feat))
        step_o = orc.buffer_insert(buf, step_o, torch.from_numpy(obs), torch.zeros(N, 1), action.cpu(), logp.cpu(), value.cpu(),
                                   torch.from_numpy(rew).unsqueeze(1), masks, bad_masks, torch.Tensor(feat))
    # the feeder
    rs_b = fresh()
    torch.cuda.manual_seed(8)
    feeder = sg.RolloutFeeder(sp, rs_b)
    acts_b = [torch.from_numpy(feeder.begin().copy())]
    for t in range(T):
        obs, rew, done, bad, feat = env[t]
        acts_b.append(torch.from_numpy(feeder.step(obs, rew, done, bad, feat).copy()))
    for a, b in zip(acts_a, acts_b):
        assert torch.equal(a, b)
    for k in ("obs", "obs_feat", "recurrent_hidden_states", "rewards", "actions", "action_log_probs", "masks", "bad_masks"):
        assert torch.equal(getattr(rs_a, k), getattr(rs_b, k)), k
        assert torch.equal(getattr(rs_b, k).cpu(), buf[k]), k
    assert torch.equal(rs_a.value_preds[:-1], rs_b.value_preds[:-1])
    with torch.no_grad():
        nv = sp.get_value(rs_a.obs[-1], None, None)
    assert torch.equal(rs_b.value_preds[-1], nv)             # slot T already holds next_value
    assert rs_b.step == rs_a.step == step_o == 0
