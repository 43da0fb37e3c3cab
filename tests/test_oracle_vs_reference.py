"""Pins oracle/ppo_gail_oracle.py against the UNMODIFIED reference modules (dev container only;
skipped where /root/reference is not mounted -- tests/test_oracle_golden.py covers that case)."""
import numpy as np
import pytest
import torch

from oracle import ppo_gail_oracle as orc
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")

O, A, H, F_, HD = 14, 7, 64, 25, 100


def _ref_policy(ref, seed, o=O, a=A, h=H):
    torch.manual_seed(seed)
    return ref.model.Policy((o,), ref_shim.BoxSpace(a), base_kwargs={"recurrent": False, "hidden_size": h})


def _ref_params(pol):
    b, d = pol.base, pol.dist
    return dict(aw1=b.actor[0].weight, ab1=b.actor[0].bias, aw2=b.actor[2].weight, ab2=b.actor[2].bias,
                cw1=b.critic[0].weight, cb1=b.critic[0].bias, cw2=b.critic[2].weight, cb2=b.critic[2].bias,
                vw=b.critic_linear.weight, vb=b.critic_linear.bias, mw=d.fc_mean.weight, mb=d.fc_mean.bias,
                logstd=d.logstd._bias)


def _fill_ref_storage(ref, buf, T, N):
    rs = ref.storage.RolloutStorage(T, N, (O,), ref_shim.BoxSpace(A), 1, F_)
    for k, v in buf.items():
        getattr(rs, k).copy_(v)
    return rs


def test_policy_init_and_forward_bitexact():
    ref = ref_shim.load()
    pol = _ref_policy(ref, 7)
    after_ref = torch.rand(2)
    torch.manual_seed(7)
    p = orc.init_policy(O, H, A)
    after = torch.rand(2)
    assert torch.equal(after, after_ref)          # RNG stream left in the same place
    for k, v in _ref_params(pol).items():
        assert torch.equal(p[k], v.data), k
    x = torch.randn(33, O)
    act = torch.randn(33, A)
    v_r, lp_r, ent_r, _ = pol.evaluate_actions(x, None, None, act)
    v, lp, ent = orc.policy_evaluate(p, x, act)
    assert torch.equal(v, v_r) and torch.equal(lp, lp_r) and torch.equal(ent, ent_r)
    torch.manual_seed(3)
    vr, ar, lr_, _ = pol.act(x, None, None)
    torch.manual_seed(3)
    vo, ao, lo = orc.policy_act(p, x)
    assert torch.equal(ar, ao) and torch.equal(lr_, lo) and torch.equal(vr, vo)


@pytest.mark.parametrize("use_gae,proper", [(True, True), (True, False), (False, True), (False, False)])
def test_compute_returns_bitexact(use_gae, proper):
    ref = ref_shim.load()
    T, N = 37, 5
    torch.manual_seed(0)
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, F_, p, seed=1, ep_len=6.0)
    buf["bad_masks"][3, 1] = 0.0
    rs = _fill_ref_storage(ref, buf, T, N)
    nv = torch.randn(N, 1)
    rs.compute_returns(nv, use_gae, 0.99, 0.95, proper)
    orc.compute_returns(buf, nv, use_gae, 0.99, 0.95, proper)
    assert torch.equal(buf["returns"], rs.returns)
    assert torch.equal(buf["value_preds"], rs.value_preds)


def test_sampler_and_generator_bitexact():
    ref = ref_shim.load()
    T, N = 16, 4
    torch.manual_seed(0)
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, F_, p, seed=2)
    rs = _fill_ref_storage(ref, buf, T, N)
    adv = torch.randn(T, N, 1)
    torch.manual_seed(11)
    ref_batches = list(rs.feed_forward_generator(adv, num_mini_batch=5))
    torch.manual_seed(11)
    my_batches = list(orc.feed_forward_batches(buf, adv, num_mini_batch=5))
    assert len(ref_batches) == len(my_batches) == 5
    for rb, mb in zip(ref_batches, my_batches):
        for x, y in zip(rb, mb):
            assert torch.equal(x, y)
    torch.manual_seed(12)
    ref_b = list(rs.feed_forward_generator(None, mini_batch_size=24))
    torch.manual_seed(12)
    my_b = list(orc.feed_forward_batches(buf, None, mini_batch_size=24))
    assert len(ref_b) == len(my_b) == 2 and ref_b[0][7] is None and my_b[0][7] is None
    assert torch.equal(ref_b[1][-1], my_b[1][-1])


def test_insert_after_update():
    ref = ref_shim.load()
    T, N = 3, 2
    rs = ref.storage.RolloutStorage(T, N, (O,), ref_shim.BoxSpace(A), 1, F_)
    buf = orc.new_buffer(T, N, O, A, F_)
    step = 0
    g = torch.Generator().manual_seed(0)
    for _ in range(4):
        args = [torch.randn(N, O, generator=g), torch.zeros(N, 1), torch.randn(N, A, generator=g),
                torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g),
                torch.ones(N, 1), torch.ones(N, 1), torch.randn(N, F_, generator=g)]
        rs.insert(*args)
        step = orc.buffer_insert(buf, step, *args)
        assert step == rs.step
    rs.after_update()
    orc.buffer_after_update(buf)
    for k, v in buf.items():
        assert torch.equal(v, getattr(rs, k)), k


def test_ppo_update_bitexact():
    ref = ref_shim.load()
    T, N = 32, 4
    pol = _ref_policy(ref, 5)
    torch.manual_seed(5)
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, F_, p, seed=3, ep_len=10.0)
    orc.compute_returns(buf, torch.zeros(N, 1), True, 0.99, 0.95, True)
    rs = _fill_ref_storage(ref, buf, T, N)
    hyper = orc.PPOHyper(ppo_epoch=3, num_mini_batch=4)
    agent = ref.ppo.PPO(pol, hyper.clip_param, hyper.ppo_epoch, hyper.num_mini_batch, hyper.value_loss_coef,
                        hyper.entropy_coef, lr=hyper.lr, eps=hyper.eps, max_grad_norm=hyper.max_grad_norm)
    mine = orc.PPOOracle(p, hyper)
    torch.manual_seed(21)
    out_ref = agent.update(rs)
    torch.manual_seed(21)
    out = mine.update(buf)
    assert out == out_ref
    for k, v in _ref_params(pol).items():
        assert torch.equal(mine.p[k].data, v.data), k


def test_disc_update_and_relabel_bitexact():
    ref = ref_shim.load()
    from torch.utils.data import DataLoader, TensorDataset
    T, N = 64, 4
    torch.manual_seed(9)
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, F_, p, seed=4, ep_len=9.0)
    rs = _fill_ref_storage(ref, buf, T, N)
    expert = torch.randn(300, F_) + 0.5
    torch.manual_seed(13)
    disc_ref = ref.gail.Discriminator(F_, HD, torch.device("cpu"))
    torch.manual_seed(13)
    disc = orc.DiscOracle(orc.init_disc(F_, HD))
    loader = DataLoader(TensorDataset(expert), batch_size=32, shuffle=True, drop_last=True)
    torch.manual_seed(17)
    outs_ref = [disc_ref.update_gail_dyn(loader, rs) for _ in range(2)]
    torch.manual_seed(17)
    outs = [disc.update_epoch(expert, buf, batch_size=32, drop_last=True) for _ in range(2)]
    assert outs == outs_ref
    for k, t in zip(orc.DISC_KEYS, disc_ref.trunk.parameters()):
        assert torch.equal(disc.d[k].data, t.data), k

    # relabel loop, verbatim semantics of main_gail_dyn_ppo.py:258-297
    gail_tar_length = 87.8
    r_sa = orc.alive_bonus_offset(buf["masks"], T, N, gail_tar_length)
    n_done = (1.0 - rs.masks).sum().cpu().numpy() + N / 2
    d_sa = 1 - n_done / (n_done + (T * N) / gail_tar_length)
    assert r_sa == float(np.log(d_sa) - np.log(1 - d_sa))
    rms_ref = ref.rms.RunningMeanStd(shape=())
    rms = orc.RunningMeanStd(shape=())
    for _rep in range(2):        # the running return and the RMS persist across iterations
        ref_means = []
        for step in range(T):
            rs.rewards[step], returns = disc_ref.predict_reward_combined(rs.obs_feat[step + 1], 0.99,
                                                                         rs.masks[step], offset=-r_sa)
            rms_ref.update(returns.view(-1).cpu().numpy())
            rews = rs.rewards[step].view(-1).cpu().numpy()
            rews = np.clip(rews / np.sqrt(rms_ref.var + 1e-7), -10.0, 10.0)
            rs.rewards[step] = torch.FloatTensor(rews).view(-1, 1)
            ref_means.append(float(torch.mean(returns)))
        means = orc.relabel_rewards(disc, rms, buf, 0.99, -r_sa)
        assert means == ref_means
        assert torch.equal(buf["rewards"], rs.rewards)
        assert rms.var == rms_ref.var and rms.mean == rms_ref.mean and rms.count == rms_ref.count


def test_running_mean_std_kat():
    """The reference's only known-answer test (running_mean_std.py:110-124), on the oracle class."""
    rng = np.random.RandomState(0)
    for shapes in [((3,), (4,), (5,)), ((3, 2), (4, 2), (5, 2))]:
        xs = [rng.randn(*s) for s in shapes]
        rms = orc.RunningMeanStd(epsilon=0.0, shape=xs[0].shape[1:])
        x = np.concatenate(xs, axis=0)
        for xi in xs:
            rms.update(xi)
        np.testing.assert_allclose([x.mean(axis=0), x.var(axis=0)], [rms.mean, rms.var])


@pytest.mark.parametrize("name,n_rows,s_dim,a_dim", [("hopper_new11_deform_n200_3.pkl", 17555, 11, 3)])
def test_expert_loader_matches_reference_merge(name, n_rows, s_dim, a_dim):
    import os
    ref = ref_shim.load()
    path = os.path.join(ref_shim.REF_ROOT, name)
    torch.manual_seed(0)
    cols = orc.load_sas_wpast(path, downsample_freq=1, load_num_trajs=200)
    assert len(cols) == 21 and cols[0].shape == (n_rows, s_dim) and cols[10].shape == (n_rows, a_dim)
    merged = orc.merge_sas(cols)
    merged_ref = ref.env_utils.select_and_merge_sas(cols, s_idx=np.array([0]), a_idx=np.array([0]))
    assert merged.shape == (n_rows, 2 * s_dim + a_dim)
    assert np.array_equal(merged, merged_ref)
    one = [list(c[5]) for c in cols]
    assert np.array_equal(orc.merge_sas(one), ref.env_utils.select_and_merge_sas(one, s_idx=np.array([0]),
                                                                                 a_idx=np.array([0])))
