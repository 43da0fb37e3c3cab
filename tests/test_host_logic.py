"""Host-side logic of the reference-shaped classes (no GPU): RNG-stream emulation, Adam scalars, CPU policy
path used by env workers, construction order, pickling / module aliasing, loud failure without CUDA."""
import io
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset
from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler

import simgan_b200 as sg
from simgan_b200 import _lib, compat, dist as sg_dist
from simgan_b200.algo.adam import FusedAdam
from oracle import ppo_gail_oracle as orc
from oracle import ref_shim
from oracle.ref_shim import BoxSpace


def test_adam_schedule_matches_torch_scalars():
    opt = FusedAdam([], lr=3e-4, betas=(0.9, 0.999), eps=1e-5)
    opt.step_count = 7
    sch = opt.schedule(5)
    for i in range(5):
        t = 8 + i
        assert sch[0, i] == np.float32(3e-4 / (1 - 0.9 ** t))
        assert sch[1, i] == np.float32((1 - 0.999 ** t) ** 0.5)
    # one real torch.optim.Adam step against the kernel's op order restated on the host
    p = torch.tensor([0.3, -1.2, 2.0]); g = torch.tensor([0.5, -0.25, 1e-3])
    ref = p.clone().requires_grad_(True)
    o = torch.optim.Adam([ref], lr=3e-4, eps=1e-5)
    ref.grad = g.clone()
    o.step()
    opt2 = FusedAdam([], lr=3e-4, eps=1e-5)
    ss, bc2 = opt2.schedule(1)[:, 0]
    m = torch.zeros(3).lerp(g, 1 - 0.9)
    v = torch.zeros(3).mul(0.999).addcmul(g, g, value=1 - 0.999)
    mine = p + (-float(ss) * m) / (v.sqrt() / float(bc2) + 1e-5)
    assert torch.equal(mine, ref.detach())


@pytest.mark.parametrize("n_expert,S,B", [(1000, 512, 128), (300, 2048, 128), (130, 128, 32)])
def test_disc_epoch_index_stream_matches_dataloader_zip(n_expert, S, B):
    """draw_epoch_indices == the draws zip(DataLoader(shuffle), feed_forward_generator) + rand(B,1) make
    (A2C/algo/gail.py:157-163, :72), including where the CPU generator is left afterwards."""
    data = torch.arange(n_expert, dtype=torch.float32).unsqueeze(1)
    loader = DataLoader(TensorDataset(data), batch_size=B, shuffle=True, drop_last=True)
    torch.manual_seed(5)
    ref_e, ref_p, ref_a = [], [], []

    def rollout_gen():           # generator body runs at the first next(), like the reference's generator
        sampler = BatchSampler(SubsetRandomSampler(range(S)), B, drop_last=True)
        for idx in sampler:
            yield idx
    for eb, pb in zip(loader, rollout_gen()):
        ref_e.append(eb[0][:, 0].long())
        ref_p.append(torch.tensor(pb))
        ref_a.append(torch.rand(B, 1)[:, 0])
    after_ref = torch.rand(3)
    torch.manual_seed(5)
    e, p, a = sg.Discriminator.draw_epoch_indices(n_expert, B, True, S)
    after = torch.rand(3)
    n = len(ref_e)
    assert e.shape == (n, B) and p.shape == (n, B) and a.shape == (n, B)
    assert torch.equal(e, torch.stack(ref_e)) and torch.equal(p, torch.stack(ref_p))
    assert torch.equal(a, torch.stack(ref_a))
    if n_expert // B <= S // B:
        # zip stops on the loader side first: the streams leave the generator in the same place
        assert torch.equal(after, after_ref)


def test_ppo_permutations_are_the_batch_sampler_stream():
    pol = sg.Policy((14,), BoxSpace(7), base_kwargs={"recurrent": False, "hidden_size": 64})
    agent = sg.PPO(pol, 0.2, 3, 4, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    S = 203
    torch.manual_seed(9)
    perm = agent.draw_permutations(S).clone()
    torch.manual_seed(9)
    for e in range(3):
        chunks = list(BatchSampler(SubsetRandomSampler(range(S)), S // 4, drop_last=True))
        assert len(chunks) == 4
        flat = torch.tensor([i for c in chunks for i in c])
        assert torch.equal(perm[e, :flat.numel()].long(), flat)


def test_construction_order_and_cpu_policy_path_match_oracle():
    torch.manual_seed(3)
    pol = sg.Policy((14,), BoxSpace(7), base_kwargs={"recurrent": False, "hidden_size": 64})
    d = sg.Discriminator(25, 100, torch.device("cpu"))
    torch.manual_seed(3)
    p = orc.init_policy(14, 64, 7)
    dp = orc.init_disc(25, 100)
    for q, k in zip(pol.hot_path_parameters(), orc.POLICY_KEYS):
        assert torch.equal(q.data.reshape(-1), p[k].reshape(-1)), k
    for q, k in zip(d.hot_path_parameters(), orc.DISC_KEYS):
        assert torch.equal(q.data.reshape(-1), dp[k].reshape(-1)), k
    assert [n for n, _ in pol.named_parameters()] == [
        "base.actor.0.weight", "base.actor.0.bias", "base.actor.2.weight", "base.actor.2.bias",
        "base.critic.0.weight", "base.critic.0.bias", "base.critic.2.weight", "base.critic.2.bias",
        "base.critic_linear.weight", "base.critic_linear.bias", "dist.fc_mean.weight", "dist.fc_mean.bias",
        "dist.logstd._bias"]
    # batch-1 CPU inference, the way env workers call it (hopper_env_combined_policy.py:213-216)
    x = torch.randn(1, 14)
    with torch.no_grad():
        v, a, lp, _ = pol.act(x, torch.zeros(1, 1), torch.ones(1, 1), deterministic=True)
        v2, lp2, ent, _ = pol.evaluate_actions(x, None, None, a)
    vo, ao, lpo = orc.policy_act(p, x, deterministic=True)
    assert torch.equal(v, vo) and torch.equal(a, ao) and torch.equal(lp, lpo)
    assert torch.equal(v2, vo) and torch.equal(lp2, lpo)
    assert pol.is_recurrent is False and pol.recurrent_hidden_state_size == 1
    pol.reset_variance(BoxSpace(7), -1.5)
    assert float(pol.dist.logstd._bias.detach()[0]) == -1.5
    pol.reset_critic((14,))
    assert pol.base.critic[0].out_features == 64


def test_whole_object_pickles_round_trip():
    pol = sg.Policy((11,), BoxSpace(3), base_kwargs={"recurrent": False, "hidden_size": 64})
    d = sg.Discriminator(25, 100, torch.device("cpu"))
    buf, buf_d = io.BytesIO(), io.BytesIO()
    torch.save([pol, None], buf)            # main_gail_dyn_ppo.py:307-316 saves [actor_critic, ob_rms]
    torch.save(d, buf_d)
    buf.seek(0)
    buf_d.seek(0)
    pol2, _ = torch.load(buf, weights_only=False)
    d2 = torch.load(buf_d, weights_only=False)
    x = torch.randn(2, 11)
    with torch.no_grad():
        assert torch.equal(pol.act(x, None, None, deterministic=True)[1], pol2.act(x, None, None, deterministic=True)[1])
    for a, b in zip(d.parameters(), d2.parameters()):
        assert torch.equal(a, b)


def test_alias_modules_resolve_to_this_package():
    compat.install()
    try:
        import third_party.a2c_ppo_acktr.model as m
        from third_party.a2c_ppo_acktr import algo, utils
        from third_party.a2c_ppo_acktr.algo import gail
        from third_party.a2c_ppo_acktr.storage import RolloutStorage
        from third_party.a2c_ppo_acktr.baselines.common.running_mean_std import RunningMeanStd
        assert m.Policy is sg.Policy and algo.PPO is sg.PPO and gail.Discriminator is sg.Discriminator
        assert RolloutStorage is sg.RolloutStorage and RunningMeanStd is sg.RunningMeanStd
        assert utils.update_linear_schedule is sg.utils.update_linear_schedule
    finally:
        compat.uninstall()


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_reference_checkpoint_loads_into_our_classes():
    """The reference's shipped behaviour policy (a whole-object pickle of ITS Policy class) un-pickles into
    simgan_b200.Policy once the aliases are installed, and acts identically on the CPU."""
    path = os.path.join(ref_shim.REF_ROOT, "trained_models_hopper_bullet_new11", "ppo", "HopperURDFEnv-v3.pt")
    ref = ref_shim.load()
    saved = {k: v for k, v in __import__("sys").modules.items() if k.startswith("third_party")}
    try:
        import sys
        sys.modules.update(ref.modules)
        theirs, _ = torch.load(path, weights_only=False, map_location="cpu")
        for k in list(sys.modules):
            if k.startswith("third_party") or k in ("pybullet",) or k.startswith("my_pybullet_envs"):
                del sys.modules[k]
        compat.install()
        ours, _ = torch.load(path, weights_only=False, map_location="cpu")
    finally:
        compat.uninstall()
        import sys
        for k in list(sys.modules):
            if k.startswith("third_party"):
                del sys.modules[k]
        sys.modules.update(saved)
    assert type(ours) is sg.Policy and type(theirs) is not sg.Policy
    x = torch.randn(5, 11)
    with torch.no_grad():
        vo, ao, lo, _ = ours.act(x, torch.zeros(5, 1), torch.ones(5, 1), deterministic=True)
        vt, at, lt, _ = theirs.act(x, torch.zeros(5, 1), torch.ones(5, 1), deterministic=True)
    assert torch.equal(vo, vt) and torch.equal(ao, at) and torch.equal(lo, lt)
    assert (ours.obs_dim, ours.hidden_size, ours.act_dim) == (11, 64, 3)


def test_discriminator_pickles_after_a_data_parallel_attach(tmp_path):
    """main_gail_dyn_ppo.py:317-320 does torch.save(discr) right after the first update; with data parallelism on, the
    object then holds ctypes pointers to CUDA IPC mappings (dist.DataParallel._ctx) that must not travel."""
    import ctypes as C
    from simgan_b200 import dist as sg_dist

    d = sg.Discriminator(25, 100, torch.device("cpu"))
    dp = sg_dist.DataParallel.__new__(sg_dist.DataParallel)
    dp.__dict__.update(group=None, rank=0, world=2, transport="p2p", n_allreduce=0, _cb=None, _ctx={"disc": C.c_void_p(1234)})
    d.dp = dp
    d.last_trace = torch.zeros(3, 3)
    path = str(tmp_path / "d.pt")
    torch.save(d, path)
    d2 = torch.load(path, weights_only=False)
    assert d2.dp is None and d2.last_trace is None and d.dp is dp
    with pytest.raises(TypeError, match="per-process handles"):
        import pickle
        pickle.dumps(dp)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_reference_discriminator_checkpoint_loads_into_our_class(tmp_path):
    """A whole-object ``*_D.pt`` written by the REFERENCE's Discriminator (after it has stepped its torch Adam) loads
    through the aliases into simgan_b200's class with the kernel-side attributes defaulted and the optimizer replaced
    by a FusedAdam that carries lr / betas / eps / step and, once bound to the flat vector, the moment estimates."""
    import sys
    from simgan_b200.algo.adam import FusedAdam
    ref = ref_shim.load()
    saved = {k: v for k, v in sys.modules.items() if k.startswith("third_party")}
    path = str(tmp_path / "ref_D.pt")
    try:
        sys.modules.update(ref.modules)
        torch.manual_seed(3)
        theirs = ref.gail.Discriminator(9, 16, torch.device("cpu"))
        x = torch.randn(32, 9)
        loss = theirs.trunk(x).pow(2).mean()
        theirs.optimizer.zero_grad(); loss.backward(); theirs.optimizer.step()
        torch.save(theirs, path)
        for k in list(sys.modules):
            if k.startswith("third_party") or k in ("pybullet",) or k.startswith("my_pybullet_envs"):
                del sys.modules[k]
        compat.install()
        ours = torch.load(path, weights_only=False, map_location="cpu")
    finally:
        compat.uninstall()
        for k in list(sys.modules):
            if k.startswith("third_party"):
                del sys.modules[k]
        sys.modules.update(saved)
    assert type(ours) is sg.Discriminator and ours.dp is None and ours.kernel_mode == 0 and ours.last_trace is None
    assert isinstance(ours.optimizer, FusedAdam)
    g = ours.optimizer.param_groups[0]
    assert (g["lr"], tuple(g["betas"]), g["eps"]) == (1e-3, (0.9, 0.999), 1e-8)
    for a, b in zip(ours.trunk.parameters(), theirs.trunk.parameters()):
        assert torch.equal(a, b)
    # moments fold into flat buffers laid out like the parameter vector (simulated binding on the CPU)
    ps = list(ours.trunk.parameters())
    flat = torch.zeros(sum(p.numel() for p in ps) + 8)
    off = 0
    for p_ in ps:
        flat[off:off + p_.numel()].copy_(p_.data.reshape(-1)); p_.data = flat[off:off + p_.numel()].view(p_.shape); off += p_.numel()
    m, v = ours.optimizer.ensure_state(flat)
    assert ours.optimizer.step_count == 1
    off = 0
    for p_, q in zip(ps, theirs.trunk.parameters()):
        st = theirs.optimizer.state[q]
        assert torch.equal(m[off:off + p_.numel()], st["exp_avg"].reshape(-1))
        assert torch.equal(v[off:off + p_.numel()], st["exp_avg_sq"].reshape(-1))
        off += p_.numel()


def test_update_gail_dyn_refuses_loaders_it_cannot_emulate():
    from torch.utils.data import SequentialSampler
    d = sg.Discriminator(5, 8, torch.device("cpu"))
    x = torch.zeros(64, 5)
    with pytest.raises(NotImplementedError, match="RandomSampler"):
        d._check_loader(DataLoader(TensorDataset(x), batch_size=16, shuffle=False))
    with pytest.raises(NotImplementedError, match="generator"):
        d._check_loader(DataLoader(TensorDataset(x), batch_size=16, shuffle=True, generator=torch.Generator()))
    d._check_loader(DataLoader(TensorDataset(x), batch_size=16, shuffle=True, drop_last=True))
    # unrunnable configurations are refused before the generator is touched
    torch.manual_seed(0)
    before = torch.get_rng_state()
    with pytest.raises(ZeroDivisionError):
        sg.Discriminator.draw_epoch_indices(64, 128, True, 1000)
    with pytest.raises(NotImplementedError):
        sg.Discriminator.draw_epoch_indices(94, 64, False, 1000)
    assert torch.equal(torch.get_rng_state(), before)


def test_hot_path_refuses_to_run_without_cuda():
    pol = sg.Policy((14,), BoxSpace(7), base_kwargs={"recurrent": False, "hidden_size": 64})
    agent = sg.PPO(pol, 0.2, 1, 2, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    rs = sg.RolloutStorage(8, 2, (14,), BoxSpace(7), 1, 25)
    with pytest.raises(_lib.SgError, match="no CPU fallback"):
        rs.compute_returns(torch.zeros(2, 1), True, 0.99, 0.95)
    with pytest.raises(_lib.SgError, match="no CPU fallback"):
        agent.update(rs)
    with pytest.raises(_lib.SgError, match="no CPU fallback"):
        list(rs.feed_forward_generator(None, num_mini_batch=2))
    d = sg.Discriminator(25, 100, torch.device("cpu"))
    with pytest.raises(_lib.SgError):
        d.predict_reward_combined(torch.zeros(2, 25), 0.99, torch.ones(2, 1))
    loader = DataLoader(TensorDataset(torch.zeros(256, 25)), batch_size=128, shuffle=True, drop_last=True)
    with pytest.raises(_lib.SgError):
        d.update_gail_dyn(loader, rs)


def test_rollout_storage_host_staging_and_shapes():
    T, N, O, A, F = 5, 3, 14, 7, 25
    rs = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F)
    buf = orc.new_buffer(T, N, O, A, F)
    assert {k: tuple(getattr(rs, k).shape) for k in buf} == {k: tuple(v.shape) for k, v in buf.items()}
    g = torch.Generator().manual_seed(0)
    step = 0
    for _ in range(T + 2):       # wraps around like the reference (storage.py:84)
        args = [torch.randn(N, O, generator=g), torch.zeros(N, 1), torch.randn(N, A, generator=g),
                torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g),
                torch.ones(N, 1), torch.ones(N, 1), torch.randn(N, F, generator=g)]
        rs.insert(*args)
        step = orc.buffer_insert(buf, step, *args)
    rs.after_update()
    orc.buffer_after_update(buf)
    for k, v in buf.items():
        assert torch.equal(getattr(rs, k), v), k


@pytest.mark.parametrize("n,world", [(1024, 8), (128, 3), (7, 7), (1000, 6)])
def test_shard_bounds_partition(n, world):
    edges = [sg_dist.shard_bounds(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (a, b), (c, d) in zip(edges, edges[1:]):
        assert b == c and b > a
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


def test_expert_loaders_match_oracle_and_golden():
    from golden_util import GOLDEN_DIR
    path = os.path.join(GOLDEN_DIR, "mini_expert.pkl")
    torch.manual_seed(0)
    cols = sg.load_sas_wpast_from_pickle(path, downsample_freq=2, load_num_trajs=3)
    after = torch.rand(2)
    torch.manual_seed(0)
    cols_o = orc.load_sas_wpast(path, downsample_freq=2, load_num_trajs=3)
    assert torch.equal(torch.rand(2), after)                      # same generator consumption
    assert len(cols) == len(cols_o) == 21
    for a, b in zip(cols, cols_o):
        assert np.array_equal(a, b)
    merged = sg.select_and_merge_sas(cols)
    assert np.array_equal(merged, np.load(os.path.join(GOLDEN_DIR, "mini_expert_merged.npy")))
    # single-transition form used per env step by the caller (main_gail_dyn_ppo.py:220-226)
    one = [c[5].tolist() for c in cols]
    assert np.array_equal(sg.select_and_merge_sas(one), merged[5])
    # longer history windows
    m2 = sg.select_and_merge_sas(cols, s_idx=np.array([0, 2]), a_idx=np.array([0, 1]))
    assert np.array_equal(m2, orc.merge_sas(cols_o, s_idx=(0, 2), a_idx=(0, 1)))


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_expert_tensor_reproduces_the_committed_fixture():
    from golden_util import GOLDEN_DIR
    torch.manual_seed(0)
    x = sg.expert_tensor(os.path.join(ref_shim.REF_ROOT, "hopper_new11_deform_n200_3.pkl"), "cpu")
    assert np.array_equal(x.numpy(), np.load(os.path.join(GOLDEN_DIR, "hopper_expert_sas_f32.npy")))


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_laikago_expert_tensor_reproduces_the_committed_fixture():
    """laika_70_deform_n200_0.pkl (BASELINE configs[2]/[3]): our loader == the reference's merge stored by
    oracle/make_golden.py --laika, (15678, 86), s_dim 37 / a_dim 12 (SURVEY section 8 glossary)."""
    from golden_util import GOLDEN_DIR
    torch.manual_seed(0)
    x = sg.expert_tensor(os.path.join(ref_shim.REF_ROOT, "laika_70_deform_n200_0.pkl"), "cpu", load_num_trajs=200)
    ref = np.load(os.path.join(GOLDEN_DIR, "laika_expert_sas_f32.npy"))
    assert x.shape == (15678, 86) and np.array_equal(x.numpy(), ref)


def test_laikago_expert_fixture_shape_and_content():
    """The committed Laikago fixture is what the bench and the full-size tests feed the discriminator (no reference
    tree needed): [s_t (37) | a_t (12) | s_{t+1} (37)], finite, consecutive rows of one trajectory chain s_{t+1} -> s_t."""
    from golden_util import GOLDEN_DIR
    x = np.load(os.path.join(GOLDEN_DIR, "laika_expert_sas_f32.npy"))
    assert x.shape == (15678, 86) and x.dtype == np.float32 and np.isfinite(x).all()
    chained = np.all(x[1:, :37] == x[:-1, 49:], axis=1)
    assert chained.mean() > 0.95          # breaks only at the 200 trajectory boundaries


class _FakeConsumer(object):
    def __init__(self, n):
        self.n, self.slot, self.early = n, None, 0

    def _speculate(self):
        from simgan_b200 import _spec
        if not _spec.still_valid(self.slot, self.n):
            self.slot = _spec.predraw(self.n, lambda: torch.randperm(self.n))
            self.early += 1

    def call(self):
        from simgan_b200 import _spec
        _spec.consumed(self)
        got, self.slot = _spec.take(self.slot, self.n), None
        out = got if got is not None else torch.randperm(self.n)
        _spec.host_idle()
        return out, got is not None


def test_early_draws_leave_the_generator_stream_unchanged():
    """_spec: draws made early + rewind are indistinguishable from late draws, the predictor learns the
    D x k -> PPO -> D ... order, and a caller that touches the generator in between just voids the early draw."""
    from simgan_b200 import _spec

    def run(enabled, perturb):
        _spec.reset()
        _spec.enabled = enabled
        torch.manual_seed(5)
        d, p = _FakeConsumer(37), _FakeConsumer(101)
        outs, hits = [], 0
        for it in range(4):
            for _ in range(3):
                o, h = d.call(); outs.append(o); hits += h
            if perturb and it == 2:
                outs.append(torch.rand(3))
            o, h = p.call(); outs.append(o); hits += h
        return outs, hits, torch.get_rng_state()

    try:
        base, hits0, st0 = run(False, False)
        spec, hits1, st1 = run(True, False)
        assert hits0 == 0 and hits1 >= 10                  # steady state: every draw after the first iteration is early
        assert torch.equal(st0, st1) and all(torch.equal(a, b) for a, b in zip(base, spec))
        base, _, st0 = run(False, True)
        spec, _, st1 = run(True, True)
        assert torch.equal(st0, st1) and all(torch.equal(a, b) for a, b in zip(base, spec))
    finally:
        _spec.enabled = True
        _spec.reset()


def test_predraw_swallows_errors_and_rewinds():
    from simgan_b200 import _spec
    torch.manual_seed(1)
    before = torch.get_rng_state()

    def boom():
        torch.rand(4)
        raise ZeroDivisionError("as gail.py:193 would")
    assert _spec.predraw("k", boom) is None
    assert torch.equal(torch.get_rng_state(), before)
    slot = _spec.predraw("k", lambda: torch.rand(2))
    assert torch.equal(torch.get_rng_state(), before)
    assert _spec.take(slot, "other") is None and torch.equal(torch.get_rng_state(), before)
    got = _spec.take(slot, "k")
    torch.set_rng_state(before)
    assert torch.equal(got, torch.rand(2))


# ---- host sampler: the torch.randperm stream of the CPU default generator, produced by csrc/sg_host.cu ----------------------
@pytest.mark.parametrize("n,k", [(1 << 17, 3), (300001, 2), (1, 2), (5, 4)])
def test_permutation_stream_equals_torch_randperm(n, k):
    """Values AND generator state: feed_forward_generator's sampler draws (A2C/storage.py:158-162) must stay bit-exact."""
    from simgan_b200 import host_sampler as hs
    assert hs.usable()
    torch.manual_seed(11)
    torch.rand(3)                                   # some earlier consumer
    start = torch.get_rng_state()
    want = [torch.randperm(n) for _ in range(k)]
    tail_want = torch.rand(5)                       # a later consumer sees the same stream
    torch.set_rng_state(start)
    out = torch.empty(k, n, dtype=torch.int32)
    old_min, hs.MIN_ELEMENTS = hs.MIN_ELEMENTS, 1   # force the C stream for the tiny sizes too
    try:
        ps = hs.PermutationStream(n, k, out)
        for e in range(k):
            got = ps.wait(e)
            assert torch.equal(got.long(), want[e])
        ps.finish()
    finally:
        hs.MIN_ELEMENTS = old_min
    assert torch.equal(torch.rand(5), tail_want)


def test_permutation_stream_small_sizes_use_torch():
    from simgan_b200 import host_sampler as hs
    torch.manual_seed(3)
    want = [torch.randperm(100) for _ in range(2)]
    after = torch.get_rng_state()
    torch.manual_seed(3)
    out = torch.empty(2, 100, dtype=torch.int32)
    ps = hs.PermutationStream(100, 2, out)
    assert ps._h is None
    ps.finish()
    assert torch.equal(out[0].long(), want[0]) and torch.equal(out[1].long(), want[1])
    assert torch.equal(torch.get_rng_state(), after)
    torch.manual_seed(4)
    a = hs.randperm_i32(1 << 17)
    torch.manual_seed(4)
    assert a.dtype == torch.int32 and torch.equal(a.long(), torch.randperm(1 << 17))


@pytest.mark.parametrize("n,m", [(1 << 17, 16384), (200003, 200003), (1 << 17, 0), (150000, 1)])
def test_randperm_prefix_equals_torch(n, m):
    from simgan_b200 import host_sampler as hs
    torch.manual_seed(21)
    want = torch.randperm(n)[:m]
    tail = torch.rand(4)
    torch.manual_seed(21)
    got = hs.randperm_prefix(n, m)
    assert got.dtype == torch.int64 and torch.equal(got, want)
    assert torch.equal(torch.rand(4), tail)


def test_return_normalizer_restated():
    """ReturnNormalizer vs a direct restatement of vec_normalize.py:50-58 (runs on the GPU box too, where the reference's
    class is not available; tests/test_twin_vs_reference.py holds it against the real class)."""
    n, gamma, eps, clip = 6, 0.99, 1e-8, 10.0
    ours = sg.ReturnNormalizer(n, gamma=gamma)
    ret = np.zeros(n)
    mean, var, count = 0.0, 1.0, 1e-4
    rs = np.random.RandomState(1)
    for t in range(100):
        rews = rs.standard_normal(n).astype(np.float32) * 3
        news = rs.rand(n) < 0.15
        ret = ret * gamma + rews
        bm, bv, bc = np.mean(ret, axis=0), np.var(ret, axis=0), ret.shape[0]
        delta = bm - mean
        tot = count + bc
        new_mean = mean + delta * bc / tot
        m2 = var * count + bv * bc + np.square(delta) * count * bc / tot
        mean, var, count = new_mean, m2 / tot, tot
        want = np.clip(rews / np.sqrt(var + eps), -clip, clip)
        ret[news] = 0.0
        got = ours(rews, news)
        assert np.array_equal(got, want), t
    assert ours.ret_rms.count == count and ours.ret_rms.var == var


def test_permutation_stream_owned_subset():
    """The ranks of one node split the walks of an update's permutations: a rank builds only the ones it owns, yet its
    generator ends where all the draws leave it."""
    from simgan_b200 import host_sampler as hs
    n = 1 << 17
    torch.manual_seed(7)
    want = [torch.randperm(n) for _ in range(5)]
    after = torch.get_rng_state()
    torch.manual_seed(7)
    out = torch.full((5, n), -1, dtype=torch.int32)
    ps = hs.PermutationStream(n, 5, out, owned=[True, False, True, False, False])
    for e in range(5):
        ps.wait(e)
    ps.finish()
    assert torch.equal(out[0].long(), want[0]) and torch.equal(out[2].long(), want[2])
    assert bool((out[1] == -1).all()) and bool((out[3] == -1).all()) and bool((out[4] == -1).all())
    assert torch.equal(torch.get_rng_state(), after)
