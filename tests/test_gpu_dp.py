"""Data-parallel (minibatch-sharded) updates on >= 2 GPUs: G-GPU result == 1-GPU result within fp32
reassociation tolerance (SURVEY.md section 8e -- there is no reference counterpart for multi-GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _spawn(fn, args_of_port, nprocs):
    """mp.spawn with a fresh rendezvous port; a port that got taken between the probe and the bind (EADDRINUSE) is retried."""
    last = None
    for _ in range(4):
        try:
            mp.spawn(fn, args=args_of_port(_free_port()), nprocs=nprocs, join=True)
            return
        except Exception as e:      # noqa: BLE001
            if "EADDRINUSE" not in str(e) and "address already in use" not in str(e):
                raise
            last = e
    raise last


def _run_case(g, dp_on, dev, transport=None):
    import gpu_util as gu
    import simgan_b200 as sg
    from simgan_b200 import dist as sg_dist
    from torch.utils.data import DataLoader, TensorDataset
    gu.DEV = dev
    pol = gu.make_policy(g.policy(), g.O, g.H, g.A)
    agent = sg.PPO(pol, 0.2, g.ppo_epoch, g.nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    d = gu.make_disc(g.disc(), g.F, g.HD)
    buf = g.buffer()
    buf["rewards"].copy_(g.t("relabel_rewards"))
    buf["returns"].copy_(g.t("gae_returns"))
    buf["value_preds"][-1] = g.t("next_value")
    rs = gu.make_storage(buf, g.O, g.A, g.F)
    expert = g.t("expert").to(dev)
    loader = DataLoader(TensorDataset(expert), batch_size=g.gail_batch, shuffle=True, drop_last=len(expert) > g.gail_batch)
    if dp_on:
        dp = sg_dist.attach(ppo=agent, disc=d, transport=transport)
        assert dp.transport == transport
    else:
        agent.kernel_mode = 1
        d.kernel_mode = 1
    dl = d.update_gail_dyn(loader, rs, replay=g.disc_replay(0))
    pl = agent.update(rs, permutations=g.t("ppo_perm"))
    return (np.array(dl), np.array(pl), d.last_trace.clone(), agent.last_trace.clone(),
            pol.flat_params().cpu().clone(), d.flat_params().cpu().clone())


def _worker(rank, world, port, case, out, transport):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        from golden_util import Golden
        torch.set_num_threads(1)
        g = Golden(case)
        res = _run_case(g, True, dev, transport)
        torch.cuda.synchronize()
        # every rank ends with identical parameters (same clip+Adam epilogue on the same reduced gradient)
        for t in (res[4], res[5]):
            ref = t.to(dev).clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref.cpu(), t), "ranks diverged"
        if rank == 0:
            dist.barrier()
            one = _run_case(g, False, dev)
            for a, b in zip(res[:2], one[:2]):
                assert np.all(np.abs(a - b) <= 1e-5 * np.maximum(np.abs(b), 0.05)), (a, b)
            assert torch.allclose(res[2], one[2], rtol=2e-5, atol=1e-6)
            tr, tr1 = res[3], one[3]
            scale = tr1.abs().max(dim=0).values.clamp_min(0.5)
            assert bool(((tr - tr1).abs() <= 2e-5 * scale).all())
            assert torch.allclose(res[4], one[4], rtol=1e-4, atol=2e-6)
            assert torch.allclose(res[5], one[5], rtol=1e-4, atol=2e-6)
            open(out, "w").write("ok")
        else:
            dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("case", ["hopper_cfg1_seed0.npz", "laika_dims_seed2.npz"])
def test_data_parallel_matches_single_gpu(case, transport, tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = min(n, int(os.environ.get("SG_DP_WORLD", "2")))
    out = str(tmp_path / "ok")
    _spawn(_worker, lambda port: (world, port, case, out, transport), world)
    assert os.path.exists(out)


def _split_case(g, dp_on, dev):
    import gpu_util as gu
    import simgan_b200 as sg
    from oracle import ppo_gail_oracle as orc
    from oracle.ref_shim import BoxSpace
    from simgan_b200 import dist as sg_dist
    gu.DEV = dev
    sp = sg.SplitPolicy((g.O,), BoxSpace(g.A), base_kwargs={"hidden_size": g.H, "num_feet": g.feet})
    for q, k in zip(sp.parameters(), orc.SPLIT_KEYS):
        q.data.copy_(g.t("sp0_%s" % k).reshape(q.shape))
    sp.to(dev)
    agent = sg.PPO(sp, 0.2, g.ppo_epoch, g.nmb, 0.5, g.entropy_coef, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    if dp_on:
        sg_dist.attach(ppo=agent, transport="p2p")
    rs = gu.make_storage(g.buffer(), g.O, g.A, 3)
    out = agent.update(rs, permutations=g.t("ppo_perm"))
    return np.array(out), agent.last_trace.clone(), sp.flat_params().cpu().clone()


def _split_worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        from test_split_oracle import SplitGolden
        torch.set_num_threads(1)
        g = SplitGolden(case)
        if (g.T * g.N // g.nmb) % world:
            if rank == 0:
                open(out, "w").write("skipped: minibatch not divisible by the world size")
            return
        res = _split_case(g, True, dev)
        torch.cuda.synchronize()
        ref = res[2].to(dev).clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref.cpu(), res[2]), "ranks diverged"
        dist.barrier()
        if rank == 0:
            one = _split_case(g, False, dev)
            assert np.all(np.abs(res[0] - one[0]) <= 1e-5 * np.maximum(np.abs(one[0]), 0.05)), (res[0], one[0])
            scale = one[1].abs().max(dim=0).values.clamp_min(0.5)
            assert bool(((res[1] - one[1]).abs() <= 2e-5 * scale).all())
            assert torch.allclose(res[2], one[2], rtol=1e-4, atol=2e-6)
            open(out, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_split_policy_data_parallel_matches_single_gpu(tmp_path):
    """SplitPolicy's PPO update with the in-kernel peer-memory exchange (p2p transport)."""
    from test_split_oracle import SPLIT_CASES
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = min(n, int(os.environ.get("SG_DP_WORLD", "2")))
    for case in SPLIT_CASES:
        out = str(tmp_path / ("ok_" + case))
        mp.spawn(_split_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
        assert os.path.exists(out)


# ---- large rollout: the ranks split the permutation walks of an update and broadcast them (PPO._shares_permutations) -----------
def _big_case(dev, dp_on):
    import simgan_b200 as sg
    from simgan_b200 import dist as sg_dist
    from oracle.ref_shim import BoxSpace
    T, N, O, A, H = 64, 2048, 16, 4, 64              # S = 131072 = host_sampler.MIN_ELEMENTS, 4 minibatches of 32768 rows
    torch.manual_seed(5)
    pol = sg.Policy((O,), BoxSpace(A), base_kwargs={"recurrent": False, "hidden_size": H})
    pol.to(dev)
    rs = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, 3)
    g = torch.Generator().manual_seed(9)
    rs.obs.copy_(torch.randn(rs.obs.shape, generator=g))
    rs.actions.copy_(torch.randn(rs.actions.shape, generator=g))
    rs.value_preds.copy_(torch.randn(rs.value_preds.shape, generator=g))
    rs.returns.copy_(torch.randn(rs.returns.shape, generator=g))
    rs.action_log_probs.fill_(-float(A))
    rs.to(dev)
    agent = sg.PPO(pol, 0.2, 3, 4, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    if dp_on:
        sg_dist.attach(ppo=agent, policy="auto")
    torch.manual_seed(77)
    outs = [agent.update(rs) for _ in range(2)]      # second call: rank 0 takes its early draw, the others do not have one
    torch.cuda.synchronize()
    return np.array(outs), agent.last_trace.clone(), agent._perm_dev.cpu().clone(), torch.get_rng_state(), agent


def _big_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        torch.set_num_threads(1)
        res = _big_case(dev, True)
        assert res[4].last_sharded and res[4]._shares_permutations(64 * 2048)
        # every rank holds the same permutations and the same generator state
        ref = res[2].to(dev).clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref.cpu(), res[2]), "permutations differ between ranks"
        st = res[3].to(dev).clone()
        dist.broadcast(st, 0)
        assert torch.equal(st.cpu(), res[3]), "generator states differ between ranks"
        dist.barrier()
        if rank == 0:
            one = _big_case(dev, False)
            assert torch.equal(one[2], res[2]) and torch.equal(one[3], res[3])          # same index streams as one GPU
            assert np.all(np.abs(res[0] - one[0]) <= 1e-4 * np.maximum(np.abs(one[0]), 0.05)), (res[0], one[0])
            open(out, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_large_rollout_ranks_split_the_permutation_walks(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    out = str(tmp_path / "ok_big")
    _spawn(_big_worker, lambda port: (2, port, out), 2)
    assert os.path.exists(out)


# ---- large rollout: the relabel's D forward is split by rows over the ranks (Discriminator.SHARD_RELABEL_FROM) ---------------
def _relabel_case(dev, dp_on):
    import simgan_b200 as sg
    from simgan_b200 import dist as sg_dist
    from oracle.ref_shim import BoxSpace
    T, N, F = 512, 2048 + 3, 25                      # T*N just above 2^20, not divisible by the world size
    torch.manual_seed(3)
    d = sg.Discriminator(F, 100, torch.device(dev))
    rs = sg.RolloutStorage(T, N, (4,), BoxSpace(2), 1, F)
    g = torch.Generator().manual_seed(4)
    rs.obs_feat.copy_(torch.randn(rs.obs_feat.shape, generator=g))
    rs.masks.copy_((torch.rand(rs.masks.shape, generator=g) > 0.02).float())
    rs.to(dev)
    if dp_on:
        sg_dist.attach(disc=d, policy="auto")
    rms = sg.RunningMeanStd(shape=())
    outs = []
    for _ in range(2):                               # the second pass starts from carried returns / statistics
        mr = d.relabel_rollout(rs, 0.99, -0.3, rms)
        outs.append((rs.rewards.cpu().clone(), d.returns.cpu().clone(), mr.cpu().clone(), float(rms.mean), float(rms.var), rms.count))
    return outs


def _relabel_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    torch.cuda.set_device(rank)
    dev = "cuda:%d" % rank
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    try:
        torch.set_num_threads(1)
        res = _relabel_case(dev, True)
        ref = res[1][0].to(dev).clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref.cpu(), res[1][0]), "ranks diverged"
        dist.barrier()
        if rank == 0:
            one = _relabel_case(dev, False)
            for a, b in zip(res, one):
                assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])      # bit-identical
                assert a[3:] == b[3:]
            open(out, "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_large_rollout_relabel_forward_is_split_over_the_ranks(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    out = str(tmp_path / "ok_relabel")
    _spawn(_relabel_worker, lambda port: (2, port, out), 2)
    assert os.path.exists(out)
