"""World-size-2 gloo tests (CPU) of the data-parallel host logic: shard bounds,
identical sampler streams on every rank, and the allreduce callback the C library invokes between the
gradient reduction and the clip+Adam epilogue."""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import simgan_b200 as sg
        from simgan_b200 import dist as sg_dist
        from oracle.ref_shim import BoxSpace
        torch.manual_seed(0)
        pol = sg.Policy((14,), BoxSpace(7), base_kwargs={"recurrent": False, "hidden_size": 64})
        agent = sg.PPO(pol, 0.2, 2, 4, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
        disc = sg.Discriminator(25, 100, torch.device("cpu"))
        dp = sg_dist.attach(ppo=agent, disc=disc)
        assert dp is not None and agent.dp is dp and disc.dp is dp
        assert (dp.rank, dp.world) == (rank, world)
        # 1. shards of a 1024-row minibatch tile [0,1024) exactly
        b, e = dp.shard(1024)
        edges = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(edges, torch.tensor([b, e]))
        assert edges[0][0] == 0 and edges[-1][1] == 1024
        assert all(int(edges[i][1]) == int(edges[i + 1][0]) for i in range(world - 1))
        with pytest.raises(ValueError):
            dp.shard(1)                     # fewer rows than ranks
        # 2. every rank derives the same index streams from its identically seeded CPU generator
        torch.manual_seed(123)
        perm = agent.draw_permutations(512).clone().long()
        ei, pi, al = sg.Discriminator.draw_epoch_indices(1000, 128, True, 512)
        for t in (perm, ei, pi, al):
            ref = t.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, t)
        # 3. the C-side allreduce hook: sums the flat gradient region of the workspace in place
        ws = torch.zeros(4096, dtype=torch.uint8)
        cb = dp.make_callback(ws)
        n = 300
        view = ws[256:256 + 4 * n].view(torch.float32)
        view.copy_(torch.arange(n, dtype=torch.float32) * (rank + 1))
        rc = cb(C.c_void_p(ws.data_ptr() + 256), n, None)
        assert rc == 0 and dp.n_allreduce == 1
        expect = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
        assert torch.equal(view, expect)
        # a pointer outside the workspace is refused instead of corrupting memory
        assert cb(C.c_void_p(ws.data_ptr() + 4096), n, None) == 2
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_attach_is_a_noop_without_process_group():
    from simgan_b200 import dist as sg_dist
    assert sg_dist.attach() is None


def _walk_worker(rank, world, port, out_dir):
    """The ranks of a node split the walks of an update's permutations (PPO._shares_permutations): rank e % world builds
    permutation e, broadcasts it, and every rank ends with the generator where ppo_epoch torch.randperm calls leave it."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from simgan_b200 import host_sampler as hs
        n, k = 1 << 17, 5
        torch.manual_seed(31)
        want = [torch.randperm(n) for _ in range(k)]
        tail = torch.rand(3)
        torch.manual_seed(31)
        stage = torch.full((k, n), -1, dtype=torch.int32)
        ps = hs.PermutationStream(n, k, stage, owned=[e % world == rank for e in range(k)])
        for e in range(k):
            ps.wait(e)
            dist.broadcast(stage[e], e % world)
            assert torch.equal(stage[e].long(), want[e]), e
        ps.finish()
        assert torch.equal(torch.rand(3), tail)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_ranks_split_the_permutation_walks_world2(tmp_path):
    world = 2
    mp.spawn(_walk_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
