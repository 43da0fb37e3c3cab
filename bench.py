#!/usr/bin/env python
"""PPO+GAIL update-steps/sec benchmark (BASELINE.json metric).

One "step" of this benchmark = ONE outer-iteration update phase of main_gail_dyn_ppo.py:238-304 on a
synthetic rollout buffer:  next_value -> gail_epoch x Discriminator.update_gail_dyn -> reward relabel
(T x predict_reward_combined + RunningMeanStd + clip) -> RolloutStorage.compute_returns (GAE) ->
PPO.update -> after_update.  The metric counts OPTIMIZER steps (PPO minibatches + discriminator
minibatches) per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl reference]

* default arm: the sm_100a CUDA path through the reference-shaped classes of ``simgan_b200``.
  ``value``  = rollout buffer already resident in HBM when the timed region starts;
  ``e2e``    = same call sequence, but every step first copies the step's rollout tensors from pinned
               HOST memory (where the host-side envs leave them) and reads the losses back.
* ``--impl reference``: the CPU restatement of the reference's own eager-PyTorch path (``oracle/``; the
  reference tree itself cannot travel to the GPU box) on the host cores, same workload and metric.
  Each step is a bounded 1/5 sample of the workload (see ``reference_sample``).

Timing: CUDA events on the launching (current) stream around every step, L2 flushed between steps
(outside the events), device times summed over the K steps, MAX over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_EXPERT = os.path.join(ROOT, "tests", "golden", "hopper_expert_sas_f32.npy")
GOLDEN_EXPERT_LAIKA = os.path.join(ROOT, "tests", "golden", "laika_expert_sas_f32.npy")

# name -> sizes (SURVEY.md section 8 per-config table).  cfg2 is the configuration the metric is quoted on.
CONFIGS = {
    "cfg1": dict(T=128, N=4, O=14, A=7, H=64, F=25, HD=100, expert="hopper", ep_len=88.0,
                 label="HopperCombinedEnv-v1 sizes num_processes=4 num_steps=128 hidden=64"),
    "cfg2": dict(T=2048, N=16, O=14, A=7, H=64, F=25, HD=100, expert="hopper", ep_len=88.0,
                 label="HopperCombinedEnv-v1 sizes num_processes=16 num_steps=2048 hidden=64 (BASELINE configs[1])"),
    "cfg3": dict(T=2048, N=16, O=64, A=28, H=256, F=86, HD=100, expert="laika", ep_len=78.0,
                 label="LaikagoCombinedEnv-v1 sizes num_processes=16 num_steps=2048 hidden=256"),
    "cfg4": dict(T=2048, N=128, O=64, A=28, H=256, F=86, HD=100, expert="laika", ep_len=78.0,
                 label="LaikagoCombinedEnv-v1 sizes num_processes=128 num_steps=2048 hidden=256"),
    "cfg5": dict(T=1024, N=4096, O=111, A=12, H=64, F=86, HD=100, expert=16384, ep_len=78.0,
                 label="synthetic rollout 1024x4096 obs_dim=111"),
    # what train_hopper_deform.sh:5 actually runs: SplitPolicy, hidden 100, 8 envs x 1000 steps, 16 minibatches
    "shipped": dict(T=1000, N=8, O=14, A=7, H=100, F=25, HD=100, expert="hopper", ep_len=88.0, split=True,
                    hyper=dict(num_mini_batch=16, entropy_coef=0.0),
                    label="train_hopper_deform.sh sizes: SplitPolicy hidden=100 num_processes=8 num_steps=1000 num_mini_batch=16"),
}
HYPER = dict(gamma=0.99, gae_lambda=0.95, clip_param=0.2, ppo_epoch=10, num_mini_batch=32, value_loss_coef=0.5,
             entropy_coef=0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5, gail_epoch=5, gail_batch=128,
             gail_tar_length=87.8)     # argparse defaults of A2C/arguments.py:33-215


class Box(object):
    """gym.spaces.Box stand-in: the path only reads __class__.__name__ and .shape (A2C/model.py:56-58)."""

    def __init__(self, dim):
        self.shape = (int(dim),)


Box.__name__ = "Box"


def expert_rows(c, seed):
    if c["expert"] == "hopper":
        return torch.from_numpy(np.load(GOLDEN_EXPERT))          # real hopper_new11_deform_n200_3.pkl rows (17555,25)
    if c["expert"] == "laika":
        return torch.from_numpy(np.load(GOLDEN_EXPERT_LAIKA))    # real laika_70_deform_n200_0.pkl rows (15678,86)
    g = torch.Generator().manual_seed(7000 + seed)
    return torch.randn(int(c["expert"]), c["F"], generator=g)


def host_rollout(c, seed, expert):
    """Seeded synthetic rollout contents on the HOST: what the env workers would have
    produced.  actions / value_preds / action_log_probs are filled in by the policy itself later."""
    g = torch.Generator().manual_seed(1000 + seed)
    T, N, O, A, F = c["T"], c["N"], c["O"], c["A"], c["F"]
    obs = torch.randn(T + 1, N, O, generator=g)
    pick = torch.randint(0, expert.shape[0], ((T + 1) * N,), generator=g)
    feat = (expert[pick] + 0.1 * torch.randn((T + 1) * N, F, generator=g)).view(T + 1, N, F)
    masks = (torch.rand(T + 1, N, 1, generator=g) >= 1.0 / c["ep_len"]).float()
    bad = torch.where((masks == 0) & (torch.rand(T + 1, N, 1, generator=g) < 1.0 / 500.0), 0.0, 1.0)
    noise = torch.randn(T * N, A, generator=g)
    return dict(obs=obs, obs_feat=feat, masks=masks, bad_masks=bad), noise


def algorithmic_work(c, n_disc_batches):
    """Bytes / FLOPs per outer iteration (SURVEY.md 8d formulas)."""
    h = HYPER
    S = c["T"] * c["N"]
    s_used = h["num_mini_batch"] * (S // h["num_mini_batch"])
    O, A, H, F, HD = c["O"], c["A"], c["H"], c["F"], c["HD"]
    w = {}
    w["gae_bytes"] = 20 * S
    w["ppo_bytes"] = h["ppo_epoch"] * s_used * 4 * (O + A + 4)
    if c.get("split"):     # three trunks, heads (8f + 6f + 1) x H; first-layer dX skipped
        w["ppo_flops"] = h["ppo_epoch"] * s_used * 2 * (3 * (3 * H * H + H * (2 * A + 1)) + 2 * (3 * O * H))
    else:
        w["ppo_flops"] = h["ppo_epoch"] * s_used * 2 * (3 * (2 * H * H + H * A + H) + 2 * (2 * O * H))
    w["disc_bytes"] = h["gail_epoch"] * n_disc_batches * h["gail_batch"] * 4 * F * 2
    w["disc_flops"] = h["gail_epoch"] * n_disc_batches * h["gail_batch"] * 12 * 2 * (F * HD + HD * HD + HD)
    w["relabel_bytes"] = S * (4 * F + 8)
    w["relabel_flops"] = S * 2 * (F * HD + HD * HD + HD)
    return w


# ------------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------------
class Workload(object):
    def __init__(self, c, seed, device, dp=False):
        import simgan_b200 as sg
        from simgan_b200 import dist as sg_dist
        from torch.utils.data import DataLoader, TensorDataset
        self.c, self.device, self.sg = c, device, sg
        h = HYPER
        torch.manual_seed(seed)
        T, N, O, A, H, F, HD = c["T"], c["N"], c["O"], c["A"], c["H"], c["F"], c["HD"]
        # construction order of main_gail_dyn_ppo.py:71-162: policy -> PPO -> expert -> D
        if c.get("split"):
            self.policy = sg.SplitPolicy((O,), Box(A), base_kwargs={"hidden_size": H, "num_feet": A // 7})
        else:
            self.policy = sg.Policy((O,), Box(A), base_kwargs={"recurrent": False, "hidden_size": H})
        self.policy.to(device)
        self.agent = sg.PPO(self.policy, h["clip_param"], h["ppo_epoch"], h["num_mini_batch"], h["value_loss_coef"],
                            h["entropy_coef"], lr=h["lr"], eps=h["eps"], max_grad_norm=h["max_grad_norm"])
        expert = expert_rows(c, seed)
        self.expert = expert.to(device)
        self.loader = DataLoader(TensorDataset(self.expert), batch_size=h["gail_batch"], shuffle=True,
                                 drop_last=len(expert) > h["gail_batch"])
        self.disc = sg.Discriminator(F, HD, device)
        self.rollouts = sg.RolloutStorage(T, N, (O,), Box(A), self.policy.recurrent_hidden_state_size, F)
        self.rollouts.to(device)
        self.ret_rms = sg.RunningMeanStd(shape=())
        self.S = T * N
        self.n_disc_batches = min(len(expert) // h["gail_batch"], self.S // h["gail_batch"])
        self.opt_steps = h["ppo_epoch"] * h["num_mini_batch"] + h["gail_epoch"] * self.n_disc_batches
        if dp:
            # "auto": shard an update only where that shortens its critical path (dist.DataParallel); SIMGAN_DP=always forces it
            sg_dist.attach(ppo=self.agent, disc=self.disc, policy=os.environ.get("SIMGAN_DP", "auto"))

        host, noise = host_rollout(c, seed, expert)
        # policy-produced fields (one batched act() over obs[:-1], as HOT LOOP A would have)
        with torch.no_grad():
            obs_dev = host["obs"].to(device)
            flat_obs = obs_dev[:-1].reshape(T * N, O)
            value, action, logp, _ = self.policy._forward_cuda(flat_obs, noise=noise.to(device))
        host["value_preds"] = torch.zeros(T + 1, N, 1)
        host["value_preds"][:-1] = value.view(T, N, 1).cpu()
        host["actions"] = action.view(T, N, A).cpu()
        host["action_log_probs"] = logp.view(T, N, 1).cpu()
        self.host = {k: v.contiguous().pin_memory() for k, v in host.items()}
        self.h2d_bytes = sum(v.numel() * 4 for v in self.host.values())
        self.upload()
        torch.cuda.synchronize()

    def upload(self):
        """Host (pinned) -> device copy of everything the env side of the loop produces."""
        r = self.rollouts
        for k, v in self.host.items():
            getattr(r, k).copy_(v, non_blocking=True)

    def update_phase(self, from_host, dropin_relabel=False):
        """One outer-iteration update phase through the public classes; returns the 6 loss floats.
        ``dropin_relabel``: run the reward relabel exactly as the unmodified caller does (main_gail_dyn_ppo.py:275-297:
        T x predict_reward_combined + host RunningMeanStd + clip, three host syncs per step) instead of the fused
        ``relabel_rollout`` entry point."""
        h, r = HYPER, self.rollouts
        if from_host:
            self.upload()
        with torch.no_grad():
            next_value = self.policy.get_value(r.obs[-1], r.recurrent_hidden_states[-1], r.masks[-1]).detach()
        for _ in range(h["gail_epoch"]):
            dl = self.disc.update_gail_dyn(self.loader, r)
        from simgan_b200.algo.gail import alive_bonus_offset
        r_sa = alive_bonus_offset(r.masks, self.c["T"], self.c["N"], h["gail_tar_length"])
        if dropin_relabel:
            gail_rewards = []
            for step in range(self.c["T"]):
                r.rewards[step], returns = self.disc.predict_reward_combined(r.obs_feat[step + 1], h["gamma"], r.masks[step],
                                                                             offset=-r_sa)
                self.ret_rms.update(returns.view(-1).cpu().numpy())
                rews = r.rewards[step].view(-1).cpu().numpy()
                rews = np.clip(rews / np.sqrt(self.ret_rms.var + 1e-7), -10.0, 10.0)
                r.rewards[step] = torch.Tensor(rews).view(-1, 1)
                gail_rewards.append(torch.mean(returns).cpu().data)
        else:
            self.disc.relabel_rollout(r, h["gamma"], -r_sa, self.ret_rms)
        r.compute_returns(next_value, True, h["gamma"], h["gae_lambda"], True)
        pl = self.agent.update(r)
        r.after_update()
        return tuple(float(x) for x in dl) + tuple(pl)


def fake_env_collection(w):
    """Collection loop (HOT LOOP A, main_gail_dyn_ppo.py:209-236) against a FAKE vec-env: the per-step env outputs
    are pre-generated NumPy arrays on the host (PyBullet / gym are not installed; no physics is simulated), so this
    measures only the feed boundary: staging + H2D + insert/act kernel + D2H + the per-step sync."""
    import simgan_b200 as sg
    c, r = w.c, w.rollouts
    T, N = c["T"], c["N"]
    steps = min(T, 512)
    obs = w.host["obs"].numpy(); feat = w.host["obs_feat"].numpy()
    done = (w.host["masks"].numpy()[:, :, 0] == 0); bad = (w.host["bad_masks"].numpy()[:, :, 0] == 0)
    rew = np.zeros(N, dtype=np.float32)
    feeder = sg.RolloutFeeder(w.policy, r)
    feeder.begin()
    for t in range(8):
        feeder.step(obs[t + 1], rew, done[t + 1], bad[t + 1], feat[t + 1])
    torch.cuda.synchronize()
    r.step = 0
    t0 = time.perf_counter()
    feeder.begin()
    for t in range(steps):
        feeder.step(obs[t + 1], rew, done[t + 1], bad[t + 1], feat[t + 1])
    torch.cuda.synchronize()
    fused = time.perf_counter() - t0
    # the same loop through the drop-in calls (Policy.act + RolloutStorage.insert), as an unmodified caller would run it
    r.step = 0
    t0 = time.perf_counter()
    for t in range(steps):
        with torch.no_grad():
            value, action, logp, hxs = w.policy.act(r.obs[t], r.recurrent_hidden_states[t], r.masks[t])
        action.cpu().numpy()
        o = torch.from_numpy(obs[t + 1]).float().to(w.device)
        m = torch.from_numpy(1.0 - done[t + 1].astype(np.float32)).unsqueeze(1)
        b = torch.from_numpy(1.0 - bad[t + 1].astype(np.float32)).unsqueeze(1)
        r.insert(o, hxs, action, logp, value, torch.from_numpy(rew).unsqueeze(1), m, b, torch.from_numpy(feat[t + 1]))
    torch.cuda.synchronize()
    dropin = time.perf_counter() - t0
    r.step = 0
    w.upload()
    torch.cuda.synchronize()
    return {"value": steps * N / fused, "unit": "env steps/s", "us_per_vec_step": 1e6 * fused / steps,
            "dropin_act_insert_value": steps * N / dropin, "dropin_us_per_vec_step": 1e6 * dropin / steps,
            "num_envs": N, "vec_steps_timed": steps,
            "note": "FAKE vec-env (pre-generated host arrays, no PyBullet physics): feed-boundary cost only"}


def sample_clocks_start():
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    try:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=f, stderr=subprocess.DEVNULL)
    except OSError:
        return None, f
    return p, f


def sample_clocks_stop(p, f, gpu_index):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if p is None:
        return out
    time.sleep(0.15)
    p.terminate()
    try:
        p.wait(timeout=5)
    except subprocess.TimeoutExpired:
        p.kill()
    f.flush()
    f.seek(0)
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for line in f.read().splitlines():
        parts = [x.strip() for x in line.split(",")]
        if len(parts) < 9:
            continue
        try:
            if int(parts[0]) != gpu_index:
                continue
            sm.append(float(parts[1]))
            mx.append(float(parts[2]))
        except ValueError:
            continue
        for nm, val in zip(names, parts[5:9]):
            if val.lower().startswith("active"):
                reasons.add(nm)
    f.close()
    os.unlink(f.name)
    if sm:
        out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
    return out


def run_cuda(args):
    import torch.distributed as dist
    from simgan_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON record): library banners (e.g. "NCCL version ...") printed while
    # the benchmark runs are routed to stderr at the file-descriptor level.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    torch.set_num_threads(1)                               # main_gail_dyn_ppo.py:64
    c = CONFIGS[args.config]
    w = Workload(c, args.seed, device, dp=world > 1)
    lib = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, from_host, with_timer, dropin_relabel=False):
        """Sum of per-step device times (ms) over n steps; L2 flushed between steps outside the events."""
        total = 0.0
        each = []
        _lib.timer.enabled = with_timer
        _lib.timer.records = []
        barrier()
        t_wall = time.perf_counter()
        for _ in range(n):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            losses = w.update_phase(from_host, dropin_relabel)
            e.record()
            torch.cuda.synchronize()
            each.append(s.elapsed_time(e))
            total += each[-1]
        timed.last_each = each
        barrier()
        wall = time.perf_counter() - t_wall
        _lib.timer.enabled = False
        kt = _lib.timer.drain()
        if world > 1:
            t = torch.tensor([total], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, wall, kt, losses

    for _ in range(max(args.warmup, 3)):
        w.update_phase(False)
    torch.cuda.synchronize()

    # ---- timed region 1: buffer resident in HBM --------------------------------------------------------
    cp, cf = sample_clocks_start() if rank == 0 else (None, None)
    l0 = lib.sg_launch_count()
    ms_total, wall, _, losses = timed(args.steps, False, False)            # per-kernel event timer OFF for the headline
    launches = lib.sg_launch_count() - l0
    ms_each = [round(x, 3) for x in timed.last_each]
    clocks = sample_clocks_stop(cp, cf, local_rank) if rank == 0 else {}
    # ---- timed region 2: end to end from pinned host buffers ------------------------------------------
    w.update_phase(True)
    ms_e2e, _, _, _ = timed(args.steps, True, False)
    # ---- separate pass with the per-kernel event timer on: kernel shares / durations for the roofline ----------------
    n_kt = max(3, args.steps // 4)
    ms_kt, _, kt, _ = timed(n_kt, False, True)
    # ---- the update phase as an UNMODIFIED caller runs it (per-step predict_reward_combined relabel loop) ---------
    n_di = max(2, min(args.steps, 5))
    ms_dropin, _, _, _ = timed(n_di, True, False, dropin_relabel=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    env_steps = fake_env_collection(w) if world == 1 else None
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))      # kernel timed inside a long step
    peak_src = "measured" if peaks else "fallback"

    ms_per_step = ms_total / args.steps
    work = algorithmic_work(c, w.n_disc_batches)
    kernels = {}
    for name, arr in kt.items():
        kernels[name] = dict(launches_per_step=len(arr) / n_kt, ms_per_launch=float(np.mean(arr)),
                             ms_per_step=float(np.sum(arr)) / n_kt)
    ms_per_step_kt = ms_kt / n_kt
    # ---- roofline of EVERY kernel; the dominant one (largest share of the step) is the record's ``roofline`` ---------
    sm_mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    sms = int(lib.sg_device_sm_count()) or 148
    fma_peak = sms * 128 * 2 * sm_mhz * 1e6 / 1e12
    tf32_peak = tc_peak / 2.0                   # kind::tf32 issues at half the bf16 rate; 3xTF32 spends 3 MMAs per product
    ppo_on_tensor_cores = bool(getattr(w.agent, "last_kernel", "") == "tensor")
    per_launch = {"ppo_update": (work["ppo_flops"], work["ppo_bytes"], HYPER["ppo_epoch"], HYPER["num_mini_batch"]),
                  "disc_update": (work["disc_flops"], work["disc_bytes"], HYPER["gail_epoch"], w.n_disc_batches),
                  "disc_relabel": (work["relabel_flops"], work["relabel_bytes"], 1, 1),
                  "compute_returns": (0, work["gae_bytes"], 1, 1)}
    traffic_all = {}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        traffic_all = json.load(open(tr_path)).get(args.config, {})
    roofs = {}
    for name, kd in kernels.items():
        if name not in per_launch:
            continue
        fl, by, nl, steps_in_launch = per_launch[name]
        fl, by = fl / world, by / world             # data parallel: a rank processes 1/world of every minibatch
        t_s = kd["ms_per_launch"] * 1e-3
        if name == "compute_returns":
            ach = by / nl / t_s / 1e9
            r = dict(bound="hbm", achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak,
                     engine="streaming scan (cp.async ring, one serial recurrence per env column)")
        else:
            ach = fl / nl / t_s / 1e12
            on_tc = name == "ppo_update" and ppo_on_tensor_cores
            r = dict(bound="tensor", achieved=ach, peak=tc_peak, unit="TFLOP/s", frac=ach / tc_peak,
                     hbm_achieved_gbs=by / nl / t_s / 1e9, hbm_peak_gbs=hbm_peak, hbm_frac=by / nl / t_s / 1e9 / hbm_peak,
                     engine=("tcgen05 kind::tf32, 3xTF32 operand splitting (3 MMAs per fp32 product), accumulators in TMEM"
                             if on_tc else "fp32 FMA on CUDA cores (no tensor instruction issued), latency-bound serial chain"),
                     fp32_fma_peak_tflops=fma_peak, frac_of_fp32_fma_peak=ach / fma_peak)
            if on_tc:
                r.update(tf32_3x_peak_tflops=tf32_peak / 3.0, frac_of_3xtf32_peak=ach / (tf32_peak / 3.0))
        r.update(kernel=name, peak_source=peak_src, traffic=traffic_all.get(name),
                 share_of_step=kd["ms_per_step"] / ms_per_step_kt,
                 algorithmic_bytes_per_launch=by / nl, algorithmic_flops_per_launch=fl / nl,
                 us_per_optimizer_step=1e3 * kd["ms_per_launch"] / steps_in_launch)
        roofs[name] = r
    dom = max(roofs, key=lambda k: kernels[k]["ms_per_step"]) if roofs else None
    roof = None
    if dom is not None:
        roof = dict(roofs[dom])
        roof["note"] = ("the bound key keeps the contract's vocabulary; what the arithmetic runs on is in `engine`. "
                        "One outer iteration is a serial chain of dependent optimizer steps (grid barriers between phases): "
                        "at cfg1-3 neither roofline binds, see serial_chain")
        try:
            if dom in ("disc_update", "ppo_update") and not c.get("split"):
                ph = dict(zip(["image", "tile", "bar1", "reduce_adam", "bar2"], w.disc.phase_cycles())) if dom == "disc_update" else \
                    dict(zip(["image", "tile", "bar1", "reduce_ssq", "bar2", "clip_adam", "bar3"], w.agent.phase_cycles()))
                sync = sum(v for k, v in ph.items() if k.startswith("bar") or k == "image")
                steps_in_launch = per_launch[dom][3]
                roof["serial_chain"] = dict(steps_per_launch=steps_in_launch,
                                            us_sync_and_refresh_per_step=sync / steps_in_launch / sm_mhz,
                                            us_tile_phase_per_step=ph["tile"] / steps_in_launch / sm_mhz)
        except Exception:
            pass
    out = {
        "metric": "PPO+GAIL update-steps/sec", "value": w.opt_steps * args.steps / (ms_total * 1e-3),
        "unit": "optimizer steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic rollout (seeded N(0,1) obs, expert-resampled obs_feat) + "
                                + ({"hopper": "real Hopper expert rows (hopper_new11_deform_n200_3.pkl, tests/golden fixture)",
                                   "laika": "real Laikago expert rows (laika_70_deform_n200_0.pkl, tests/golden fixture)"}.get(
                                       c["expert"], "N(0,1) expert rows")),
        "config": {"workload": c["label"], "name": args.config, "T": c["T"], "N": c["N"], "obs_dim": c["O"], "act_dim": c["A"],
                   "hidden": c["H"], "feat_dim": c["F"], "disc_hidden": c["HD"], "ppo_epoch": HYPER["ppo_epoch"],
                   "num_mini_batch": HYPER["num_mini_batch"], "gail_epoch": HYPER["gail_epoch"],
                   "gail_batch": HYPER["gail_batch"], "optimizer_steps_per_step": w.opt_steps,
                   "parallelism": "single GPU" if world == 1 else
                   "dp%d (%s policy): PPO update %s, discriminator update %s; gradient exchange inside the step kernel over NVLink peer "
                   "memory (%s transport)" % (world, w.agent.dp.policy,
                                              "sharded by minibatch rows" if getattr(w.agent, "last_sharded", False) else
                                              "replicated (its tiles already run concurrently on one GPU)",
                                              "sharded" if w.disc.__dict__.get("last_sharded") else
                                              "replicated (128 row triples = 128 tiles <= SMs)", w.agent.dp.transport),
                   "l2": "flushed between timed steps (256 MiB write)", "timing": "cuda events per step, max over ranks"},
        "iters_per_s": args.steps / (ms_total * 1e-3),
        "samples_per_s": HYPER["ppo_epoch"] * w.S * args.steps / (ms_total * 1e-3),
        "e2e": {"value": w.opt_steps * args.steps / (ms_e2e * 1e-3), "unit": "optimizer steps/s",
                "h2d_bytes_per_step": w.h2d_bytes + 4 * (HYPER["ppo_epoch"] * w.S + HYPER["gail_epoch"] * w.n_disc_batches * HYPER["gail_batch"] * 3),
                "d2h_bytes_per_step": 4 * (4 * HYPER["ppo_epoch"] * HYPER["num_mini_batch"] + 3 * HYPER["gail_epoch"] * w.n_disc_batches) + 24 + 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "env_steps": env_steps,
        "kernels": kernels, "roofline": roof, "roofline_all": roofs, "clocks": clocks,
        "update_phase_dropin": {"value": w.opt_steps * n_di / (ms_dropin * 1e-3), "unit": "optimizer steps/s",
                                "ms_per_step": ms_dropin / n_di, "steps_timed": n_di,
                                "note": "same update phase from pinned host buffers, reward relabel through the UNMODIFIED caller loop "
                                        "(main_gail_dyn_ppo.py:275-297: T x predict_reward_combined + host RunningMeanStd, 3 host syncs "
                                        "per step); `e2e` uses the fused relabel_rollout entry point instead"},
        "losses_last_step": [float(x) for x in losses],
        "phase_cycles_last_launch": {"ppo": dict(zip(["image", "tile", "bar1", "reduce_ssq", "bar2", "clip_adam", "bar3"],
                                                     w.agent.phase_cycles())),
                                     "disc": dict(zip(["image", "tile", "bar1", "reduce_adam", "bar2"], w.disc.phase_cycles()))},
        "host_wall_ms_per_step": 1e3 * wall / args.steps,
        "ms_each_timed_step_rank0": ms_each,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = run_reference(args, quiet=True, budget_s=20.0)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(out))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# reference arm (CPU): the oracle restatement of the reference's eager path
# ------------------------------------------------------------------------------------------------------
def reference_sample(c, seed, threads):
    """Build the same workload on the CPU and return a callable that runs ONE bounded sample:
    1 of 5 discriminator epochs + the whole relabel + GAE + 2 of 10 PPO epochs, i.e. exactly 1/5 of the
    optimizer steps of a full iteration in the workload's own D:PPO proportion.  relabel+GAE are needed
    in full to make the PPO inputs valid, so their time is scaled by 1/5 in the returned seconds."""
    from oracle import ppo_gail_oracle as orc
    h = HYPER
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    T, N, O, A, H, F, HD = c["T"], c["N"], c["O"], c["A"], c["H"], c["F"], c["HD"]
    split = bool(c.get("split"))
    pol = orc.init_split_policy(O, H, A // 7) if split else orc.init_policy(O, H, A)
    expert = expert_rows(c, seed)
    dparams = orc.init_disc(F, HD)
    host, noise = host_rollout(c, seed, expert)
    buf = orc.new_buffer(T, N, O, A, F)
    for k, v in host.items():
        buf[k].copy_(v)
    v, a, lp = (orc.split_act if split else orc.policy_act)(pol, buf["obs"][:-1].reshape(T * N, O), noise=noise)
    buf["value_preds"][:-1] = v.view(T, N, 1)
    buf["actions"].copy_(a.view(T, N, A))
    buf["action_log_probs"].copy_(lp.view(T, N, 1))
    hyper = orc.PPOHyper(ppo_epoch=2, num_mini_batch=h["num_mini_batch"], entropy_coef=h["entropy_coef"])
    ppo = (orc.PPOOracle(pol, hyper, keys=orc.SPLIT_KEYS, evaluate=orc.split_evaluate) if split else orc.PPOOracle(pol, hyper))
    fwd = orc.split_forward if split else orc.policy_forward
    disc = orc.DiscOracle(dparams)
    rms = orc.RunningMeanStd(shape=())
    n_db = min(len(expert) // h["gail_batch"], T * N // h["gail_batch"])
    steps = n_db + 2 * h["num_mini_batch"]

    def one():
        t0 = time.perf_counter()
        nv = fwd(ppo.params(), buf["obs"][-1])[0]
        disc.update_epoch(expert, buf, batch_size=h["gail_batch"], drop_last=len(expert) > h["gail_batch"])
        t1 = time.perf_counter()
        r_sa = orc.alive_bonus_offset(buf["masks"], T, N, h["gail_tar_length"])
        orc.relabel_rewards(disc, rms, buf, h["gamma"], -r_sa)
        orc.compute_returns(buf, nv, True, h["gamma"], h["gae_lambda"], True)
        t2 = time.perf_counter()
        ppo.update(buf)
        t3 = time.perf_counter()
        return (t1 - t0) + (t2 - t1) / 5.0 + (t3 - t2), dict(disc=t1 - t0, relabel_gae=t2 - t1, ppo=t3 - t2)

    return one, steps


def reference_sample_real(c, seed, threads):
    """The same bounded sample as ``reference_sample`` driven through the UNMODIFIED reference modules
    (third_party/a2c_ppo_acktr/{model,storage,algo/ppo,algo/gail}.py imported through oracle/ref_shim.py).  Only possible
    where a reference tree is mounted (the dev container; never the GPU box)."""
    from oracle import ref_shim
    from torch.utils.data import DataLoader, TensorDataset
    ref = ref_shim.load()
    h = HYPER
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    T, N, O, A, H, F, HD = c["T"], c["N"], c["O"], c["A"], c["H"], c["F"], c["HD"]
    space = ref_shim.BoxSpace(A)
    if c.get("split"):
        pol = ref.model_split.SplitPolicy((O,), space, base_kwargs={"hidden_size": H, "num_feet": A // 7})
    else:
        pol = ref.model.Policy((O,), space, base_kwargs={"recurrent": False, "hidden_size": H})
    agent = ref.ppo.PPO(pol, h["clip_param"], 2, h["num_mini_batch"], h["value_loss_coef"], h["entropy_coef"], lr=h["lr"],
                        eps=h["eps"], max_grad_norm=h["max_grad_norm"])
    expert = expert_rows(c, seed)
    loader = DataLoader(TensorDataset(expert), batch_size=h["gail_batch"], shuffle=True, drop_last=len(expert) > h["gail_batch"])
    disc = ref.gail.Discriminator(F, HD, torch.device("cpu"))
    rollouts = ref.storage.RolloutStorage(T, N, (O,), space, pol.recurrent_hidden_state_size, F)
    host, _ = host_rollout(c, seed, expert)
    for k, v in host.items():
        getattr(rollouts, k).copy_(v)
    with torch.no_grad():
        value, action, logp, _ = pol.act(rollouts.obs[:-1].reshape(T * N, O), rollouts.recurrent_hidden_states[:-1].reshape(T * N, -1),
                                         rollouts.masks[:-1].reshape(T * N, 1))
    rollouts.value_preds[:-1].copy_(value.view(T, N, 1))
    rollouts.actions.copy_(action.view(T, N, A))
    rollouts.action_log_probs.copy_(logp.view(T, N, 1))
    ret_rms = ref.rms.RunningMeanStd(shape=())
    n_db = min(len(expert) // h["gail_batch"], T * N // h["gail_batch"])
    steps = n_db + 2 * h["num_mini_batch"]

    def one():
        t0 = time.perf_counter()
        with torch.no_grad():
            next_value = pol.get_value(rollouts.obs[-1], rollouts.recurrent_hidden_states[-1], rollouts.masks[-1]).detach()
        disc.update_gail_dyn(loader, rollouts)
        t1 = time.perf_counter()
        n_done = (1.0 - rollouts.masks).sum().cpu().numpy() + N / 2              # main_gail_dyn_ppo.py:258-271
        d_sa = 1 - n_done / (n_done + (T * N) / h["gail_tar_length"])
        r_sa = np.log(d_sa) - np.log(1 - d_sa)
        for step in range(T):                                                    # main_gail_dyn_ppo.py:275-297
            rollouts.rewards[step], returns = disc.predict_reward_combined(rollouts.obs_feat[step + 1], h["gamma"],
                                                                           rollouts.masks[step], offset=-r_sa)
            ret_rms.update(returns.view(-1).cpu().numpy())
            rews = rollouts.rewards[step].view(-1).cpu().numpy()
            rews = np.clip(rews / np.sqrt(ret_rms.var + 1e-7), -10.0, 10.0)
            rollouts.rewards[step] = torch.Tensor(rews).view(-1, 1)
        rollouts.compute_returns(next_value, True, h["gamma"], h["gae_lambda"], True)
        t2 = time.perf_counter()
        agent.update(rollouts)
        t3 = time.perf_counter()
        return (t1 - t0) + (t2 - t1) / 5.0 + (t3 - t2), dict(disc=t1 - t0, relabel_gae=t2 - t1, ppo=t3 - t2)

    return one, steps


def run_reference(args, quiet=False, budget_s=None):
    c = CONFIGS[args.config]
    ncores = os.cpu_count() or 1
    kind = "port"
    sampler = reference_sample
    try:
        from oracle import ref_shim
        if ref_shim.available() and os.environ.get("SIMGAN_BENCH_PORT", "0") != "1":
            kind, sampler = "reference", reference_sample_real
    except Exception:
        pass
    # the reference itself runs with one intra-op thread (main_gail_dyn_ppo.py:64); try all cores too and
    # keep whichever is faster for this workload
    best = None
    for threads in sorted({1, ncores}):
        one, steps = sampler(c, args.seed, threads)
        one()                             # warm-up
        t = min(one()[0], one()[0])       # probe
        if best is None or t < best[0]:
            best = (t, threads)
    threads = best[1]
    one, steps = sampler(c, args.seed, threads)
    n_warm = 1 if quiet else max(1, min(args.warmup, 3))
    n_steps = args.steps
    if budget_s is not None:
        n_steps = max(1, min(args.steps, int(budget_s / max(best[0], 1e-3))))
    for _ in range(n_warm):
        one()
    tot, parts = 0.0, []
    for _ in range(n_steps):
        t, p = one()
        tot += t
        parts.append(p)
    value = steps * n_steps / tot
    sample = ("1/5 outer iteration per step (1 of 5 D epochs + relabel + GAE + 2 of 10 PPO epochs = %d optimizer steps; "
              "relabel+GAE time scaled by 1/5), %d steps timed, torch CPU eager, %d intra-op thread(s) "
              "(faster of 1 and %d); %s" % (steps, n_steps, threads, ncores,
                                           "the reference's own modules imported through oracle/ref_shim.py" if kind == "reference"
                                           else "oracle port (no reference tree on this box)"))
    cpu = {"value": value, "unit": "optimizer steps/s", "cores": threads, "kind": kind, "sample": sample,
           "host_cores": ncores, "s_per_sample": tot / n_steps,
           "phase_s": {k: float(np.mean([p[k] for p in parts])) for k in parts[0]}}
    if quiet:
        return cpu
    out = {"impl": "reference", "metric": "PPO+GAIL update-steps/sec", "value": value, "unit": "optimizer steps/s",
           "n_gpus": args.gpus, "steps": n_steps, "warmup": n_warm, "ms_per_step": 1e3 * tot / n_steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
           "data": "same synthetic workload as the CUDA arm",
           "config": {"workload": c["label"], "name": args.config, "optimizer_steps_per_step": steps},
           "cpu_baseline": cpu,
           "e2e": {"value": value, "unit": "optimizer steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    HYPER.update(CONFIGS[args.config].get("hyper", {}))
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference(args)
        return
    run_cuda(args)


if __name__ == "__main__":
    main()
