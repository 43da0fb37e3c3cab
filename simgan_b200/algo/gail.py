"""GAIL discriminator with the reference's surface (third_party/a2c_ppo_acktr/algo/gail.py:34-217).

``update_gail_dyn`` / ``predict_reward_combined`` keep their signatures and return types; the
minibatch loop (gather expert + policy rows, three trunk forwards, two BCE-with-logits terms, the
gradient penalty with a hand-derived double backward, Adam) runs in the sm_100a kernel behind
``sg_disc_update``.  Index streams and the mixup alphas are drawn on the host from the CPU default
generator in exactly the order ``zip(DataLoader, feed_forward_generator)`` consumes it
(SURVEY.md section 8g-2), then shipped to the device once per epoch.

``relabel_rollout`` is the whole-rollout form of the caller's T-step relabel loop
(main_gail_dyn_ppo.py:275-297): one device pass instead of T launches with three host syncs each.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lazy, _lib, _spec, host_sampler
from ..running_mean_std import RunningMeanStd
from .adam import FusedAdam


class Discriminator(nn.Module):
    def __init__(self, input_dim, hidden_dim, device):
        super().__init__()
        self.device = device
        # default nn.Linear init, three layers in this order (gail.py:40-43)
        self.trunk = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.Tanh(),
                                   nn.Linear(hidden_dim, hidden_dim), nn.Tanh(),
                                   nn.Linear(hidden_dim, 1)).to(device)
        self.trunk.train()
        self.optimizer = FusedAdam(self.trunk.parameters())       # Adam defaults lr 1e-3, eps 1e-8 (gail.py:48)
        self.returns = None
        self.ret_rms = RunningMeanStd(shape=())
        self.kernel_mode = 0
        self.dp = None
        self.lazy_losses = True          # update_gail_dyn returns LazyFloat scalars (no blocking read-back per call)
        self.__dict__["_pending"] = None
        self.__dict__["_ws"] = None

    @property
    def last_trace(self):
        """(n_steps, 3) per-step {loss, expert_loss, policy_loss} of the last update call (host tensor)."""
        pend = self.__dict__.get("_pending")
        return None if pend is None else pend.trace()

    @last_trace.setter
    def last_trace(self, value):
        self.__dict__["_pending"] = None if value is None else _lazy.HostTrace(value)

    # ---- flat parameter buffer ---------------------------------------------------------------------------
    @property
    def feat_dim(self):
        return self.trunk[0].in_features

    @property
    def hidden_dim(self):
        return self.trunk[0].out_features

    def hot_path_parameters(self):
        t = self.trunk
        return [t[0].weight, t[0].bias, t[2].weight, t[2].bias, t[4].weight, t[4].bias]

    def flat_params(self):
        ps = self.hot_path_parameters()
        dev = ps[0].device
        if dev.type != "cuda":
            raise _lib.SgError("the GAIL discriminator hot path needs a CUDA device (got %s); no CPU fallback" % dev)
        offs, total = _lib.disc_layout(self.feat_dim, self.hidden_dim)
        flat = self.__dict__.get("_flat")
        bound = (flat is not None and flat.device == dev and flat.numel() == total and
                 all(p.dtype == torch.float32 and p.is_contiguous() and p.data_ptr() == flat.data_ptr() + 4 * o
                     for p, o in zip(ps, offs)))
        if not bound:
            flat = torch.zeros(total, device=dev, dtype=torch.float32)
            for p, o in zip(ps, offs):
                n = p.numel()
                flat[o:o + n].copy_(p.data.reshape(-1))
                p.data = flat[o:o + n].view(p.shape)
            self.__dict__["_flat"] = flat
        return flat

    def __getstate__(self):
        """Whole-object ``torch.save(discr)`` (main_gail_dyn_ppo.py:319-320): only what the reference's object holds
        travels.  Staging buffers, workspaces and the data-parallel handle (ctypes pointers to CUDA IPC mappings, a
        process group) are per-process state and are dropped; ``dist.attach`` re-creates the handle after a load."""
        state = dict(self.__dict__)
        state.pop("_flat", None)
        state["_ws"] = None
        state["dp"] = None
        state["_pending"] = None
        state.pop("last_trace", None)
        state.pop("_prof_view", None)
        state.pop("_rms_dev", None)
        state.pop("_predraw", None)
        for k in ("_stage", "_stage_cur", "_side_upload", "_copy_stream", "_last_key", "_relabel_ws", "_relabel_raw"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        """Also accepts a pickle written by the REFERENCE's Discriminator (loaded through ``compat.install()``): its
        ``__dict__`` has none of the kernel-side attributes and its optimizer is a ``torch.optim.Adam``, which is
        replaced by a FusedAdam carrying over lr / betas / eps and, when present, the moment estimates and step."""
        self.__dict__.update(state)
        d = self.__dict__
        d.setdefault("kernel_mode", 0)
        d.setdefault("dp", None)
        d.pop("last_trace", None)
        d.setdefault("_pending", None)
        d.setdefault("lazy_losses", True)
        d.setdefault("_ws", None)
        d.setdefault("returns", None)
        if "ret_rms" not in d:
            d["ret_rms"] = RunningMeanStd(shape=())
        opt = d.get("optimizer")
        if opt is not None and not isinstance(opt, FusedAdam):
            d["optimizer"] = FusedAdam.from_torch_adam(opt, list(self.trunk.parameters()))

    # ---- update --------------------------------------------------------------------------------------------
    @staticmethod
    def draw_epoch_indices(n_expert, batch_size, drop_last, n_rollout):
        """Host RNG emulation of one ``zip(expert_loader, feed_forward_generator)`` pass.

        Order of CPU-default-generator draws (torch 2.x): DataLoader base seed (one int64 random_()),
        RandomSampler seed (one int64 random_()), expert permutation on a PRIVATE generator, the rollout
        sampler's randperm(S), then one rand(B,1) per zipped minibatch (gail.py:72).
        Returns (expert_idx (n,B) int64, policy_idx (n,B) int64, alpha (n,B) fp32)."""
        # configurations that cannot run are refused BEFORE anything is drawn, so a caller that catches the error
        # finds the generator where it left it
        n_e = n_expert // batch_size
        if not drop_last and n_expert % batch_size:
            raise NotImplementedError(
                "expert set of %d rows with gail_batch_size=%d and drop_last=False yields a ragged expert batch that "
                "the reference's mixup pairs with a full policy batch (SURVEY.md section 8g-1); not a runnable "
                "configuration" % (n_expert, batch_size))
        if n_e == 0:
            raise ZeroDivisionError("no full expert minibatch (gail.py:193 divides by n=0)")
        n = min(n_e, n_rollout // batch_size)
        if n == 0:
            raise ZeroDivisionError("rollout smaller than gail_batch_size (gail.py:193 divides by n=0)")
        torch.empty((), dtype=torch.int64).random_()
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        gen = torch.Generator()
        gen.manual_seed(seed)
        e_perm = torch.randperm(n_expert, generator=gen)
        p_perm = host_sampler.randperm_prefix(n_rollout, n * batch_size)     # == torch.randperm(n_rollout)[:n*B], same consumption
        alpha = torch.rand(n * batch_size)          # == n consecutive torch.rand(B,1) draws (tests pin this)
        return (e_perm[:n * batch_size].view(n, batch_size), p_perm.view(n, batch_size),
                alpha.view(n, batch_size))

    # ---- early index draws (simgan_b200/_spec.py) -----------------------------------------------------------------
    # The caller runs gail_epoch back-to-back epochs (main_gail_dyn_ppo.py:255-256).  While a kernel runs, the host
    # draws the streams the NEXT call would draw, stages them in the idle one of two pinned blocks and uploads that
    # block on a side stream -- then rewinds the CPU generator.  The next call uses the staged block only if the
    # generator is still exactly where the early draw started and the Adam schedule it baked in is still the right
    # one; otherwise it is discarded and the call draws / stages as usual.
    @_lib.on_device(lambda self, *a, **k: self.trunk[0].weight.device)
    def _speculate(self):
        key = self.__dict__.get("_last_key")
        flat = self.__dict__.get("_flat")
        if key is None or flat is None or _spec.still_valid(self.__dict__.get("_predraw"), key):
            return
        slot = _spec.predraw(key, lambda: self.draw_epoch_indices(*key))
        if slot is None:
            return
        e_idx, p_idx, alpha = slot[3]
        staged = self._fill_stage(flat.device, e_idx, p_idx, alpha, side_stream=True)
        self.__dict__["_predraw"] = slot[:3] + ((e_idx, p_idx, alpha, staged),)

    def _take_or_draw(self, key):
        _spec.consumed(self)
        self.__dict__["_last_key"] = key
        got = _spec.take(self.__dict__.pop("_predraw", None), key)
        return got if got is not None else self.draw_epoch_indices(*key) + (None,)

    def _sched_sig(self, n):
        g = self.optimizer.param_groups[0]
        return (self.optimizer.step_count + 1, n, g["lr"], tuple(g["betas"]))

    def _fill_stage(self, dev, e_idx, p_idx, alpha, side_stream=False):
        """Fill the idle staging pair and start its upload:
        int32 words [expert_idx (n,B) | policy_idx (n,B) | alpha bits (n,B) | Adam step sizes (n) | sqrt(1-beta2^t) (n)].
        Returns (pair index, schedule signature, upload-done event or None)."""
        n, B = e_idx.shape
        nb = n * B
        words = 3 * nb + 2 * n
        pairs = self.__dict__.setdefault("_stage", [None, None])
        which = 1 - self.__dict__.get("_stage_cur", 1)
        pending = self.__dict__.pop("_side_upload", None)
        if pending is not None:
            pending.synchronize()       # an abandoned early upload may still be reading / writing this pair
        st = pairs[which]
        if st is None or st[0].numel() != words or st[1].device != dev:
            st = (torch.empty(words, dtype=torch.int32).pin_memory(), torch.empty(words, dtype=torch.int32, device=dev))
            pairs[which] = st
        stage, stage_dev = st
        stage[:nb].view(n, B).copy_(e_idx)
        stage[nb:2 * nb].view(n, B).copy_(p_idx)
        stage[2 * nb:3 * nb].view(torch.float32).view(n, B).copy_(alpha)
        stage[3 * nb:].view(torch.float32).copy_(torch.from_numpy(self.optimizer.schedule(n)).reshape(-1))
        event = None
        if side_stream:
            cs = self.__dict__.get("_copy_stream")
            if cs is None or cs.device != dev:
                cs = torch.cuda.Stream(dev)
                self.__dict__["_copy_stream"] = cs
            with torch.cuda.stream(cs):
                stage_dev.copy_(stage, non_blocking=True)
                event = cs.record_event()
            self.__dict__["_side_upload"] = event
        else:
            stage_dev.copy_(stage, non_blocking=True)
        return which, self._sched_sig(n), event

    @_lib.on_device(lambda self, *a, **k: self.trunk[0].weight.device)
    def _run_update(self, expert, policy_feat, e_idx, p_idx, alpha, staged=None):
        flat = self.flat_params()
        dev = flat.device
        n, B = e_idx.shape
        opt = self.optimizer
        m, v = opt.ensure_state(flat)
        g = opt.param_groups[0]
        cfg = _lib.DiscConfig()
        cfg.feat_dim, cfg.hidden, cfg.batch_size, cfg.n_steps = self.feat_dim, self.hidden_dim, B, n
        cfg.gp_lambda = 10.0
        cfg.beta1, cfg.beta2, cfg.adam_eps = g["betas"][0], g["betas"][1], g["eps"]
        cfg.first_adam_step = opt.step_count + 1
        dp = self.dp if (self.dp is not None and self.dp.shards(B, 1)) else None        # "auto": replicate small batches
        self.__dict__["last_sharded"] = dp is not None
        cfg.row_begin, cfg.row_end = (0, B) if dp is None else dp.shard(B)
        p2p = dp is not None and dp.p2p_ok(B)
        cfg.mode = self.kernel_mode if (dp is None or p2p) else 1
        if cfg.mode == 1 and p2p:
            cfg.mode = 0
        cfg.dp_ctx = dp.context("disc", flat.numel()) if p2p else None
        lib = _lib.lib()
        need = lib.sg_disc_workspace_bytes(C.byref(cfg))
        if need < 0:
            _lib.check(1, "sg_disc_workspace_bytes")
        ws = self.__dict__.get("_ws")
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
            self.__dict__["_ws"] = ws
        # ONE pinned staging block and ONE async H2D copy per call -- already done when the draw was made early
        pairs = self.__dict__.get("_stage")
        if (staged is not None and staged[1] == self._sched_sig(n) and pairs is not None
                and pairs[staged[0]] is not None and pairs[staged[0]][1].device == dev):
            torch.cuda.current_stream().wait_event(staged[2])
            which = staged[0]
        else:
            which = self._fill_stage(dev, e_idx, p_idx, alpha)[0]
        self.__dict__["_stage_cur"] = which
        stage_dev = self.__dict__["_stage"][which][1]
        nb = n * B
        idx_dev = stage_dev[:2 * nb].view(2, n, B)
        alpha_dev = stage_dev[2 * nb:3 * nb].view(torch.float32)
        sched = stage_dev[3 * nb:].view(torch.float32).view(2, n)
        trace = torch.empty(n, 3, device=dev)
        cb, user = _lib.NULL_ALLREDUCE, None
        if dp is not None and not p2p:
            cb = dp.make_callback(ws)
        tok = _lib.timer.start("disc_update")
        rc = lib.sg_disc_update(C.byref(cfg), _lib.ptr(flat), _lib.ptr(m), _lib.ptr(v), _lib.ptr(expert),
                                _lib.ptr(policy_feat), _lib.ptr(idx_dev[0]), _lib.ptr(idx_dev[1]), _lib.ptr(alpha_dev),
                                _lib.ptr(sched[0]), _lib.ptr(sched[1]), _lib.ptr(trace), _lib.ptr(ws), cb, user,
                                _lib.current_stream())
        _lib.timer.stop(tok)
        _lib.check(rc, "sg_disc_update")
        opt.step_count += n
        self.__dict__["_prof_view"] = (ws, int(lib.sg_disc_phase_cycles_offset(C.byref(cfg))))
        if p2p:
            dp.sum_trace_(trace, 3)           # all three loss columns are per-rank partial sums
        # an earlier call whose trace has landed by now is checked here (non-finite losses raise), without waiting
        prev = self.__dict__.get("_pending")
        if prev is not None and prev.done():
            prev.means()
        pend = _lazy.PendingTrace(trace, 3, lambda msg: _lib.SgError("sg_disc_update: " + msg))
        self.__dict__["_pending"] = pend
        _spec.host_idle()                     # the next consumer of the CPU generator draws while the GPU runs this epoch
        out = tuple(_lazy.LazyFloat(pend, c) for c in range(3))
        if not self.lazy_losses:
            out = tuple(float(x) for x in out)
        return out

    def phase_cycles(self):
        """Diagnostics: per-phase SM-clock totals of CTA 0 of the last persistent launch
        {param image, tile phase, barrier 1, reduce+Adam, barrier 2}."""
        ws, off = self.__dict__["_prof_view"]
        return ws[off:off + 64].view(torch.int64)[:5].cpu().tolist()

    def phase_cycles_all(self, n_ctas):
        ws, off = self.__dict__["_prof_view"]
        return ws[off:off + 64 * n_ctas].view(torch.int64).view(n_ctas, 8).cpu()

    def update_gail_dyn(self, expert_loader, rollouts, replay=None):
        """gail.py:154-193.  ``expert_loader`` is the caller's DataLoader over TensorDataset(expert (N_exp,F))
        (main_gail_dyn_ppo.py:165-175); its dataset tensor is used in place on the device and its
        batch_size / drop_last drive the index emulation.  ``replay`` = (expert_idx, policy_idx, alpha)
        bypasses the generator (parity tests)."""
        if not self.training:
            self.train()
        self._check_loader(expert_loader)
        expert = expert_loader.dataset.tensors[0]
        if not expert.is_cuda or not rollouts.obs_feat.is_cuda:
            raise _lib.SgError("update_gail_dyn needs the expert set and the rollout buffer on the CUDA device")
        expert = expert if (expert.dtype == torch.float32 and expert.is_contiguous()) else expert.float().contiguous()
        T, N = rollouts.rewards.shape[:2]
        S = T * N
        F = rollouts.obs_feat.shape[-1]
        assert expert.shape[1] == F == self.feat_dim
        policy_feat = rollouts.obs_feat[1:].reshape(S, F)       # next_obs_feat rows (storage.py:172, gail.py:166)
        staged = None
        if replay is None:
            key = (expert.shape[0], expert_loader.batch_size, bool(expert_loader.drop_last), S)
            e_idx, p_idx, alpha, staged = self._take_or_draw(key)
        else:
            e_idx, p_idx, alpha = [x if torch.is_tensor(x) else torch.stack([torch.as_tensor(r).reshape(-1) for r in x])
                                   for x in replay]
        return self._run_update(expert, policy_feat, e_idx, p_idx, alpha, staged)

    @staticmethod
    def _check_loader(expert_loader):
        """The index emulation reproduces exactly one loader shape -- the caller's DataLoader(shuffle=True) on the default
        generator (main_gail_dyn_ppo.py:170-175).  Anything else would silently get that stream too: refuse it."""
        from torch.utils.data import RandomSampler
        sampler = getattr(expert_loader, "sampler", None)
        if not isinstance(sampler, RandomSampler) or getattr(sampler, "replacement", False):
            raise NotImplementedError("update_gail_dyn emulates DataLoader(shuffle=True) (RandomSampler without replacement); "
                                      "got sampler %r" % type(sampler).__name__)
        if getattr(expert_loader, "generator", None) is not None or getattr(sampler, "generator", None) is not None:
            raise NotImplementedError("update_gail_dyn emulates a DataLoader on the DEFAULT CPU generator; a loader with its own "
                                      "generator would consume a different stream")

    def update(self, expert_loader, rollouts, obsfilt=None, is_gail_dyn=False, a_dim=None):
        """Legacy (state, action)-split variant (gail.py:91-152): the D input is cat([state, action]) on
        both sides, so it maps onto the same kernel over concatenated row matrices."""
        if obsfilt is not None:
            raise NotImplementedError("obsfilt (VecNormalize observation filter) is not used by the GAIL-dyn path")
        self.train()
        self._check_loader(expert_loader)
        es, ea = expert_loader.dataset.tensors[:2]
        expert = torch.cat([es, ea], dim=1).float().contiguous()
        T, N = rollouts.rewards.shape[:2]
        S = T * N
        fr = rollouts.flat_rows()
        if not is_gail_dyn:
            policy_feat = torch.cat([fr["obs"], fr["actions"]], dim=1).contiguous()
        else:
            policy_feat = torch.cat([fr["obs_feat"], fr["obs"][:, -a_dim:], fr["next_obs_feat"]], dim=1).contiguous()
        assert expert.shape[1] == policy_feat.shape[1] == self.feat_dim
        e_idx, p_idx, alpha = self.draw_epoch_indices(expert.shape[0], expert_loader.batch_size,
                                                      bool(expert_loader.drop_last), S)
        return self._run_update(expert, policy_feat, e_idx, p_idx, alpha)

    # ---- differentiable penalty (API completeness; the update kernels derive it by hand) -------------------
    def compute_grad_pen_combined(self, expert_combined, policy_combined, lambda_=10.):
        alpha = torch.rand(expert_combined.size(0), 1).expand_as(expert_combined).to(expert_combined.device)
        mix = (alpha * expert_combined + (1 - alpha) * policy_combined).detach().requires_grad_(True)
        out = self.trunk(mix)
        (grad,) = torch.autograd.grad(out, mix, torch.ones_like(out), create_graph=True, retain_graph=True)
        return lambda_ * (grad.norm(2, dim=1) - 1).pow(2).mean()

    def compute_grad_pen(self, expert_state, expert_action, policy_state, policy_action, lambda_=10.):
        return self.compute_grad_pen_combined(torch.cat([expert_state, expert_action], dim=1),
                                              torch.cat([policy_state, policy_action], dim=1), lambda_)

    # ---- rewards ---------------------------------------------------------------------------------------------
    @_lib.on_device(lambda self, *a, **k: self.trunk[0].weight.device)
    def predict_reward_combined(self, d_in, gamma, masks, offset=0.0):
        """gail.py:201-210: (reward (N,1), running returns (N,1)); ``self.returns`` persists across calls."""
        flat = self.flat_params()
        x = d_in.detach()
        if not x.is_cuda:
            raise _lib.SgError("predict_reward_combined needs CUDA tensors")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        n = x.shape[0]
        reward = torch.empty(n, 1, device=x.device)
        has = self.returns is not None
        if not has:
            self.returns = torch.empty(n, 1, device=x.device)
        else:
            self.returns = self.returns.clone()     # the reference rebinds a fresh tensor each call
        mk = masks.detach().float().contiguous()
        rc = _lib.lib().sg_disc_predict_reward(_lib.ptr(flat), self.feat_dim, self.hidden_dim, _lib.ptr(x), n,
                                               float(gamma), _lib.ptr(mk), float(offset), int(has), _lib.ptr(reward),
                                               _lib.ptr(self.returns), _lib.current_stream())
        _lib.check(rc, "sg_disc_predict_reward")
        return reward, self.returns

    def predict_reward(self, state, action, gamma, masks, offset=0.0):
        self.eval()
        return self.predict_reward_combined(torch.cat([state, action], dim=1), gamma, masks, offset)

    def predict_prob_single_step(self, state, action):
        with torch.no_grad():
            self.eval()
            return torch.sigmoid(self.trunk(torch.cat([state, action], dim=1)))

    SHARD_RELABEL_FROM = 1 << 20        # rows (T*N) from which a data-parallel relabel splits its D forward over the ranks

    @_lib.on_device(lambda self, *a, **k: self.trunk[0].weight.device)
    def relabel_rollout(self, rollouts, gamma, offset, ret_rms, sync=True):
        """The caller's whole relabel loop (main_gail_dyn_ppo.py:275-297) in one device pass:
        for every step t, reward_t = predict_reward_combined(obs_feat[t+1], gamma, masks[t], offset),
        ret_rms.update(returns_t), rewards[t] = clip(reward_t / sqrt(ret_rms.var + 1e-7), +-10).
        Updates ``rollouts.rewards``, ``self.returns`` and (when ``sync``) ``ret_rms`` in place and returns the
        (T,) device tensor of per-step mean(returns) -- what the caller appends to ``gail_rewards``."""
        flat = self.flat_params()
        if not rollouts.obs_feat.is_cuda:
            raise _lib.SgError("relabel_rollout needs the rollout buffer on the CUDA device")
        dev = flat.device
        T, N = rollouts.rewards.shape[:2]
        has = self.returns is not None
        if not has:
            self.returns = torch.zeros(N, 1, device=dev)
        lib = _lib.lib()
        need = int(lib.sg_relabel_workspace_bytes(T, N))
        ws = self.__dict__.get("_relabel_ws")
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self.__dict__["_relabel_ws"] = ws
        rms_dev = torch.tensor([float(ret_rms.mean), float(ret_rms.var), float(ret_rms.count)], dtype=torch.float64,
                               device=dev)
        mean_returns = torch.empty(T, device=dev)
        tok = _lib.timer.start("disc_relabel")
        dp = self.dp
        rows = T * N
        if dp is not None and dp.world > 1 and rows >= self.SHARD_RELABEL_FROM:
            # data parallel over a large rollout: the D forward over the T*N rows (all of the relabel's arithmetic weight) is
            # split by rows over the ranks and all-gathered; a row's reward does not depend on the split, and the serial,
            # bit-exact part (running-return scan, moments, RunningMeanStd chain, normalise + clip) then runs on every rank
            # from the same raw rewards -- the replicas stay bit-identical
            import torch.distributed as dist
            chunk = (rows + dp.world - 1) // dp.world
            raw = self.__dict__.get("_relabel_raw")
            if raw is None or raw.numel() != chunk * dp.world or raw.device != dev:
                raw = torch.empty(chunk * dp.world, device=dev)
                self.__dict__["_relabel_raw"] = raw
            b, e = dp.rank * chunk, min(rows, (dp.rank + 1) * chunk)
            feat = rollouts.obs_feat[1:].reshape(rows, self.feat_dim)          # reward_t reads obs_feat[t+1] (:278)
            if e > b:
                rc = lib.sg_disc_predict_reward(_lib.ptr(flat), self.feat_dim, self.hidden_dim, _lib.ptr(feat[b:e]), e - b,
                                                float(gamma), None, float(offset), 0, _lib.ptr(raw[b:e]), None,
                                                _lib.current_stream())
                _lib.check(rc, "sg_disc_predict_reward")
            dist.all_gather_into_tensor(raw, raw[dp.rank * chunk:(dp.rank + 1) * chunk], group=dp.group)
            rc = lib.sg_relabel_normalize(_lib.ptr(raw), _lib.ptr(rollouts.masks), _lib.ptr(rollouts.rewards), T, N,
                                          float(gamma), _lib.ptr(self.returns), int(has), _lib.ptr(rms_dev),
                                          _lib.ptr(mean_returns), _lib.ptr(ws), _lib.current_stream())
        else:
            rc = lib.sg_disc_relabel(_lib.ptr(flat), self.feat_dim, self.hidden_dim, _lib.ptr(rollouts.obs_feat),
                                     _lib.ptr(rollouts.masks), _lib.ptr(rollouts.rewards), T, N, float(gamma), float(offset),
                                     _lib.ptr(self.returns), int(has), _lib.ptr(rms_dev), _lib.ptr(mean_returns),
                                     _lib.ptr(ws), _lib.current_stream())
        _lib.timer.stop(tok)
        _lib.check(rc, "sg_disc_relabel")
        self.__dict__["_rms_dev"] = rms_dev
        if sync:
            _spec.host_idle()
            self.sync_rms(ret_rms)
        return mean_returns

    def sync_rms(self, ret_rms):
        """Copy the device-side running statistics of the last relabel back into the host object."""
        st = self.__dict__.get("_rms_dev")
        if st is not None:
            mean, var, count = st.cpu().tolist()
            ret_rms.mean = np.array(mean, dtype=np.float64)
            ret_rms.var = np.array(var, dtype=np.float64)
            ret_rms.count = count


def alive_bonus_offset(masks, num_steps, num_processes, gail_tar_length, no_alive_bonus=False):
    """r_sa of main_gail_dyn_ppo.py:258-271; the relabel is called with offset=-r_sa."""
    if no_alive_bonus:
        return 0.0
    n_done = (1.0 - masks).sum().cpu().numpy() + num_processes / 2
    n_expert_done = (num_steps * num_processes) / gail_tar_length
    d_sa = 1 - n_done / (n_done + n_expert_done)
    return float(np.log(d_sa) - np.log(1 - d_sa))
