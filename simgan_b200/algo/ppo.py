"""PPO with the reference's surface (third_party/a2c_ppo_acktr/algo/ppo.py:29-157).

``update(rollouts)`` keeps the reference's contract -- same constructor, three Python floats back,
``optimizer.param_groups[i]['lr']`` honoured at every call, sampler indices drawn from the CPU default
generator in the reference's order -- but the whole ``ppo_epoch x num_mini_batch`` loop (gather ->
actor/critic forward -> Gaussian log-prob -> clipped surrogate + clipped value loss -> backward ->
global-norm clip -> Adam) runs inside the sm_100a kernel behind ``sg_ppo_update``; the three
``.item()`` syncs per minibatch of the reference (ppo.py:147-149) become one device->host read of a
per-step trace at the end.
"""
import ctypes as C

import torch

from .. import _lib, _spec, host_sampler
from .adam import FusedAdam


class PPO(object):
    MMA_MODE = 4        # kernel_mode: tcgen05 3xTF32 tensor-core tiles (sg_ppo_config.mode 4)
    # From this many rows per minibatch on, the rows of every minibatch are visited in ascending sample order.  The sampler's
    # partition of the rollout into minibatches (which samples, which minibatch: A2C/storage.py:158-185) is untouched; only the
    # order in which a minibatch's rows are summed changes (every loss is a batch mean), like the tiling itself already does.
    # A 128-row tile then reads rows that lie within a few MB of each other instead of 128 random places of a multi-GB
    # buffer: the gather at the head of a tile was bound by address-translation misses, not by bytes.
    SORT_ROWS_FROM = 1 << 14

    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef,
                 symmetry_coef=0, lr=None, eps=None, max_grad_norm=None, use_clipped_value_loss=True,
                 mirror_obs=None, mirror_act=None):
        if mirror_obs and symmetry_coef > 0:
            raise NotImplementedError("the symmetry-loss branch (ppo.py:111-136) is outside the hot path "
                                      "(--loss-sym defaults to 0)")
        self.actor_critic = actor_critic
        self.clip_param = clip_param
        self.ppo_epoch = ppo_epoch
        self.num_mini_batch = num_mini_batch
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.optimizer = FusedAdam(actor_critic.parameters(), lr=lr, eps=eps)
        self.symmetry_coef = symmetry_coef
        self.mirror_obs = mirror_obs
        self.mirror_act = mirror_act
        self.is_cuda = next(actor_critic.parameters()).is_cuda
        self.kernel_mode = 0            # 0: automatic, 1: one launch per phase, 2/3: CUDA-core tiles, 4: tensor-core tiles
        self.dp = None                  # simgan_b200.dist.DataParallel or None
        self.last_trace = None          # (n_steps, 4) {value_loss, action_loss, entropy, grad_norm}
        self.last_kernel = None         # "tensor" (tcgen05 tiles) or "cuda-core": which tile phase the last update ran
        self._ws = None
        self._stage = None
        self._perm_dev = None
        self._predraw = None            # early draw of the next update's first sampler permutation (_spec.py)
        self._last_S = None
        self._sched_cache = None

    # ---- helpers -------------------------------------------------------------------------------------
    def _workspace(self, cfg, dev, split=False):
        lib = _lib.lib()
        need = (lib.sg_split_ppo_workspace_bytes if split else lib.sg_ppo_workspace_bytes)(C.byref(cfg))
        if need < 0:
            _lib.check(1, "sg_ppo_workspace_bytes")
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(int(need), dtype=torch.uint8, device=dev)
        return self._ws

    def draw_permutations(self, S):
        """ppo_epoch x torch.randperm(S) on the CPU default generator -- the sampler stream of
        feed_forward_generator (storage.py:158-162), one draw per epoch (ppo.py:79-80)."""
        n = self.ppo_epoch
        out = torch.empty(n, S, dtype=torch.int32)
        for e in range(n):
            out[e].copy_(torch.randperm(S))
        return out

    def _speculate(self):
        """Draw the next update's first sampler permutation early and rewind the generator (_spec.py)."""
        S = self._last_S
        if S is None or _spec.still_valid(self._predraw, S):
            return
        if self._shares_permutations(S) and self.dp.rank != 0:
            return          # rank 0 builds (and broadcasts) the first permutation; see update()
        self._predraw = _spec.predraw(S, lambda: host_sampler.randperm_i32(S))
        self._schedule(self.ppo_epoch * self.num_mini_batch)      # pure host arithmetic; cached by its inputs

    def _shares_permutations(self, S):
        """Sharded data-parallel updates over a large rollout: every rank would walk the SAME ppo_epoch permutations of S
        elements (identical seeds) -- tens of ms of host time each, more than an epoch's kernel once the minibatch is split
        over several GPUs.  The ranks split the walks instead (permutation e is built by rank e % world and broadcast on the
        device); every rank still runs the mt19937 engine through all draws, so the generators stay identical."""
        dp = self.dp
        return (dp is not None and dp.world > 1 and getattr(self, "last_sharded", False) and dp.transport == "p2p"
                and S >= host_sampler.MIN_ELEMENTS and host_sampler.usable())

    def _schedule(self, n_steps):
        """Adam step-size scalars of the next n_steps (FusedAdam.schedule), cached by everything they depend on."""
        g = self.optimizer.param_groups[0]
        sig = (self.optimizer.step_count, n_steps, g["lr"], tuple(g["betas"]))
        if self._sched_cache is None or self._sched_cache[0] != sig:
            self._sched_cache = (sig, torch.from_numpy(self.optimizer.schedule(n_steps)))
        return self._sched_cache[1]

    def phase_cycles(self):
        """Diagnostics: per-phase SM-clock totals of CTA 0 of the last persistent launch
        {param image, tile phase, barrier 1, reduce+ssq, barrier 2, clip+Adam, barrier 3}."""
        ws, off = self._prof_view
        return ws[off:off + 64].view(torch.int64)[:7].cpu().tolist()

    def phase_cycles_all(self, n_ctas):
        """(n_ctas, 8) per-CTA phase totals of the last persistent launch."""
        ws, off = self._prof_view
        return ws[off:off + 64 * n_ctas].view(torch.int64).view(n_ctas, 8).cpu()

    @_lib.on_device(lambda self, *a, **k: next(self.actor_critic.parameters()).device)
    def update(self, rollouts, permutations=None):
        """PPO.update (ppo.py:65-157).  ``permutations`` (ppo_epoch, S) overrides the sampler draw (used by
        parity tests to replay a recorded index stream)."""
        ac = self.actor_critic
        flat = ac.flat_params()
        dev = flat.device
        if not rollouts.rewards.is_cuda:
            raise _lib.SgError("PPO.update needs the rollout buffer on the policy's CUDA device")
        T, N = rollouts.rewards.shape[:2]
        S = T * N
        assert S >= self.num_mini_batch, (
            "PPO requires the number of processes ({}) * number of steps ({}) = {} to be greater than or equal to "
            "the number of PPO mini batches ({}).".format(N, T, S, self.num_mini_batch))
        mbs = S // self.num_mini_batch
        n_steps = self.ppo_epoch * self.num_mini_batch
        opt = self.optimizer
        m, v = opt.ensure_state(flat)
        g = opt.param_groups[0]

        is_split = type(ac).__name__ == "SplitPolicy"
        cfg = _lib.PpoConfig()
        cfg.obs_dim, cfg.hidden, cfg.act_dim = ac.obs_dim, ac.hidden_size, ac.act_dim
        cfg.T, cfg.N = T, N
        cfg.ppo_epoch, cfg.num_mini_batch, cfg.mini_batch_size = self.ppo_epoch, self.num_mini_batch, mbs
        cfg.clip_param, cfg.value_loss_coef, cfg.entropy_coef = self.clip_param, self.value_loss_coef, self.entropy_coef
        cfg.max_grad_norm = self.max_grad_norm
        cfg.beta1, cfg.beta2, cfg.adam_eps = g["betas"][0], g["betas"][1], g["eps"]
        cfg.use_clipped_value_loss = int(bool(self.use_clipped_value_loss))
        cfg.first_adam_step = opt.step_count + 1
        dp = self.dp if (self.dp is not None and self.dp.shards(mbs, 8)) else None      # "auto": replicate small minibatches
        self.last_sharded = dp is not None
        cfg.row_begin, cfg.row_end = (0, mbs) if dp is None else dp.shard(mbs)
        p2p = dp is not None and dp.p2p_ok(mbs)
        if is_split and dp is not None and not p2p:
            # SplitPolicy has no phased kernel for the nccl callback: where the in-kernel exchange cannot run (nccl transport, or
            # a minibatch the world size does not divide) every rank computes the whole minibatch instead -- identical seeds
            # and deterministic kernels keep the replicas bit-identical, exactly as under the "auto" policy
            dp, self.last_sharded = None, False
            cfg.row_begin, cfg.row_end = 0, mbs
        cfg.mode = self.kernel_mode if (dp is None or p2p) else 1
        if cfg.mode == 1 and p2p:
            cfg.mode = 0
        cfg.dp_ctx = dp.context("ppo", flat.numel()) if p2p else None

        lib = _lib.lib()
        stream = _lib.current_stream()
        ws = self._workspace(cfg, dev, is_split)
        # advantage statistics (ppo.py:66-68); the normalisation itself is fused into the tile phase
        stats = torch.empty(2, device=dev)
        stat_ws = torch.empty(int(lib.sg_adv_stats_workspace_bytes(S)), dtype=torch.uint8, device=dev)
        _lib.check(lib.sg_adv_stats(_lib.ptr(rollouts.returns), _lib.ptr(rollouts.value_preds), S, _lib.ptr(stats),
                                    _lib.ptr(stat_ws), stream), "sg_adv_stats")
        # Adam scalars -> device; sampler index stream: one torch.randperm(S) per epoch on the CPU default generator
        # (storage.py:158-162).  Epochs are launched one by one so that the host draws epoch e+1 while the GPU
        # runs epoch e (the launches are asynchronous; nothing below synchronises until the trace is read).
        sched = self._schedule(n_steps).to(dev)
        trace = torch.empty(n_steps, 4, device=dev)
        nmb = self.num_mini_batch
        if self._stage is None or self._stage.shape != (self.ppo_epoch, S):
            self._stage = torch.empty(self.ppo_epoch, S, dtype=torch.int32)
            if torch.cuda.is_available():
                self._stage = self._stage.pin_memory()
        if self._perm_dev is None or self._perm_dev.shape != (self.ppo_epoch, S) or self._perm_dev.device != dev:
            self._perm_dev = torch.empty(self.ppo_epoch, S, dtype=torch.int32, device=dev)
        if permutations is not None:
            permutations = torch.as_tensor(permutations).to(torch.int32).reshape(self.ppo_epoch, S)

        cb, user = _lib.NULL_ALLREDUCE, None
        if dp is not None and not p2p:
            cb = dp.make_callback(ws)
        first = None
        if permutations is None:
            _spec.consumed(self)
            self._last_S = S
            first, self._predraw = _spec.take(self._predraw, S), None
        cfg.ppo_epoch = 1
        pstream, drawn0 = None, 0
        share = permutations is None and dp is not None and self._shares_permutations(S)
        if permutations is None:
            # the remaining epochs' permutations are produced on helper threads while the epochs' kernels run
            # (host_sampler.PermutationStream: the same mt19937 stream torch.randperm would consume)
            if first is not None:
                self._stage[0].copy_(first)
                drawn0 = 1
            owned = [e % dp.world == dp.rank for e in range(drawn0, self.ppo_epoch)] if share else None
            pstream = host_sampler.PermutationStream(S, self.ppo_epoch - drawn0, self._stage[drawn0:], owned=owned)
        for e in range(self.ppo_epoch):
            if permutations is not None:
                self._stage[e].copy_(permutations[e])
            elif e >= drawn0:
                pstream.wait(e - drawn0)
            if share:
                import torch.distributed as dist
                owner = e % dp.world
                if owner == dp.rank:
                    self._perm_dev[e].copy_(self._stage[e], non_blocking=True)
                dist.broadcast(self._perm_dev[e], src=dist.get_global_rank(dp.group, owner) if dp.group is not None else owner,
                               group=dp.group)
            else:
                self._perm_dev[e].copy_(self._stage[e], non_blocking=True)
            if mbs >= self.SORT_ROWS_FROM:
                rows = self._perm_dev[e][:nmb * mbs].view(nmb, mbs)
                rows.copy_(torch.sort(rows, dim=1).values)
            cfg.first_adam_step = opt.step_count + 1 + e * nmb
            tok = _lib.timer.start("ppo_update")
            args = (C.byref(cfg), _lib.ptr(flat), _lib.ptr(m), _lib.ptr(v), _lib.ptr(rollouts.obs),
                    _lib.ptr(rollouts.actions), _lib.ptr(rollouts.value_preds), _lib.ptr(rollouts.returns),
                    _lib.ptr(rollouts.action_log_probs), _lib.ptr(stats), _lib.ptr(self._perm_dev[e]),
                    _lib.ptr(sched[0, e * nmb:]), _lib.ptr(sched[1, e * nmb:]), _lib.ptr(trace[e * nmb:]), _lib.ptr(ws))
            rc = lib.sg_split_ppo_update(*args, stream) if is_split else lib.sg_ppo_update(*args, cb, user, stream)
            _lib.timer.stop(tok)
            _lib.check(rc, "sg_ppo_update")
        if pstream is not None:
            pstream.finish()             # the CPU default generator is now where ppo_epoch torch.randperm(S) calls leave it
        cfg.ppo_epoch = self.ppo_epoch
        opt.step_count += n_steps

        self._prof_view = (ws, 0 if is_split else int(lib.sg_ppo_phase_cycles_offset(C.byref(cfg))))
        self.last_kernel = "cuda-core" if is_split or lib.sg_ppo_uses_tensor_cores(C.byref(cfg)) != 1 else "tensor"
        if p2p:
            # value / action loss columns (and SplitPolicy's state-dependent entropy) are per-rank partial sums
            dp.sum_trace_(trace, 3 if is_split else 2)
        _spec.host_idle()           # the next consumer of the CPU generator draws while the last epoch runs
        tr = _lib.read_back(trace)  # the one host sync of the update
        if not bool(torch.isfinite(tr).all()):
            raise _lib.SgError("sg_ppo_update produced non-finite losses (grid barrier timeout or diverged update)")
        self.last_trace = tr
        vl = al = ent = 0.0
        for row in tr.tolist():     # Python-double running sums, like the reference's += .item()
            vl += row[0]
            al += row[1]
            ent += row[2]
        return vl / n_steps, al / n_steps, ent / n_steps
