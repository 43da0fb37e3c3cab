"""CPU oracle for the PPO+GAIL inner loop  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch (CPU, fp32, eager + autograd) restatement of the reference's hot path, written
functionally over explicit tensors.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; nothing under
``simgan_b200/`` does.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` runs every function below against the
unmodified reference modules (imported through ``oracle/ref_shim.py``) in the dev container and
requires bit-identical results at one intra-op thread; ``tests/golden/*.npz`` (written by
``oracle/make_golden.py`` from the real reference) pin it again on the GPU box where
``/root/reference`` is absent.  The arithmetic below the reference (``nn.Linear``, ``tanh``,
``Normal.log_prob``, ``binary_cross_entropy_with_logits``, ``autograd``, ``optim.Adam``,
``clip_grad_norm_``, ``randperm``) is PyTorch itself (reference pins torch 1.5, README.md:33; this
image has 2.11), so parity is pinned relative to torch 2.11 semantics.

All ``file:line`` citations are relative to /root/reference; A2C = third_party/a2c_ppo_acktr.
"""
from __future__ import annotations

import math
import pickle
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------
# actor-critic:  A2C/model.py:233-264 (MLPBase), A2C/distributions.py:91-118 (DiagGaussian)
# ----------------------------------------------------------------------------------------------
POLICY_KEYS = ("aw1", "ab1", "aw2", "ab2", "cw1", "cb1", "cw2", "cb2", "vw", "vb", "mw", "mb", "logstd")


def _ortho_linear(n_out: int, n_in: int, gain: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.Linear construction followed by orthogonal_/zero-bias (A2C/utils.py:75-78).

    nn.Linear's own reset_parameters() draws from the CPU generator *before* the orthogonal
    re-initialisation, so it is kept to leave the RNG stream where the reference leaves it."""
    lin = torch.nn.Linear(n_in, n_out)
    torch.nn.init.orthogonal_(lin.weight.data, gain=gain)
    torch.nn.init.constant_(lin.bias.data, 0)
    return lin.weight.data.clone(), lin.bias.data.clone()


def init_policy(obs_dim: int, hidden: int, act_dim: int) -> Dict[str, torch.Tensor]:
    """Parameter set of Policy(MLPBase)+DiagGaussian in the reference's construction order
    (actor L1, L2, critic L1, L2, critic_linear: A2C/model.py:243-251; fc_mean gain 1 then /50 and
    logstd=-0.5 with shape (A,1): A2C/distributions.py:95-104, A2C/utils.py:57)."""
    g = math.sqrt(2.0)
    p = {}
    p["aw1"], p["ab1"] = _ortho_linear(hidden, obs_dim, g)
    p["aw2"], p["ab2"] = _ortho_linear(hidden, hidden, g)
    p["cw1"], p["cb1"] = _ortho_linear(hidden, obs_dim, g)
    p["cw2"], p["cb2"] = _ortho_linear(hidden, hidden, g)
    p["vw"], p["vb"] = _ortho_linear(1, hidden, g)
    mw, mb = _ortho_linear(act_dim, hidden, 1.0)
    p["mw"], p["mb"] = mw / 50.0, mb / 50.0
    p["logstd"] = (torch.ones(act_dim) * -0.5).unsqueeze(1)
    return p


def policy_forward(p: Dict[str, torch.Tensor], x: torch.Tensor):
    """-> value (B,1), mean (B,A), logstd row (1,A).  A2C/model.py:255-264, distributions.py:109-118."""
    hc = torch.tanh(F.linear(torch.tanh(F.linear(x, p["cw1"], p["cb1"])), p["cw2"], p["cb2"]))
    ha = torch.tanh(F.linear(torch.tanh(F.linear(x, p["aw1"], p["ab1"])), p["aw2"], p["ab2"]))
    value = F.linear(hc, p["vw"], p["vb"])
    mean = F.linear(ha, p["mw"], p["mb"])
    logstd = torch.zeros_like(mean) + p["logstd"].t().view(1, -1)
    return value, mean, logstd


def gaussian_logp_entropy(mean, logstd, action):
    """Sum over the action dim of Normal.log_prob (keepdim) and Normal.entropy
    (A2C/distributions.py:51-56)."""
    dist = torch.distributions.Normal(mean, logstd.exp())
    return dist.log_prob(action).sum(-1, keepdim=True), dist.entropy().sum(-1)


def policy_evaluate(p, x, action):
    """-> value (B,1), logp (B,1), mean entropy ().   A2C/model.py:107-114."""
    value, mean, logstd = policy_forward(p, x)
    logp, ent = gaussian_logp_entropy(mean, logstd, action)
    return value, logp, ent.mean()


def policy_act(p, x, deterministic=False, noise: Optional[torch.Tensor] = None):
    """-> value, action, logp.  A2C/model.py:89-101.  ``noise`` (standard normal, (B,A)) replaces
    Normal.sample()'s internal draw when given (action = mean + std*noise, which is what
    torch.normal computes)."""
    with torch.no_grad():
        value, mean, logstd = policy_forward(p, x)
        if deterministic:
            action = mean
        elif noise is not None:
            action = mean + logstd.exp() * noise
        else:
            action = torch.distributions.Normal(mean, logstd.exp()).sample()
        logp, _ = gaussian_logp_entropy(mean, logstd, action)
    return value, action, logp


# ----------------------------------------------------------------------------------------------
# SplitPolicy: A2C/model_split.py:39-95, 157-238 (what the shipped train_*.sh scripts use, --use-split-pi)
# ----------------------------------------------------------------------------------------------
SPLIT_KEYS = ("c_w1", "c_b1", "c_w2", "c_b2", "a_w1", "a_b1", "a_w2", "a_b2",
              "v_w1", "v_b1", "v_w2", "v_b2", "v_w3", "v_b3",
              "cm_w", "cm_b", "am_w", "am_b", "cl_w", "cl_b", "al_w", "al_b")


def _ortho_linear_bias(n_out: int, n_in: int, gain: float, bias: float):
    lin = torch.nn.Linear(n_in, n_out)
    torch.nn.init.orthogonal_(lin.weight.data, gain=gain)
    torch.nn.init.constant_(lin.bias.data, bias)
    return lin.weight.data.clone(), lin.bias.data.clone()


def init_split_policy(obs_dim: int, hidden: int, num_feet: int = 1) -> Dict[str, torch.Tensor]:
    """Parameters of SplitPolicy in construction (= RNG) order: SplitPolicyBaseNew (contact trunk, actuator
    trunk, critic trunk incl. its gain-1 final layer, model_split.py:172-185), then StateDiagGaussianNew
    (contact_mean, actuator_mean gain 0.02; contact_logstd, actuator_logstd gain 1 / bias -0.5, :220-224)."""
    g = math.sqrt(2.0)
    p = {}
    p["c_w1"], p["c_b1"] = _ortho_linear(hidden, obs_dim, g)
    p["c_w2"], p["c_b2"] = _ortho_linear(hidden, hidden, g)
    p["a_w1"], p["a_b1"] = _ortho_linear(hidden, obs_dim, g)
    p["a_w2"], p["a_b2"] = _ortho_linear(hidden, hidden, g)
    p["v_w1"], p["v_b1"] = _ortho_linear(hidden, obs_dim, g)
    p["v_w2"], p["v_b2"] = _ortho_linear(hidden, hidden, g)
    p["v_w3"], p["v_b3"] = _ortho_linear(1, hidden, 1.0)
    p["cm_w"], p["cm_b"] = _ortho_linear_bias(4 * num_feet, hidden, 0.02, 0.0)
    p["am_w"], p["am_b"] = _ortho_linear_bias(3 * num_feet, hidden, 0.02, 0.0)
    p["cl_w"], p["cl_b"] = _ortho_linear_bias(4 * num_feet, hidden, 1.0, -0.5)
    p["al_w"], p["al_b"] = _ortho_linear_bias(3 * num_feet, hidden, 1.0, -0.5)
    return p


def split_forward(p: Dict[str, torch.Tensor], x: torch.Tensor):
    """-> value (B,1), mean (B,A), logstd (B,A) -- state-dependent log-std (model_split.py:187-198, 226-238)."""
    def trunk(pre):
        return torch.tanh(F.linear(torch.tanh(F.linear(x, p[pre + "_w1"], p[pre + "_b1"])), p[pre + "_w2"], p[pre + "_b2"]))
    value = F.linear(trunk("v"), p["v_w3"], p["v_b3"])
    y1, y2 = trunk("c"), trunk("a")
    mean = torch.cat((F.linear(y1, p["cm_w"], p["cm_b"]), F.linear(y2, p["am_w"], p["am_b"])), 1)
    logstd = torch.cat((F.linear(y1, p["cl_w"], p["cl_b"]), F.linear(y2, p["al_w"], p["al_b"])), 1)
    return value, mean, logstd


def split_evaluate(p, x, action):
    """-> value, logp (B,1), batch-mean entropy (model_split.py:87-95)."""
    value, mean, logstd = split_forward(p, x)
    logp, ent = gaussian_logp_entropy(mean, logstd, action)
    return value, logp, ent.mean()


def split_act(p, x, deterministic=False, noise: Optional[torch.Tensor] = None):
    """-> value, action, logp (model_split.py:69-81)."""
    with torch.no_grad():
        value, mean, logstd = split_forward(p, x)
        if deterministic:
            action = mean
        elif noise is not None:
            action = mean + logstd.exp() * noise
        else:
            action = torch.distributions.Normal(mean, logstd.exp()).sample()
        logp, _ = gaussian_logp_entropy(mean, logstd, action)
    return value, action, logp


# ----------------------------------------------------------------------------------------------
# rollout buffer: A2C/storage.py
# ----------------------------------------------------------------------------------------------
def new_buffer(T: int, N: int, obs_dim: int, act_dim: int, feat_len: int) -> Dict[str, torch.Tensor]:
    """Ten fp32 time-major tensors, masks/bad_masks initialised to one (A2C/storage.py:34-53)."""
    z = torch.zeros
    return dict(obs=z(T + 1, N, obs_dim), obs_feat=z(T + 1, N, feat_len), recurrent_hidden_states=z(T + 1, N, 1),
                rewards=z(T, N, 1), value_preds=z(T + 1, N, 1), returns=z(T + 1, N, 1),
                action_log_probs=z(T, N, 1), actions=z(T, N, act_dim),
                masks=torch.ones(T + 1, N, 1), bad_masks=torch.ones(T + 1, N, 1))


def buffer_insert(buf, step, obs, hxs, actions, logp, value, rewards, masks, bad_masks, obs_feat=None) -> int:
    """Slot step+1 for obs/feat/hxs/masks/bad_masks, slot step for the rest; returns the next
    step index (A2C/storage.py:70-84)."""
    buf["obs"][step + 1] = obs
    if obs_feat is not None:
        buf["obs_feat"][step + 1] = obs_feat
    buf["recurrent_hidden_states"][step + 1] = hxs
    buf["actions"][step] = actions
    buf["action_log_probs"][step] = logp
    buf["value_preds"][step] = value
    buf["rewards"][step] = rewards
    buf["masks"][step + 1] = masks
    buf["bad_masks"][step + 1] = bad_masks
    return (step + 1) % buf["rewards"].shape[0]


def buffer_after_update(buf) -> None:
    """Slot T -> slot 0 for the five 'T+1' tensors the next rollout starts from (A2C/storage.py:96-101)."""
    for k in ("obs", "obs_feat", "recurrent_hidden_states", "masks", "bad_masks"):
        buf[k][0] = buf[k][-1]


def compute_returns(buf, next_value, use_gae: bool, gamma: float, gae_lambda: float,
                    use_proper_time_limits: bool = True) -> None:
    """All four branches of A2C/storage.py:103-142, in place on buf['returns'] (and
    buf['value_preds'][-1] in the GAE branches).  The elementwise op order inside each time step is
    the reference's, so results are bit-identical."""
    r, v, m, b, ret = buf["rewards"], buf["value_preds"], buf["masks"], buf["bad_masks"], buf["returns"]
    T = r.shape[0]
    if use_gae:
        v[-1] = next_value
        run = 0
        for t in range(T - 1, -1, -1):
            delta = r[t] + gamma * v[t + 1] * m[t + 1] - v[t]
            run = delta + gamma * gae_lambda * m[t + 1] * run
            if use_proper_time_limits:
                run = run * b[t + 1]
            ret[t] = run + v[t]
    else:
        ret[-1] = next_value
        for t in range(T - 1, -1, -1):
            if use_proper_time_limits:
                ret[t] = (ret[t + 1] * gamma * m[t + 1] + r[t]) * b[t + 1] + (1 - b[t + 1]) * v[t]
            else:
                ret[t] = ret[t + 1] * gamma * m[t + 1] + r[t]


def sampler_chunks(n_samples: int, mini_batch_size: int) -> List[torch.Tensor]:
    """Index stream of BatchSampler(SubsetRandomSampler(range(S)), mb, drop_last=True)
    (A2C/storage.py:158-162): ONE torch.randperm(S) on the CPU default generator, cut into
    consecutive chunks of mb, ragged tail dropped."""
    perm = torch.randperm(n_samples)
    n_full = n_samples // mini_batch_size
    return [perm[i * mini_batch_size:(i + 1) * mini_batch_size] for i in range(n_full)]


def flat_views(buf) -> Dict[str, torch.Tensor]:
    """The (S, D) row views the sampler indexes with flat id t*N+n (A2C/storage.py:169-181)."""
    def fl(x):
        return x.reshape(-1, x.shape[-1])
    return dict(obs=fl(buf["obs"][:-1]), next_obs=fl(buf["obs"][1:]), obs_feat=fl(buf["obs_feat"][:-1]),
                next_obs_feat=fl(buf["obs_feat"][1:]), hxs=fl(buf["recurrent_hidden_states"][:-1]),
                actions=fl(buf["actions"]), value_preds=fl(buf["value_preds"][:-1]),
                returns=fl(buf["returns"][:-1]), masks=fl(buf["masks"][:-1]),
                action_log_probs=fl(buf["action_log_probs"]))


def feed_forward_batches(buf, advantages, num_mini_batch=None, mini_batch_size=None):
    """Generator of the reference's 10-tuples (A2C/storage.py:144-192)."""
    T, N = buf["rewards"].shape[:2]
    S = T * N
    if mini_batch_size is None:
        assert S >= num_mini_batch
        mini_batch_size = S // num_mini_batch
    fv = flat_views(buf)
    adv = None if advantages is None else advantages.reshape(-1, 1)
    for idx in sampler_chunks(S, mini_batch_size):
        yield (fv["obs"][idx], fv["hxs"][idx], fv["actions"][idx], fv["value_preds"][idx], fv["returns"][idx],
               fv["masks"][idx], fv["action_log_probs"][idx], None if adv is None else adv[idx],
               fv["obs_feat"][idx], fv["next_obs_feat"][idx])


# ----------------------------------------------------------------------------------------------
# PPO: A2C/algo/ppo.py:65-157
# ----------------------------------------------------------------------------------------------
@dataclass
class PPOHyper:
    clip_param: float = 0.2
    ppo_epoch: int = 10
    num_mini_batch: int = 32
    value_loss_coef: float = 0.5
    entropy_coef: float = 0.01
    lr: float = 3e-4
    eps: float = 1e-5
    max_grad_norm: float = 0.5
    use_clipped_value_loss: bool = True


def normalized_advantages(buf) -> torch.Tensor:
    """(returns-value_preds)[:-1], centred, divided by the UNBIASED std + 1e-5 (A2C/algo/ppo.py:66-68)."""
    adv = buf["returns"][:-1] - buf["value_preds"][:-1]
    return (adv - adv.mean()) / (adv.std() + 1e-5)


def ppo_losses(p, hyper: PPOHyper, obs, actions, value_preds, returns, old_logp, adv, evaluate=None):
    """value_loss, action_loss, entropy for one minibatch (A2C/algo/ppo.py:88-108)."""
    values, logp, entropy = (evaluate or policy_evaluate)(p, obs, actions)
    ratio = torch.exp(logp - old_logp)
    s1 = ratio * adv
    s2 = torch.clamp(ratio, 1.0 - hyper.clip_param, 1.0 + hyper.clip_param) * adv
    action_loss = -torch.min(s1, s2).mean()
    if hyper.use_clipped_value_loss:
        v_clip = value_preds + (values - value_preds).clamp(-hyper.clip_param, hyper.clip_param)
        value_loss = 0.5 * torch.max((values - returns).pow(2), (v_clip - returns).pow(2)).mean()
    else:
        value_loss = 0.5 * (returns - values).pow(2).mean()
    return value_loss, action_loss, entropy


class PPOOracle:
    """Holds leaf parameters + torch.optim.Adam exactly as A2C/algo/ppo.py:57 does."""

    def __init__(self, params: Dict[str, torch.Tensor], hyper: PPOHyper, keys=None, evaluate=None):
        self.hyper = hyper
        self.keys = tuple(keys or POLICY_KEYS)          # parameter order = nn.Module.parameters() order
        self.evaluate = evaluate                        # policy_evaluate (Policy) or split_evaluate (SplitPolicy)
        self.p = {k: params[k].clone().requires_grad_(True) for k in self.keys}
        self.optimizer = torch.optim.Adam([self.p[k] for k in self.keys], lr=hyper.lr, eps=hyper.eps)

    def step(self, obs, actions, value_preds, returns, old_logp, adv):
        """One optimizer step; returns the three loss scalars and the pre-clip gradient norm."""
        h = self.hyper
        vl, al, ent = ppo_losses(self.p, h, obs, actions, value_preds, returns, old_logp, adv, self.evaluate)
        self.optimizer.zero_grad()
        (vl * h.value_loss_coef + al - ent * h.entropy_coef).backward()
        gn = torch.nn.utils.clip_grad_norm_([self.p[k] for k in self.keys], h.max_grad_norm)
        self.optimizer.step()
        return vl.item(), al.item(), ent.item(), float(gn)

    def update(self, buf, index_chunks: Optional[Sequence[Sequence[torch.Tensor]]] = None, trace=None):
        """Full PPO.update (A2C/algo/ppo.py:65-157).  ``index_chunks[e]`` overrides the sampler of
        epoch e (used to replay a recorded index stream); otherwise the CPU generator is drawn."""
        h = self.hyper
        adv = normalized_advantages(buf).reshape(-1, 1)
        fv = flat_views(buf)
        S = adv.shape[0]
        tot = [0.0, 0.0, 0.0]
        for e in range(h.ppo_epoch):
            chunks = index_chunks[e] if index_chunks is not None else sampler_chunks(S, S // h.num_mini_batch)
            for idx in chunks:
                vl, al, ent, gn = self.step(fv["obs"][idx], fv["actions"][idx], fv["value_preds"][idx],
                                            fv["returns"][idx], fv["action_log_probs"][idx], adv[idx])
                tot[0] += vl
                tot[1] += al
                tot[2] += ent
                if trace is not None:
                    trace.append((vl, al, ent, gn))
        n = h.ppo_epoch * h.num_mini_batch
        return tot[0] / n, tot[1] / n, tot[2] / n

    def params(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach().clone() for k, v in self.p.items()}


# ----------------------------------------------------------------------------------------------
# GAIL discriminator: A2C/algo/gail.py
# ----------------------------------------------------------------------------------------------
DISC_KEYS = ("w1", "b1", "w2", "b2", "w3", "b3")


def init_disc(feat_dim: int, hidden: int) -> Dict[str, torch.Tensor]:
    """Three default-initialised nn.Linear layers, in order (A2C/algo/gail.py:40-43)."""
    l1, l2, l3 = torch.nn.Linear(feat_dim, hidden), torch.nn.Linear(hidden, hidden), torch.nn.Linear(hidden, 1)
    return dict(w1=l1.weight.data.clone(), b1=l1.bias.data.clone(), w2=l2.weight.data.clone(),
                b2=l2.bias.data.clone(), w3=l3.weight.data.clone(), b3=l3.bias.data.clone())


def disc_logit(d, x):
    return F.linear(torch.tanh(F.linear(torch.tanh(F.linear(x, d["w1"], d["b1"])), d["w2"], d["b2"])),
                    d["w3"], d["b3"])


def grad_penalty(d, expert, policy, alpha, lambda_=10.0):
    """lambda * mean((||d D(x^)/d x^||_2 - 1)^2) on x^ = a*e + (1-a)*p (A2C/algo/gail.py:67-89).
    ``alpha`` is the (B,1) uniform draw the reference takes from the CPU generator (gail.py:72)."""
    a = alpha.expand_as(expert)
    mix = (a * expert + (1 - a) * policy).detach().requires_grad_(True)
    out = disc_logit(d, mix)
    (g,) = torch.autograd.grad(out, mix, torch.ones_like(out), create_graph=True, retain_graph=True)
    return lambda_ * (g.norm(2, dim=1) - 1).pow(2).mean()


def expert_loader_indices(n_expert: int, batch_size: int, drop_last: bool) -> List[torch.Tensor]:
    """Index batches an ``iter(DataLoader(TensorDataset(X), batch, shuffle=True, drop_last))`` yields
    (A2C/main_gail_dyn_ppo.py:170-175), reproduced with the same CPU-generator consumption as torch
    2.11: one int64 random_() for the loader's base seed, one for RandomSampler's seed, then a
    randperm on a PRIVATE generator."""
    torch.empty((), dtype=torch.int64).random_()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    perm = torch.randperm(n_expert, generator=g)
    n_full = n_expert // batch_size
    out = [perm[i * batch_size:(i + 1) * batch_size] for i in range(n_full)]
    if not drop_last and n_expert % batch_size:
        out.append(perm[n_full * batch_size:])
    return out


class DiscOracle:
    """Parameters + Adam(lr 1e-3, eps 1e-8 defaults) as A2C/algo/gail.py:48; stateful running
    return as gail.py:50, 206-209."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.d = {k: params[k].clone().requires_grad_(True) for k in DISC_KEYS}
        self.optimizer = torch.optim.Adam([self.d[k] for k in DISC_KEYS])
        self.returns = None

    def step(self, expert, policy, alpha):
        """One minibatch of update_gail_dyn (A2C/algo/gail.py:165-188) -> (total, expert, policy) losses."""
        pd = disc_logit(self.d, policy)
        ed = disc_logit(self.d, expert)
        el = F.binary_cross_entropy_with_logits(ed, torch.ones_like(ed))
        pl = F.binary_cross_entropy_with_logits(pd, torch.zeros_like(pd))
        gp = grad_penalty(self.d, expert, policy, alpha)
        total = el + pl + gp
        self.optimizer.zero_grad()
        total.backward()
        self.optimizer.step()
        return total.item(), el.item(), pl.item()

    def update_epoch(self, expert_set: torch.Tensor, buf, batch_size=128, drop_last=True, replay=None, trace=None):
        """update_gail_dyn (A2C/algo/gail.py:154-193) with the RNG order of zip(loader, generator):
        loader seeds -> randperm(S) -> one rand(B,1) per zipped batch.  ``replay`` =
        (expert_idx list, policy_idx list, alpha list) bypasses the generator."""
        T, N = buf["rewards"].shape[:2]
        S = T * N
        nxt = buf["obs_feat"][1:].reshape(S, -1)
        if replay is None:
            e_idx = expert_loader_indices(expert_set.shape[0], batch_size, drop_last)
            p_idx = None
        else:
            e_idx, p_idx, alphas = replay
        tot = [0.0, 0.0, 0.0]
        n = 0
        for i, ei in enumerate(e_idx):
            if replay is None:
                if p_idx is None:
                    p_idx = sampler_chunks(S, batch_size)   # generator body starts at first next()
                if i >= len(p_idx):
                    break
                alpha = torch.rand(ei.shape[0], 1)
            else:
                if i >= len(p_idx):
                    break
                alpha = alphas[i]
            lt, le, lp = self.step(expert_set[ei], nxt[p_idx[i]], alpha)
            tot[0] += lt
            tot[1] += le
            tot[2] += lp
            n += 1
            if trace is not None:
                trace.append((lt, le, lp))
        return tot[0] / n, tot[1] / n, tot[2] / n

    def predict_reward(self, d_in, gamma, masks, offset=0.0):
        """log D - log(1-D) + offset and the stateful discounted return (A2C/algo/gail.py:201-210)."""
        with torch.no_grad():
            s = torch.sigmoid(disc_logit(self.d, d_in))
            reward = (s + 1e-7).log() - (1 - s + 1e-7).log() + offset
            if self.returns is None:
                self.returns = reward.clone()
            else:
                self.returns = self.returns * gamma * masks + reward
            return reward, self.returns

    def params(self):
        return {k: v.detach().clone() for k, v in self.d.items()}


# ----------------------------------------------------------------------------------------------
# running return statistics: A2C/baselines/common/running_mean_std.py:27-56
# ----------------------------------------------------------------------------------------------
class RunningMeanStd:
    """float64 Chan/Welford merge; batch moments are numpy's (float32 in -> float32 moments)."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        self.merge(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def merge(self, b_mean, b_var, b_count):
        delta = b_mean - self.mean
        tot = self.count + b_count
        new_mean = self.mean + delta * b_count / tot
        m2 = self.var * self.count + b_var * b_count + np.square(delta) * self.count * b_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


def alive_bonus_offset(masks: torch.Tensor, num_steps: int, num_processes: int, gail_tar_length: float,
                       no_alive_bonus: bool = False) -> float:
    """r_sa of A2C/main_gail_dyn_ppo.py:258-271 (the relabel passes offset=-r_sa)."""
    if no_alive_bonus:
        return 0.0
    n_done = (1.0 - masks).sum().cpu().numpy() + num_processes / 2
    n_exp_done = (num_steps * num_processes) / gail_tar_length
    d_sa = 1 - n_done / (n_done + n_exp_done)
    return float(np.log(d_sa) - np.log(1 - d_sa))


def relabel_rewards(disc: DiscOracle, ret_rms: RunningMeanStd, buf, gamma: float, offset: float) -> List[float]:
    """The per-step reward overwrite loop of A2C/main_gail_dyn_ppo.py:275-297.  Returns the list of
    per-step mean(returns) values the caller pushes into its ``gail_rewards`` deque."""
    T = buf["rewards"].shape[0]
    means = []
    for t in range(T):
        reward, returns = disc.predict_reward(buf["obs_feat"][t + 1], gamma, buf["masks"][t], offset=offset)
        buf["rewards"][t] = reward
        ret_rms.update(returns.view(-1).cpu().numpy())
        rews = buf["rewards"][t].view(-1).cpu().numpy()
        rews = np.clip(rews / np.sqrt(ret_rms.var + 1e-7), -10.0, 10.0)
        buf["rewards"][t] = torch.as_tensor(rews, dtype=torch.float32).view(-1, 1)
        means.append(float(torch.mean(returns)))
    return means


# ----------------------------------------------------------------------------------------------
# expert trajectories: my_pybullet_envs/utils.py:170-199, 233-263
# ----------------------------------------------------------------------------------------------
def load_sas_wpast(pathname: str, downsample_freq: int = 1, load_num_trajs: Optional[int] = None) -> List[np.ndarray]:
    """pkl dict[int -> list[row]], row = 2W+1 lists [s_t..s_{t-W+1}, a_t..a_{t-W+1}, s_{t+1}] ->
    list of 2W+1 (N_exp, dim) arrays.  Consumes one torch.randint(0, freq, (n_trajs,)) from the CPU
    generator like my_pybullet_envs/utils.py:178-179.  The reference's ragged ``np.array(sas)``
    (utils.py:193) raises on numpy>=1.24; columns are stacked one by one instead."""
    with open(pathname, "rb") as fh:
        trajs = pickle.load(fh)
    start = torch.randint(0, downsample_freq, size=(len(trajs),)).long()
    rows = []
    for k, traj in trajs.items():
        rows.extend(traj[int(start[k])::downsample_freq])
        if load_num_trajs and k >= load_num_trajs - 1:
            break
    width = len(rows[0])
    return [np.array([row[c] for row in rows]) for c in range(width)]


def merge_sas(sas: Sequence, s_idx=(0,), a_idx=(0,)) -> np.ndarray:
    """[s_{t-i} for i in s_idx | a_{t-j} for j in a_idx | s_{t+1}], float64, for either a list of
    (N,dim) arrays or a single window of 1-D lists (my_pybullet_envs/utils.py:233-263)."""
    one = np.asarray(sas[0]).ndim == 1
    cols = [np.asarray(c, dtype=np.float64)[None, :] if one else np.asarray(c, dtype=np.float64) for c in sas]
    w = (len(sas) - 1) // 2
    parts = [cols[i] for i in s_idx] + [cols[w + j] for j in a_idx] + [cols[-1]]
    out = np.concatenate(parts, axis=1)
    return out[0] if one else out


# ----------------------------------------------------------------------------------------------
# seeded synthetic workload shared by tests and bench
# ----------------------------------------------------------------------------------------------
def synth_rollout(T, N, obs_dim, act_dim, feat_len, policy_params, seed=0, ep_len=88.0, feat_bank=None):
    """Deterministic synthetic rollout buffer: obs ~ N(0,1); obs_feat ~ N(0,1) or rows resampled from
    ``feat_bank`` + N(0,0.1); masks ~ Bernoulli(1-1/ep_len); bad_masks ~ Bernoulli(1-1/500) only where
    mask==0; actions/value_preds/action_log_probs from the given policy's act() on obs[:-1]; rewards
    N(0,1) placeholders (the relabel overwrites them)."""
    g = torch.Generator().manual_seed(1000 + seed)
    buf = new_buffer(T, N, obs_dim, act_dim, feat_len)
    buf["obs"].copy_(torch.randn(T + 1, N, obs_dim, generator=g))
    if feat_bank is None:
        buf["obs_feat"].copy_(torch.randn(T + 1, N, feat_len, generator=g))
    else:
        pick = torch.randint(0, feat_bank.shape[0], ((T + 1) * N,), generator=g)
        buf["obs_feat"].copy_((feat_bank[pick] + 0.1 * torch.randn((T + 1) * N, feat_len, generator=g))
                              .view(T + 1, N, feat_len))
    m = (torch.rand(T + 1, N, 1, generator=g) >= 1.0 / ep_len).float()
    bad = torch.where((m == 0) & (torch.rand(T + 1, N, 1, generator=g) < 1.0 / 500.0), 0.0, 1.0)
    buf["masks"].copy_(m)
    buf["bad_masks"].copy_(bad)
    noise = torch.randn(T * N, act_dim, generator=g)
    v, a, lp = policy_act(policy_params, buf["obs"][:-1].reshape(T * N, obs_dim), noise=noise)
    buf["value_preds"][:-1] = v.view(T, N, 1)
    buf["actions"].copy_(a.view(T, N, act_dim))
    buf["action_log_probs"].copy_(lp.view(T, N, 1))
    buf["rewards"].copy_(torch.randn(T, N, 1, generator=g))
    return buf
