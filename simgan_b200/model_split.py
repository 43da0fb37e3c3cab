"""SplitPolicy with the reference's surface (third_party/a2c_ppo_acktr/model_split.py:39-95, 157-238) -- the policy
class the shipped ``train_*.sh`` scripts train (``--use-split-pi``): contact-actor, actuator-actor and critic
trunks on the same observation, and a diagonal Gaussian whose mean and log-std are linear heads of the actor
trunks (state-dependent log-std, contact actions first).

Module tree, parameter names/order and initialisation (= CPU RNG consumption) are the reference's, so state_dicts
and whole-object pickles stay interchangeable; on a CUDA device the 22 parameters are views into one flat fp32
vector laid out as ``sg_split_layout`` wants it, ``act`` / ``get_value`` / ``evaluate_actions`` run the sm_100a kernel
``sg_split_forward`` and ``PPO.update`` runs ``sg_split_ppo_update``.  CPU tensors take the plain torch path that
un-pickled policies inside env workers need (my_pybullet_envs/hopper_env_combined_policy.py:213-216).
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .distributions import FixedNormal
from .utils import init


class SplitPolicyBaseNew(nn.Module):
    def __init__(self, num_inputs, hidden_size=64, num_feet=1):
        super().__init__()
        init_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0), np.sqrt(2))    # noqa: E731
        init_final_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0))           # noqa: E731
        self.actor_contact = nn.Sequential(init_(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
                                           init_(nn.Linear(hidden_size, hidden_size)), nn.Tanh())
        self.actor_actuator = nn.Sequential(init_(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
                                            init_(nn.Linear(hidden_size, hidden_size)), nn.Tanh())
        self.critic_full = nn.Sequential(init_(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
                                         init_(nn.Linear(hidden_size, hidden_size)), nn.Tanh(),
                                         init_final_(nn.Linear(hidden_size, 1)))
        self.train()

    def forward(self, inputs, rnn_hxs, masks):
        value = self.critic_full(inputs)
        return value, torch.cat((self.actor_contact(inputs), self.actor_actuator(inputs)), 1), rnn_hxs


class StateDiagGaussianNew(nn.Module):
    def __init__(self, num_outputs, hidden_size=64, num_feet=1):
        super().__init__()
        assert num_outputs == (4 + 3) * num_feet          # contact 4, actuator 3 per foot (model_split.py:205)
        self.hidden_size = hidden_size
        init_mean_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0), gain=0.02)     # noqa: E731
        init_logstd_ = lambda m: init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, -0.5), gain=1.0)  # noqa: E731
        self.contact_mean = init_mean_(nn.Linear(hidden_size, 4 * num_feet))
        self.actuator_mean = init_mean_(nn.Linear(hidden_size, 3 * num_feet))
        self.contact_logstd = init_logstd_(nn.Linear(hidden_size, 4 * num_feet))
        self.actuator_logstd = init_logstd_(nn.Linear(hidden_size, 3 * num_feet))

    def forward(self, x):
        c, a = x[:, :self.hidden_size], x[:, self.hidden_size:]
        mean = torch.cat((self.contact_mean(c), self.actuator_mean(a)), 1)
        logstd = torch.cat((self.contact_logstd(c), self.actuator_logstd(a)), 1)
        return FixedNormal(mean, logstd.exp())


class SplitPolicy(nn.Module):
    def __init__(self, obs_shape, action_space, base_kwargs=None):
        super().__init__()
        base_kwargs = dict(base_kwargs or {})
        num_outputs = action_space.shape[0]
        self.base = SplitPolicyBaseNew(obs_shape[0], **base_kwargs)
        self.dist = StateDiagGaussianNew(num_outputs, **base_kwargs)

    @property
    def is_recurrent(self):
        return False

    @property
    def recurrent_hidden_state_size(self):
        return 1

    def forward(self, inputs, rnn_hxs, masks):
        raise NotImplementedError

    # ---- reference surface --------------------------------------------------------------------------------------
    def act(self, inputs, rnn_hxs, masks, deterministic=False):
        if inputs.is_cuda:
            noise = None
            if not deterministic:
                noise = torch.randn(inputs.shape[0], self.act_dim, device=inputs.device, dtype=torch.float32)
            value, action, logp, _ = self._forward_cuda(inputs, noise=noise)
            return value, action, logp, rnn_hxs
        value, feat, rnn_hxs = self.base(inputs, rnn_hxs, masks)
        dist = self.dist(feat)
        action = dist.mode() if deterministic else dist.sample()
        return value, action, dist.log_probs(action), rnn_hxs

    def get_value(self, inputs, rnn_hxs, masks):
        if inputs.is_cuda:
            return self._forward_cuda(inputs, want=("value",))[0]
        return self.base(inputs, rnn_hxs, masks)[0]

    def evaluate_actions(self, inputs, rnn_hxs, masks, action):
        if inputs.is_cuda and not torch.is_grad_enabled():
            value, _, logp, ent = self._forward_cuda(inputs, actions_in=action, want=("value", "logp", "entropy"))
            return value, logp, ent.mean(), rnn_hxs
        value, feat, rnn_hxs = self.base(inputs, rnn_hxs, masks)
        dist = self.dist(feat)
        return value, dist.log_probs(action), dist.entropy().mean(), rnn_hxs

    # ---- flat parameter buffer ------------------------------------------------------------------------------------
    @property
    def obs_dim(self):
        return self.base.actor_contact[0].in_features

    @property
    def hidden_size(self):
        return self.base.actor_contact[0].out_features

    @property
    def num_feet(self):
        return self.dist.contact_mean.out_features // 4

    @property
    def act_dim(self):
        return 7 * self.num_feet

    def hot_path_parameters(self):
        """The 22 parameters in nn.Module.parameters() order (== the offsets table of sg_split_layout)."""
        return list(self.parameters())

    def flat_params(self):
        ps = self.hot_path_parameters()
        dev = ps[0].device
        if dev.type != "cuda":
            raise _lib.SgError("the PPO+GAIL hot path needs the policy on a CUDA device (got %s); there is no CPU fallback" % dev)
        offs, total = _lib.split_layout(self.obs_dim, self.hidden_size, self.num_feet)
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
        state = dict(self.__dict__)
        state.pop("_flat", None)
        return state

    @_lib.on_device(lambda self, inputs, *a, **k: inputs.device)
    def _forward_cuda(self, inputs, noise=None, actions_in=None, want=("value", "action", "logp")):
        flat = self.flat_params()
        x = inputs.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        B, A, dev = x.shape[0], self.act_dim, x.device
        value = torch.empty(B, 1, device=dev) if "value" in want else None
        action = torch.empty(B, A, device=dev) if "action" in want else None
        logp = torch.empty(B, 1, device=dev) if "logp" in want else None
        ent = torch.empty(B, device=dev) if "entropy" in want else None
        if actions_in is not None:
            actions_in = actions_in.detach().float().contiguous()
        rc = _lib.lib().sg_split_forward(_lib.ptr(flat), self.obs_dim, self.hidden_size, self.num_feet, _lib.ptr(x), B,
                                         _lib.ptr(noise), _lib.ptr(actions_in), _lib.ptr(value), _lib.ptr(action),
                                         _lib.ptr(logp), _lib.ptr(ent), _lib.current_stream())
        _lib.check(rc, "sg_split_forward")
        return value, action, logp, ent
