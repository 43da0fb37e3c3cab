"""Actor-critic with the reference's surface (third_party/a2c_ppo_acktr/model.py:37-114, 233-264).

``Policy`` keeps the reference's module tree (``base.actor``, ``base.critic``, ``base.critic_linear``,
``dist.fc_mean``, ``dist.logstd._bias``) so state_dicts and whole-object pickles stay interchangeable,
but on a CUDA device all thirteen parameters are VIEWS into one flat fp32 buffer laid out as the C ABI
expects (include/simgan_b200.h, sg_policy_layout).  ``act`` / ``get_value`` / ``evaluate_actions`` on
CUDA tensors run the sm_100a kernel ``sg_policy_forward``; PPO.update consumes the same flat buffer.

CPU tensors take a plain torch path.  That path exists only because the reference's env workers
un-pickle whole policies and call ``act`` at batch 1 on the host
(my_pybullet_envs/hopper_env_combined_policy.py:213-216); it is not part of the measured hot path.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .distributions import DiagGaussian
from .utils import init


def _ortho(m):
    return init(m, nn.init.orthogonal_, lambda x: nn.init.constant_(x, 0), np.sqrt(2))


class MLPBase(nn.Module):
    """Separate 2x(Linear+Tanh) actor and critic trunks + critic_linear (model.py:233-264).
    Construction order (= CPU RNG order): actor L1, L2, critic L1, L2, critic_linear."""

    def __init__(self, num_inputs, recurrent=False, hidden_size=64):
        super().__init__()
        if recurrent:
            raise NotImplementedError("recurrent (GRU) policies are outside the PPO+GAIL hot path "
                                      "(reference default --recurrent-policy False, arguments.py:157-161)")
        self._hidden_size = hidden_size
        self._recurrent = False
        self.actor = nn.Sequential(_ortho(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
                                   _ortho(nn.Linear(hidden_size, hidden_size)), nn.Tanh())
        self.critic = nn.Sequential(_ortho(nn.Linear(num_inputs, hidden_size)), nn.Tanh(),
                                    _ortho(nn.Linear(hidden_size, hidden_size)), nn.Tanh())
        self.critic_linear = _ortho(nn.Linear(hidden_size, 1))
        self.train()

    @property
    def is_recurrent(self):
        return self._recurrent

    @property
    def recurrent_hidden_state_size(self):
        return self._hidden_size if self._recurrent else 1

    @property
    def output_size(self):
        return self._hidden_size

    def forward(self, inputs, rnn_hxs, masks):
        return self.critic_linear(self.critic(inputs)), self.actor(inputs), rnn_hxs


class Policy(nn.Module):
    def __init__(self, obs_shape, action_space, base=None, base_kwargs=None):
        super().__init__()
        base_kwargs = dict(base_kwargs or {})
        if base is None:
            if len(obs_shape) != 1:
                raise NotImplementedError("only flat observations are on the hot path "
                                          "(main_gail_dyn_ppo.py:139 asserts it)")
            base = MLPBase
        self.base = base(obs_shape[0], **base_kwargs)
        if action_space.__class__.__name__ != "Box":
            raise NotImplementedError("only Box action spaces (DiagGaussian head) are on the hot path")
        self.dist = DiagGaussian(self.base.output_size, action_space.shape[0])

    # ---- reference surface -----------------------------------------------------------------------
    @property
    def is_recurrent(self):
        return self.base.is_recurrent

    @property
    def recurrent_hidden_state_size(self):
        return self.base.recurrent_hidden_state_size

    def forward(self, inputs, rnn_hxs, masks):
        raise NotImplementedError

    def reset_variance(self, action_space, log_std):
        self.dist.reset_variance(action_space.shape[0], log_std)

    def reset_critic(self, obs_shape):
        """Fresh critic of width 64 -- the reference hard-codes 64 here (model.py:80-87)."""
        self.base.critic = nn.Sequential(_ortho(nn.Linear(obs_shape[0], 64)), nn.Tanh(),
                                         _ortho(nn.Linear(64, 64)), nn.Tanh())
        self.base.critic_linear = _ortho(nn.Linear(64, 1))

    def act(self, inputs, rnn_hxs, masks, deterministic=False):
        if inputs.is_cuda:
            noise = None
            if not deterministic:
                # same CUDA-generator draw Normal.sample() makes in the reference (model.py:96)
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
            return value, logp, ent, rnn_hxs
        # differentiable form for external callers; PPO.update never comes through here
        value, feat, rnn_hxs = self.base(inputs, rnn_hxs, masks)
        dist = self.dist(feat)
        return value, dist.log_probs(action), dist.entropy().mean(), rnn_hxs

    # ---- flat parameter buffer ----------------------------------------------------------------------
    @property
    def obs_dim(self):
        return self.base.actor[0].in_features

    @property
    def hidden_size(self):
        return self.base.actor[0].out_features

    @property
    def act_dim(self):
        return self.dist.fc_mean.out_features

    def hot_path_parameters(self):
        """The 13 parameters in C-ABI segment order (== nn.Module.parameters() order)."""
        b, d = self.base, self.dist
        return [b.actor[0].weight, b.actor[0].bias, b.actor[2].weight, b.actor[2].bias,
                b.critic[0].weight, b.critic[0].bias, b.critic[2].weight, b.critic[2].bias,
                b.critic_linear.weight, b.critic_linear.bias, d.fc_mean.weight, d.fc_mean.bias, d.logstd._bias]

    def flat_params(self):
        """Flat CUDA parameter vector in the sg_policy_layout order; (re)binds the nn.Parameters as
        views of it whenever they have been moved or replaced (``.to()``, ``reset_critic`` ...)."""
        ps = self.hot_path_parameters()
        dev = ps[0].device
        if dev.type != "cuda":
            raise _lib.SgError("the PPO+GAIL hot path needs the policy on a CUDA device (got %s); "
                               "there is no CPU fallback" % dev)
        b = self.base
        if b.critic[0].out_features != self.hidden_size or b.critic[0].in_features != self.obs_dim:
            raise NotImplementedError("actor and critic trunks of different shapes (after reset_critic with "
                                      "hidden_size != 64) are not supported by the fused kernels")
        offs, total = _lib.policy_layout(self.obs_dim, self.hidden_size, self.act_dim)
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
        state.pop("_flat", None)       # rebuilt on demand; parameters carry the data
        return state

    @_lib.on_device(lambda self, inputs, *a, **k: inputs.device)
    def _forward_cuda(self, inputs, noise=None, actions_in=None, want=("value", "action", "logp")):
        flat = self.flat_params()
        x = inputs.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        B, A = x.shape[0], self.act_dim
        dev = x.device
        value = torch.empty(B, 1, device=dev) if "value" in want else None
        action = torch.empty(B, A, device=dev) if "action" in want else None
        logp = torch.empty(B, 1, device=dev) if "logp" in want else None
        ent = torch.empty((), device=dev) if "entropy" in want else None
        if actions_in is not None:
            actions_in = actions_in.detach().float().contiguous()
        rc = _lib.lib().sg_policy_forward(_lib.ptr(flat), self.obs_dim, self.hidden_size, A, _lib.ptr(x), B,
                                          _lib.ptr(noise), _lib.ptr(actions_in), _lib.ptr(value), _lib.ptr(action),
                                          _lib.ptr(logp), _lib.ptr(ent), _lib.current_stream())
        _lib.check(rc, "sg_policy_forward")
        return value, action, logp, ent
