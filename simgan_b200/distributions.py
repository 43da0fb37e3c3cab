"""Diagonal Gaussian action head (third_party/a2c_ppo_acktr/distributions.py:51-59, 91-118)."""
import torch
import torch.nn as nn

from .utils import AddBias, init


class FixedNormal(torch.distributions.Normal):
    """Normal whose log-prob / entropy are summed over the action dimension (distributions.py:51-59)."""

    def log_probs(self, actions):
        return super().log_prob(actions).sum(-1, keepdim=True)

    def entropy(self):
        return super().entropy().sum(-1)

    def mode(self):
        return self.mean


class DiagGaussian(nn.Module):
    """Linear mean (orthogonal gain 1, zero bias, then weights/50) + state-independent log-std
    parameter of shape (A,1) initialised to -0.5 (distributions.py:91-118)."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.fc_mean = init(nn.Linear(num_inputs, num_outputs), nn.init.orthogonal_,
                            lambda x: nn.init.constant_(x, 0))
        self.logstd = AddBias(torch.ones(num_outputs) * -0.5)
        for p in self.fc_mean.parameters():
            p.data = p.data / 50.0

    def reset_variance(self, num_outputs, log_std):
        self.logstd = AddBias(torch.ones(num_outputs) * log_std)

    def forward(self, x):
        mean = self.fc_mean(x)
        logstd = self.logstd(torch.zeros_like(mean))
        return FixedNormal(mean, logstd.exp())
