"""AddBias / init / linear LR schedule with the reference's contracts
(third_party/a2c_ppo_acktr/utils.py:54-78)."""
import torch.nn as nn


class AddBias(nn.Module):
    """Learnable per-channel bias stored with shape (C, 1) (reference utils.py:54-65); the
    DiagGaussian log-std lives here so state_dicts / pickles stay interchangeable."""

    def __init__(self, bias):
        super().__init__()
        self._bias = nn.Parameter(bias.unsqueeze(1))

    def forward(self, x):
        shape = (1, -1) if x.dim() == 2 else (1, -1, 1, 1)
        return x + self._bias.t().view(*shape)


def init(module, weight_init, bias_init, gain=1):
    """Apply weight_init(weight, gain=gain) and bias_init(bias); returns the module (utils.py:75-78)."""
    weight_init(module.weight.data, gain=gain)
    bias_init(module.bias.data)
    return module


def update_linear_schedule(optimizer, epoch, total_num_epochs, initial_lr):
    """lr <- initial_lr * (1 - epoch/total) written into every param group (utils.py:68-72)."""
    lr = initial_lr - (initial_lr * (epoch / float(total_num_epochs)))
    for group in optimizer.param_groups:
        group["lr"] = lr


def get_vec_normalize(venv):
    """Walk a wrapper chain for an object exposing ``ob_rms`` (utils.py:44-50); env plumbing itself is
    out of scope, so this only duck-types."""
    while venv is not None:
        if type(venv).__name__ == "VecNormalize":
            return venv
        venv = getattr(venv, "venv", None)
    return None
