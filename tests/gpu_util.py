"""Helpers for the -m gpu parity tests: build simgan_b200 objects on cuda:0 from golden / oracle data."""
import torch

from oracle import ppo_gail_oracle as orc
from oracle.ref_shim import BoxSpace

import simgan_b200 as sg
from simgan_b200.algo import gail as sg_gail

DEV = "cuda:0"


def make_policy(params, O, H, A):
    """simgan_b200.Policy on the GPU holding the given oracle/golden parameter dict."""
    pol = sg.Policy((O,), BoxSpace(A), base_kwargs={"recurrent": False, "hidden_size": H})
    for p, k in zip(pol.hot_path_parameters(), orc.POLICY_KEYS):
        p.data.copy_(params[k].reshape(p.shape))
    pol.to(DEV)
    return pol


def policy_params(pol):
    return {k: p.detach().cpu().reshape(-1) for p, k in zip(pol.hot_path_parameters(), orc.POLICY_KEYS)}


def make_disc(params, F, HD):
    d = sg_gail.Discriminator(F, HD, torch.device(DEV))
    for p, k in zip(d.hot_path_parameters(), orc.DISC_KEYS):
        p.data.copy_(params[k].reshape(p.shape).to(DEV))
    return d


def disc_params(d):
    return {k: p.detach().cpu().reshape(-1) for p, k in zip(d.hot_path_parameters(), orc.DISC_KEYS)}


def make_storage(buf, O, A, F):
    T, N = buf["rewards"].shape[:2]
    rs = sg.RolloutStorage(T, N, (O,), BoxSpace(A), 1, F)
    for k, v in buf.items():
        getattr(rs, k).copy_(v)
    rs.to(DEV)
    return rs


def rel_err(a, b, floor=1e-30):
    """max |a-b| relative to max |b|; ``floor`` is the natural scale of the quantity (keeps a batch whose
    values happen to cancel to ~0 from turning fp32 rounding into a huge relative error)."""
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))
