"""The tensor-core PPO tiles (sg_ppo_config.mode 4: tcgen05 kind::tf32 MMAs with 3xTF32 operand splitting,
csrc/sg_ppo_mma.cuh) against the CPU oracle: per-step losses within the contract's 1e-4 relative, parameters after
the update, and agreement with the CUDA-core tiles.  Sizes cover both job heights (128 rows at hidden 64, 64 rows at
hidden 128/256), both nets' head widths, an observation width that is not a multiple of 4 (scalar gathers, zero-padded
K), several jobs per CTA (partial gradients accumulated across jobs) and a ragged last tile."""
import numpy as np
import pytest
import torch

from oracle import ppo_gail_oracle as orc

import gpu_util as gu
import simgan_b200 as sg

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-4


def _run(O, H, A, T, N, nmb, epochs, mode, seed=4):
    torch.manual_seed(1)
    p = orc.init_policy(O, H, A)
    buf = orc.synth_rollout(T, N, O, A, 3, p, seed=seed, ep_len=50.0)
    nv = orc.policy_forward(p, buf["obs"][-1])[0]
    orc.compute_returns(buf, nv, True, 0.99, 0.95, True)
    hyper = orc.PPOHyper(ppo_epoch=epochs, num_mini_batch=nmb)
    ora = orc.PPOOracle(p, hyper)
    torch.manual_seed(7)
    trace = []
    ora.update(buf, trace=trace)
    pol = gu.make_policy(p, O, H, A)
    agent = sg.PPO(pol, 0.2, epochs, nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    agent.kernel_mode = mode
    torch.manual_seed(7)
    agent.update(gu.make_storage(buf, O, A, 3))
    return np.array(trace), agent.last_trace.double().numpy(), ora, pol


@pytest.mark.parametrize("O,H,A,T,N,nmb", [
    (14, 64, 7, 100, 8, 4),          # 200-row minibatches: 2 tiles of 128 rows, the second ragged
    (111, 64, 12, 300, 128, 1),      # 38400 rows: 300 tiles x 2 nets on 148 CTAs -> 4-5 jobs per CTA, O % 4 != 0
    (64, 128, 28, 250, 16, 2),       # hidden 128 -> 64-row jobs, one M=128 block for the weight gradients
    (64, 256, 28, 251, 8, 2),        # hidden 256 -> 64-row jobs, two M=128 blocks, ragged tail (1004 rows)
    (20, 256, 5, 1251, 8, 2),        # 5004-row minibatches at hidden 256
])
def test_tensor_core_tiles_vs_oracle(O, H, A, T, N, nmb):
    tr_o, tr, ora, pol = _run(O, H, A, T, N, nmb, 2, sg.PPO.MMA_MODE)
    scale = np.abs(tr_o).max(axis=0)
    scale[1] = max(scale[1], 0.5)
    assert np.all(np.abs(tr[0] - tr_o[0]) <= 1e-5 * scale + 1e-7), (tr[0], tr_o[0])
    assert np.all(np.abs(tr - tr_o) <= LOSS_RTOL * scale + 1e-6), np.abs(tr - tr_o).max(axis=0) / scale
    pm, po = gu.policy_params(pol), ora.params()
    for k in orc.POLICY_KEYS:
        assert torch.allclose(pm[k], po[k].reshape(-1), rtol=1e-3, atol=2e-5), k


def test_tensor_core_and_cuda_core_tiles_agree():
    a = _run(14, 64, 7, 512, 16, 4, 2, sg.PPO.MMA_MODE)
    b = _run(14, 64, 7, 512, 16, 4, 2, 3)
    assert np.allclose(a[1], b[1], rtol=2e-5, atol=1e-6)
    assert torch.allclose(a[3].flat_params(), b[3].flat_params(), rtol=1e-3, atol=1e-5)


def test_unsupported_sizes_are_rejected():
    p = orc.init_policy(11, 100, 3)
    pol = gu.make_policy(p, 11, 100, 3)
    buf = orc.synth_rollout(16, 4, 11, 3, 3, p, seed=1, ep_len=50.0)
    agent = sg.PPO(pol, 0.2, 1, 2, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
    agent.kernel_mode = sg.PPO.MMA_MODE
    with pytest.raises(sg.SgError):
        agent.update(gu.make_storage(buf, 11, 3, 3))
