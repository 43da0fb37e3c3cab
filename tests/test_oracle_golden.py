"""Oracle vs golden vectors produced by the REAL reference (oracle/make_golden.py).  Runs anywhere
(no GPU, no /root/reference): this is what pins the oracle on the GPU box."""
import os

import numpy as np
import pytest
import torch

from golden_util import CASES, GOLDEN_DIR, Golden
from oracle import ppo_gail_oracle as orc


@pytest.mark.parametrize("case", CASES)
def test_full_update_phase_replay(case):
    g = Golden(case)
    buf = g.buffer()
    expert = g.t("expert")
    disc = orc.DiscOracle(g.disc())
    # D epochs, consuming the CPU generator from the recorded state (checks RNG-order emulation)
    torch.set_rng_state(g.t("rng_before_disc"))
    losses = [disc.update_epoch(expert, buf, batch_size=g.gail_batch, drop_last=len(expert) > g.gail_batch)
              for _ in range(g.gail_epoch)]
    assert np.array_equal(np.array(losses), g.z["disc_losses"])
    for k, v in g.disc("disc1").items():
        assert torch.equal(disc.d[k].data, v), k
    # D epochs again from explicit index streams
    disc2 = orc.DiscOracle(g.disc())
    losses2 = [disc2.update_epoch(expert, buf, batch_size=g.gail_batch, replay=g.disc_replay(e))
               for e in range(g.gail_epoch)]
    assert np.array_equal(np.array(losses2), g.z["disc_losses"])
    # relabel
    r_sa = orc.alive_bonus_offset(buf["masks"], g.T, g.N, float(g.z["gail_tar_length"]))
    assert r_sa == float(g.z["r_sa"])
    rms = orc.RunningMeanStd(shape=())
    means = orc.relabel_rewards(disc, rms, buf, 0.99, -r_sa)
    assert np.array_equal(np.array(means), g.z["relabel_mean_returns"])
    assert torch.equal(buf["rewards"], g.t("relabel_rewards"))
    assert torch.equal(disc.returns, g.t("relabel_disc_returns"))
    assert np.array_equal(np.array([float(rms.mean), float(rms.var), float(rms.count)]), g.z["relabel_rms"])
    # GAE
    orc.compute_returns(buf, g.t("next_value"), True, 0.99, 0.95, True)
    assert torch.equal(buf["returns"], g.t("gae_returns"))
    adv = buf["returns"][:-1] - buf["value_preds"][:-1]
    assert [float(adv.mean()), float(adv.std())] == list(g.z["adv_mean_std"])
    # PPO
    ppo = orc.PPOOracle(g.policy(), g.hyper())
    torch.set_rng_state(g.t("rng_before_ppo"))
    out = ppo.update(buf)
    assert np.array_equal(np.array(out), g.z["ppo_losses"])
    for k, v in g.policy("pol1").items():
        assert torch.equal(ppo.p[k].data, v), k
    ppo2 = orc.PPOOracle(g.policy(), g.hyper())
    assert np.array_equal(np.array(ppo2.update(buf, index_chunks=g.ppo_chunks())), g.z["ppo_losses"])


def test_next_value_and_policy_forward():
    g = Golden(CASES[0])
    buf = g.buffer()
    v, _, _ = orc.policy_forward(g.policy(), buf["obs"][-1])
    assert torch.equal(v, g.t("next_value"))


def test_mini_expert_pkl():
    torch.manual_seed(0)
    cols = orc.load_sas_wpast(os.path.join(GOLDEN_DIR, "mini_expert.pkl"), downsample_freq=2, load_num_trajs=3)
    merged = orc.merge_sas(cols)
    assert np.array_equal(merged, np.load(os.path.join(GOLDEN_DIR, "mini_expert_merged.npy")))
    assert merged.shape[1] == 25


def test_full_expert_fixture_shape():
    x = np.load(os.path.join(GOLDEN_DIR, "hopper_expert_sas_f32.npy"))
    assert x.shape == (17555, 25) and x.dtype == np.float32


def test_running_mean_std_kat():
    """Known-answer test carried over from running_mean_std.py:110-124 of the reference."""
    rng = np.random.RandomState(0)
    for shapes in [((3,), (4,), (5,)), ((3, 2), (4, 2), (5, 2))]:
        xs = [rng.randn(*s) for s in shapes]
        rms = orc.RunningMeanStd(epsilon=0.0, shape=xs[0].shape[1:])
        for xi in xs:
            rms.update(xi)
        x = np.concatenate(xs, axis=0)
        np.testing.assert_allclose([x.mean(axis=0), x.var(axis=0)], [rms.mean, rms.var])
