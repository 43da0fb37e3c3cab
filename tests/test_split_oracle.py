"""SplitPolicy (third_party/a2c_ppo_acktr/model_split.py) oracle: pinned against the real reference when it is
mounted, and against golden vectors generated from it (oracle/make_golden.py --split) anywhere."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from oracle import ppo_gail_oracle as orc
from oracle import ref_shim

SPLIT_CASES = ["split_hopper_seed3.npz", "split_feet4_seed4.npz"]


class SplitGolden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name))
        (self.T, self.N, self.O, self.A, self.H, self.feet, self.ppo_epoch, self.nmb, self.seed) = [int(v) for v in self.z["meta_dims"]]
        self.S = self.T * self.N
        self.entropy_coef = float(self.z["entropy_coef"])

    def t(self, key):
        return torch.from_numpy(np.array(self.z[key]))

    def params(self, which="sp0"):
        return {k: self.t("%s_%s" % (which, k)) for k in orc.SPLIT_KEYS}

    def buffer(self):
        keys = ("obs", "obs_feat", "recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs",
                "actions", "masks", "bad_masks")
        return {k: self.t("buf_" + k).clone() for k in keys}

    def hyper(self):
        return orc.PPOHyper(ppo_epoch=self.ppo_epoch, num_mini_batch=self.nmb, entropy_coef=self.entropy_coef)

    def chunks(self):
        mb = self.S // self.nmb
        perm = self.t("ppo_perm")
        return [[perm[e, i * mb:(i + 1) * mb] for i in range(self.nmb)] for e in range(self.ppo_epoch)]


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_split_oracle_replays_golden(case):
    g = SplitGolden(case)
    p, buf = g.params(), g.buffer()
    v, _, _ = orc.split_forward(p, buf["obs"][-1])
    assert torch.equal(v, g.t("next_value"))
    ve, lpe, ente = orc.split_evaluate(p, buf["obs"][3], buf["actions"][3])
    assert torch.equal(ve, g.t("eval_value")) and torch.equal(lpe, g.t("eval_logp")) and float(ente) == float(g.z["eval_entropy"])
    ora = orc.PPOOracle(p, g.hyper(), keys=orc.SPLIT_KEYS, evaluate=orc.split_evaluate)
    torch.set_rng_state(g.t("rng_before_ppo"))
    out = ora.update(buf)
    assert np.array_equal(np.array(out), g.z["ppo_losses"])
    for k, val in g.params("sp1").items():
        assert torch.equal(ora.p[k].data, val), k
    ora2 = orc.PPOOracle(p, g.hyper(), keys=orc.SPLIT_KEYS, evaluate=orc.split_evaluate)
    assert np.array_equal(np.array(ora2.update(buf, index_chunks=g.chunks())), g.z["ppo_losses"])


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_split_oracle_matches_reference_bitexact():
    ref = ref_shim.load()
    torch.manual_seed(5)
    sp = ref.model_split.SplitPolicy((11,), ref_shim.BoxSpace(14), base_kwargs={"hidden_size": 32, "num_feet": 2})
    torch.manual_seed(5)
    p = orc.init_split_policy(11, 32, 2)
    assert [n for n, _ in sp.named_parameters()] == [
        "base.actor_contact.0.weight", "base.actor_contact.0.bias", "base.actor_contact.2.weight", "base.actor_contact.2.bias",
        "base.actor_actuator.0.weight", "base.actor_actuator.0.bias", "base.actor_actuator.2.weight", "base.actor_actuator.2.bias",
        "base.critic_full.0.weight", "base.critic_full.0.bias", "base.critic_full.2.weight", "base.critic_full.2.bias",
        "base.critic_full.4.weight", "base.critic_full.4.bias", "dist.contact_mean.weight", "dist.contact_mean.bias",
        "dist.actuator_mean.weight", "dist.actuator_mean.bias", "dist.contact_logstd.weight", "dist.contact_logstd.bias",
        "dist.actuator_logstd.weight", "dist.actuator_logstd.bias"]
    for (n, q), k in zip(sp.named_parameters(), orc.SPLIT_KEYS):
        assert torch.equal(q.data, p[k]), (n, k)
    assert torch.equal(torch.rand(2), (torch.manual_seed(5), ref.model_split.SplitPolicy(
        (11,), ref_shim.BoxSpace(14), base_kwargs={"hidden_size": 32, "num_feet": 2}), torch.rand(2))[2])
    x, a = torch.randn(6, 11), torch.randn(6, 14)
    with torch.no_grad():
        v, lp, ent, _ = sp.evaluate_actions(x, None, None, a)
        vd, ad, lpd, _ = sp.act(x, None, None, deterministic=True)
    vo, lpo, ento = orc.split_evaluate(p, x, a)
    assert torch.equal(v, vo) and torch.equal(lp, lpo) and torch.equal(ent, ento)
    vo2, ao2, lpo2 = orc.split_act(p, x, deterministic=True)
    assert torch.equal(vd, vo2) and torch.equal(ad, ao2) and torch.equal(lpd, lpo2)


# ---- the product-side SplitPolicy class (host logic; CUDA parity is in tests/test_gpu_split.py) -------------------
def test_split_policy_class_matches_oracle_on_cpu():
    import simgan_b200 as sg
    from simgan_b200 import compat
    torch.manual_seed(8)
    sp = sg.SplitPolicy((14,), ref_shim.BoxSpace(7), base_kwargs={"hidden_size": 100, "num_feet": 1})
    after = torch.rand(3)
    torch.manual_seed(8)
    p = orc.init_split_policy(14, 100, 1)
    assert torch.equal(torch.rand(3), after)                     # same CPU generator consumption as the reference
    for (n, q), k in zip(sp.named_parameters(), orc.SPLIT_KEYS):
        assert torch.equal(q.data, p[k]), (n, k)
    x, a = torch.randn(5, 14), torch.randn(5, 7)
    with torch.no_grad():
        v, lp, ent, _ = sp.evaluate_actions(x, None, None, a)
        vd, ad, lpd, _ = sp.act(x, None, None, deterministic=True)
    vo, lpo, ento = orc.split_evaluate(p, x, a)
    assert torch.equal(v, vo) and torch.equal(lp, lpo) and torch.equal(ent, ento)
    vo2, ao2, lpo2 = orc.split_act(p, x, deterministic=True)
    assert torch.equal(vd, vo2) and torch.equal(ad, ao2) and torch.equal(lpd, lpo2)
    assert (sp.obs_dim, sp.hidden_size, sp.num_feet, sp.act_dim) == (14, 100, 1, 7)
    assert not sp.is_recurrent and sp.recurrent_hidden_state_size == 1
    compat.install()
    try:
        from third_party.a2c_ppo_acktr.model_split import SplitPolicy
        assert SplitPolicy is sg.SplitPolicy
    finally:
        compat.uninstall()
    import io
    buf = io.BytesIO()
    torch.save([sp, None], buf)
    buf.seek(0)
    sp2, _ = torch.load(buf, weights_only=False)
    with torch.no_grad():
        assert torch.equal(sp2.act(x, None, None, deterministic=True)[1], ad)


def test_split_layout_table():
    from simgan_b200 import _lib
    offs, total = _lib.split_layout(14, 100, 1)
    sizes = [1400, 100, 10000, 100] * 2 + [1400, 100, 10000, 100, 100, 1, 400, 4, 300, 3, 400, 4, 300, 3]
    assert len(offs) == 22 and total >= sum(sizes)
    spans = sorted((o, o + n) for o, n in zip(offs, sizes))
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))          # no overlap
    # mean and log-std heads of each actor are adjacent: one (2n, H) matrix / (2n) bias per net
    assert offs[18] == offs[14] + 400 and offs[19] == offs[15] + 4        # contact
    assert offs[20] == offs[16] + 300 and offs[21] == offs[17] + 3        # actuator
