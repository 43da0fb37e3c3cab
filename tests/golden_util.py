"""Helpers to load tests/golden/*.npz (written by oracle/make_golden.py from the real reference)."""
import os

import numpy as np
import torch

from oracle import ppo_gail_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["hopper_cfg1_seed0.npz", "ragged_seed1.npz", "laika_dims_seed2.npz"]
BUF_KEYS = ("obs", "obs_feat", "recurrent_hidden_states", "rewards", "value_preds", "returns",
            "action_log_probs", "actions", "masks", "bad_masks")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name))
        (self.T, self.N, self.O, self.A, self.H, self.F, self.HD, self.gail_epoch, self.gail_batch,
         self.ppo_epoch, self.nmb, self.seed) = [int(v) for v in self.z["meta_dims"]]
        self.S = self.T * self.N

    def t(self, key):
        return torch.from_numpy(np.array(self.z[key]))

    def policy(self, which="pol0"):
        return {k: self.t("%s_%s" % (which, k)) for k in orc.POLICY_KEYS}

    def disc(self, which="disc0"):
        return {k: self.t("%s_%s" % (which, k)) for k in orc.DISC_KEYS}

    def buffer(self):
        return {k: self.t("buf_" + k).clone() for k in BUF_KEYS}

    def hyper(self):
        return orc.PPOHyper(ppo_epoch=self.ppo_epoch, num_mini_batch=self.nmb)

    def ppo_chunks(self):
        mb = self.S // self.nmb
        perm = self.t("ppo_perm")
        return [[perm[e, i * mb:(i + 1) * mb] for i in range(self.nmb)] for e in range(self.ppo_epoch)]

    def disc_replay(self, e):
        ei, pi, al = self.t("disc_expert_idx")[e], self.t("disc_policy_idx")[e], self.t("disc_alpha")[e]
        return list(ei), list(pi), list(al)
