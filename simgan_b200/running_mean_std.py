"""float64 running mean/variance (third_party/a2c_ppo_acktr/baselines/common/running_mean_std.py:27-56).

Host-side object with the reference's attributes (mean, var, count).  The whole-rollout relabel
(Discriminator.relabel_rollout) carries the same three numbers through the device kernel
sg_disc_relabel and writes them back here."""
import numpy as np


class RunningMeanStd(object):
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        self.mean, self.var, self.count = update_mean_var_count_from_moments(
            self.mean, self.var, self.count, batch_mean, batch_var, batch_count)


def update_mean_var_count_from_moments(mean, var, count, batch_mean, batch_var, batch_count):
    """Chan et al. parallel-variance merge, same evaluation order as the reference (:45-56)."""
    delta = batch_mean - mean
    total = count + batch_count
    merged_mean = mean + delta * batch_count / total
    m2 = var * count + batch_var * batch_count + np.square(delta) * count * batch_count / total
    return merged_mean, m2 / total, total
