"""simgan_b200 -- B200-native PPO+GAIL inner loop behind the reference's Python surface.

Classes mirror third_party/a2c_ppo_acktr of jyf588/SimGAN (Policy, PPO, gail.Discriminator,
RolloutStorage); the minibatch math runs in hand-written sm_100a kernels behind the C ABI declared
in include/simgan_b200.h (loaded by simgan_b200._lib).
"""
from ._lib import SgError  # noqa: F401
from .model import Policy, MLPBase  # noqa: F401
from .model_split import SplitPolicy  # noqa: F401
from .storage import RolloutStorage  # noqa: F401
from .running_mean_std import RunningMeanStd  # noqa: F401
from . import algo  # noqa: F401
from .algo import PPO  # noqa: F401
from .algo.gail import Discriminator  # noqa: F401
from .feed import RolloutFeeder, ReturnNormalizer  # noqa: F401
from .expert_data import load_sas_wpast_from_pickle, select_and_merge_sas, expert_tensor  # noqa: F401

__version__ = "0.1.0"
