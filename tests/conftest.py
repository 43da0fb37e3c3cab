import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(autouse=True)
def _one_thread():
    # the reference forces one intra-op thread (third_party/a2c_ppo_acktr/main_gail_dyn_ppo.py:64);
    # oracle results are bit-reproducible only at a fixed thread count.
    import torch
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(old)
