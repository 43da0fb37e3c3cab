"""One PPO epoch with the tensor-core tiles at a chosen size (driver for ncu captures / phase-cycle reads).
usage: python profiles/ppo_mma_probe.py O H A T N nmb [mode]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import simgan_b200 as sg

O, H, A, T, N, nmb = [int(x) for x in sys.argv[1:7]]
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)


class Box:
    def __init__(self, n):
        self.shape = (n,)


pol = sg.Policy((O,), Box(A), base_kwargs={"recurrent": False, "hidden_size": H}).to(dev)
rs = sg.RolloutStorage(T, N, (O,), Box(A), 1, 3)
rs.to(dev)
rs.obs.normal_()
rs.actions.normal_()
rs.rewards.normal_()
rs.value_preds.normal_()
rs.returns.normal_()
rs.action_log_probs.fill_(-float(A))
agent = sg.PPO(pol, 0.2, 1, nmb, 0.5, 0.01, lr=3e-4, eps=1e-5, max_grad_norm=0.5)
agent.kernel_mode = mode
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = agent.update(rs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("update %d: %.3f ms, %.1f us per step" % (it, dt * 1e3, dt * 1e6 / nmb), out)
