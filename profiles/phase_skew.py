#!/usr/bin/env python
"""Per-CTA phase-cycle distribution of the persistent kernels on the cfg2 workload (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

c = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
w = bench.Workload(c, 0, torch.device("cuda:0"))
for _ in range(3):
    w.update_phase(False)
torch.cuda.synchronize()
for name, obj, n, steps, labels in (("ppo", w.agent, 128, 32, ["image", "tile", "bar1", "reduce", "bar2", "adam", "bar3"]),
                                    ("disc", w.disc, 128, w.n_disc_batches, ["image", "tile", "bar1", "reduce_adam", "bar2"])):
    pc = obj.phase_cycles_all(n).double() / steps
    print(name, "cycles per optimizer step: min / median / max over CTAs, and CTA 0")
    for i, lab in enumerate(labels):
        col = pc[:, i]
        print("  %-12s %8.0f %8.0f %8.0f   cta0 %8.0f  argmax %d" % (lab, col.min(), col.median(), col.max(), col[0], int(col.argmax())))
    print("  total        %8.0f" % pc[0, :len(labels)].sum())
