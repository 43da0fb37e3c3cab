"""Device timeline of one update phase at a BASELINE config: for every hot-path launch its start offset, duration and
the idle gap before it (CUDA events on the launch stream).  Shows where the step time that is not inside a kernel goes.
Usage: python profiles/gap_timeline.py [cfg2]"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

import bench
from simgan_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.Workload(bench.CONFIGS[name], 0, torch.device("cuda:0"))
for _ in range(4):
    w.update_phase(False)
torch.cuda.synchronize()
_lib.timer.enabled = True
for it in range(3):
    _lib.timer.records = []
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    w.update_phase(False)
    t1.record()
    torch.cuda.synchronize()
    recs = _lib.timer.records
    total = t0.elapsed_time(t1)
    prev_end = t0
    busy = 0.0
    lines = []
    for nm, s, e in recs:
        gap = prev_end.elapsed_time(s)
        dur = s.elapsed_time(e)
        busy += dur
        lines.append("%-16s start %8.3f  dur %7.3f  gap_before %7.3f" % (nm, t0.elapsed_time(s), dur, gap))
        prev_end = e
    tail = prev_end.elapsed_time(t1)
    if it == 2:
        print("\n".join(lines))
    print("iteration %d: total %.3f ms, in timed launches %.3f ms, tail %.3f ms" % (it, total, busy, tail))
