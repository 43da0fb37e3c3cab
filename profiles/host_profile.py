import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import torch, bench
c = bench.CONFIGS["cfg2"]
w = bench.Workload(c, 0, torch.device("cuda:0"))
for _ in range(3): w.update_phase(False)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0=time.perf_counter()
pr.enable()
for _ in range(10): w.update_phase(False)
torch.cuda.synchronize()
pr.disable()
print("wall per iter ms", (time.perf_counter()-t0)*100)
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
