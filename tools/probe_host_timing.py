import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, bench, torch
x, box = bench.tiled((4,5,4))
sim, kw = bench.build_sim(x, box, 0, "R1")
sim.run(1, 150)
torch.cuda.synchronize()
for mode in ("per-step", "block"):
    t0 = time.perf_counter()
    if mode == "per-step":
        for s in range(50): sim.run(200+s, 1)
    else:
        sim.run(300, 50)
    torch.cuda.synchronize()
    print(mode, "ms/step", (time.perf_counter()-t0)*1e3/50)
# per-hook host timing
import collections
acc = collections.Counter()
for s in range(50):
    for name, fn in (("nve_i", sim.nve_initial), ("build", sim.build_cells), ("pre", sim.pre_exchange), ("pbc", sim.pbc), ("pair", lambda: sim.pair_compute(400+s)), ("post", sim.post_force), ("nve_f", sim.nve_final)):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); acc[name] += time.perf_counter()-t0
        if name == "pre": acc["rounds"] += r[0].usher_rounds; acc["trials"] += r[0].ntrials
print({k: round(v*1e3/50, 4) for k, v in acc.items()})
