#!/usr/bin/env python
"""Times PairDPD alone (obmd_pair_compute through the C ABI) on a tiled box for a list of
(OBMD_B200_TILE, OBMD_B200_ROWS, OBMD_B200_ZC, OBMD_B200_PT[, OBMD_B200_COLS]) settings.  usage: python tools/pair_bench.py [reps x y z] [iters]
Prints one line per setting: mean kernel-family time from the library's CUDA-event timers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common_cases as cc  # noqa: E402


def main():
    reps = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 5, 4)
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    settings = os.environ.get("PAIR_BENCH_SETTINGS", "1:0:0:0 2:0:0:0:0 2:4:5:0:4 2:4:4:0:4 2:4:6:0:4 2:4:7:0:2 2:4:11:0:2 2:4:9:0:2 2:2:11:0:2 2:2:11:0:4 2:4:11:0:1 2:2:15:0:4 2:4:5:256:4 2:4:5:288:4").split()
    tag, x, box = cc.tiled(reps)
    v = np.random.default_rng(9).normal(0, 1, x.shape)
    n = len(x)
    for s in settings:
        tile, rows, zc, pt, cols = (s.split(":") + ["0"])[:5]
        os.environ["OBMD_B200_TILE"], os.environ["OBMD_B200_ROWS"], os.environ["OBMD_B200_COLS"] = tile, rows, cols
        os.environ["OBMD_B200_ZC"], os.environ["OBMD_B200_PT"] = zc, pt
        sim = cc.device_sim(box, fix=False)
        sim.upload(x, v, np.ones(n, np.int32), tag)
        sim.build_cells(); sim.pbc()
        for k in range(3):
            sim.pair_compute(k)
        sim.timing(True); sim.timing_reset()
        for k in range(iters):
            sim.pair_compute(10 + k)
        ms, nl = sim.timing_get("pair")
        f = sim.download(("f",))["f"]
        print("tile=%s rows=%s zc=%s threads=%s cols=%s : %.1f us / launch-family  (%d launches)  %.1f GB/s algorithmic  |f|max %.3f  sum %s"
              % (tile, rows, zc, pt, cols, 1e3 * ms / iters, nl, 80.0 * n / (ms / iters * 1e-3) / 1e9, np.abs(f).max(),
                 np.array2string(f.sum(0), precision=3)), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
