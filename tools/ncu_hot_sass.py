#!/usr/bin/env python
"""Per-SASS-instruction hot spots of one kernel in an .ncu-rep (needs --set full --import-source on).
usage: python tools/ncu_hot_sass.py rep.ncu-rep <kernel substring> [top N]
Prints total warp instructions, the instruction mix by opcode, and the top-N instructions by stall samples."""
import csv
import subprocess
import sys
from collections import Counter


def main(path, kname, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if kname not in b["name"]:
            continue
        hdr = b["rows"][0]
        src, smp, ie, tie = (hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed",
                                                     "Thread Instructions Executed"))
        data = [r for r in b["rows"][1:] if len(r) > tie]
        tot_i = sum(int(r[ie]) for r in data)
        tot_t = sum(int(r[tie]) for r in data)
        tot_s = sum(int(r[smp]) for r in data)
        print("== %s\nwarp instructions %d, thread instructions %d (%.1f lanes), samples %d, SASS lines %d" %
              (b["name"][:90], tot_i, tot_t, tot_t / max(tot_i, 1), tot_s, len(data)))
        mix, mixs = Counter(), Counter()
        for r in data:
            toks = r[src].split()
            op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
            op = op.split(".")[0]
            mix[op] += int(r[ie])
            mixs[op] += int(r[smp])
        print("opcode mix (share of warp instructions | share of samples):")
        for op, n in mix.most_common(25):
            print("  %-10s %6.2f %%   %6.2f %%" % (op, 100.0 * n / tot_i, 100.0 * mixs[op] / max(tot_s, 1)))
        print("top instructions by samples:")
        for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][smp]))[:top]:
            print("  #%4d %6.2f %% smp  %9d inst  %s" % (k, 100.0 * int(r[smp]) / max(tot_s, 1), int(r[ie]), r[src].strip()))
        break


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
