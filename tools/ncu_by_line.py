#!/usr/bin/env python
"""Per-source-line profile of one kernel: joins `nvdisasm -g` line info of the cubin inside a .so
with the per-SASS-instruction counts of an .ncu-rep (ncu --set full --import-source on).
usage: python tools/ncu_by_line.py rep.ncu-rep lib.so <mangled-name substring> <ncu kernel substring> [top N] [launch #]
The n-th SASS instruction of the function in the cubin is the n-th row of ncu's source page."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, mangled):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    out = []
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
        infun, cur = False, ("?", 0)
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                infun = mangled in ln
                continue
            if not infun:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                out.append(cur)
        if out:
            break
    return out


def ncu_rows(rep, kname, which=0):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--launch-skip", str(which), "--launch-count", "1"], capture_output=True, text=True).stdout
    which = 0
    rows = list(csv.reader(txt.splitlines()))
    cur, blocks = None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    hits = [b for b in blocks if kname in b["name"]]
    if which >= len(hits):
        raise SystemExit("kernel not found (%d launches match)" % len(hits))
    hdr = hits[which]["rows"][0]
    return hdr, [r for r in hits[which]["rows"][1:] if len(r) == len(hdr)]


def main():
    rep, so, mangled, kname = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    which = int(sys.argv[6]) if len(sys.argv) > 6 else 0  # n-th matching launch in the report
    lines = sass_lines(so, mangled)
    hdr, rows = ncu_rows(rep, kname, which)
    ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    if len(lines) != len(rows):
        print("warning: %d SASS lines in cubin vs %d in report (different build?)" % (len(lines), len(rows)))
    agg = defaultdict(lambda: [0, 0])
    for k, r in enumerate(rows):
        key = lines[k] if k < len(lines) else ("?", 0)
        agg[key][0] += int(r[ie])
        agg[key][1] += int(r[smp])
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("total warp instructions %d, samples %d" % (ti, ts))
    src = {}
    print("%-22s %8s %8s" % ("file:line", "inst %", "smp %"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        f, l = key
        if f not in src:
            for root in (os.path.dirname(os.path.abspath(so)) + "/csrc", "."):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    src[f] = open(p).read().splitlines()
                    break
            else:
                src[f] = []
        text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
        print("%-22s %7.2f%% %7.2f%%  %s" % ("%s:%d" % (f, l), 100.0 * v[0] / ti, 100.0 * v[1] / max(ts, 1), text[:100]))


if __name__ == "__main__":
    main()
