"""Where a kernel loses lanes: per source line, the warp instructions executed and the threads
that were active in them, from the source page of an ncu report captured with
`--set full --import-source on` (and a build with -lineinfo).

    python scripts/ncu_lost_lanes.py gpurun_out/prof_lnl_v7.ncu-rep [kernel-regex] [launch-index]

Prints the average active lanes per warp instruction, the share of instructions per function of
tri_model.cuh, and the source lines ranked by lost lane-instructions (32 x instructions - thread
instructions).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "lnl_kernel"
launch = sys.argv[3] if len(sys.argv) > 3 else "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                      "--kernel-id", "::regex:%s:%s" % (kernel, launch)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, cur_file, last_line, last_src = None, None, None, ""
per_line = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        if r[0]:
            last_line, last_src = int(r[0]), r[1]
        if not r[2]:
            continue
        try:
            inst, tinst = int(r[7]), int(r[8])
        except ValueError:
            continue
        p = per_line[(os.path.basename(cur_file), last_line)]
        p[0] += inst
        p[1] += tinst
        p[2] = last_src.strip()[:90]
tot_i = sum(p[0] for p in per_line.values())
tot_t = sum(p[1] for p in per_line.values())
print("warp instructions %.3e, average active lanes %.2f" % (tot_i, tot_t / tot_i))

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
model = open(os.path.join(root, "triceratops_b200", "csrc", "tri_model.cuh")).read().splitlines()
funcs = [(i, m.group(1)) for i, l in enumerate(model, 1)
         for m in [re.match(r'^TRI_HD\s+(?:inline\s+)?[\w:<>\s\*&]+?\s+(\w+)\s*\(', l)] if m]


def fn_of(line):
    name = "?"
    for i, n in funcs:
        if i <= line:
            name = n
    return name


agg = collections.Counter()
for (f, l), p in per_line.items():
    agg["model:" + fn_of(l) if f == "tri_model.cuh" else f] += p[0]
print("\ninstruction share by function")
for k, v in agg.most_common(14):
    print("  %-32s %5.1f %%" % (k, 100 * v / tot_i))

lost = sorted(((32 * p[0] - p[1], k, p) for k, p in per_line.items()), reverse=True)
tl = sum(l for l, _, _ in lost)
print("\nlost lane-instructions: %.1f %% of 32 x instructions; by source line" % (100 * tl / (32 * tot_i)))
for l, k, p in lost[:16]:
    print("  %5.2f %% of the loss | %5.2f %% of inst | %4.1f lanes | %s:%d | %s"
          % (100 * l / tl, 100 * p[0] / tot_i, p[1] / max(p[0], 1), k[0], k[1], p[2]))
