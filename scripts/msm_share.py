#!/usr/bin/env python3
"""From an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: isolate the device-resident 2^24 MSM
steps (one msm_coarse_hist ... msm_final sequence whose accumulate launch exceeds 25 ms) and print the mean per-kernel
time and share of such a step -- the figure to hold against bench.py's live `kernel_share_of_step`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
    seq.append((name, v))
inst, cur = [], None
for name, v in seq:
    if name == "msm_coarse_hist_kernel":
        cur = []
        inst.append(cur)
    if cur is not None and (name.startswith("msm_") or name.startswith("scan_") or name.startswith("task_")):
        cur.append((name, v))
        if name == "msm_final_kernel":
            cur = None
full = [i for i in inst if i and i[-1][0] == "msm_final_kernel" and any(n == "msm_accumulate_kernel" and v > 25.0 for n, v in i)
        and sum(1 for n, _ in i if n == "msm_accumulate_kernel") == 1]
agg = collections.OrderedDict()
for i in full:
    for n, v in i:
        agg[n] = agg.get(n, 0.0) + v
tot = sum(agg.values())
print(f"{len(full)} device-resident 2^24 MSM steps in the list; mean step (sum of kernel times, serialised, cold cache) {tot / len(full):.3f} ms")
groups = {"accumulate": ("msm_accumulate_kernel",), "reduce": ("msm_merge_kernel", "msm_clear_empty_kernel", "msm_reduce_kernel", "msm_window_kernel", "msm_final_kernel")}
for n, v in agg.items():
    print(f"  {n:28s} {v / len(full):8.3f} ms  {100 * v / tot:5.1f} %")
acc = sum(agg[k] for k in groups["accumulate"])
red = sum(agg.get(k, 0.0) for k in groups["reduce"])
print(f"shares: accumulate {100 * acc / tot:.1f} %, sort+tasks {100 * (tot - acc - red) / tot:.1f} %, reduce {100 * red / tot:.1f} %")
