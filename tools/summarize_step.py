"""Summarise ONE complete train step out of an ncu launch list (csv or csv.gz): the launches between two consecutive
prep_input kernels (the first launch of every step).

    python tools/summarize_step.py gpurun_out/launches_<tag>.csv.gz <tag> [which_step=1]   -> profiles/launch_summary_<tag>.json
"""
import collections
import csv
import gzip
import json
import os
import sys


def main():
    path, tag = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.DictReader(l for l in op(path, "rt") if not l.startswith("==")))
    starts = [int(r["ID"]) for r in rows if r["Metric Name"] == "gpu__time_duration.sum" and "prep_input" in r["Kernel Name"]]
    last = int(rows[-1]["ID"]) + 1
    lo, hi = starts[which], (starts[which + 1] if which + 1 < len(starts) else last)
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows:
        if not lo <= int(r["ID"]) < hi:
            continue
        name = r["Kernel Name"].split("(")[0][:80]
        v, u, m = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[u]
            cnt[name] += 1
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        agg[name][m] += v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    out = [dict(kernel=k, ms_per_step=a["gpu__time_duration.sum"] / 1e3, launches_per_step=cnt[k],
                share=a["gpu__time_duration.sum"] / tot, dram_read_GB_per_step=a["dram__bytes_read.sum"] / 1e9,
                dram_write_GB_per_step=a["dram__bytes_write.sum"] / 1e9)
           for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "profiles", f"launch_summary_{tag}.json")
    json.dump(dict(source="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                          "--csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e; ONE complete train step (launches "
                          f"between prep_input #{which} and #{which + 1}). Per-launch times under ncu are cold-cache and serialised: "
                          "the SHARE of each kernel is the meaningful number, not the absolute.",
                   total_ms_per_step=tot / 1e3, launches_per_step=sum(cnt.values()),
                   dram_GB_per_step=sum(o["dram_read_GB_per_step"] + o["dram_write_GB_per_step"] for o in out), kernels=out),
              open(dst, "w"), indent=1)
    for o in out[:18]:
        print("%7.2f ms %4d  %5.1f%%  dram %6.2f+%6.2f GB  %s" % (o["ms_per_step"], o["launches_per_step"], 100 * o["share"],
              o["dram_read_GB_per_step"], o["dram_write_GB_per_step"], o["kernel"]))
    print("total %.2f ms, %d launches, %.1f GB DRAM -> %s" % (tot / 1e3, sum(cnt.values()),
          sum(o["dram_read_GB_per_step"] + o["dram_write_GB_per_step"] for o in out), dst))


if __name__ == "__main__":
    main()
