"""Summarise an ncu launch list (csv) into profiles/launch_summary_<tag>.json.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv \
        --log-file gpurun_out/launches_<tag>.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
    python tools/summarize_launches.py gpurun_out/launches_<tag>.csv <tag> [steps_in_capture=3]
(bench.py --steps 1 --warmup 1 runs 3 train steps: warm-up, timed, roofline pass.)  Per-launch times under ncu are
cold-cache and serialised: the SHARE of each kernel is the meaningful number, not the absolute.
"""
import collections
import csv
import json
import os
import sys


def main():
    path, tag = sys.argv[1], sys.argv[2]
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][:80]
        v = float(row["Metric Value"].replace(",", ""))
        unit, m = row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}[unit]
            cnt[name] += 1
        else:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        agg[name][m] += v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    out = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = a["gpu__time_duration.sum"]
        out.append(dict(kernel=k, ms_per_step=t / nsteps / 1e3, launches_per_step=cnt[k] // nsteps, share=t / tot,
                        dram_read_GB_per_step=a["dram__bytes_read.sum"] / nsteps / 1e9,
                        dram_write_GB_per_step=a["dram__bytes_write.sum"] / nsteps / 1e9))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "profiles", f"launch_summary_{tag}.json")
    json.dump(dict(source=__doc__.split("\n\n")[1], total_ms_per_step=tot / nsteps / 1e3, kernels=out), open(dst, "w"), indent=1)
    for o in out[:16]:
        print(f"{o['ms_per_step']:7.2f} ms/step {o['launches_per_step']:4d}/step {100 * o['share']:5.1f}%  "
              f"dram {o['dram_read_GB_per_step']:6.2f}+{o['dram_write_GB_per_step']:6.2f} GB/step  {o['kernel']}")
    print("total", tot / nsteps / 1e3, "ms/step ->", dst)


if __name__ == "__main__":
    main()
