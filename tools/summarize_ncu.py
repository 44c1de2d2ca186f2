"""Condense `ncu --page raw --csv` output (one row per profiled launch) into a small JSON of the metrics the roofline
discussion uses: duration, DRAM bytes / throughput, tensor-pipe activity, achieved occupancy, registers.

    python tools/summarize_ncu.py <raw.csv> <out.json> [hbm_peak_GBs]
"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed": "smem_tc_wavefront_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3,
         "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    path, out = sys.argv[1], sys.argv[2]
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6584.8
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0][:70]}
        for m, name in WANT.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                d[name] = round(v * SCALE.get(units[idx[m]], 1.0), 3)
        if "duration_us" in d and "dram_read_MB" in d:
            gbs = (d["dram_read_MB"] + d.get("dram_write_MB", 0.0)) * 1e-3 / (d["duration_us"] * 1e-6)
            d["dram_GBs"] = round(gbs, 1)
            d["dram_frac_of_measured_copy_peak"] = round(gbs / peak, 3)
        res.append(d)
    json.dump({"source": "ncu --set full --clock-control none (one launch per kernel, tools/profile_kernels.py); "
                         "dram_frac = (dram read + write bytes / duration) / measured copy bandwidth %.1f GB/s" % peak,
               "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print("%-60s %9.1f us  dram %7.1f GB/s (%4.0f%%)  tensor %5s%%" % (d["kernel"], d.get("duration_us", 0), d.get("dram_GBs", 0),
              100 * d.get("dram_frac_of_measured_copy_peak", 0), d.get("tensor_pipe_pct", "-")))


if __name__ == "__main__":
    main()
