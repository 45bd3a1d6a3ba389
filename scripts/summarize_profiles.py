#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launch lists into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> <tiles.ncu-rep> <launches.csv> [regex.ncu-rep]"""
import csv
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def summarize(rep, dst):
    hdr, units, rows = raw(rep)
    keys = [k for k in KEYS if k in hdr]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(keys)
        w.writerow([units[hdr.index(k)] for k in keys])
        for r in rows:
            w.writerow([r[hdr.index(k)] for k in keys])
    return hdr, units, rows


def main():
    tag, tiles_rep, launches = sys.argv[1:4]
    hdr, units, rows = summarize(tiles_rep, f"profiles/{tag}_ncu_full_k_fixed_tiles.csv")
    rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    total = sum(to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]) for r in rows)
    json.dump({"dram_bytes_per_step": total, "launches": len(rows),
               "source": f"profiles/{tag}_ncu_full_k_fixed_tiles.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                         f"{len(rows)} k_fixed_tiles launches of one step (ncu --set full, 100 M rows)"},
              open("profiles/traffic.json", "w"), indent=1)
    lines = [l for l in open(launches) if not l.startswith("==")]
    open(f"profiles/{tag}_launches_bench.csv", "w").writelines(lines)
    if len(sys.argv) > 4:
        summarize(sys.argv[4], f"profiles/{tag}_ncu_full_k_regex_tiles.csv")
    print("dram bytes per step", total)


if __name__ == "__main__":
    main()
