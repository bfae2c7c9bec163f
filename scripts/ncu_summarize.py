"""Summaries of the ncu outputs of scripts/gpu_round.sh for profiles/ (run here, where the .ncu-rep / .csv files land).

    python scripts/ncu_summarize.py launches gpurun_out/<tag>/ncu_launches_bench.csv "<command>" > profiles/..._summary.txt
    python scripts/ncu_summarize.py full gpurun_out/<tag>/ncu_full_stream_r10.ncu-rep [more .ncu-rep] > profiles/..._summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"]


def launches(path, command):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = OrderedDict()
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        per.setdefault(r["Kernel Name"], []).append(ns / 1e3)
    total = sum(sum(v) for v in per.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none: %s" % command)
    print("# per-launch device time (us), cold-cache and serialised under ncu: compare shares, not absolutes (%d launches)"
          % sum(len(v) for v in per.values()))
    print("%-68s %5s %10s %10s %7s" % ("kernel", "n", "avg_us", "total_us", "share"))
    for k, v in per.items():
        print("%-68s %5d %10.1f %10.1f %6.1f%%" % (k[:66], len(v), sum(v) / len(v), sum(v), 100.0 * sum(v) / total))


def full(paths):
    print("# ncu --set full --clock-control none --import-source on; first launch of each kernel shown")
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(out)))
        header, units = rd[0], rd[1]
        seen = set()
        for row in rd[2:]:
            name = row[header.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            print("## %s  %s" % (path.split("/")[-1], name))
            for m in KEEP:
                if m in header:
                    i = header.index(m)
                    print("   %-72s %s %s" % (m, row[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[2:])
