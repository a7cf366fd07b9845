#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few counters the roofline
argument needs. Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [...] > profiles/xyz.md"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print("## %s — %s" % (path.split("/")[-1], d.get("Kernel Name", "?")[:90]))
            print("grid %s block %s" % (d.get("Grid Size"), d.get("Block Size")))
            print()
            print("| metric | value | unit |")
            print("|---|---|---|")
            for k in KEYS:
                if k in d:
                    print("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
            try:
                rd = float(d["dram__bytes_read.sum"].replace(",", ""))
                wr = float(d["dram__bytes_write.sum"].replace(",", ""))
                t = float(d["gpu__time_duration.sum"].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
                rd *= scale.get(u["dram__bytes_read.sum"], 1)
                wr *= scale.get(u["dram__bytes_write.sum"], 1)
                t *= {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u["gpu__time_duration.sum"], 1)
                print("| **dram traffic (read+write)** | %.4g | byte |" % (rd + wr))
                print("| **dram GB/s under ncu** | %.1f | GB/s |" % ((rd + wr) / t / 1e9))
            except Exception as e:
                print("| (traffic parse failed: %r) | | |" % (e,))
            print()


if __name__ == "__main__":
    main()
