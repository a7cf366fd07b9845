#!/usr/bin/env python
"""Per-opcode executed-instruction histogram, stall-reason totals and the hottest SASS lines of one
.ncu-rep (source page; needs -lineinfo builds). usage: tools/ncu_stalls.py report.ncu-rep [top=20]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, v = rows[0], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    print("kernel:", v[h.index("Kernel Name")][:100])
    for k in want:
        if k in h:
            print("  %-70s %s %s" % (k, v[h.index(k)], rows[1][h.index(k)]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h, data = rows[1], rows[2:]
    iS, iE, iN, iT = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    op, thr, stalls, tot = Counter(), Counter(), Counter(), 0
    full = Counter()   # full mnemonic (with modifiers) of the integer / move instructions
    for r in data:
        try:
            e = int(r[iE])
        except ValueError:
            continue
        s = r[iS].split()
        o = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        op[o] += e
        thr[o] += int(r[iT] or 0)
        if o in ("IMAD", "ISETP", "LEA", "VIADD", "IADD3", "MOV", "LOP3", "PLOP3", "SEL", "FSEL", "LDC", "UISETP"):
            full[s[1] if s[0].startswith("@") else s[0]] += e
        tot += e
        for i in stall_cols:
            stalls[h[i]] += int(r[i] or 0)
    print("warp instructions executed: %d" % tot)
    for k, n in op.most_common(18):
        print("  %-10s %6.2f %%   avg active threads %.1f" % (k, 100.0 * n / tot, thr[k] / max(n, 1)))
    print("integer / move instructions by full mnemonic:")
    for k, n in full.most_common(14):
        print("  %-22s %6.2f %%" % (k, 100.0 * n / tot))
    ssum = sum(stalls.values())
    print("stall samples:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * n / ssum) for k, n in stalls.most_common(8)))
    print("hottest lines (samples, executed, sass):")
    for r in sorted(data, key=lambda r: -int(r[iN] or 0))[:top]:
        print("  %6s %10s  %s" % (r[iN], r[iE], r[iS]))


if __name__ == "__main__":
    main()
