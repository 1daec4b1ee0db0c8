#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep OUT.md -- condense an `ncu --set full` report into the few
per-launch numbers DESIGN.md / bench.py cite (duration, DRAM bytes, hit rates, pipe use, stalls)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not selected"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "samples: selected (issuing)"),
    ("smsp__pcsamp_warps_issue_stalled_membar", "stall samples: membar"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
    ("smsp__pcsamp_sample_count", "samples total"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary of %s\n\n" % rep.split("/")[-1])
        f.write("| metric | unit | " + " | ".join("launch %d" % k for k in range(len(data))) + " |\n")
        f.write("|---|---|" + "---|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join(d[idx["Kernel Name"]].replace("|", "/") for d in data) + " |\n")
        f.write("| grid | | " + " | ".join(d[idx["Grid Size"]] for d in data) + " |\n")
        for k, name in KEYS:
            if k in idx:
                f.write("| %s (`%s`) | %s | %s |\n" % (name, k, units[idx[k]], " | ".join(d[idx[k]] for d in data)))
    print("wrote", out)


if __name__ == "__main__":
    main()
