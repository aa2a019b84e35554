#!/usr/bin/env python
"""Text summary of .ncu-rep captures for profiles/ (runs here, reads reports brought back in gpurun_out/).

    python tools/ncu_summarize.py gpurun_out/prof_*_r01b.ncu-rep > profiles/r01_ncu_family_kernels.txt
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def main():
    for f in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print(f"== {f}: no kernel captured\n")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            print(f"== {f}\n   kernel: {name}   (ncu --set full --clock-control none, one launch; absolute times are cold-cache/serialised)")
            for w in WANT:
                if w in hdr:
                    print(f"   {w:74s} {vals[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
            stalls = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                    try:
                        v = float(vals[i])
                    except ValueError:
                        continue
                    if v >= 0.3:
                        stalls.append((v, h.split("stalled_")[1].split("_per")[0]))
            print("   warp stall reasons (cycles per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))
            print()


if __name__ == "__main__":
    main()
