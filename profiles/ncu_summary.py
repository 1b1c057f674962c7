#!/usr/bin/env python3
"""Summarise .ncu-rep captures (ncu --set full) into a markdown table: one row per captured launch.
usage: python profiles/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/r02/ncu_summary.md   (runs where ncu is installed; no GPU needed)"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__inst_executed.sum", "warp instr"), ("sass__inst_executed_local_stores", "local stores")]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle", "not_selected", "branch_resolving", "no_instruction"]


def main():
    print("| capture | kernel | " + " | ".join(n for _, n in WANT) + " | top stalls (warps per issue) |")
    print("|" + "---|" * (len(WANT) + 3))
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print(f"| {path} | (no launches) |"); continue
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
            cells = []
            for key, _ in WANT:
                i = ix.get(key)
                cells.append("-" if i is None else f"{r[i]} {units[i]}".strip())
            st = []
            for s in STALLS:
                i = ix.get(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio")
                if i is not None:
                    try:
                        st.append((float(r[i]), s))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print(f"| {path.split('/')[-1]} | {name} | " + " | ".join(cells) + " | " + ", ".join(f"{s} {v:.2f}" for v, s in st[:4]) + " |")


if __name__ == "__main__":
    main()
