"""Condense an .ncu-rep (via `ncu -i ... --page raw --csv`) into the handful of numbers the roofline argument uses."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_branch.sum", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "?")[:120])
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(v), k[len(STALLS):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                         if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        print("  stall reasons (warps per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
