"""Extracts the metrics DESIGN.md / profiles/ quote from an .ncu-rep (run where ncu is installed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tma.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none summary of {rep}")
    for r in data:
        print(f"\n## {r[name_i]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:75s} {r[i]:>20s} {units[i]}")
        stalls = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio")]
        stalls = sorted(((h, float(v)) for h, v in stalls if v not in ("", "n/a")), key=lambda x: -x[1])[:8]
        for h, v in stalls:
            print(f"{h:95s} {v:8.3f}")


if __name__ == "__main__":
    main()
