"""Summarise an .ncu-rep (read on the CPU box) into the handful of counters DESIGN.md / profiles/ cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "note" > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
    'launch__block_size', 'launch__grid_size', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__thread_inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
]


def main():
    rep = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    print(f"# {rep}: {note}")
    for v in rows[2:]:
        print(f"## kernel {v[h.index('Kernel Name')]}")
        for i, n in enumerate(h):
            if n in KEEP or n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio'):
                print(f"{n} = {v[i]} {u[i]}")


if __name__ == "__main__":
    main()
