"""Pipe / issue utilisation summary of `ncu --set full` captures (one launch each): python tools/ncu_pipes.py a.ncu-rep b.ncu-rep ..."""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers / thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64 per SM)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe wavefronts %"),
        ("l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "L1 hit rate, global loads %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("smsp__inst_executed.sum", "warp instructions")]
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u, v = rows[0], rows[1], rows[-1]
    name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
    print(f"== {path}: {name}")
    for key, label in KEYS:
        if key in h:
            i = h.index(key)
            print(f"   {label:34s} {v[i]:>16s} {u[i]}")
