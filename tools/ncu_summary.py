"""Summarise an ncu --page raw --csv dump: key metrics + warp stall reasons per captured launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum',
        'smsp__inst_executed.sum']
for w in want:
    if w in idx:
        print(f"{w[:66]:66s} [{units[idx[w]][:8]:8s}]", [d[idx[w]][:14] for d in data])
st = [h for h in hdr if 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
print("--- stall reasons (warp cycles per issued instruction) ---")
for h in sorted(st):
    vals = [d[idx[h]] for d in data]
    try:
        if max(float(v) for v in vals) < 0.3:
            continue
    except ValueError:
        pass
    print(f"{h.split('issue_stalled_')[1][:40]:40s}", [v[:5] for v in vals])
