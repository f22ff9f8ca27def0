"""Key metrics per kernel from an `ncu --page raw --csv` export of a --set full capture."""
import csv, sys, re
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__inst_executed.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, units = rows[hi], rows[hi + 1]
col = {h: i for i, h in enumerate(hdr)}
pat = sys.argv[2] if len(sys.argv) > 2 else None
stall = [h for h in hdr if h.startswith('smsp__average_warp') and 'per_issue_active' in h and h.endswith('.ratio') and 'not_issued' not in h]
if not stall:
    stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
for r in rows[hi + 2:]:
    if len(r) < len(hdr): continue
    name = r[col['Kernel Name']]
    if pat and not re.search(pat, name): continue
    print('==', name[:100], 'id', r[col['ID']])
    for k in KEYS:
        if k in col: print(f'   {k:75s} {r[col[k]]:>16s} {units[col[k]]}')
    st = []
    for h in stall:
        try: st.append((float(r[col[h]].replace(',', '')), h))
        except ValueError: pass
    st.sort(reverse=True)
    for v, h in st[:7]:
        print(f'   stall {h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")[:60]:60s} {v:10.2f}')
