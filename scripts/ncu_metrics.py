"""Key counters of one ncu report (raw page): usage python scripts/ncu_metrics.py report.ncu-rep [kernel index, default 0]
(a report of the split strategy pipeline holds two kernels: 0 = the pair producer, 1 = the rule kernel)"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h, u = rows[0], rows[1]
if 2 + k >= len(rows):
    sys.exit(0)
v = rows[2 + k]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg.per_second', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']
ki = h.index("Kernel Name") if "Kernel Name" in h else None
if ki is not None:
    print(f'kernel,,"{v[ki][:90]}"')
for i, n in enumerate(h):
    if n in want or ('issue_stalled' in n and 'per_issue_active' in n):
        try:
            if float(v[i].replace(',', '')) > 0.05: print(f'{n},{u[i]},{v[i]}')
        except ValueError:
            pass
