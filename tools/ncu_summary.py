"""Per-kernel summary of an ncu report: python tools/ncu_summary.py file.ncu-rep  (reads it with `ncu -i ... --page raw --csv`)"""
import csv, subprocess, sys
KEYS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct_active'),
        ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'hmma_pct'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'), ('lts__t_bytes.sum', 'l2_bytes'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_throughput_pct'),
        ('l1tex__m_xbar2l1tex_read_bytes.sum', 'l2_to_sm_read_bytes'), ('lts__t_sectors_op_read.sum', 'l2_read_sectors'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'mufu_pipe_pct'),
        ('sm__inst_executed.avg.per_cycle_elapsed', 'ipc'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'),
        ('launch__shared_mem_per_block_dynamic', 'dyn_smem')]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d.get('Kernel Name', '?')[:110]
    print(name)
    for k, short in KEYS:
        if k in d and d[k] != '':
            print(f'    {short:24s} {d[k]:>14s} {u.get(k, "")}')
