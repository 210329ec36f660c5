"""Condenses an `ncu --page raw --csv` export into one row per kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'blk'), ('launch__registers_per_thread', 'regs'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankcf'),
        ('smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'longsb'), ('smsp__pcsamp_warps_issue_stalled_barrier', 'bar'),
        ('smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'shortsb'), ('smsp__pcsamp_warps_issue_stalled_mio_throttle', 'mio'),
        ('smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'math'), ('smsp__pcsamp_warps_issue_stalled_wait', 'wait'),
        ('smsp__pcsamp_warps_issue_stalled_selected', 'sel'), ('smsp__pcsamp_warps_issue_stalled_not_selected', 'notsel'),
        ('smsp__pcsamp_warps_issue_stalled_lg_throttle', 'lg'), ('smsp__pcsamp_warps_issue_stalled_no_instructions', 'noinst'),
        ('smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'disp')]
cols = [c for c in cols if c[0] in idx]
print(" ".join("%7s" % c[1] for c in cols), "kernel")
for r in data:
    out = []
    for c, n in cols:
        v = r[idx[c]].replace(',', '')
        try:
            f = float(v)
            if n == 'us':
                f *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3}.get(units[idx[c]], 1)
            if n in ('rdMB', 'wrMB'):
                f *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(units[idx[c]], 1)
            v = "%.4g" % f
        except ValueError:
            pass
        out.append("%7s" % v[:7])
    print(" ".join(out), r[idx['Kernel Name']].replace('void unnamed>::', '').replace('unnamed>::', '')[:24])
