"""Condenses gpurun_out/raw_<tag>.csv (ncu --set full, one chunk of frames = 21 launches) and
gpurun_out/launches_<tag>.csv into tracked files under profiles/: a per-launch summary CSV and the
per-launch DRAM traffic table bench.py quotes in its `roofline.traffic` field."""
import csv, json, sys
tag, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open('gpurun_out/raw_%s.csv' % tag)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
keep = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_selected']
keep = [k for k in keep if k in idx]
w = csv.writer(open('profiles/%s_ncu_full_summary.csv' % out, 'w'))
w.writerow(['launch'] + keep)
w.writerow([''] + [units[idx[k]] for k in keep])
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
traffic = []
# the capture window may start mid-chunk: rotate so that launch 0 is the chunk's first kernel (k_letterbox), as bench.py numbers them
first = [i for i, r in enumerate(data) if 'k_letterbox' in r[idx['Kernel Name']]]
if first:
    data = data[first[0]:] + data[:first[0]]
for i, r in enumerate(data):
    w.writerow([i] + [r[idx[k]] for k in keep])
    rd = float(r[idx['dram__bytes_read.sum']].replace(',', '')) * scale[units[idx['dram__bytes_read.sum']]]
    wr = float(r[idx['dram__bytes_write.sum']].replace(',', '')) * scale[units[idx['dram__bytes_write.sum']]]
    traffic.append({'launch': i, 'kernel': r[idx['Kernel Name']].split('(')[0].replace('void fdt::<unnamed>::', '').replace('fdt::<unnamed>::', '').replace('void unnamed>::', '').replace('unnamed>::', ''),
                    'dram_bytes_per_launch': rd + wr, 'images_per_launch': int(sys.argv[3]) if len(sys.argv) > 3 else 512})
json.dump({'source': 'ncu --set full --clock-control none, tools/prof_target.py (one chunk of frames), capture %s' % tag,
           'launches': traffic}, open('profiles/%s_traffic.json' % out, 'w'), indent=1)
import shutil
shutil.copy('gpurun_out/launches_%s.csv' % tag, 'profiles/%s_launches.csv' % out)
print(open('profiles/%s_traffic.json' % out).read()[:600])
