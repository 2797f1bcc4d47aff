"""Short per-kernel digest of an `ncu --set full` report: time, DRAM / L2 / SM throughput, occupancy, issue rate, stall mix.
    python tools/ncu_brief.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
W = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
     'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
     'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
     'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
     'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
     'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:100])
    for w in W:
        if w in hdr:
            i = hdr.index(w)
            print('  %-66s %-10s %s' % (w, units[i], r[i]))
    st = [(float(r[i]), h.split('issue_stalled_')[1].split('_per_issue')[0]) for i, h in enumerate(hdr)
          if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
    tot = sum(v for v, _ in st)
    print('  stalls (warps per issue): ' + ', '.join('%s %.2f' % (n, v) for v, n in sorted(st, reverse=True)[:7]) + '  | total %.2f' % tot)
