"""Summaries for profiles/: (1) launch list csv from `ncu --metrics gpu__time_duration.sum` -> per-kernel count / time /
share; (2) selected raw metrics of a `--set full` report.
    python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rN_launches.txt
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep > profiles/rN_<kernel>_full.txt"""
import collections, csv, io, re, subprocess, sys

mode, path = sys.argv[1], sys.argv[2]
if mode == 'launches':
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 10]
    hdr = rows[0]
    iK, iV = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r'\(.*', '', r[iK]).replace('void ', '')
        ns = float(r[iV].replace(',', ''))
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + ns)
    tot = sum(t for _, t in agg.values())
    print("# %d launches, %.3f ms of kernel time (ncu replay: cold caches, serialised -- compare shares, not absolutes)" % (len(rows) - 1, tot / 1e6))
    print("%-44s %8s %12s %8s" % ("kernel", "launches", "total_us", "share"))
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %12.1f %7.1f%%" % (name, c, t / 1e3, 100 * t / tot))
else:
    # 'full': an .ncu-rep report; 'csv': the `ncu -i report --page raw --csv` export made on the GPU box (reports above a few
    # tens of MB do not travel back)
    txt = open(path, errors='ignore').read() if mode == 'csv' else subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
            'launch__occupancy_limit_shared_mem', 'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
    for v in vals:
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                print("%-62s %-14s %s" % (w, units[i], v[i]))
        st = [(float(v[i]), h.split('issue_stalled_')[1].split('_per_issue')[0]) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and v[i] not in ('', 'n/a')]
        if st:
            print("%-62s %-14s %s" % ('stalls (warps per issue, top 6)', '', ', '.join('%s %.2f' % (n, x) for x, n in sorted(st, reverse=True)[:6])))
        print()
