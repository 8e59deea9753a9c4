"""Summarise ncu outputs (run here, no GPU needed).
    python tools/ncu_summary.py launches gpurun_out/launches_X.csv [skip_first_n]
    python tools/ncu_summary.py rep gpurun_out/prof_X.ncu-rep
"""
import collections, csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def launches(path, skip=0):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        n += 1
        if n <= skip:
            continue
        k = row['Kernel Name'][:90]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'{path}: {n - skip} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)')
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{a[1]:10.1f} us {a[0]:5d} x {a[1] / a[0]:9.1f} us {100 * a[1] / tot:5.1f}%  {k}')


def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('==', r[ki][:100])
        for i, h in enumerate(hdr):
            if any(h == k or (k in h and k.endswith('.sum') is False and h.startswith(k)) for k in KEYS):
                print(f'   {h:78s} {units[i]:14s} {r[i]}')


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        rep(sys.argv[2])
