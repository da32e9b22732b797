"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel time (and DRAM traffic) of one train step, the GEMM launches in step order, and -- with --json -- the
per-step DRAM traffic of the GEMM kernel class that bench.py reports as roofline.traffic."""
import collections
import csv
import json
import re
import sys

GEMM = ('tc_gemm_kernel', 'tc_conv_reuse_kernel', 'gemm_simt_kernel')


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    H = rows[0]
    ki, mi, vi, ui, gi, ii = (H.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'ID'))
    recs = collections.OrderedDict()
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    for r in rows[1:]:
        d = recs.setdefault(r[ii], {'name': short(r[ki]), 'grid': r[gi]})
        d[r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)
    return list(recs.values())


def short(n):
    n = re.sub(r'^void ', '', n).replace('emb::', '').replace('<unnamed>::', '')
    return n.split('(')[0]


def main(path, json_out=None):
    L = load(path)
    ends = [i for i, d in enumerate(L) if d['name'].startswith('opt_step')]
    s, e = ends[-2] + 1, ends[-1]
    step = L[s:e + 1]
    T = 'gpu__time_duration.sum'
    has_dram = 'dram__bytes_read.sum' in step[0]
    agg, tot = collections.OrderedDict(), 0.0
    for d in step:
        a = agg.setdefault(d['name'], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d[T]
        a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
        tot += d[T]
    print(f'one train step = launches {s}..{e} of the capture ({e - s + 1} kernels, {tot / 1e3:.2f} ms summed, cold-cache serialised)')
    for n, (c, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        extra = f'  {b / 1e6:9.1f} MB DRAM  {b / max(t, 1e-9) / 1e6:6.2f} TB/s' if has_dram else ''
        print(f'{t:10.1f} us  x{c:3d}  {100 * t / tot:5.1f}%  {n}{extra}')
    print('\nGEMM-class launches in step order (grid, us' + (', DRAM MB' if has_dram else '') + '):')
    g_t = g_b = g_n = 0
    for d in step:
        if d['name'] in GEMM:
            b = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
            g_t += d[T]; g_b += b; g_n += 1
            print(f"  {d['name']:22s} {d['grid']:>14s} {d[T]:9.1f}" + (f' {b / 1e6:9.1f}' if has_dram else ''))
    print(f'GEMM class: {g_n} launches, {g_t:.1f} us, {g_b / 1e6:.1f} MB DRAM traffic per step')
    if json_out:
        json.dump({'source': path, 'gemm_launches_per_step': g_n, 'gemm_us_per_step_under_ncu': g_t, 'gemm_dram_bytes_per_step': g_b,
                   'gemm_dram_bytes_per_launch': g_b / max(g_n, 1), 'step_us_under_ncu': tot}, open(json_out, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
