"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of one train step."""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    H = rows[0]
    ki, vi, gi = H.index('Kernel Name'), H.index('Metric Value'), H.index('Grid Size')
    data = rows[1:]

    def short(n):
        n = re.sub(r'^void ', '', n)
        n = n.replace('emb::', '').replace('<unnamed>::', '')
        return n.split('(')[0]
    names = [short(r[ki]) for r in data]
    ends = [i for i, n in enumerate(names) if n.startswith('opt_step')]
    s, e = ends[0] + 1, ends[1]
    agg, tot = collections.OrderedDict(), 0.0
    for r, n in zip(data[s:e + 1], names[s:e + 1]):
        t = float(r[vi].replace(',', '')) / 1e3
        agg.setdefault(n, [0, 0.0])
        agg[n][0] += 1
        agg[n][1] += t
        tot += t
    print(f'one train step = launches {s}..{e} of the capture ({e - s + 1} kernels, {tot / 1e3:.2f} ms summed, cold-cache serialised)')
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{t:10.1f} us  x{c:3d}  {100 * t / tot:5.1f}%  {n}')
    print('\nGEMM launches in step order (grid, us):')
    for r, n in zip(data[s:e + 1], names[s:e + 1]):
        if 'gemm' in n:
            print(f'  {n:20s} {r[gi]:>18s} {float(r[vi].replace(",", "")) / 1e3:9.1f}')


if __name__ == '__main__':
    main(sys.argv[1])
