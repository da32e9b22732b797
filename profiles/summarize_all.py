"""Per-kernel totals of an ncu launch list (any workload): python profiles/summarize_all.py <csv> [first_marker_kernel]
Groups the launches of the LAST complete repetition (between the last two launches of `first_marker_kernel`, default: the
kernel that is launched first in the file) by kernel name: count, time, DRAM bytes, GB/s."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_launches import load   # noqa: E402


def main(path, marker=None):
    L = load(path)
    marker = marker or L[0]['name']
    idx = [i for i, d in enumerate(L) if d['name'].startswith(marker)]
    step = L[idx[-2]:idx[-1]] if len(idx) >= 2 else L
    T = 'gpu__time_duration.sum'
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(d['name'], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d[T]
        a[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    tot = sum(a[1] for a in agg.values())
    print(f'{len(step)} launches, {tot:.1f} us under ncu (cold cache, serialised); marker {marker}\n')
    print(f"{'kernel':64s}{'n':>4s}{'us':>10s}{'share':>7s}{'DRAM MB':>10s}{'GB/s':>8s}")
    for name, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{name[:63]:64s}{n:4d}{t:10.1f}{t / tot:7.2f}{b / 1e6:10.1f}{(b / t / 1e3 if t else 0):8.0f}')


if __name__ == '__main__':
    main(*sys.argv[1:])
