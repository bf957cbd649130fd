"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv).

    python tools/launch_summary.py profiles/r2k_launches.csv
"""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, mi, vi, idi = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
    per_launch = collections.defaultdict(dict)
    for r in rows[1:]:
        per_launch[(r[idi], r[ki].split('(')[0])][r[mi]] = float(r[vi].replace(',', ''))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (_, k), m in per_launch.items():
        a = agg[k]
        a[0] += 1
        a[1] += m['gpu__time_duration.sum'] / 1e3
        a[2] += m.get('smsp__inst_executed.sum', 0.0)
    total = sum(a[1] for a in agg.values())
    out = []
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append('%-46s %3d x %7.1f us  %5.1f %%  %6.1f M warp inst' % (k[-46:], a[0], a[1] / a[0], 100 * a[1] / total,
                                                                        a[2] / a[0] / 1e6))
    out.append('total %.1f us over %d launches' % (total, sum(a[0] for a in agg.values())))
    return '\n'.join(out)


if __name__ == '__main__':
    print(summarize(sys.argv[1]))
