"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('iod::', '')
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('| kernel | launches | total us | share | avg us |')
    print('|---|---|---|---|---|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.1f | %.1f%% | %.1f |' % (k, a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
    print('\ntotal device time in the capture: %.1f us' % tot)


if __name__ == '__main__':
    main(sys.argv[1])
