"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    seq = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = row['Kernel Name'].split('(')[0].replace('stb::<unnamed>::', '').replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
        seq.append((name, v))
    return seq


if __name__ == '__main__':
    seq = load(sys.argv[1])
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for n, v in seq:
        tot[n] = tot.get(n, 0) + v
        cnt[n] += 1
    T = sum(tot.values())
    print(f'total {T:.1f} us over {len(seq)} launches')
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f'{v:10.1f} us {100 * v / T:5.1f}% n={cnt[k]:4d}  {k[:100]}')
    if len(sys.argv) > 2:
        for n, v in seq:
            if sys.argv[2] in n:
                print(f'{v:9.1f}', n[:90])
