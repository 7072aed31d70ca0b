#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into profiles/<name>.md + compact csv."""
import collections
import csv
import re
import sys


def main(src, dst_prefix, passes):
    lines = [l for l in open(src) if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'<unnamed>::|void |\(anonymous namespace\)::', '', name)
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        rows.append((int(row['ID']), name[:70], row['Grid Size'], row['Block Size'], v))
    with open(dst_prefix + '.csv', 'w') as f:
        f.write('id,kernel,grid,block,time_us\n')
        for r in rows:
            f.write('%d,"%s","%s","%s",%.2f\n' % r)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, name, _, _, v in rows:
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst_prefix + '.md', 'w') as f:
        f.write(f'# ncu launch list summary: {src}\n\n')
        f.write(f'{len(rows)} launches over {passes} forward passes (cold-cache, serialised by ncu: compare SHARES, '
                f'not absolutes); total {tot / 1e3:.2f} ms = {tot / 1e3 / passes:.2f} ms per forward.\n\n')
        f.write('| kernel | launches/fwd | us/fwd | avg us | share |\n|---|---|---|---|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %.1f | %.1f | %.1f | %.1f%% |\n' % (k, v[0] / passes, v[1] / passes, v[1] / v[0], 100 * v[1] / tot))
    print(open(dst_prefix + '.md').read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
