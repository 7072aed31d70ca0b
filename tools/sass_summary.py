#!/usr/bin/env python
"""Blackwell evidence from the built library: per kernel of libdi_b200.so the count of tensor-core / TMA / TMEM SASS
mnemonics (cuobjdump -sass) -> markdown.   python tools/sass_summary.py [lib] > profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'SYNCS', 'HMMA', 'LDGSTS', 'MUFU.EX2']


def main(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    cur, counts, size = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or '/*' not in line:
            continue
        size[cur] += 1
        for mn in MNEMONICS:
            if re.search(r'\b' + re.escape(mn) + r'\b', line) or (mn + '.') in line:
                counts[cur][mn] += 1
    demangle = subprocess.run(['c++filt'] + list(counts), capture_output=True, text=True).stdout.splitlines()
    print('# SASS mnemonics per kernel of %s (cuobjdump -sass, sm_100a)\n' % os.path.relpath(lib, ROOT))
    print('`UTCHMMA` = tcgen05.mma (kind::f16 / tf32), `LDTM` / `STTM` = tcgen05.ld / st (tensor memory), `UTMALDG` / '
          '`UTMASTG` = TMA tensor load / store, `HMMA` = legacy mma.sync, `LDGSTS` = cp.async.\n')
    print('| kernel | instr | ' + ' | '.join(MNEMONICS) + ' |')
    print('|---|---|' + '---|' * len(MNEMONICS))
    for (fn, c), name in zip(counts.items(), demangle):
        if not any(c[m] for m in MNEMONICS[:9]):
            continue
        name = re.sub(r'\(anonymous namespace\)::|\(CUtensorMap_st.*', '', name)[:64]
        print('| `%s` | %d | ' % (name, size[fn]) + ' | '.join(str(c[m]) if c[m] else '' for m in MNEMONICS) + ' |')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'deepinteraction_b200', 'libdi_b200.so'))
