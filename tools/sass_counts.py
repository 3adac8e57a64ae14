"""Counts the Blackwell-specific SASS mnemonics per kernel of the built library
(cuobjdump -sass): UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st),
UTMALDG / UTMASTG (TMA), UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async).

    python tools/sass_counts.py > profiles/r02_sass.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'preworld_b200', 'lib', 'libpreworld_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'UTCATOMSWS',
        'SYNCS', 'LDGSTS', 'FFMA2', 'HMMA', 'FFMA']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace('(anonymous namespace)::', '').replace('void ', '')
            name = re.sub(r'\(.*', '', name)
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m:
            op = m.group(1)
            cur['total'] += 1
            for k in KEYS:
                if op == k or op.startswith(k + '.'):
                    cur[k] += 1
    print('# SASS mnemonics per kernel of `preworld_b200/lib/libpreworld_b200.so` (sm_100a)\n')
    print('`python tools/sass_counts.py` (cuobjdump -sass).  UTCHMMA = tcgen05.mma, LDTM / STTM = '
          'tcgen05.ld / .st,\nUTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier '
          'ops, FFMA2 = packed fp32x2.\n')
    cols = [k for k in KEYS if any(c[k] for c in counts.values())]
    print('| kernel | instr | ' + ' | '.join(cols) + ' |')
    print('|---|---:|' + '---:|' * len(cols))
    for name, c in counts.items():
        if not any(c[k] for k in ('UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG')):
            continue
        print(f'| `{name}` | {c["total"]} | ' + ' | '.join(str(c[k]) for k in cols) + ' |')
    rest = [n for n, c in counts.items()
            if not any(c[k] for k in ('UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG'))]
    print(f'\n{len(rest)} other kernels (SIMT: gather / streaming / reduction kernels) carry none of '
          'the tensor-core / TMA mnemonics.')


if __name__ == '__main__':
    main()
