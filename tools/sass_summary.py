#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA), HMMA (legacy mma.sync), from `cuobjdump -sass` of the shipped library.

    python tools/sass_summary.py [path/to/libmsmd_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'ubisoft-laforge-msmd_b200', 'libmsmd_b200.so')
MNEMONICS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDGSTS', 'MUFU']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            if op in MNEMONICS:
                counts[cur][op] += 1
    names = demangle(list(counts))
    print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)   tools/sass_summary.py')
    print(f'# {"kernel":100s} ' + ' '.join(f'{m:>8s}' for m in MNEMONICS))
    rows = []
    for k, c in counts.items():
        short = re.sub(r'\s+', ' ', names.get(k, k))
        short = re.sub(r'^void ', '', short)
        short = re.sub(r'\((?:[^()]|\([^()]*\))*\)$', '', short)[:100]
        rows.append((short, c))
    for short, c in sorted(rows):
        print(f'  {short:100s} ' + ' '.join(f'{c.get(m, 0):8d}' for m in MNEMONICS))
    tc = sum(1 for _, c in rows if c.get('UTCHMMA', 0) or c.get('UTCQMMA', 0))
    legacy = [s for s, c in rows if c.get('HMMA', 0)]
    print(f'# {len(rows)} kernels; {tc} issue tcgen05.mma (UTC*MMA); kernels with legacy HMMA (mma.sync): {len(legacy)}')
    for s in legacy:
        print(f'#   HMMA: {s}')


if __name__ == '__main__':
    main()
