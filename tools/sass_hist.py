"""SASS instruction histogram per kernel of libv2ce_b200.so (cuobjdump -sass): the mnemonics that prove the Blackwell
paths (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDTM = tcgen05.ld, SYNCS = mbarrier) plus
the ten most frequent opcodes.  python tools/sass_hist.py > profiles/sass_histogram_rN.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'v2ce_toolbox_b200', 'libv2ce_b200.so')
KEY = ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'SYNCS', 'UTCBAR', 'UTCATOMSWS', 'REDUX', 'MATCH', 'VOTE', 'SHFL', 'ATOMS', 'ATOMG', 'RED')


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f'# {os.path.basename(LIB)}: {len(kernels)} kernels, {sum(sum(c.values()) for c in kernels.values())} SASS instructions')
    tot = collections.Counter()
    for c in kernels.values():
        tot.update({k: v for k, v in c.items() if k in KEY})
    print('# whole library: ' + ', '.join(f'{k} {tot[k]}' for k in KEY if tot[k]))
    for (name, c), pretty in zip(kernels.items(), demangle):
        short = re.sub(r'\(.*', '', pretty)[:110]
        n = sum(c.values())
        keys = ', '.join(f'{k} {c[k]}' for k in KEY if c[k])
        top = ', '.join(f'{k} {v}' for k, v in c.most_common(8))
        print(f'{short}\n    {n} instructions | {keys or "-"} | top: {top}')


if __name__ == '__main__':
    sys.exit(main())
