"""Compare the SASS of every kernel in two object files (cuobjdump -sass), kernel by kernel.

    python tools/sass_diff.py old.o new.o [substring ...]

Used to show that a refactor left a hardware-verified kernel untouched when no GPU is at hand: "opcodes identical" means
the same instruction sequence (only registers / constant-bank parameter offsets may differ), "text identical" the same
disassembly.  Template arguments appended on one side are ignored when pairing (``emit_kernel<4, unsigned int, false>``
pairs with ``emit_kernel<4, unsigned int, false, false>``)."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', cur).replace('void ', '')
            res[cur] = []
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(.*?);', line)
        if m and cur:
            res[cur].append(re.sub(r'\s+', ' ', m.group(1)))
    return res


def opcode(ins):
    parts = ins.split(' ')
    return parts[1] if parts[0].startswith('@') else parts[0]


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    wanted = sys.argv[3:]
    for a in sorted(old):
        if wanted and not any(w in a for w in wanted):
            continue
        stem = a[:-1] if a.endswith('>') else a
        match = [b for b in new if b == a or (b.startswith(stem + ',') and b.count(', false') > a.count(', false'))]
        match = [b for b in match if b == a] or match
        if not match:
            print(f'{a:70s} -> no counterpart')
            continue
        ia, ib = old[a], new[match[0]]
        same_ops = [opcode(i) for i in ia] == [opcode(i) for i in ib]
        print(f'{a:70s} {len(ia):5d} {len(ib):5d}  ' + ('opcodes identical' if same_ops else 'OPCODES DIFFER') +
              (' | text identical' if ia == ib else ' | text differs'))


if __name__ == '__main__':
    main()
