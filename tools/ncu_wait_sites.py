"""Where the warps of one captured launch wait: python tools/ncu_wait_sites.py report.ncu-rep <launch index> [top]
Reads `ncu --page source --csv` (SASS view, PC sampling) and lists the instructions with the most samples, with the
mbarrier try-wait sites (SYNCS...TRYWAIT) and the tcgen05 / TMEM instructions marked, plus the share of all samples that
falls between consecutive UTCHMMA (the MMA issue loop)."""
import csv
import io
import subprocess
import sys


def main():
    rep, launch = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(launch), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1]
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) > ix['# Samples'] and r[0].startswith('0x')]
    seen = set()
    uniq = []
    for r in body:                                  # the page lists the function once per view: keep the first
        if r[0] in seen:
            break
        seen.add(r[0])
        uniq.append(r)
    body = uniq
    samples = [int(r[ix['# Samples']] or 0) for r in body]
    execd = [int(r[ix['Instructions Executed']] or 0) for r in body]
    total = sum(samples)
    print(f'# {name[:100]}')
    print(f'# launch {launch}: {total} samples, {sum(execd)} warp instructions, {len(body)} SASS instructions')
    mma = [i for i, r in enumerate(body) if 'UTCHMMA' in r[ix['Source']]]
    if mma:
        lo, hi = mma[0], mma[-1]
        inside = sum(samples[lo:hi + 1])
        n_mma = sum(execd[i] for i in mma)
        n_inst = sum(execd[lo:hi + 1])
        print(f'# MMA issue region (first..last UTCHMMA): {100 * inside / max(total, 1):.1f} % of samples; '
              f'{n_inst / max(n_mma, 1):.2f} executed warp instructions per executed UTCHMMA ({n_mma} MMAs)')
    order = sorted(range(len(body)), key=lambda i: -samples[i])[:top]
    for i in order:
        src = body[i][ix['Source']].strip()
        mark = 'WAIT ' if 'TRYWAIT' in src else 'MMA  ' if 'UTCHMMA' in src else 'TMEM ' if 'LDTM' in src else '     '
        print(f'{mark}{100 * samples[i] / max(total, 1):5.1f} %  exec {execd[i]:9d}  +{i:5d}  {src[:90]}')


if __name__ == '__main__':
    main()
