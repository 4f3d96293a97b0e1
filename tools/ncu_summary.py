"""Compact per-launch table from an `ncu --set full` report: python tools/ncu_summary.py report.ncu-rep [title]
(duration, DRAM / L2 bytes, instructions, IPC, busiest pipes, shared-memory wavefronts, top stall reasons)."""
import csv
import io
import subprocess
import sys

COLS = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('lts__t_bytes.sum', 'l2_bytes'), ('smsp__inst_executed.sum', 'warp_inst'),
        ('sm__inst_executed.avg.per_cycle_elapsed', 'ipc'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor%'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'),
        ('sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'adu%'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu%'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsu_wavefronts%'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wavefronts'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_conflicts'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%')]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    if len(sys.argv) > 2:
        print('# ' + sys.argv[2])
    print(f'# source: ncu --set full --clock-control none ({rep.split("/")[-1]}); one line per captured launch')
    for r in data:
        name = r[idx['Kernel Name']]
        name = name[:name.find('(')] if '(' in name else name
        parts = []
        for col, label in COLS:
            if col in idx and r[idx[col]] not in ('', 'n/a'):
                v = r[idx[col]]
                try:
                    f = float(v.replace(',', ''))
                    v = f'{f:.4g}'
                except ValueError:
                    pass
                parts.append(f'{label}={v}{units[idx[col]] if label in ("time", "dram_rd", "dram_wr", "l2_bytes") else ""}')
        stalls = []
        for h in hdr:
            if 'pcsamp_warps_issue_stalled_' in h and 'not_issued' not in h:
                try:
                    stalls.append((float(r[idx[h]].replace(',', '')), h.split('stalled_')[1]))
                except ValueError:
                    pass
        tot = sum(s for s, _ in stalls) or 1.0
        top = ', '.join(f'{n} {100 * s / tot:.0f}%' for s, n in sorted(stalls, reverse=True)[:5])
        print(f'{name}\n    ' + ' '.join(parts) + f'\n    stall samples: {top}')


if __name__ == '__main__':
    main()
