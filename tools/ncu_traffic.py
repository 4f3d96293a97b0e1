"""DRAM / L2 bytes and duration of every launch of ONE forward from an ncu CSV:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum \
        --clock-control none --csv --log-file traffic.csv python tools/layer_times.py 4 1
    python tools/ncu_traffic.py traffic.csv > profiles/forward_traffic_rN.txt
The CSV holds several forwards (warm-up, timed, untimed); the last complete one (from its sn_wtu_kernel launch) is listed."""
import collections
import csv
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0,
        'nsecond': 1e-3, 'msecond': 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= ix['Metric Value']:
            continue
        key = int(r[ix['ID']])
        name = r[ix['Kernel Name']]
        name = name[:name.find('(')] if '(' in name else name
        d = launches.setdefault(key, {'name': name.replace('void ', '').replace('v2ce::', '')})
        try:
            v = float(r[ix['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        d[r[ix['Metric Name']]] = v * UNIT.get(r[ix['Metric Unit']], 1.0)
    keys = list(launches)
    starts = [k for k in keys if 'sn_wtu_kernel' in launches[k]['name']]
    first = starts[-1]
    nxt = [k for k in keys if k > first and 'sn_wtu_kernel' in launches[k]['name']]
    sel = [k for k in keys if k >= first and (not nxt or k < nxt[0])]
    tot = collections.Counter()
    print('# one V2ce3d forward, batch 4 x 16 x 260 x 346 (cold-cache, serialised launches)')
    for k in sel:
        d = launches[k]
        t, rd, wr, l2 = (d.get(m, 0.0) for m in ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                                                 'lts__t_bytes.sum'))
        tot.update(t=t, rd=rd, wr=wr, l2=l2)
        print(f'{k:4d} {d["name"][:52]:52s} {t:8.1f} us  read {rd / 1e6:8.1f} MB  write {wr / 1e6:8.1f} MB  L2 {l2 / 1e6:8.1f} MB')
    print(f'# total: {len(sel)} launches, {tot["t"] / 1e3:.3f} ms, DRAM read {tot["rd"] / 1e9:.2f} GB + write {tot["wr"] / 1e9:.2f} GB = '
          f'{(tot["rd"] + tot["wr"]) / 1e9:.2f} GB, L2 {tot["l2"] / 1e9:.1f} GB')


if __name__ == '__main__':
    main()
