#!/usr/bin/env python
"""profiles/roofline_traffic.json from an `ncu --page raw --csv` export: DRAM bytes (read + write) per launch
of the kernels matching a regex, keyed the way bench.py looks them up ('<workload>/<model>/aggregation').

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > profiles/r2_spmm_raw.csv
    python tools/ncu_traffic.py profiles/r2_spmm_raw.csv ml-25m/gcn/aggregation 'csr_rows_kernel|csr_chunk_kernel'
"""
import csv
import json
import os
import re
import sys


def main():
    path, key, pattern = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
    rows = list(csv.reader(open(path)))
    head = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    cols = {name: j for j, name in enumerate(rows[head])}
    rd, wr, dur = cols['dram__bytes_read.sum'], cols['dram__bytes_write.sum'], cols.get('gpu__time_duration.sum')
    units = rows[head + 1]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    total, n, per = 0.0, 0, []
    for r in rows[head + 2:]:
        if len(r) <= max(rd, wr) or not pattern.search(r[cols['Kernel Name']]):
            continue
        b = float(r[rd].replace(',', '')) * scale.get(units[rd], 1.0) + float(r[wr].replace(',', '')) * scale.get(units[wr], 1.0)
        per.append({'kernel': r[cols['Kernel Name']][:80], 'dram_bytes': b,
                    'duration': (r[dur] + ' ' + units[dur]) if dur is not None else None})
        total += b
        n += 1
    out_path = os.path.join(os.path.dirname(os.path.abspath(path)), 'roofline_traffic.json')
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    table[key] = {'dram_bytes_per_launch': total / max(n, 1), 'launches': n, 'source': os.path.relpath(path), 'per_launch': per}
    json.dump(table, open(out_path, 'w'), indent=1)
    print(key, table[key]['dram_bytes_per_launch'], 'bytes per launch over', n, 'launches')


if __name__ == '__main__':
    main()
