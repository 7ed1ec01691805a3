#!/usr/bin/env python3
"""Per-family DRAM traffic of the 1x1-conv GEMM launches from an ncu metrics pass over one eager batch:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \\
        --clock-control none -k regex:gemm_ --csv --log-file gpurun_out/gemm_dram.csv python tools/profile_step.py 57
    python tools/traffic_from_ncu.py gpurun_out/gemm_dram.csv gpurun_out/profile_step_labels.json profiles/r1h_traffic.json

The ncu rows (launch order) are matched one to one with the engine's launch-ordered family labels written by profile_step.py.
"""
import csv
import json
import sys


def main():
    rows_csv, labels_json, out = sys.argv[1:4]
    lines = [l for l in open(rows_csv) if l.startswith('"')]
    per = {}
    for d in csv.DictReader(lines):
        if not d['Kernel Name'].startswith('gemm_') and 'gemm_' not in d['Kernel Name']:
            continue
        per.setdefault(int(d['ID']), {'kernel': d['Kernel Name']})[d['Metric Name']] = float(d['Metric Value'].replace(',', ''))
    launches = [per[k] for k in sorted(per)]
    meta = json.load(open(labels_json))
    labels = meta['gemm_launches']
    if len(launches) != len(labels):
        raise SystemExit('ncu saw %d gemm launches, the engine recorded %d' % (len(launches), len(labels)))
    fam = {}
    for (lab, nbytes), l in zip(labels, launches):
        f = fam.setdefault(lab, dict(launches=0, dram_read_bytes=0.0, dram_write_bytes=0.0, alg_bytes=0.0, ns=0.0))
        f['launches'] += 1
        f['dram_read_bytes'] += l['dram__bytes_read.sum']
        f['dram_write_bytes'] += l['dram__bytes_write.sum']
        f['alg_bytes'] += nbytes
        f['ns'] += l.get('gpu__time_duration.sum', 0.0)
    for f in fam.values():
        f['dram_bytes_per_launch'] = (f['dram_read_bytes'] + f['dram_write_bytes']) / f['launches']
        f['alg_bytes_per_launch'] = f['alg_bytes'] / f['launches']
    json.dump(dict(source='ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one eager batch '
                          '(bench.py --ncu-step: one 169-clip video), FigureSkatingComp_small %s' % meta['precision'],
                   clips_per_batch=meta['clips_per_batch'], path=meta.get('path', 'engine'), clips_per_step=meta.get('clips_per_step'),
                   families=fam), open(out, 'w'), indent=1)
    for k, f in fam.items():
        print(k, f['launches'], 'dram/launch %.1f MB' % (f['dram_bytes_per_launch'] / 1e6), 'alg/launch %.1f MB' % (f['alg_bytes_per_launch'] / 1e6))


if __name__ == '__main__':
    main()
