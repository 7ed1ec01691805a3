"""Last (warm) record per (kernel, grid) of an `ncu --csv --metrics ...` log: python tools/ncu_last.py log.csv"""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; recs=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    key=(int(d['ID']),d['Kernel Name'].split('(')[0][:40],d['Grid Size'])
    recs.setdefault(key,{})[d['Metric Name']]=d['Metric Value']
seen=collections.OrderedDict()
for k,m in recs.items():
    seen[(k[1],k[2])]=m
for k,m in seen.items():
    print(k, {a.split('.')[0][-18:]:b for a,b in m.items()})
