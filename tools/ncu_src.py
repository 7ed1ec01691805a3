"""Hot source lines of an `ncu -i rep --page source --csv --print-source cuda,sass` export (share of stall samples and of executed
instructions per CUDA line): python tools/ncu_src.py export.csv [min_share]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.012
hdr=None; agg=[]; curfile=''
for r in rows:
    if len(r)>8 and r[0]=='Line No' and '# Samples' in r:
        hdr=r; continue
    if hdr is None or len(r)!=len(hdr):
        if len(r)==2 and r[0]=='File Path': curfile=r[1].split('/')[-1]
        continue
    if r[0]!='':
        agg.append((curfile, r[0], r[1], int(r[hdr.index('# Samples')] or 0), int(r[hdr.index('Instructions Executed')] or 0)))
ts=sum(a[3] for a in agg); ti=sum(a[4] for a in agg)
print(ts,ti)
for f,l,src,s_,n in agg:
    if s_>ts*thr or n>ti*thr:
        print(f'{f[:12]:12} {l:>4} {100*s_/ts:5.1f}% smp {100*n/ti:5.1f}% inst  {src[:100]}')
