"""Instruction / stall-sample share per kernel phase (source lines that start with `// ----`) from the same export as ncu_src.py:
python tools/ncu_phase.py export.csv path/to/kernel.cu"""
import csv,sys,re
rows=list(csv.reader(open(sys.argv[1])))
src=open(sys.argv[2]).read().split('\n')
# phase markers: lines starting with "  // ----" inside the kernel
hdr=None; agg=[]; curfile=''
for r in rows:
    if len(r)>8 and r[0]=='Line No' and '# Samples' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr):
        if len(r)==2 and r[0]=='File Path': curfile=r[1].split('/')[-1]
        continue
    if r[0]!='': agg.append((curfile,int(r[0]),int(r[hdr.index('# Samples')] or 0),int(r[hdr.index('Instructions Executed')] or 0)))
ti=sum(a[3] for a in agg); ts=sum(a[2] for a in agg)
fn=sys.argv[2].split('/')[-1]
marks=[(i+1,l.strip()[:60]) for i,l in enumerate(src) if l.strip().startswith('// ----')]
lo=min(a[1] for a in agg if a[0]==fn); hi=max(a[1] for a in agg if a[0]==fn)
marks=[(lo,'(start)')]+[m for m in marks if lo<m[0]<=hi]
for j,(ln,name) in enumerate(marks):
    end=marks[j+1][0]-1 if j+1<len(marks) else hi
    i=sum(x[3] for x in agg if x[0]==fn and ln<=x[1]<=end); s_=sum(x[2] for x in agg if x[0]==fn and ln<=x[1]<=end)
    print(f'{ln:4}-{end:4} inst {100*i/ti:5.1f}%  smp {100*s_/ts:5.1f}%  {name}')
print('total inst',ti,'samples',ts)
