#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name for the LAST forward.
usage: python scripts/ncu_launches.py launches.csv [launches_per_forward]"""
import csv, sys, collections, re
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None; data=[]
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr and r[0].isdigit(): data.append(r)
iname=hdr.index('Kernel Name'); ival=hdr.index('Metric Value'); iunit=hdr.index('Metric Unit')
# one forward = the launches from one patch-im2col kernel (first kernel of the encoder) to the next one
marks=[i for i,r in enumerate(data) if 'im2col_patch' in r[iname]]
if len(sys.argv)>2:
    n=int(sys.argv[2]); last=data[-n:]
elif len(marks)>=2:
    last=data[marks[-2]:marks[-1]]; n=len(last)
else:
    n=len(data)//3; last=data[-n:]
agg=collections.OrderedDict()
for r in last:
    name=re.sub(r'\(.*','',r[iname]); name=name.replace('void ','').replace('dclip::','')
    v=float(r[ival].replace(',','')); u=r[iunit]
    v = v/1000 if u in ('ns','nsecond') else v   # -> us
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print(f"{n} launches, total {tot/1000:.3f} ms (serialised, cold-cache: compare shares)")
for k,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{t/1000:9.3f} ms {100*t/tot:5.1f}%  x{c:<4d} {k[:110]}")
