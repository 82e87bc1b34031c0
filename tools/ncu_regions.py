#!/usr/bin/env python
"""Sample share and stall mix between consecutive BAR.SYNCs of a kernel (ncu source page; no GPU needed).
usage: ncu_regions.py report.ncu-rep kernel_regex"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu","-i",sys.argv[1],"--page","source","--csv","--kernel-name","regex:"+sys.argv[2],"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[1]
ia,ie,isamp=hdr.index('Source'),hdr.index('Instructions Executed'),hdr.index('# Samples')
cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[r for r in rows[2:] if len(r)>isamp and r[ie].isdigit()]
ts=sum(int(r[isamp]) for r in data)
bars=[i for i,r in enumerate(data) if 'BAR.SYNC' in r[ia] or 'EXIT' in r[ia]]
prev=0
for b in bars+[len(data)-1]:
    seg=data[prev:b+1]
    if not seg: continue
    sm=sum(int(r[isamp]) for r in seg)
    st={hdr[c][6:]:sum(int(r[c] or 0) for r in seg) for c in cols}
    st={k:v for k,v in st.items() if v>sm*0.08 and v>5}
    ex=max(int(r[ie]) for r in seg); wi=sum(int(r[ie]) for r in seg)
    if sm/ts>0.003: print(f'[{prev}:{b}] n={len(seg)} samples={sm/ts:.3f} maxexec={ex} warpinstr={wi}',st)
    prev=b+1
