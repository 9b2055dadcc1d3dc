#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` into instructions / stall samples / shared-memory
wavefronts per CUDA source line (top N).  Usage: ncu_source_summary.py <csv> [N]"""
import csv,sys
def toi(x):
    try: return int(x)
    except: return 0
f=sys.argv[1]
rows=list(csv.reader(open(f)))
cur=None; hdr=None; out=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and r and r[0]!="" and len(r)>10:
        d=dict(zip(hdr[4:],r[4:]))
        out.append((cur,int(r[0]),r[1].strip()[:90],toi(d["Instructions Executed"]),toi(d["Thread Instructions Executed"]),toi(d["# Samples"]), toi(d.get("L1 Wavefronts Shared",0)), toi(d.get("L1 Wavefronts Shared Ideal",0))))
tot=sum(o[3] for o in out); tots=sum(o[5] for o in out); totw=sum(o[6] for o in out)
print("total warp instr",tot,"samples",tots,"smem wavefronts",totw)
out.sort(key=lambda o:-o[3])
for o in out[:int(sys.argv[2]) if len(sys.argv)>2 else 60]:
    print(f"{o[0][:18]:18s} {o[1]:5d} {100*o[3]/tot:5.1f}% thr/warp {o[4]/max(o[3],1):5.1f} smp {100*o[5]/tots:5.1f}% wf {100*o[6]/max(totw,1):5.1f}% (ideal {100*o[7]/max(totw,1):4.1f}) | {o[2]}")
