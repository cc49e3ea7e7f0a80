#!/usr/bin/env python3
"""Instructions executed and stall samples per SOURCE line of one kernel in an .ncu-rep (needs -lineinfo and
`ncu --import-source on`): ncu_lines.py rep kernel-regex [topN]"""
import csv,subprocess,sys,io,collections
rep,rx=sys.argv[1],sys.argv[2]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+rx,"--launch-count","1","--print-source","cuda,sass"],capture_output=True,text=True).stdout
lines=raw.splitlines()
cur_file=None; agg=collections.OrderedDict(); tot_i=0; tot_s=0
i=0
hdr=None
for l in lines:
    if l.startswith('"File Path"'):
        cur_file=l.split(',')[1].strip('"').split('/')[-1]; hdr=None; continue
    if l.startswith('"Line No"'):
        hdr=next(csv.reader([l])); continue
    if hdr is None: continue
    r=next(csv.reader([l]))
    if len(r)!=len(hdr) or r[0]=="" : continue
    try:
        ln=int(r[0])
    except: continue
    ins=int(r[7] or 0); smp=int(r[6] or 0)
    key=(cur_file,ln,r[1].strip()[:90])
    a=agg.setdefault(key,[0,0]); a[0]+=ins; a[1]+=smp; tot_i+=ins; tot_s+=smp
print("total warp-instr",tot_i,"samples",tot_s)
for (f,ln,src),(ins,smp) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 30]:
    print(f"{100*ins/tot_i:5.1f}% ins {100*smp/max(tot_s,1):5.1f}% smp  {f}:{ln}  {src}")
