"""Per-source-line instruction and stall-sample profile of one ncu capture taken with --import-source on:
   python profiles/source_profile.py capture.ncu-rep UNITS TOP   (UNITS = what to divide the instruction counts by,
   e.g. warp-steps of the launch; TOP = number of lines to print)"""
import csv, sys, subprocess, collections
rep, div, top = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
txt = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
fp=None; agg={}; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": fp=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0]!="" and len(r)>=8:
        try: ie=int(r[7]); sm=int(r[6])
        except: continue
        a=agg.setdefault((fp,int(r[0])),[0,0,r[1]]); a[0]+=ie; a[1]+=sm
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
print("total instr/unit", tot/div, "samples", ts)
byfile=collections.defaultdict(float)
for k,v in agg.items(): byfile[k[0]]+=v[0]/div
print(dict(byfile))
for (f,l),v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{v[0]/div:7.1f} {100*v[1]/ts:5.1f}% {f}:{l}  {v[2].strip()[:105]}")
