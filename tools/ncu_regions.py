import csv,subprocess,sys
rep=sys.argv[1]; binsz=int(sys.argv[2]) if len(sys.argv)>2 else 50
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
lines=src.splitlines()
start=next(i for i,l in enumerate(lines) if l.startswith('"Address"'))
rd=csv.DictReader(lines[start:])
recs=[]
for r in rd:
    try: recs.append((int(r["# Samples"]),r["Source"].strip(),r))
    except: pass
tot=sum(s for s,_,_ in recs)
stall_cols=[c for c in rd.fieldnames if c.startswith("stall_") and "Not Issued" not in c]
print("total",tot,"n",len(recs))
for b in range(0,len(recs),binsz):
    chunk=recs[b:b+binsz]
    s=sum(x[0] for x in chunk)
    if s< tot*0.01: continue
    agg={}
    for _,_,r in chunk:
        for c in stall_cols:
            agg[c]=agg.get(c,0)+int(r[c] or 0)
    top=sorted(agg.items(),key=lambda kv:-kv[1])[:3]
    ops={}
    for _,t,_ in chunk:
        o=t.split()[0] if not t.startswith('@') else t.split()[1]
        ops[o]=ops.get(o,0)+1
    topops=sorted(ops.items(),key=lambda kv:-kv[1])[:5]
    print(f"{b:5d}-{b+binsz:5d} {100*s/tot:5.1f}%  {top}  {topops}")
