import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {n:i for i,n in enumerate(hdr)}
seg=[]; cur=dict(n=0,s=0,w=0,ops=Counter(),first=None,cnt=0)
tot=0; tots=0
for r in rows[2:]:
    if len(r)<len(hdr): continue
    src=r[ci["Source"]].strip(); toks=src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks)>1 else toks[0]
    op=op.rstrip(";")
    n=int(float(r[ci["Instructions Executed"]] or 0)); s=int(float(r[ci["# Samples"]] or 0))
    w=r[ci["L1 Wavefronts Shared"]]; w=int(float(w)) if w else 0
    if cur["first"] is None: cur["first"]=r[ci["Address"]]
    cur["n"]+=n; cur["s"]+=s; cur["w"]+=w; cur["ops"][op.split('.')[0]]+=n; cur["cnt"]+=1
    tot+=n; tots+=s
    if op.startswith("BAR") or op.startswith("EXIT"):
        seg.append(cur); cur=dict(n=0,s=0,w=0,ops=Counter(),first=None,cnt=0)
seg.append(cur)
print("total", tot, "samples", tots)
for i,c in enumerate(seg):
    if c["n"]==0: continue
    top=", ".join("%s %.0f%%"%(k,100*v/c["n"]) for k,v in c["ops"].most_common(7))
    print("seg %2d sass %4d  inst %5.1f%%  samples %5.1f%%  smem-wavefronts %10d | %s"%(i,c["cnt"],100*c["n"]/tot,100*c["s"]/tots,c["w"],top))
