import csv, io, subprocess, sys, collections
path=sys.argv[1]; binsz=int(sys.argv[2]) if len(sys.argv)>2 else 10
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur=None; hdr=None; agg=collections.Counter(); thr=collections.Counter()
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if r[0] in ('Function Name',) or hdr is None: continue
    if r[0] != '':
        d=dict(zip(hdr,r))
        try: ie=int(d["Instructions Executed"]); te=int(d["Thread Instructions Executed"]); ln=int(r[0])
        except (ValueError, KeyError): continue
        key=(cur, ln//binsz*binsz); agg[key]+=ie; thr[key]+=te
tot=sum(agg.values())
for k,c in sorted(agg.items()):
    if c/tot>0.004: print(f"{k[0]:24s} {k[1]:5d}  {100*c/tot:5.1f}%  thr/inst {thr[k]/c:5.1f}")
