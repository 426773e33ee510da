import csv, collections, sys, subprocess, io
def mix(path, top=14):
    out=subprocess.run(["ncu","-i",path,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
    rows=list(csv.reader(io.StringIO(out)))
    hdr=None; agg=collections.Counter(); thr=collections.Counter()
    for r in rows:
        if r and r[0]=='Address': hdr=r; continue
        if hdr is None or len(r)<len(hdr): continue
        d=dict(zip(hdr,r))
        try: ie=int(d['Instructions Executed']); te=int(d['Thread Instructions Executed'])
        except: continue
        toks=d['Source'].split()
        op=toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
        op=op.split('.')[0] if not op.startswith('IMAD') else '.'.join(op.split('.')[:2])
        agg[op]+=ie; thr[op]+=te
    tot=sum(agg.values())
    print(path.split('/')[-1], f"total {tot:.3e}")
    print('  '+'  '.join(f"{op} {100*c/tot:.1f}" for op,c in agg.most_common(top)))
for p in sys.argv[1:]: mix(p)
