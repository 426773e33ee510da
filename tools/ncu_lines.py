#!/usr/bin/env python
"""Per-source-line instruction counts from an .ncu-rep (needs -lineinfo): where do the warp instructions go?"""
import csv, io, subprocess, sys, collections

def lines(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr = None, None
    agg = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] in ("Function Name",) or hdr is None:
            continue
        if r[0] != "":  # a source line row
            d = dict(zip(hdr, r))
            try:
                ie = int(d["Instructions Executed"]); te = int(d["Thread Instructions Executed"]); smp = int(d["# Samples"])
            except (ValueError, KeyError):
                continue
            if ie:
                key = (cur_file, int(r[0]))
                a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:100]])
                a[0] += ie; a[1] += te; a[2] += smp
    tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
    print(f"total warp-inst {tot:.3e}, samples {tots}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*a[0]/tot:5.1f}% inst {100*a[2]/max(1,tots):5.1f}% smp  thr/inst {a[1]/a[0]:5.1f}  {f}:{ln}  {a[3]}")

if __name__ == "__main__":
    lines(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
