"""Developer tool: instructions executed / stall samples per CUDA source line of an ncu capture.
ncu -i rep --page source --csv --print-source cuda,sass > f.csv ; python scripts/ncu_lines.py f.csv [lo hi]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
fname = ""
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; iex = r.index("Instructions Executed") - len(r); ismp = r.index("# Samples") - len(r); continue   # from the end: ncu does not escape quotes inside source text
    if hdr is None or r[0] in ("Function Name",): continue
    if r[0] != "":      # a source line: its totals
        try: ln = int(r[0])
        except ValueError: continue
        key = (fname, ln)
        num = lambda v: int(v) if v not in ('', '-') else 0
        pe, ps = agg.get(key, ('', 0, 0))[1:]
        agg[key] = (r[1].strip(), pe + num(r[iex]), ps + num(r[ismp]))
tot_e = sum(v[1] for v in agg.values()); tot_s = sum(v[2] for v in agg.values())
print(f"total executed {tot_e}, samples {tot_s}")
for (f, ln), (src, e, s) in agg.items():
    if lo <= ln <= hi and (e > tot_e * 0.002 or s > tot_s * 0.004):
        print(f"{f}:{ln:5d} {100*e/tot_e:5.1f}% ex {100*s/max(tot_s,1):5.1f}% smp | {src[:110]}")
