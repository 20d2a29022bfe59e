"""Developer tool: hottest SASS regions of an ncu capture (needs `ncu -i rep --page source --csv > file`).
python scripts/ncu_hot.py file.csv [top]: prints instruction-executed and stall-sample totals by opcode, and the
top instructions by samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(f"instructions executed {tot_inst}, samples {tot_samp}, sass lines {len(data)}")
by_op = collections.Counter(); by_op_s = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    by_op[op] += int(r[ix["Instructions Executed"]]); by_op_s[op] += int(r[ix["# Samples"]])
print("by opcode (inst %, samples %):")
for op, n in by_op.most_common(25):
    print(f"  {op:10s} {100*n/tot_inst:5.1f} %  {100*by_op_s[op]/max(tot_samp,1):5.1f} %")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
print("stall reasons (all samples):")
tot = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, n in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:24s} {100*n/max(tot_samp,1):5.1f} %")
print("top instructions by samples:")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    best = max(stalls, key=lambda s: int(r[ix[s]]))
    print(f"  #{i:5d} {int(r[ix['# Samples']]):7d} smp {int(r[ix['Instructions Executed']]):9d} ex  {best:18s} {r[ix['Source']].strip()[:90]}")
