"""Summarise an exported ncu source page (scripts/gpu.sh: *_source_N.csv.gz): executed instructions per cell by opcode, stall-sample
shares, and the top source lines.   python scripts/ncu_source_summary.py file.csv.gz <cells per launch> [top]"""
import collections, csv, gzip, io, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
cells = float(sys.argv[2]) / 32
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
print(rows[0][1])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:          # the export repeats the listing (one per source view): keep each SASS address once
    if len(r) >= len(hdr) and r[ix["Instructions Executed"]].isdigit() and r[ix["Address"]] not in seen:
        seen.add(r[ix["Address"]])
        data.append(r)
ex = lambda r: int(r[ix['Instructions Executed']])
sm = lambda r: int(r[ix['# Samples']])
tot_exec, tot_samp = sum(map(ex, data)), sum(map(sm, data))
print(f"warp instructions executed {tot_exec}  = {tot_exec / cells:.1f} per cell; samples {tot_samp}; SASS lines {len(data)}")
c, cs = collections.Counter(), collections.Counter()
for r in data:
    t = r[ix['Source']].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += ex(r); cs[op] += sm(r)
fp64 = sum(v for k, v in c.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print(f"FP64 per cell {fp64 / cells:.1f}; other {(tot_exec - fp64) / cells:.1f}")
for k, v in c.most_common(30):
    print(f"  {k:10s} {v / cells:7.2f} per cell   {100 * cs[k] / tot_samp:5.1f} % of samples")
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("stall reasons (% of all samples):", {h[6:]: round(100 * sum(int(r[ix[h]]) for r in data) / tot_samp, 1) for h in st})
print(f"top {top} SASS lines by samples:")
for r in sorted(data, key=sm, reverse=True)[:top]:
    dom = max(st, key=lambda h: int(r[ix[h]]))
    print(f"  {100 * sm(r) / tot_samp:5.2f} %  exec/cell {ex(r) / cells:6.2f}  {dom[6:]:14s} {r[ix['Source']].strip()}")
