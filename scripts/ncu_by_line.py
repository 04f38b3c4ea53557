"""Join an exported ncu source page (SASS, scripts/gpu.sh) with `nvdisasm -g` output of the same kernel (instruction order) and aggregate
executed instructions / stall samples per source line.
  python scripts/ncu_by_line.py source.csv.gz lines.sass '<substring of the mangled kernel name>' <cells per launch> [top]"""
import collections, csv, gzip, io, re, sys
src, sass, key, cells = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]) / 32
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(io.TextIOWrapper(gzip.open(src))))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:
    if len(r) >= len(hdr) and r[ix["Instructions Executed"]].isdigit() and r[ix["Address"]] not in seen:
        seen.add(r[ix["Address"]]); data.append(r)
# nvdisasm: find the function's text section
lines = open(sass).read().split("\n")
ins, cur, infn = [], ("?", 0), False
for l in lines:
    if l.startswith(".text.") or l.startswith("\t.section\t.text."):
        infn = key in l
        continue
    if l.startswith("\t.section") or l.startswith(".section"):
        infn = False
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((cur, m.group(2).strip()))
print(f"ncu SASS lines {len(data)}, nvdisasm instructions {len(ins)}")
n = min(len(data), len(ins))
mism = sum(1 for i in range(n) if data[i][ix['Source']].split()[-1 if False else 0].strip('@!P0123456789U') == 'x')
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_e = tot_s = 0
for i in range(n):
    (f, ln), text = ins[i]
    r = data[i]
    e, s = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
    t = text.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    fp = op in ('DFMA', 'DMUL', 'DADD', 'DSETP')
    a = agg[(f, ln)]
    a[0] += e; a[1] += s; a[2] += e if fp else 0; a[3][op] += e
    tot_e += e; tot_s += s
srcs = {}
def srcline(f, ln):
    import glob
    if f not in srcs:
        p = glob.glob(f"/root/repo/oldoceananigans.jl_b200/csrc/{f}")
        srcs[f] = open(p[0]).read().split("\n") if p else []
    return srcs[f][ln - 1].strip()[:110] if 0 < ln <= len(srcs[f]) else ""
print(f"total exec/cell {tot_e / cells:.1f}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ops = ",".join(f"{k}:{v / cells:.1f}" for k, v in a[3].most_common(4))
    print(f"{a[0] / cells:7.2f}/cell fp64 {a[2] / cells:6.2f} samp {100 * a[1] / tot_s:5.1f}%  {f}:{ln:4d}  {srcline(f, ln)}\n          [{ops}]")
