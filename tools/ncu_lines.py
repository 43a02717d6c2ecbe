"""Join an ncu SASS source-page csv with nvdisasm -g line info: per source line instruction and
stall-sample shares.  usage: ncu_lines.py <ncu_source.csv> <api.sass> <mangled-substring> [topN]"""
import collections
import csv
import re
import sys

csv.field_size_limit(10**9)
src, sassf, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
sass = open(sassf).read().split("\n")
start = [i for i, l in enumerate(sass) if ".text." in l and pat in l and l.strip().startswith("//---")]
i0 = start[0]
cur, inst = None, []
for l in sass[i0 + 1:]:
    if l.startswith("//--------------------- .text.") and inst:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        inst.append((cur, m.group(2).strip()))
rows = list(csv.reader(open(src)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
data = rows[2:]
print("sass", len(inst), "ncu", len(data))
agg, aggs = collections.Counter(), collections.Counter()
tot = stot = 0
for k in range(min(len(inst), len(data))):
    n = int(data[k][ix["Instructions Executed"]] or 0)
    s = int(data[k][ix["# Samples"]] or 0)
    agg[inst[k][0]] += n
    aggs[inst[k][0]] += s
    tot += n
    stot += s
print("total instr", tot, "samples", stot)
byfile = collections.Counter()
byfiles = collections.Counter()
for k, n in agg.items():
    byfile[k[0] if k else None] += n
    byfiles[k[0] if k else None] += aggs[k]
print({k: (round(v / tot * 100, 1), round(byfiles[k] / stot * 100, 1)) for k, v in byfile.items()})
for k, n in sorted(aggs.items(), key=lambda kv: -kv[1])[:top]:
    print(k, f"{agg[k] / tot * 100:5.2f}% instr {n / stot * 100:5.2f}% samples")
