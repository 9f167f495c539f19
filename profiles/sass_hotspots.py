"""Join an `ncu --page source --csv` (SASS view) dump with `nvdisasm -g` line info and print the hottest
source lines: share of executed warp instructions, share of stall samples, average active threads.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libseir_b200.so; nvdisasm -g -c seir_capi.sm_100a.cubin > all.sass
    python profiles/sass_hotspots.py src.csv all.sass <kernel mangled name fragment>
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, frag = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(sass).read().split("\n")
# restrict to the kernel's section
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and frag in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//-----")), len(lines))
loc = {}
cur = ("?", 0)
inl = []
for l in lines[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        loc[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(data[0][0], 16)
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
tot_i = tot_s = 0.0
for r in data:
    off = int(r[0], 16) - base
    key = loc.get(off, ("?", 0))
    inst = float(r[ix["Instructions Executed"]] or 0)
    thr = float(r[ix["Thread Instructions Executed"]] or 0)
    smp = float(r[ix["# Samples"]] or 0)
    a = agg[key]
    a[0] += inst; a[1] += thr; a[2] += smp
    tot_i += inst; tot_s += smp
print("total warp instructions %.0f, samples %.0f" % (tot_i, tot_s))
top = sorted(agg.items(), key=lambda kv: -kv[1][2])[: int(sys.argv[4]) if len(sys.argv) > 4 else 40]
for (f, ln), (inst, thr, smp) in top:
    print("%-18s:%4d  inst %5.1f%%  stall-samples %5.1f%%  avg-threads %4.1f" % (f, ln, 100 * inst / tot_i, 100 * smp / tot_s, thr / inst if inst else 0))
