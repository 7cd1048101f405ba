"""Top SASS instructions of an ncu source-page export by warp samples, with stall reason:
   ncu -i rep --page source --csv --print-source sass > x.csv ; python ncu_sass_top.py x.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = rows[1]
ix = {}
for i, h in enumerate(hdr): ix.setdefault(h, i)
stall = [h for h in hdr if h.startswith('stall_') and '(' not in h]
data = [r for r in rows[2:] if len(r) > 30]
def I_(x):
    try: return int(x)
    except ValueError: return 0
S = sum(I_(r[ix["# Samples"]]) for r in data)
I = sum(I_(r[ix["Instructions Executed"]]) for r in data)
print("samples", S, "instr", I)
tot = {h: sum(I_(r[ix[h]]) for r in data) for h in stall}
print(" ".join(f"{h[6:]}={100*v/S:.1f}%" for h, v in sorted(tot.items(), key=lambda x: -x[1])[:10]))
order = sorted(range(len(data)), key=lambda i: -I_(data[i][ix["# Samples"]]))[:n]
for i in sorted(order):
    r = data[i]
    top = sorted(((h, I_(r[ix[h]])) for h in stall), key=lambda x: -x[1])[:2]
    print(f"{i:5d} {100*I_(r[ix['# Samples']])/S:5.2f}%smp {100*I_(r[ix['Instructions Executed']])/I:5.2f}%inst {' '.join(f'{h[6:]}={v}' for h,v in top):36s} | {r[1].strip()[:90]}")
