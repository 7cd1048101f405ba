"""Per-CUDA-source-line breakdown of an ncu report (needs --import-source on at capture time):
   ncu -i rep --page source --csv --print-source cuda,sass > x.csv ; python ncu_by_line.py x.csv [min_pct]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
cur = None; hdr = None; data = []
for r in rows:
    if r and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if r and r[0] == "Function Name": continue
    if hdr and len(r) > 8 and r[0].isdigit():
        data.append((cur, int(r[0]), r))
def I_(x):
    try: return int(x)
    except ValueError: return 0
ix = {}
for i, h in enumerate(hdr): ix.setdefault(h, i)
S = sum(I_(r[ix["# Samples"]]) for _, _, r in data); I = sum(I_(r[ix["Instructions Executed"]]) for _, _, r in data)
print("total samples", S, "warp-instructions", I)
stall = [h for h in hdr if h.startswith('stall_')]
tot = collections.Counter()
for f, ln, r in data:
    for h in stall: tot[h] += I_(r[ix[h]])
print("stalls:", " ".join(f"{h[6:]}={100*v/S:.1f}%" for h, v in tot.most_common(9)))
for f, ln, r in sorted(data, key=lambda x: (x[0], x[1])):
    s = I_(r[ix["# Samples"]]); n = I_(r[ix["Instructions Executed"]])
    if s * 100 < thr * S and n * 100 < thr * I: continue
    top = sorted(((h, I_(r[ix[h]])) for h in stall), key=lambda x: -x[1])[:2]
    print(f"{f}:{ln:4d} {100*s/S:5.1f}%smp {100*n/I:5.1f}%inst {' '.join(f'{h[6:]}={v}' for h, v in top):32s} | {r[1].strip()[:100]}")
