"""Map ncu per-SASS samples to CUDA source lines (via nvdisasm -g) for tc_curve_kernel<true>."""
import re, csv, collections, subprocess, io, sys
rep, lineinfo, srcfile = sys.argv[1], sys.argv[2], sys.argv[3]
txt = open(lineinfo).read().split('\n')
infn = False; cur = None; amap = {}
for ln in txt:
    if '.text.' in ln: infn = 'tc_curve_kernelILb1' in ln
    m2 = re.search(r'//## File ".*?([A-Za-z_0-9]+\.cuh?)", line (\d+)', ln)
    if m2: cur = (m2.group(1), int(m2.group(2)))
    m3 = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m3 and infn: amap[int(m3.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
byline = collections.defaultdict(collections.Counter); S = 0; base = None
for r in data:
    try: s = int(r[ix['# Samples']]); n = int(r[ix['Instructions Executed']]); a = int(r[ix['Address']], 16)
    except Exception: continue
    if base is None: base = a
    S += s
    key = amap.get(a - base)
    byline[key]['samples'] += s; byline[key]['inst'] += n
    for c in stall_cols:
        try: byline[key][c] += int(r[ix[c]])
        except ValueError: pass
lines = open(srcfile).read().split('\n')
I = sum(v['inst'] for v in byline.values())
for k, v in sorted(byline.items(), key=lambda kv: -kv[1]['samples'])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    top = sorted(((c, v[c]) for c in stall_cols), key=lambda x: -x[1])[:3]
    tops = " ".join(f"{c[6:]}={100*n/max(v['samples'],1):.0f}%" for c, n in top)
    txtl = lines[k[1] - 1].strip()[:70] if k and k[0] == srcfile.split('/')[-1] else ''
    print(f"{100*v['samples']/S:5.1f}% smp {100*v['inst']/I:5.1f}% inst  {k}  [{tops}]  {txtl}")
