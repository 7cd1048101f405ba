"""Where does a kernel spill?  usage: spill_lines.py <lib.so> <mangled-substring>  (needs -lineinfo)"""
import re, collections, subprocess, sys, tempfile, os, glob
so, key = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
cub = glob.glob(d + "/*.cubin")[0]
txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.splitlines()
on = False; cur = None; cnt = collections.Counter()
for l in txt:
    if l.startswith(".text."):
        on = key in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.search(r'\b(STL|LDL)', l): cnt[(cur, 'STL' if 'STL' in l else 'LDL')] += 1
for (c, k), v in sorted(cnt.items(), key=lambda x: (x[0][0] or ('', 0), x[0][1])): print(c, k, v)
