"""Build libvlg_b200 variants with different -D flags into scratch/variants_build/<name>.so
(selected at run time with VLG_B200_LIB).  usage: build_variants.py name="-DFOO=1 -DBAR=2" ..."""
import subprocess, sys, shlex
from pathlib import Path
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vlg_b200
from vlg_b200 import build as B
out = Path(__file__).resolve().parent / "variants_build"
out.mkdir(exist_ok=True)
def one(arg):
    name, flags = arg.split("=", 1)
    cmd = [B._nvcc(), *B.NVCC_FLAGS, *shlex.split(flags), "-o", str(out / f"{name}.so"), *[str(B.CSRC / s) for s in B.SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-2000:]
with ThreadPoolExecutor(3) as ex:
    for name, rc, err in ex.map(one, sys.argv[1:]):
        print(name, rc, err if rc else "")
