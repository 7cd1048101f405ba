import sys, ctypes
sys.path.insert(0, ".")
import torch, vlg_b200
from vlg_b200 import _lib
lib = _lib.load()
fn = lib.vlg_selftest_mma_rate
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
for mode in (0, 1, 2, 3):
    for (N, lbo, sbo) in [(128, 2048, 128), (64, 1024, 128), (256, 4096, 128)]:
        ctas = 148
        out = torch.zeros(ctas, dtype=torch.int64, device="cuda")
        iters = 4096
        assert fn(N, iters, lbo, sbo, ctas, mode, out.data_ptr(), 0) == 0
        torch.cuda.synchronize()
        c = out.float().mean().item() / iters
        print(f"mode {mode} (1=tmem ld/st, 2=lds) N {N:3d}: {c:.1f} cycles/MMA  -> {128*N*8/c:.0f} MAC/clk/SM")
