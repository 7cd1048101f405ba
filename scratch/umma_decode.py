import sys, ctypes, torch
sys.path.insert(0, '.')
import vlg_b200
from vlg_b200 import _lib
lib = _lib.load()
fn = lib.vlg_selftest_umma_ex
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p]*4 + [ctypes.c_int]*7 + [ctypes.c_void_p]
def run(N, K, mn, lbo=0, sbo=0, kstep=0):
    KK = N if mn else K; NN = K if mn else N
    img = torch.arange(N*K, dtype=torch.float32).cuda()   # value = float index in the image
    A = torch.zeros(128, KK)
    for m in range(128): A[m, m % KK] = 1.0
    A = A.cuda(); D = torch.full((128, NN), -1.0).cuda()
    rc = fn(A.data_ptr(), img.data_ptr(), img.data_ptr(), D.data_ptr(), N, K, mn, 0, lbo, sbo, kstep, 0)
    torch.cuda.synchronize()
    return rc, D.cpu()
torch.set_printoptions(linewidth=250, sci_mode=False)
for (N,K,mn,lbo,sbo,ks) in [(16,16,0,0,0,0),(16,16,1,0,0,0),(16,16,1,256,128,0),(32,16,1,0,0,0),(16,32,1,0,0,0)]:
    rc, D = run(N,K,mn,lbo,sbo,ks)
    print("N",N,"K",K,"mn",mn,"lbo",lbo,"sbo",sbo,"rc",rc)
    print(D[:KK if (KK:=(N if mn else K))<=16 else 16, :].int())
