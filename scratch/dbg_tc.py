import sys; sys.path.insert(0, ".")
import numpy as np, torch
import vlg_b200
from tests import helpers as Hh
g = Hh.load("synth_np4_T130"); K = int(g["K"]); T = int(g["T"])
arrs = Hh.decoder_arrays(g)
dec = vlg_b200.DecoderEnsemble.from_arrays(*[arrs[k] for k in Hh.DEC_KEYS], "cuda")
t = torch.linspace(0, 1, T, device="cuda")
def model():
    return vlg_b200.GeodesicSplineBatch(*(torch.tensor(g[k], device="cuda") for k in ("a", "b", "basis", "omega_init")), int(g["n_poly"]))
for kk in (1, 2, 4):
    for M in (1, 2):
        e32 = vlg_b200.compute_energy_mc(model(), dec[:kk], t, M=M, seed=1, precision="fp32").cpu().numpy()
        etc = vlg_b200.compute_energy_mc(model(), dec[:kk], t, M=M, seed=1, precision="tf32").cpu().numpy()
        print("K", kk, "M", M, "ratio tf32/fp32", np.round(etc / e32, 4))
