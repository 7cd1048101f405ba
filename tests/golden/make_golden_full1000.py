"""Golden for the north-star statement: BASELINE config 1 at full length -- all 45 curves of
experiment/splines_init_model_seed12/spline_batch_init_euclidean_10.pt, 1000 Adam steps, K=10, M=2,
T=2000 -- run with the REFERENCE's own classes (src/optimize.py:13-75,152-162) in fp64, decoder draws
taken from the counter-based stream (oracle.counter_draws, seed below) so the GPU kernels can replay
them without storing 2.9 GB of indices.  Takes ~1 h of CPU; run once in the build container:

    python tests/golden/make_golden_full1000.py
"""
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
ROOT = OUT.parent.parent
for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
             "mpl_toolkits.axes_grid1"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, str(REF))  # the reference's `src` package first ...

import src.optimize as ref_opt  # noqa: E402
from src.train import EVAE, GaussianDecoder, GaussianEncoder, GaussianPrior, make_decoder_net, make_encoder_net  # noqa: E402

sys.path.insert(0, str(ROOT))  # ... then this repo (its own `src` is not imported here)
from oracle import geodesic_oracle as O  # noqa: E402

SEED, STEPS, T, M, K = 2024, 1000, 2000, 2, 10
torch.set_num_threads(8)


class CounterFeeder:
    """Stands in for torch.randint inside compute_energy_mc: serves oracle.counter_draws lazily."""

    def __init__(self, n):
        self.ids = np.arange(n)
        self.step, self.call, self.cur = 0, 0, None

    def __call__(self, low, high, size, device=None, **kw):
        if self.call == 0:
            self.cur = torch.from_numpy(O.counter_draws(SEED, self.ids, self.step, T, M, K))
        m, role = divmod(self.call, 2)
        out = self.cur[m, role]
        self.call += 1
        if self.call == 2 * M:
            self.call = 0
            self.step += 1
        return out


if __name__ == "__main__":
    model = EVAE(GaussianPrior(2), GaussianEncoder(make_encoder_net(50, 2)), GaussianDecoder(make_decoder_net(2, 50)), num_decoders=10)
    model.load_state_dict(torch.load(REF / "experiment/model_seed12.pt", map_location="cpu"))
    model = model.double().eval()
    for prm in model.parameters():
        prm.requires_grad_(False)  # skips only the decoder weight gradients the reference never uses
    decoders = list(model.decoder)
    blob = torch.load(REF / "experiment/splines_init_model_seed12/spline_batch_init_euclidean_10.pt", map_location="cpu",
                      weights_only=False)
    sd = blob["spline_data"]
    a = torch.stack([d["a"] for d in sd]).double()
    b = torch.stack([d["b"] for d in sd]).double()
    om = torch.stack([d["omega_init"] for d in sd]).double()
    basis = sd[0]["basis"].double()
    t_vals = torch.linspace(0, 1, T).double()
    spl = ref_opt.GeodesicSplineBatch(a, b, basis, om.clone(), 4)
    opt = torch.optim.Adam([spl.omega], lr=1e-3)
    feeder = CounterFeeder(len(sd))
    ref_opt.torch.randint = feeder
    keep = {0, 1, 10, 50, 100, 250, 500, 750, 999}
    energies = {}
    t0 = time.time()
    for step in range(STEPS):
        opt.zero_grad()
        energy = ref_opt.compute_energy_mc(spl, decoders, t_vals, M=M)
        endpoint_error = (spl(t_vals[-1:]) - b[None]) ** 2
        loss = energy + 1000 * endpoint_error.sum(dim=(0, 2))
        loss.sum().backward()
        opt.step()
        if step in keep:
            energies[step] = energy.detach().numpy().copy()
        if step % 50 == 0:
            print(f"step {step} mean energy {energy.mean().item():.3f}  ({time.time() - t0:.0f} s)", flush=True)
    np.savez_compressed(OUT / "ens_seed12_full1000.npz", seed=SEED, steps=STEPS, T=T, M=M, K=K,
                        energy_steps=np.array(sorted(energies)), energy_f64=np.stack([energies[s] for s in sorted(energies)]),
                        omega_f64=spl.omega.detach().numpy(), final_length_f64=np.sqrt(energies[STEPS - 1]))
    print("done", time.time() - t0)
