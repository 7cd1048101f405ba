"""Golden for the BENCHMARKED job: curves of bench.synthetic_workload (BASELINE config 3: 8778 pairs, the
committed eVAE seed-12 decoders, K=10, M=2, T=2000), 1000 free-running Adam steps, run with the
REFERENCE's own classes (src/optimize.py:13-75 and the loop at 152-162) in fp64 AND in fp32, decoder
draws from the counter-based stream the kernels use in production (oracle.counter_draws, seed 0, keyed on
the GLOBAL curve id) -- so the GPU run of the full 8778-pair job can be compared curve by curve.

The reference's own fp32-vs-fp64 gap after 1000 steps is stored beside the fp64 result: Adam normalises
the gradient, so rounding-level perturbations grow along badly conditioned curves, and a tolerance on final
lengths only means something next to what the reference's own arithmetic does on the same curve.

    python tests/golden/make_golden_config3.py --ids 0:48 --extra gpurun_out/worst_ids.txt --dtype f64
    python tests/golden/make_golden_config3.py ... --dtype f32
    python tests/golden/make_golden_config3.py --merge      # -> config3_synth_1000.npz

~1.5 h of CPU per dtype for 64 curves (8 threads).  /root/reference is needed (build container only).
"""
import argparse
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
ROOT = OUT.parent.parent
SEED, STEPS, T, M, K, N_POLY = 0, 1000, 2000, 2, 10, 4
KEEP = [0, 1, 10, 50, 100, 250, 500, 750, 999]


def parse_ids(spec, extra):
    ids = []
    for part in spec.split(","):
        if ":" in part:
            lo, hi = part.split(":")
            ids += list(range(int(lo), int(hi)))
        elif part:
            ids.append(int(part))
    if extra and Path(extra).exists():
        ids += [int(x) for x in Path(extra).read_text().split()]
    out = []
    for i in ids:
        if i not in out:
            out.append(i)
    return np.array(out, dtype=np.int64)


def run(ids, dtype, threads):
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, str(REF))  # the reference's `src` package first ...
    import src.optimize as ref_opt
    from src.train import EVAE, GaussianDecoder, GaussianEncoder, GaussianPrior, make_decoder_net, make_encoder_net
    sys.path.insert(0, str(ROOT))  # ... then this repo (its own `src` is not imported here)
    import bench
    from oracle import geodesic_oracle as O

    torch.set_num_threads(threads)
    tdt = torch.float64 if dtype == "f64" else torch.float32

    class CounterFeeder:
        """Stands in for torch.randint inside compute_energy_mc: serves oracle.counter_draws lazily."""

        def __init__(self):
            self.step, self.call, self.cur = 0, 0, None

        def __call__(self, low, high, size, device=None, **kw):
            if self.call == 0:
                self.cur = torch.from_numpy(O.counter_draws(SEED, ids, self.step, T, M, K))
            m, role = divmod(self.call, 2)
            out = self.cur[m, role]
            self.call += 1
            if self.call == 2 * M:
                self.call = 0
                self.step += 1
            return out

    model = EVAE(GaussianPrior(2), GaussianEncoder(make_encoder_net(50, 2)), GaussianDecoder(make_decoder_net(2, 50)),
                 num_decoders=10)
    model.load_state_dict(torch.load(REF / "experiment/model_seed12.pt", map_location="cpu"))
    model = model.to(tdt).eval()
    for prm in model.parameters():
        prm.requires_grad_(False)  # skips only the decoder weight gradients the reference never uses
    decoders = list(model.decoder)
    w, a, b, omega, weights = bench.synthetic_workload(bench.N_CURVES)
    assert weights.startswith("evae_seed12"), weights
    # the workload's decoders ARE the checkpoint's (tests/golden/evae_seed12_decoders.npz was cut from it)
    assert np.array_equal(w["W2"][3], model.decoder[3].decoder_net[2].weight.detach().float().numpy())
    sel = torch.from_numpy(ids)
    a, b, om = a[sel].to(tdt), b[sel].to(tdt), omega[sel].to(tdt)
    # the basis bench.py / the engine use (vlg_b200.construct_nullspace_basis is host torch, no GPU needed)
    import vlg_b200
    basis = vlg_b200.construct_nullspace_basis(N_POLY)[0].to(tdt)
    t_vals = torch.linspace(0, 1, T).to(tdt)   # fp32 linspace widened, like the kernels' t grid
    spl = ref_opt.GeodesicSplineBatch(a, b, basis, om.clone(), N_POLY)
    opt = torch.optim.Adam([spl.omega], lr=1e-3)
    ref_opt.torch.randint = CounterFeeder()
    energies = {}
    t0 = time.time()
    for step in range(STEPS):
        opt.zero_grad()
        energy = ref_opt.compute_energy_mc(spl, decoders, t_vals, M=M)
        endpoint_error = (spl(t_vals[-1:]) - b[None]) ** 2
        loss = energy + 1000 * endpoint_error.sum(dim=(0, 2))
        loss.sum().backward()
        opt.step()
        if step in KEEP:
            energies[step] = energy.detach().double().numpy().copy()
        if step % 25 == 0:
            print(f"[{dtype}] step {step} mean energy {energy.mean().item():.3f}  ({time.time() - t0:.0f} s)", flush=True)
    np.savez_compressed(OUT / f"_config3_part_{dtype}.npz", ids=ids, energy=np.stack([energies[s] for s in KEEP]),
                        omega=spl.omega.detach().double().numpy())
    print("done", time.time() - t0)


def merge():
    p64, p32 = (dict(np.load(OUT / f"_config3_part_{d}.npz")) for d in ("f64", "f32"))
    assert np.array_equal(p64["ids"], p32["ids"])
    l64, l32 = np.sqrt(p64["energy"][-1]), np.sqrt(p32["energy"][-1])
    gap = np.abs(l32 / l64 - 1)
    print(f"{len(l64)} curves; reference fp32-vs-fp64 final-length gap: median {np.median(gap):.2e}, max {gap.max():.2e}")
    np.savez_compressed(OUT / "config3_synth_1000.npz", seed=SEED, steps=STEPS, T=T, M=M, K=K, ids=p64["ids"],
                        energy_steps=np.array(KEEP), energy_f64=p64["energy"], energy_f32=p32["energy"].astype(np.float32),
                        omega_f64=p64["omega"], omega_f32=p32["omega"].astype(np.float32),
                        final_length_f64=l64, final_length_f32=l32.astype(np.float32))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ids", default="0:48")
    ap.add_argument("--extra", default=None, help="text file with extra global curve ids (e.g. the worst-diverging ones)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--merge", action="store_true")
    args = ap.parse_args()
    if args.merge:
        merge()
    else:
        run(parse_ids(args.ids, args.extra), args.dtype, args.threads)
