"""Known answers for BASELINE config 2 (single-decoder VAE): 64 curves of the reference's committed input
src/artifacts/spline_batch_seed123.pt with the 500-step results the reference itself committed in
src/artifacts/spline_batch_optimized_batched_seed123.pt (length_geodesic, omega_optimized), plus a re-run of
the reference's loop (src/single_decoder/optimize_energy_batched.py:88-106) in fp32 on THIS machine -- the
reference-vs-reference spread across machines is the yardstick for the 3e-3 tolerance.

    python tests/golden/make_golden_single64.py        # build container only (/root/reference)
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
             "mpl_toolkits.axes_grid1"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, str(REF))
import src.single_decoder.optimize_energy_batched as ref_sb  # noqa: E402
from src.single_decoder.vae import VAE  # noqa: E402

SEED, COUNT, STEPS, T = 123, 64, 500, 2000

if __name__ == "__main__":
    torch.set_num_threads(4)
    vae = VAE(input_dim=50, latent_dim=2)
    vae.load_state_dict(torch.load(REF / f"src/artifacts/vae_best_seed{SEED}.pth", map_location="cpu"))
    vae.eval()
    init_all = torch.load(REF / f"src/artifacts/spline_batch_seed{SEED}.pt", map_location="cpu", weights_only=False)["spline_data"]
    done_all = torch.load(REF / f"src/artifacts/spline_batch_optimized_batched_seed{SEED}.pt", map_location="cpu", weights_only=False)
    idx = np.linspace(0, len(init_all) - 1, COUNT).round().astype(int)
    init, done = [init_all[i] for i in idx], [done_all[i] for i in idx]
    for d0, d1 in zip(init, done):
        assert torch.equal(d0["a"], d1["a"]) and torch.equal(d0["b"], d1["b"])
    a = torch.stack([d["a"] for d in init]); b = torch.stack([d["b"] for d in init])
    om = torch.stack([d["omega_init"] for d in init])
    basis = done[0]["basis"]           # the basis the committed omegas refer to
    assert torch.equal(basis, init[0]["basis"]), "init and optimised files were made with the same basis"
    n_poly = int(init[0]["n_poly"])
    f32 = lambda x: x.detach().float().numpy()
    out = dict(idx=idx, a=f32(a), b=f32(b), omega_init=f32(om), basis=f32(basis), n_poly=n_poly, T=T, steps=STEPS,
               labels=np.array([f"{d['a_label']}|{d['b_label']}" for d in init]),
               committed_length_geodesic=np.array([d["length_geodesic"] for d in done]),
               committed_length_euclidean=np.array([d["length_euclidean"] for d in done]),
               committed_omega_optimized=f32(torch.stack([d["omega_optimized"] for d in done])))
    t_vals = torch.linspace(0, 1, T)
    model = ref_sb.GeodesicSplineBatch(a, b, basis, om.clone(), n_poly)
    opt = torch.optim.Adam([model.omega], lr=1e-3)
    for step in range(STEPS):
        opt.zero_grad()
        energy = ref_sb.compute_energy(model, vae.decoder, t_vals)
        endpoint_error = (model(t_vals[-1:]) - b[None]) ** 2
        loss = energy + 1000 * endpoint_error.sum(dim=(0, 2))
        loss.sum().backward()
        opt.step()
        if step % 100 == 0:
            print("step", step, float(energy.mean()), flush=True)
    out["rerun_length_f32"] = f32(ref_sb.compute_geodesic_lengths(model, vae.decoder, t_vals))
    out["rerun_omega_f32"] = f32(model.omega)
    gap = np.abs(out["rerun_length_f32"] / out["committed_length_geodesic"] - 1)
    print(f"reference re-run vs committed lengths: median {np.median(gap):.2e}, max {gap.max():.2e}")
    np.savez_compressed(OUT / "single_seed123_64.npz", **out)
