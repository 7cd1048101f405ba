"""Golden for the spline-initialisation pipeline, generated with the REFERENCE's own
functions (src/init_splines_ensemble.py:18-95 and the per-pair loop at 160-205): latent grid, Euclidean kNN
graph, entropy-weighted graph (all-decoder forward + std), KD-tree snap, Dijkstra, LBFGS spline fit.

Inputs: 2000 real 2-D latents (src/artifacts/latents_VAE_ld2_ep100_bs64_lr1e-03_seed12.npy, every 11th row),
8 "representatives" -> 28 pairs, a 40 x 40 grid (the reference uses 200 x 200; the code path is the same),
the committed eVAE seed-12 decoders.  data/tasic-pca50.npy is a missing blob, so the encoder step
(latents = encoder means) is replaced by these latents; everything after it is the reference's code.

    python tests/golden/make_golden_init.py        # build container only (/root/reference)
"""
import sys
import types
from itertools import combinations
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
             "mpl_toolkits.axes_grid1"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["mpl_toolkits.axes_grid1"].make_axes_locatable = None   # plotting only (src/plotting.py:9-10)
sys.modules["matplotlib"].colormaps = None
sys.path.insert(0, str(REF))

import src.init_splines_ensemble as ref_init  # noqa: E402
from scipy.sparse.csgraph import dijkstra  # noqa: E402
from src.single_decoder.optimize_energy import construct_nullspace_basis  # noqa: E402
from src.single_decoder.optimize_energy_batched import GeodesicSplineBatch  # noqa: E402
from src.train import EVAE, GaussianDecoder, GaussianEncoder, GaussianPrior, make_decoder_net, make_encoder_net  # noqa: E402

N_GRID, N_POLY = 40, 4

if __name__ == "__main__":
    torch.manual_seed(0)
    lat = np.load(REF / "src/artifacts/latents_VAE_ld2_ep100_bs64_lr1e-03_seed12.npy")[::11][:2000].astype(np.float32)
    rng = np.random.default_rng(3)
    reps = sorted(rng.choice(len(lat), 8, replace=False).tolist())
    pairs = [list(p) for p in combinations(reps, 2)]
    model = EVAE(GaussianPrior(2), GaussianEncoder(make_encoder_net(50, 2)), GaussianDecoder(make_decoder_net(2, 50)),
                 num_decoders=10)
    model.load_state_dict(torch.load(REF / "experiment/model_seed12.pt", map_location="cpu"))
    model.eval()
    grid, _ = ref_init.create_latent_grid_from_data(lat, n_points_per_axis=N_GRID)
    basis, _ = construct_nullspace_basis(n_poly=N_POLY, device="cpu")
    out = {"latents": lat, "pairs": np.array(pairs), "grid": grid.numpy(), "basis": basis.numpy(), "n_grid": N_GRID,
           "n_poly": N_POLY}
    # the disagreement field itself (a-11), for the CPU test of the graph builder
    with torch.no_grad():
        outs = torch.stack([d(grid).mean for d in model.decoder])
        field = outs.std(dim=0).norm(dim=1)
    out["std_field"] = field.numpy()
    for kind in ("euclidean", "entropy"):
        if kind == "entropy":
            graph, tree = ref_init.build_entropy_weighted_graph(grid, model.decoder)
        else:
            graph, tree = ref_init.build_grid_graph(grid, k=8)
        graph.sort_indices()
        out[f"{kind}_indptr"], out[f"{kind}_indices"], out[f"{kind}_data"] = graph.indptr, graph.indices, graph.data
        path_cat, path_off, omegas, ab, kept = [], [0], [], [], []
        for pi, (idx_a, idx_b) in enumerate(pairs):       # src/init_splines_ensemble.py:160-205
            start_idx = tree.query(lat[idx_a])[1]
            end_idx = tree.query(lat[idx_b])[1]
            if start_idx == end_idx:
                continue
            _, preds = dijkstra(graph, indices=start_idx, return_predecessors=True)
            path = ref_init.reconstruct_path(preds, start_idx, end_idx)
            if not path:
                continue
            target = grid[path]
            a, b = target[0], target[-1]
            spline = GeodesicSplineBatch(a.unsqueeze(0), b.unsqueeze(0), basis,
                                         omega=torch.zeros((1, basis.shape[1], a.shape[0])), n_poly=N_POLY)
            t_vals = torch.linspace(0, 1, len(target))
            optimizer = torch.optim.LBFGS([spline.omega], max_iter=50)

            def closure():
                optimizer.zero_grad()
                loss = torch.nn.functional.mse_loss(spline(t_vals).squeeze(1), target)
                loss.backward()
                return loss

            optimizer.step(closure)
            kept.append(pi)
            path_cat += [int(x) for x in path]
            path_off.append(len(path_cat))
            omegas.append(spline.omega.detach().squeeze(0).numpy().copy())
            ab.append(np.stack([a.numpy(), b.numpy()]))
        out[f"{kind}_kept"] = np.array(kept)
        out[f"{kind}_paths"] = np.array(path_cat, dtype=np.int32)
        out[f"{kind}_path_off"] = np.array(path_off, dtype=np.int32)
        out[f"{kind}_omega_init"] = np.stack(omegas)
        out[f"{kind}_ab"] = np.stack(ab)
        print(kind, "pairs kept", len(kept), "mean path length", np.diff(path_off).mean(), "nnz", graph.nnz)
    np.savez_compressed(OUT / "init_pipeline.npz", **out)
    print("wrote", OUT / "init_pipeline.npz")
