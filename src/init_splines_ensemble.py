"""python -m src.init_splines_ensemble -- initial splines from shortest paths on a latent grid
(drop-in for the reference's src/init_splines_ensemble.py:98-228: same flags, same output file).

What changed underneath: the ensemble disagreement field runs as one CUDA kernel
(vlg_ensemble_std_norm), the kNN graph is built with one vectorised KD-tree query, Dijkstra is run
once per distinct source node instead of once per pair, and the spline fit is the batched
closed-form least-squares solve (vlg_fit_splines) that the reference's LBFGS loop iterates towards.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import dijkstra
from scipy.spatial import cKDTree

import vlg_b200
from vlg_b200 import evae, formats


def create_latent_grid_from_data(latents, n_points_per_axis=150, margin=0.1):
    """Regular grid over the latent bounding box, grown by `margin` (src/init_splines_ensemble.py:21-36)."""
    latents = torch.as_tensor(latents)
    z_min, z_max = latents.min(dim=0).values, latents.max(dim=0).values
    span = z_max - z_min
    z_min, z_max = z_min - margin * span, z_max + margin * span
    gx, gy = torch.meshgrid(torch.linspace(z_min[0], z_max[0], n_points_per_axis),
                            torch.linspace(z_min[1], z_max[1], n_points_per_axis), indexing="ij")
    return torch.stack([gx, gy], dim=-1).view(-1, 2), (n_points_per_axis, n_points_per_axis)


def _knn(grid_np, k):
    tree = cKDTree(grid_np)
    dists, idx = tree.query(grid_np, k=k + 1)
    return tree, dists[:, 1:], idx[:, 1:]


def build_grid_graph(grid, k=8):
    """Directed kNN graph with Euclidean edge lengths (src/init_splines_ensemble.py:72-82)."""
    grid_np = grid.cpu().numpy() if isinstance(grid, torch.Tensor) else grid
    tree, dists, idx = _knn(grid_np, k)
    n = len(grid_np)
    rows = np.repeat(np.arange(n), k)
    return csr_matrix((dists.ravel(), (rows, idx.ravel())), shape=(n, n)), tree


def entropy_graph_from_field(grid, field, k=8, eps=1e-8):
    """Symmetric kNN graph, edge weight = mean of the end points' min-max normalised disagreement
    (src/init_splines_ensemble.py:53-68).  `field` = || std_k f_k(grid) ||, one value per grid node."""
    ent = torch.as_tensor(field, dtype=torch.float32).cpu()
    ent = ((ent - ent.min()) / (ent.max() - ent.min() + eps)).numpy()
    grid_np = grid.cpu().numpy() if isinstance(grid, torch.Tensor) else grid
    tree, _, idx = _knn(grid_np, k)
    n = len(grid_np)
    rows = np.repeat(np.arange(n), k)
    cols = idx.ravel()
    # the reference averages two Python floats taken with .item() from the fp32 field
    w = 0.5 * (ent[rows].astype(np.float64) + ent[cols].astype(np.float64))
    r2, c2, w2 = np.concatenate([rows, cols]), np.concatenate([cols, rows]), np.concatenate([w, w])
    _, first = np.unique(r2.astype(np.int64) * n + c2, return_index=True)  # an edge may be found from both ends
    return csr_matrix((w2[first], (r2[first], c2[first])), shape=(n, n)), tree


def build_entropy_weighted_graph(grid, decoders, k=8, eps=1e-8):
    """src/init_splines_ensemble.py:39-68; the disagreement field itself is the CUDA kernel
    (vlg_ensemble_std_norm)."""
    return entropy_graph_from_field(grid, vlg_b200.ensemble_std_norm(decoders, grid.to(decoders.device)), k, eps)


def reconstruct_path(predecessors, start, end):
    path, i = [], end
    while i != start:
        if i == -9999:
            return []
        path.append(i)
        i = predecessors[i]
    path.append(start)
    return path[::-1]


def shortest_paths(latents, pairs, graph, tree):
    """KD-tree snap of both end points + Dijkstra (src/init_splines_ensemble.py:160-170), one Dijkstra per
    DISTINCT source node instead of one per pair.  -> list of (pair index, node path), in pair order; pairs
    whose end points snap to the same node, or that are not connected, are skipped like in the reference."""
    starts = tree.query(latents[[p[0] for p in pairs]])[1]
    ends = tree.query(latents[[p[1] for p in pairs]])[1]
    found = []
    for src in np.unique(starts):
        _, pred = dijkstra(graph, indices=int(src), return_predecessors=True)
        for i in np.nonzero(starts == src)[0]:
            if starts[i] == ends[i]:
                continue
            path = reconstruct_path(pred, int(src), int(ends[i]))
            if path:
                found.append((int(i), path))
    return sorted(found, key=lambda x: x[0])


def initial_splines(latents, pairs, representatives, graph, tree, grid, basis, n_poly, device):
    """Shortest path per pair -> least-squares spline fit; returns the spline_data list."""
    found = shortest_paths(latents, pairs, graph, tree)
    label = {r["index"]: r["label"] for r in representatives}
    if not found:
        return []
    a, b, omega = vlg_b200.fit_splines_to_paths([grid[path] for _, path in found], basis, n_poly, device)
    a, b, omega = a.cpu(), b.cpu(), omega.cpu()
    out = []
    for j, (i, _) in enumerate(found):
        ia, ib = pairs[i]
        out.append(formats.init_spline_dict(a[j], b[j], ia, ib, label[ia], label[ib], n_poly, basis, omega[j]))
    return out


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--model-path", type=str, required=True)
    parser.add_argument("--pairfile", type=str, required=True)
    parser.add_argument("--use-entropy", action="store_true")
    parser.add_argument("--save-dir", type=str, default=None)
    parser.add_argument("--plot-latents", action="store_true", help="accepted for compatibility; plotting is not part of this package")
    parser.add_argument("--n-poly", type=int, default=4)
    parser.add_argument("--data-path", type=str, default="data/tasic-pca50.npy")
    args = parser.parse_args()

    model_name = Path(args.model_path).stem
    save_dir = Path(args.save_dir) if args.save_dir else Path("experiment") / f"splines_init_{model_name}"
    save_dir.mkdir(parents=True, exist_ok=True)
    if not torch.cuda.is_available():
        raise RuntimeError("src.init_splines_ensemble needs a CUDA (B200) device")
    device = torch.device("cuda")
    print(f"[INFO] Using device: {device}")

    sd = evae.load_state_dict(args.model_path)
    assert evae.num_decoders(sd) == 10, "[ERROR] Expected 10 decoders in ensemble."
    decoders = vlg_b200.DecoderEnsemble.from_state_dict(sd, device)
    data = torch.from_numpy(np.load(args.data_path).astype(np.float32)).to(device)
    with torch.no_grad():
        latents = evae.encoder_mean(sd, data).cpu().numpy()

    representatives, pairs = formats.load_pairs(args.pairfile)
    grid, _ = create_latent_grid_from_data(latents, n_points_per_axis=200)
    if args.use_entropy:
        print("[INFO] Building entropy-weighted graph...")
        graph, tree = build_entropy_weighted_graph(grid, decoders)
    else:
        print("[INFO] Building Euclidean graph...")
        graph, tree = build_grid_graph(grid, k=8)
    basis, _ = vlg_b200.construct_nullspace_basis(n_poly=args.n_poly, device="cpu")
    spline_data = initial_splines(latents, pairs, representatives, graph, tree, grid, basis, args.n_poly, device)

    pairname = Path(args.pairfile).stem.replace("selected_pairs_", "")
    graph_type = "entropy" if args.use_entropy else "euclidean"
    save_path = save_dir / f"spline_batch_init_{graph_type}_{pairname}.pt"
    formats.save_init_blob(spline_data, representatives, pairs, save_path)
    print(f"[✓] Saved {len(spline_data)} initialized splines to: {save_path}")
