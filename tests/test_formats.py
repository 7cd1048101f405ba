"""CPU: the reference's file layouts -- writers produce what the readers (and the reference's
src/eval.py matrix code) expect."""
import json

import numpy as np
import torch


def _blob(n_rep=5):
    from vlg_b200 import formats
    reps = [{"index": 100 + 7 * i, "label": f"type{i}"} for i in range(n_rep)]
    pairs = [[reps[i]["index"], reps[j]["index"]] for i in range(n_rep) for j in range(i + 1, n_rep)]
    basis = torch.randn(16, 5)
    data = [formats.init_spline_dict(torch.randn(2), torch.randn(2), ia, ib, f"l{ia}", f"l{ib}", 4, basis, torch.randn(5, 2))
            for ia, ib in pairs]
    return reps, pairs, data


def test_init_and_opt_blob_round_trip(tmp_path):
    from vlg_b200 import formats
    reps, pairs, data = _blob()
    p = tmp_path / "splines_init_model_seed12" / "spline_batch_init_euclidean_5.pt"
    formats.save_init_blob(data, reps, pairs, p)
    blob = formats.load_spline_blob(p)
    assert set(blob) == {"spline_data", "representatives", "pairs"}
    assert set(blob["spline_data"][0]) == {"a", "b", "a_index", "b_index", "a_label", "b_label", "n_poly", "basis", "omega_init"}
    arr = formats.splines_to_arrays(blob["spline_data"])
    assert arr["a"].shape == (10, 2) and arr["omega"].shape == (10, 5, 2) and arr["basis"].shape == (16, 5)
    lengths = torch.arange(10).float() + 1
    formats.write_back_optimized(blob["spline_data"], arr["omega"] + 1, lengths, np.arange(10) * 0.5)
    q = tmp_path / "opt.pt"
    formats.save_opt_blob(blob["spline_data"], reps, pairs, "model_seed12", "euclidean", 5, 2, 1000, q)
    out = torch.load(q, weights_only=False)
    assert out["metadata"] == {"model_name": "model_seed12", "init_type": "euclidean", "pair_count": 5, "mc_samples": 2, "steps": 1000}
    d = out["spline_data"][3]
    assert isinstance(d["geodesic_length"], float) and d["geodesic_length"] == 4.0 and d["euclidean_distance"] == 1.5
    assert torch.equal(d["omega_optimized"], arr["omega"][3] + 1)
    mat, labels, skipped = formats.distance_matrix(out, "geodesic")
    assert mat.shape == (5, 5) and skipped == 0 and np.allclose(mat, mat.T) and (np.diag(mat) == 0).all()
    assert mat[0, 1] == 1.0 and mat[3, 4] == 10.0 and labels[2] == "type2"


def test_distance_matrix_missing_pairs_are_nan_and_errors():
    import pytest
    from vlg_b200 import formats
    reps, pairs, data = _blob(4)
    for i, d in enumerate(data):
        d["geodesic_length"] = float(i)
        d["euclidean_distance"] = 0.1
    blob = {"spline_data": data[:-1], "representatives": reps, "pairs": pairs}
    mat, _, _ = formats.distance_matrix(blob)
    assert np.isnan(mat[2, 3]) and np.isnan(mat[3, 2]) and mat[0, 1] == 0.0
    with pytest.raises(ValueError):
        formats.distance_matrix({"spline_data": data, "representatives": None})


def test_json_layouts(tmp_path):
    from vlg_b200 import formats
    formats.save_pairs([{"index": 3, "label": "x"}, {"index": 9, "label": "y"}], [(3, 9)], tmp_path / "pairs.json")
    reps, pairs = formats.load_pairs(tmp_path / "pairs.json")
    assert reps[1]["label"] == "y" and pairs == [[3, 9]]
    formats.save_distance_json(12, [1, 2], np.array([[0.0, 2.5], [2.5, 0.0]]), tmp_path / "d.json")
    assert json.loads((tmp_path / "d.json").read_text()) == {"seed": 12, "cluster_ids": [1, 2], "distance_matrix": [[0.0, 2.5], [2.5, 0.0]]}
    pay = formats.cov_payload({1: 0.26, 10: 0.09}, 0.27, {1: [0.2, 0.3], 10: [0.1, 0.08]}, [0.25, 0.29], [12, 123], [1, 10], 2)
    assert list(pay) == ["avg_cov_geodesic", "avg_cov_euclidean", "raw_cov_geodesic", "raw_cov_euclidean", "seeds",
                         "decoder_counts", "num_pairs"]
    assert pay["avg_cov_geodesic"] == {"1": 0.26, "10": 0.09}


def test_encoder_mean_matches_torch_modules():
    import torch.nn as nn
    from vlg_b200 import evae
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(50, 256), nn.SiLU(), nn.LayerNorm(256), nn.Linear(256, 128), nn.SiLU(), nn.LayerNorm(128),
                        nn.Linear(128, 4))
    sd = {f"encoder.encoder_net.{k}": v for k, v in net.state_dict().items()}
    x = torch.randn(7, 50)
    assert torch.allclose(evae.encoder_mean(sd, x), net(x)[:, :2], atol=1e-6)


def test_grid_graphs_without_gpu():
    """The euclidean kNN graph of the drop-in init script (host-only part)."""
    import importlib
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    mod = importlib.import_module("src.init_splines_ensemble")
    lat = torch.rand(500, 2) * 4 - 2
    grid, shape = mod.create_latent_grid_from_data(lat, n_points_per_axis=20)
    assert grid.shape == (400, 2) and shape == (20, 20)
    graph, tree = mod.build_grid_graph(grid, k=8)
    assert graph.shape == (400, 400) and graph.nnz == 400 * 8
    from scipy.sparse.csgraph import dijkstra
    _, pred = dijkstra(graph, indices=0, return_predecessors=True)
    path = mod.reconstruct_path(pred, 0, 399)
    assert path[0] == 0 and path[-1] == 399 and 15 <= len(path) <= 40
