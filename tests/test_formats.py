"""CPU: the reference's file layouts -- writers produce what the readers (and the reference's
src/eval.py matrix code) expect."""
import json
from pathlib import Path

import numpy as np
import torch

import vlg_b200  # noqa: F401  (import shim for the hyphenated package directory)
from vlg_b200 import formats


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


# ---------------------------------------------------------------------------------------------
# Against files the REFERENCE itself wrote (tests/golden/ref_files/, copied verbatim by
# tests/golden/make_golden_formats.py) and the matrices its own plot_geodesic_matrix builds from them.
# ---------------------------------------------------------------------------------------------
REF_FILES = Path(__file__).resolve().parent / "golden" / "ref_files"


def _same_value(x, y):
    if isinstance(x, torch.Tensor):
        return isinstance(y, torch.Tensor) and x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y)
    return type(x) is type(y) and x == y


def _same_blob(b1, b2):
    assert list(b1.keys()) == list(b2.keys())
    assert b1["representatives"] == b2["representatives"] and b1["pairs"] == b2["pairs"]
    assert b1.get("metadata") == b2.get("metadata")
    assert len(b1["spline_data"]) == len(b2["spline_data"])
    for d1, d2 in zip(b1["spline_data"], b2["spline_data"]):
        assert list(d1.keys()) == list(d2.keys())
        for k in d1:
            assert _same_value(d1[k], d2[k]), k


def test_reference_opt_blob_loads_and_round_trips(tmp_path):
    ref = formats.load_spline_blob(REF_FILES / "spline_batch_opt_euclidean_10.pt")
    assert len(ref["spline_data"]) == 45 and ref["metadata"]["steps"] == 1000
    arr = formats.splines_to_arrays(ref["spline_data"])
    assert arr["a"].shape == (45, 2) and arr["omega"].shape == (45, 5, 2) and arr["basis"].shape == (16, 5) and arr["n_poly"] == 4
    # rebuild the file from structure-of-arrays with OUR writers, exactly as src/optimize.py (ours) does
    init = formats.load_spline_blob(REF_FILES / "spline_batch_init_euclidean_10.pt")
    spline_data = init["spline_data"]
    om = torch.stack([d["omega_optimized"] for d in ref["spline_data"]])
    gl = torch.tensor([d["geodesic_length"] for d in ref["spline_data"]], dtype=torch.float64)
    eu = [d["euclidean_distance"] for d in ref["spline_data"]]
    formats.write_back_optimized(spline_data, om, gl, eu)
    md = ref["metadata"]
    out = tmp_path / "spline_batch_opt_euclidean_10.pt"
    formats.save_opt_blob(spline_data, init["representatives"], init["pairs"], md["model_name"], md["init_type"],
                          md["pair_count"], md["mc_samples"], md["steps"], out)
    _same_blob(formats.load_spline_blob(out), ref)      # same keys, key order, types, dtypes, values


def test_reference_init_blob_round_trips(tmp_path):
    ref = formats.load_spline_blob(REF_FILES / "spline_batch_init_euclidean_10.pt")
    sd = [formats.init_spline_dict(d["a"], d["b"], d["a_index"], d["b_index"], d["a_label"], d["b_label"], d["n_poly"],
                                   d["basis"], d["omega_init"]) for d in ref["spline_data"]]
    formats.save_init_blob(sd, ref["representatives"], ref["pairs"], tmp_path / "i.pt")
    _same_blob(formats.load_spline_blob(tmp_path / "i.pt"), ref)


def test_distance_matrix_equals_the_references_plot_matrix():
    ref = formats.load_spline_blob(REF_FILES / "spline_batch_opt_euclidean_10.pt")
    m = np.load(REF_FILES / "matrices.npz")
    for len_type in ("geodesic", "euclidean_dist"):
        mat, labels, skipped = formats.distance_matrix(ref, len_type)
        assert skipped == 0 and mat.shape == (10, 10)
        assert np.array_equal(mat, m[len_type]) and [str(x) for x in labels] == list(m["labels"])
    blob2 = {"spline_data": [dict(d) for d in ref["spline_data"][1:]], "representatives": ref["representatives"]}
    blob2["spline_data"][0]["a_index"] = -7
    mat, _, skipped = formats.distance_matrix(blob2, "geodesic")
    assert skipped == 1 and np.array_equal(mat, m["geodesic_missing"], equal_nan=True) and np.isnan(mat).sum() == 4


def test_single_decoder_list_matches_reference_file():
    ref = torch.load(REF_FILES / "spline_batch_optimized_batched_seed12.pt", map_location="cpu", weights_only=False)
    arr = formats.splines_to_arrays(ref)
    recs = formats.single_decoder_records(arr["a"], arr["b"], [d["cluster_pair"] for d in ref], arr["n_poly"], arr["basis"],
                                          torch.stack([d["omega_init"] for d in ref]),
                                          torch.stack([d["omega_optimized"] for d in ref]),
                                          [d["length_geodesic"] for d in ref])
    assert isinstance(recs, list) and len(recs) == len(ref)
    for r, d in zip(recs, ref):
        assert list(r.keys()) == list(d.keys())
        for k in d:
            if k == "length_euclidean":      # recomputed: ||a - b|| (optimize_energy_batched.py:121)
                assert abs(r[k] - d[k]) <= 1e-6 * max(1.0, abs(d[k]))
            else:
                assert _same_value(r[k], d[k]), k


def test_json_files_are_byte_identical_to_the_references(tmp_path):
    # pairs JSON (src/select_representative_pairs.py:37-44)
    reps, pairs = formats.load_pairs(REF_FILES / "selected_pairs_10.json")
    formats.save_pairs(reps, pairs, tmp_path / "p.json")
    assert (tmp_path / "p.json").read_text() == (REF_FILES / "selected_pairs_10.json").read_text()
    # CoV JSON (src/eval.py:139-157): rebuilt from its own raw values through cov_payload
    ref = json.loads((REF_FILES / "cov_values_alldec_alldec.json").read_text())
    counts = ref["decoder_counts"]
    raw = {k: ref["raw_cov_geodesic"][str(k)] for k in counts}
    payload = formats.cov_payload({k: np.mean(raw[k]) for k in counts}, np.mean(ref["raw_cov_euclidean"]), raw,
                                  ref["raw_cov_euclidean"], ref["seeds"], counts, ref["num_pairs"])
    (tmp_path / "c.json").write_text(json.dumps(payload, indent=2))
    assert (tmp_path / "c.json").read_text() == (REF_FILES / "cov_values_alldec_alldec.json").read_text()
    # distance-matrix JSON (src/single_decoder/density_batched.py:135-142)
    d = json.loads((REF_FILES / "geodesic_distances_seed12_p12.json").read_text())
    formats.save_distance_json(d["seed"], d["cluster_ids"], np.array(d["distance_matrix"]), tmp_path / "d.json")
    assert (tmp_path / "d.json").read_text() == (REF_FILES / "geodesic_distances_seed12_p12.json").read_text()


def test_single_decoder_distance_matrix_from_reference_records(tmp_path):
    """src/single_decoder/density_batched.py (ours) on the reference's committed 45-record single-decoder file:
    points numbered by first appearance, labels from cluster_pair, symmetric, zero diagonal, every entry a
    committed length_geodesic; the JSON carries exactly the reference's keys."""
    import shutil
    from src.single_decoder import density_batched as db
    recs = torch.load(REF_FILES / "spline_batch_optimized_batched_seed12.pt", map_location="cpu", weights_only=False)
    ids, mat = db.distance_matrix_from_records(recs)
    assert len(ids) == 10 and mat.shape == (10, 10) and not np.isnan(mat).any()
    assert np.array_equal(mat, mat.T) and (np.diag(mat) == 0).all()
    assert ids[0] == recs[0]["cluster_pair"][0] and ids[1] == recs[0]["cluster_pair"][1]
    assert mat[0, 1] == recs[0]["length_geodesic"]
    assert sorted(mat[np.triu_indices(10, 1)].tolist()) == sorted(r["length_geodesic"] for r in recs)
    art = tmp_path / "artifacts"
    art.mkdir()
    shutil.copyfile(REF_FILES / "spline_batch_optimized_batched_seed12.pt", art / "spline_batch_optimized_batched_seed12_p10.pt")
    out = db.main(12, "src/artifacts/selected_pairs_10.json", artifact_dir=str(art))
    d = json.loads(Path(out).read_text())
    assert list(d.keys()) == ["seed", "cluster_ids", "distance_matrix"] and d["seed"] == 12 and d["cluster_ids"] == ids
    assert np.array_equal(np.array(d["distance_matrix"]), mat)
