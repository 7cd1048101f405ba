"""GPU: the drop-in entry points end to end on a synthetic experiment directory
(python -m src.init_splines_ensemble -> python -m src.optimize -> python -m src.eval)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _fake_experiment(tmp: Path, seeds=(12, 123)):
    import torch.nn as nn
    (tmp / "experiment" / "pairs").mkdir(parents=True)
    (tmp / "data").mkdir()
    rng = np.random.default_rng(0)
    np.save(tmp / "data" / "tasic-pca50.npy", rng.normal(size=(400, 50)).astype(np.float32) * 3)
    for seed in seeds:
        torch.manual_seed(seed)
        sd = {"prior.mean": torch.zeros(2), "prior.std": torch.ones(2)}
        enc = nn.Sequential(nn.Linear(50, 256), nn.SiLU(), nn.LayerNorm(256), nn.Linear(256, 128), nn.SiLU(),
                            nn.LayerNorm(128), nn.Linear(128, 4))
        sd.update({f"encoder.encoder_net.{k}": v for k, v in enc.state_dict().items()})
        for i in range(10):
            dec = nn.Sequential(nn.Linear(2, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 50))
            sd.update({f"decoder.{i}.decoder_net.{k}": v for k, v in dec.state_dict().items()})
        torch.save(sd, tmp / "experiment" / f"model_seed{seed}.pt")
    reps = [{"index": int(i), "label": f"c{j}"} for j, i in enumerate([3, 50, 111, 200, 377])]
    pairs = [[reps[i]["index"], reps[j]["index"]] for i in range(5) for j in range(i + 1, 5)]
    with open(tmp / "experiment" / "pairs" / "selected_pairs_5.json", "w") as f:
        json.dump({"representatives": reps, "pairs": pairs}, f)


def _run(tmp, *args):
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    res = subprocess.run([sys.executable, "-m", *args], cwd=tmp, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


def test_init_optimize_eval_pipeline(tmp_path, built_lib):
    _fake_experiment(tmp_path)
    for flag in ([], ["--use-entropy"]):
        _run(tmp_path, "src.init_splines_ensemble", "--model-path", "experiment/model_seed12.pt", "--pairfile",
             "experiment/pairs/selected_pairs_5.json", *flag)
    init = torch.load(tmp_path / "experiment/splines_init_model_seed12/spline_batch_init_entropy_5.pt", weights_only=False)
    assert len(init["spline_data"]) >= 8 and init["spline_data"][0]["omega_init"].shape == (5, 2)
    assert init["spline_data"][0]["basis"].shape == (16, 5)

    out = _run(tmp_path, "src.optimize", "--model-path", "experiment/model_seed12.pt", "--init-type", "euclidean",
               "--pair-count", "5", "--steps", "60", "--batch-size", "200")
    assert "[Step 0] Mean Energy:" in out and "[Step 50] Mean Energy:" in out
    opt = torch.load(tmp_path / "experiment/splines_opt_model_seed12/spline_batch_opt_euclidean_5.pt", weights_only=False)
    assert opt["metadata"] == {"model_name": "model_seed12", "init_type": "euclidean", "pair_count": 5, "mc_samples": 2, "steps": 60}
    d = opt["spline_data"][0]
    assert d["omega_optimized"].shape == (5, 2) and d["geodesic_length"] > 0 and d["euclidean_distance"] > 0
    assert not torch.equal(d["omega_optimized"], d["omega_init"])
    # the energy went down during the optimisation
    e0 = float(out.split("[Step 0] Mean Energy:")[1].split()[0])
    e50 = float(out.split("[Step 50] Mean Energy:")[1].split()[0])
    assert e50 < e0

    _run(tmp_path, "src.eval", "--mode", "matrix", "--len-type", "geodesic", "--init-type", "euclidean", "--pair-count", "5",
         "--seed", "12")
    mat = np.load(tmp_path / "experiment/plots/geodesic_matrix_seed12_euclidean_5.npy")
    assert mat.shape == (5, 5) and np.allclose(np.nan_to_num(mat), np.nan_to_num(mat).T) and (np.diag(mat) == 0).all()


def test_cov_driver_small(tmp_path, built_lib):
    _fake_experiment(tmp_path)
    sys.path.insert(0, str(ROOT))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        import importlib
        ev = importlib.import_module("src.eval")
        pay = ev.run_cov_analysis(seeds=[12, 123], decoder_counts=[1, 2, 10], pairfile="experiment/pairs/selected_pairs_5.json",
                                  model_dir="experiment", data_path="data/tasic-pca50.npy",
                                  output_plot="experiment/plots/cov_plot_5_alldec.png", steps=4)
    finally:
        os.chdir(cwd)
    assert pay["num_pairs"] == 10 and set(pay["avg_cov_geodesic"]) == {"1", "2", "10"}
    assert all(0 <= v < 2 for v in pay["raw_cov_geodesic"]["10"])
    assert (tmp_path / "experiment/plots/cov_values_alldec_alldec.json").exists()


def _six_seed_state_dicts():
    g = np.load(ROOT / "tests" / "golden" / "evae_six_seeds.npz")
    seeds = [int(s) for s in g["seeds"]]
    sds = []
    for s in seeds:
        pre = f"s{s}/"
        sds.append({k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)})
    return seeds, sds


def test_cov_one_launch_per_k_equals_per_seed_launches(built_lib):
    """decoder_base: all seeds' weight sets in one launch give bit-identical results to one launch per seed
    (same curve ids, hence the same draws)."""
    import vlg_b200
    from vlg_b200 import evae
    seeds, sds = _six_seed_state_dicts()
    dev, N, S, k = "cuda", 7, 3, 4
    g = torch.Generator().manual_seed(5)
    za = torch.rand(S, N, 2, generator=g) * 4 - 2
    zb = torch.rand(S, N, 2, generator=g) * 4 - 2
    basis, _ = vlg_b200.construct_nullspace_basis(4, dev)
    t = torch.linspace(0, 1, 2000, device=dev)
    packed = vlg_b200.DecoderEnsemble.from_state_dicts(sds[:S], dev)
    assert packed.K == 10 * S
    base = (torch.arange(S, dtype=torch.int32) * 10).repeat_interleave(N)
    for prec in ("fp32", "f16x3", "f16"):
        m = vlg_b200.GeodesicSplineBatch(za.reshape(-1, 2).to(dev), zb.reshape(-1, 2).to(dev), basis,
                                         torch.zeros(S * N, 5, 2, device=dev), 4)
        e_all = vlg_b200.optimize_splines(m, packed, t, 3, M=2, seed=1, precision=prec, decoder_base=base, k_active=k)
        for si in range(S):
            one = vlg_b200.DecoderEnsemble.from_state_dict(sds[si], dev)
            m1 = vlg_b200.GeodesicSplineBatch(za[si].to(dev), zb[si].to(dev), basis, torch.zeros(N, 5, 2, device=dev), 4)
            e1 = vlg_b200.optimize_splines(m1, one[:k], t, 3, M=2, seed=1, curve_id0=si * N, precision=prec)
            assert torch.equal(e1, e_all[si * N:(si + 1) * N]) and torch.equal(m1.omega, m.omega[si * N:(si + 1) * N])
    with pytest.raises(vlg_b200.VlgError):   # a base that runs past the packed buffer
        vlg_b200.optimize_splines(m, packed, t, 1, M=2, decoder_base=base + 25, k_active=k)


def test_cov_shrinks_with_more_decoders_on_the_committed_checkpoints(built_lib):
    """The reference's headline finding (experiment/plots/cov_values_alldec_alldec.json: geodesic CoV 0.262 at
    k = 1 -> 0.089 at k = 10), on the six committed checkpoints.  The cells' data rows are a missing blob, so
    stand-in cells are generated with the seed-12 ensemble (mean decoder output at random latent points) and
    encoded by every seed's own encoder, as src/eval.py:102-104 does.  Statistical: shorter runs, 16 pairs."""
    import importlib
    import vlg_b200
    from vlg_b200 import evae
    sys.path.insert(0, str(ROOT))
    ev = importlib.import_module("src.eval")
    seeds, sds = _six_seed_state_dicts()
    dev = "cuda"
    g = torch.Generator().manual_seed(11)
    z = torch.rand(32, 2, generator=g) * 4 - 2
    with torch.no_grad():
        sd = sds[0]
        xs = []
        for i in range(10):
            w = lambda l, n: sd[f"decoder.{i}.decoder_net.{l}.{n}"]
            h = torch.relu(z @ w(0, "weight").T + w(0, "bias"))
            h = torch.relu(h @ w(2, "weight").T + w(2, "bias"))
            xs.append(h @ w(4, "weight").T + w(4, "bias"))
        cells = torch.stack(xs).mean(0)
        za = torch.stack([evae.encoder_mean(s_, cells[:16]) for s_ in sds])
        zb = torch.stack([evae.encoder_mean(s_, cells[16:]) for s_ in sds])
    geo = ev.cov_lengths(sds, za, zb, [1, 3, 10], steps=100, precision=None, draw_seed=0, device=dev)
    cov = {k: float(np.mean([ev.compute_cov(geo[k][:, i]) for i in range(16)])) for k in (1, 3, 10)}
    print("mean CoV over 16 pairs:", cov)
    assert all(np.isfinite(v) and v > 0 for v in cov.values())
    assert cov[10] < cov[1]


@pytest.mark.parametrize("prec", ["fp32", "f16x3", "f16x3f"])
def test_single_decoder_dropin_reproduces_committed_lengths(tmp_path, built_lib, prec):
    """BASELINE config 2 through the drop-in CLI module: 64 curves of the reference's committed input
    (src/artifacts/spline_batch_seed123.pt), 500 steps, against the length_geodesic the reference committed in
    spline_batch_optimized_batched_seed123.pt: <= 3e-3 relative (the reference's own re-run on
    another machine only reproduces them to ~1e-3; that spread is printed)."""
    g = np.load(ROOT / "tests" / "golden" / "single_seed123_64.npz")
    w = np.load(ROOT / "tests" / "golden" / "single_seed123.npz")
    art = tmp_path / "artifacts"
    art.mkdir()
    # the VAE checkpoint layout of src/single_decoder/vae.py:29-42: last layer emits mean || log_std (100 rows)
    sd = {"decoder.decoder_net.0.weight": torch.from_numpy(w["W1"][0]), "decoder.decoder_net.0.bias": torch.from_numpy(w["b1"][0]),
          "decoder.decoder_net.2.weight": torch.from_numpy(w["W2"][0]), "decoder.decoder_net.2.bias": torch.from_numpy(w["b2"][0]),
          "decoder.decoder_net.4.weight": torch.cat([torch.from_numpy(w["W3"][0]), torch.zeros(50, 128)]),
          "decoder.decoder_net.4.bias": torch.cat([torch.from_numpy(w["b3"][0]), torch.zeros(50)])}
    torch.save(sd, art / "vae_best_seed123.pth")
    basis = torch.from_numpy(g["basis"])
    spline_data = []
    for i in range(len(g["idx"])):
        la, lb = str(g["labels"][i]).split("|")
        spline_data.append({"a": torch.from_numpy(g["a"][i]), "b": torch.from_numpy(g["b"][i]), "a_label": la, "b_label": lb,
                            "n_poly": int(g["n_poly"]), "basis": basis, "omega_init": torch.from_numpy(g["omega_init"][i])})
    torch.save({"spline_data": spline_data}, art / "spline_batch_seed123_p64.pt")
    sys.path.insert(0, str(ROOT))
    import importlib
    mod = importlib.import_module("src.single_decoder.optimize_energy_batched")
    out_path = mod.main(123, "src/artifacts/selected_pairs_64.json", steps=int(g["steps"]), precision=prec, artifact_dir=str(art))
    recs = torch.load(out_path, map_location="cpu", weights_only=False)
    assert isinstance(recs, list) and len(recs) == 64
    assert list(recs[0].keys()) == ["a", "b", "cluster_pair", "n_poly", "basis", "omega_init", "omega_optimized",
                                    "length_geodesic", "length_euclidean"]
    got = np.array([r["length_geodesic"] for r in recs])
    ref = g["committed_length_geodesic"]
    err = np.abs(got / ref - 1)
    spread = np.abs(g["rerun_length_f32"] / ref - 1)
    print(f"\nsingle decoder [{prec}], 64 curves x 500 steps: length vs committed: median {np.median(err):.2e}, max {err.max():.2e}; "
          f"reference CPU re-run vs committed: median {np.median(spread):.2e}, max {spread.max():.2e}")
    # 500 Adam steps amplify rounding-level differences on a few curves: the reference's own CPU re-run misses its
    # committed lengths by up to 4.7e-3 (2 of 64 beyond 3e-3).  Same bar for the engine: the typical curve far inside
    # 3e-3, no more outliers than the reference itself has (+2), the worst one within twice the reference's worst.
    assert np.median(err) < 5e-4 and np.quantile(err, 0.9) < 3e-3
    assert (err > 3e-3).sum() <= (spread > 3e-3).sum() + 2 and err.max() <= 2 * spread.max(), (err.max(), spread.max())
    assert np.abs(np.array([r["length_euclidean"] for r in recs]) / g["committed_length_euclidean"] - 1).max() < 1e-5
    assert not torch.equal(recs[0]["omega_init"], recs[0]["omega_optimized"])   # the reference aliases them
    # and the matrix / JSON step on top of it
    db = importlib.import_module("src.single_decoder.density_batched")
    d = json.loads(Path(db.main(123, "src/artifacts/selected_pairs_64.json", artifact_dir=str(art))).read_text())
    assert d["seed"] == 123 and len(d["cluster_ids"]) == len(d["distance_matrix"])
