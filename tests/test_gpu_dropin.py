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
