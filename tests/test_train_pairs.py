"""pair selection and eVAE training drop-ins against the reference's own outputs
(tests/golden/make_golden_pairs.py: the reference's `select_representatives` on its committed latents, and 30 Adam
steps of its EVAE).  CPU tests."""
import json
from itertools import combinations
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("n", [10, 50, 133])
def test_representatives_match_the_reference(n):
    from src import select_representative_pairs as S
    g = np.load(GOLD / "pairs_seed12.npz")
    reps = S.select_representatives(g["latents"], g["labels"], max_labels=n)
    assert [r["index"] for r in reps] == g[f"rep_index_{n}"].tolist()
    assert [r["label"] for r in reps] == g[f"rep_label_{n}"].tolist()
    assert len(list(combinations(reps, 2))) == n * (n - 1) // 2


def test_pairs_file_is_the_reference_file(tmp_path):
    """CLI, --model-type vae: byte-identical to the file the reference committed (src/artifacts/selected_pairs_10.json)."""
    from src import select_representative_pairs as S
    g = np.load(GOLD / "pairs_seed12.npz")
    np.save(tmp_path / "lat.npy", g["latents"])
    np.save(tmp_path / "lab.npy", g["labels"])
    out = S.main(["--model-type", "vae", "--vae-latent-path", str(tmp_path / "lat.npy"), "--label-path", str(tmp_path / "lab.npy"),
                  "--max-labels", "10", "--output-path", str(tmp_path / "pairs" / "selected_pairs_10.json")])
    ref = (GOLD / "ref_files" / "artifacts_selected_pairs_10.json").read_text()
    assert Path(out).read_text() == ref
    reps, pairs = S.load_pairs(out)
    assert len(reps) == 10 and len(pairs) == 45 and pairs == json.loads(ref)["pairs"]


def test_parameter_init_is_the_reference_stream():
    from src import train as T
    g = np.load(GOLD / "evae_train_30.npz")
    torch.manual_seed(7)
    sd = T.init_parameters(50, 2, 3)
    keys = [k[5:] for k in g.files if k.startswith("init/")]
    assert list(sd.keys()) == keys                      # same names, same order as the reference's state dict
    for k in keys:
        assert np.array_equal(sd[k].numpy(), g["init/" + k]), k


def test_thirty_training_steps_match_the_reference():
    from src import train as T
    g = np.load(GOLD / "evae_train_30.npz")
    data = torch.from_numpy(g["data"])
    torch.manual_seed(7)
    np.random.seed(7)
    sd = T.init_parameters(50, 2, 3)
    opt = T.make_optimizer(sd, 1e-3)
    losses = []
    for s in range(30):
        lo = (s * 64) % len(data)
        losses.append(T.train_step(sd, opt, data[lo:lo + 64], 3))
    rel = np.abs(np.array(losses) / g["losses"] - 1)
    print(f"\neVAE training, 30 steps: loss vs reference: max rel {rel.max():.2e}")
    assert rel.max() < 1e-5                             # same noise, same decoder draws, same Adam: rounding only
    worst = max(float(np.abs(sd[k].detach().numpy() - g["final/" + k]).max()) for k in sd)
    assert worst < 1e-4, worst


def test_checkpoint_feeds_the_decoder_ensemble(tmp_path):
    """A checkpoint written by the trainer has the layout every other entry point reads."""
    from src import train as T
    from vlg_b200 import evae
    rng = np.random.default_rng(0)
    np.save(tmp_path / "d.npy", rng.normal(size=(200, 50)).astype(np.float32))
    out = T.main(["--epochs", "1", "--num-decoders", "2", "--seed", "3", "--save-dir", str(tmp_path / "exp"),
                  "--data-path", str(tmp_path / "d.npy")])
    sd = evae.load_state_dict(out)
    assert evae.num_decoders(sd) == 2 and sd["decoder.1.decoder_net.4.weight"].shape == (50, 128)
    z = evae.encoder_mean(sd, torch.zeros(4, 50))
    assert z.shape == (4, 2) and torch.isfinite(z).all()
