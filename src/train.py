"""Drop-in for `python -m src.train` (reference: src/train.py:126-176): trains the 10-decoder ensemble VAE on
tasic-pca50-shaped data and writes `model_seed{seed}.pt` with the reference's state-dict keys, so that every other
entry point (`src.optimize`, `src.eval`, `src.init_splines_ensemble`, `vlg_b200.DecoderEnsemble.from_checkpoint`)
reads it unchanged.

SURVEY §8 row f-4: one-off and tiny (47 k encoder + K x 23 k decoder parameters), NOT the hot path -- it stays plain
PyTorch on whatever `--device` names (the reference's default is "cpu"; "cuda" uses the fused Adam).  The model is
held functionally -- one flat dict of tensors under the reference's key names, the ELBO written out in closed form
instead of `torch.distributions` objects (the reference's argument validation alone was 15 % of its step) -- and
reproduces the reference bit for bit on CPU: same initialisation stream (`nn.Linear` defaults in the reference's
construction order, every decoder a copy of ONE initial decoder, src/train.py:52), same reparameterisation noise
(`empty().normal_()`), same decoder draw (`np.random.choice`), same Adam.  `tests/test_train_pairs.py` checks 30 steps
against the reference's own EVAE (`tests/golden/evae_train_30.npz`).
"""
from __future__ import annotations

import argparse
import math
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

DECODER_SIGMA = 5.0          # GaussianDecoder: Normal(mean, 5)  (src/train.py:46)
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
ENCODER_LAYERS = ((0, 256), (3, 128))   # Linear index -> width; LayerNorm sits two entries later (src/train.py:71-78)


def init_parameters(input_dim: int, latent_dim: int, num_decoders: int) -> "OrderedDict[str, torch.Tensor]":
    """Parameters and buffers under the reference's state-dict keys, drawn from the global torch RNG exactly as
    the reference's constructors do (call `torch.manual_seed` first)."""
    def linear(i, o):
        m = torch.nn.Linear(i, o)      # library default init = the reference's
        return m.weight.detach().clone(), m.bias.detach().clone()

    enc = OrderedDict()
    width = input_dim
    for idx, out in ENCODER_LAYERS:
        enc[f"{idx}.weight"], enc[f"{idx}.bias"] = linear(width, out)
        enc[f"{idx + 2}.weight"], enc[f"{idx + 2}.bias"] = torch.ones(out), torch.zeros(out)
        width = out
    enc["6.weight"], enc["6.bias"] = linear(width, 2 * latent_dim)
    dec = OrderedDict()
    for idx, (i, o) in zip((0, 2, 4), ((latent_dim, 128), (128, 128), (128, input_dim))):
        dec[f"{idx}.weight"], dec[f"{idx}.bias"] = linear(i, o)
    sd = OrderedDict()
    sd["prior.mean"], sd["prior.std"] = torch.zeros(latent_dim), torch.ones(latent_dim)
    for k, v in enc.items():
        sd[f"encoder.encoder_net.{k}"] = v
    for j in range(num_decoders):
        for k, v in dec.items():
            sd[f"decoder.{j}.decoder_net.{k}"] = v.clone()
    return sd


def trainable(sd):
    return [v for k, v in sd.items() if not k.startswith("prior.")]


def encode(sd, x):
    """-> (mean, log_std) of q(z|x)  (src/train.py:30-34,71-78)."""
    g = lambda k: sd[f"encoder.encoder_net.{k}"]
    h = x
    for idx, _ in ENCODER_LAYERS:
        h = F.silu(F.linear(h, g(f"{idx}.weight"), g(f"{idx}.bias")))
        h = F.layer_norm(h, (h.shape[-1],), g(f"{idx + 2}.weight"), g(f"{idx + 2}.bias"))
    return F.linear(h, g("6.weight"), g("6.bias")).chunk(2, dim=-1)


def decode(sd, j: int, z):
    """Mean of decoder j (src/train.py:42-43,80-85)."""
    g = lambda k: sd[f"decoder.{j}.decoder_net.{k}"]
    h = F.relu(F.linear(z, g("0.weight"), g("0.bias")))
    h = F.relu(F.linear(h, g("2.weight"), g("2.bias")))
    return F.linear(h, g("4.weight"), g("4.bias"))


def neg_elbo(sd, x, j: int, eps, beta: float = 1.0):
    """-mean_b [ log p(x|z) - beta (log q(z|x) - log p(z)) ] with z = mean + std * eps and decoder j
    (src/train.py:55-64); the three Gaussian log-densities written out."""
    mean, log_std = encode(sd, x)
    std = torch.exp(log_std)
    z = mean + std * eps
    xm = decode(sd, j, z)
    log_px = (-((x - xm) ** 2) / (2 * DECODER_SIGMA ** 2) - math.log(DECODER_SIGMA) - HALF_LOG_2PI).sum(-1)
    log_q = (-((z - mean) ** 2) / (2 * std ** 2) - log_std - HALF_LOG_2PI).sum(-1)
    log_p = (-(z ** 2) / 2 - HALF_LOG_2PI).sum(-1)
    return -torch.mean(log_px - beta * (log_q - log_p))


def train_step(sd, opt, x, num_decoders: int, beta: float = 1.0) -> float:
    """One optimisation step on batch x: noise first, then the decoder draw (the reference's order)."""
    eps = torch.empty((x.shape[0], sd["prior.mean"].numel()), dtype=x.dtype, device=x.device).normal_()
    j = int(np.random.choice(num_decoders))
    loss = neg_elbo(sd, x, j, eps, beta)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.item())


def make_optimizer(sd, lr: float):
    params = trainable(sd)
    for p in params:
        p.requires_grad_(True)
    return torch.optim.Adam(params, lr=lr, fused=True) if params[0].is_cuda else torch.optim.Adam(params, lr=lr)


def evaluate(sd, loader, device, num_decoders: int, beta: float = 1.0) -> float:
    vals = []
    with torch.no_grad():
        for (x,) in loader:
            x = x.to(device)
            eps = torch.empty((x.shape[0], sd["prior.mean"].numel()), dtype=x.dtype, device=device).normal_()
            vals.append(float(neg_elbo(sd, x, int(np.random.choice(num_decoders)), eps, beta)))
    return float(np.mean(vals))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent-dim", type=int, default=2)
    ap.add_argument("--num-decoders", type=int, default=10)
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--device", type=str, default="cpu")
    ap.add_argument("--save-dir", type=str, default="experiment")
    ap.add_argument("--data-path", type=str, default="data/tasic-pca50.npy")
    args = ap.parse_args(argv)

    os.makedirs(os.path.join(args.save_dir, "plots"), exist_ok=True)
    torch.manual_seed(args.seed)
    data = torch.from_numpy(np.load(args.data_path).astype(np.float32))
    n = len(data)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(args.seed))
    n_val = int(0.1 * n)
    from torch.utils.data import DataLoader, TensorDataset
    train_loader = DataLoader(TensorDataset(data[perm[n_val:]]), batch_size=args.batch_size, shuffle=True)
    val_loader = DataLoader(TensorDataset(data[perm[:n_val]]), batch_size=args.batch_size)

    sd = init_parameters(data.shape[1], args.latent_dim, args.num_decoders)
    sd = OrderedDict((k, v.to(args.device)) for k, v in sd.items())
    opt = make_optimizer(sd, args.lr)
    print("beta = ", 1.0)
    history = {"train": [], "val": []}
    for epoch in range(args.epochs):
        losses = [train_step(sd, opt, x.to(args.device), args.num_decoders) for (x,) in train_loader]
        history["train"].append(float(np.mean(losses)))
        history["val"].append(evaluate(sd, val_loader, args.device, args.num_decoders))
        print(f"Epoch {epoch + 1:3d} | Train: {history['train'][-1]:.2f} | Val: {history['val'][-1]:.2f}")
    # the reference plots the two curves (matplotlib is not a dependency here: the numbers are saved instead)
    np.save(os.path.join(args.save_dir, "plots", f"loss_curve_seed{args.seed}.npy"),
            np.array([history["train"], history["val"]]))
    out = os.path.join(args.save_dir, f"model_seed{args.seed}.pt")
    torch.save(OrderedDict((k, v.detach().cpu()) for k, v in sd.items()), out)
    print(f"\nSaved model + {args.num_decoders} decoders.")
    with torch.no_grad():
        z = encode(sd, data.to(args.device))[0]
        print("Mean of latent z across dataset:", z.mean(dim=0))
        print("Std of latent z across dataset:", z.std(dim=0))
    return out


if __name__ == "__main__":
    main()
