"""CPU port of the reference's optimisation loop in PyTorch (autograd + torch.optim.Adam).

TEST / BASELINE INFRASTRUCTURE ONLY: used by ``bench.py`` for the ``cpu_baseline`` leg and
``--impl reference`` (the reference itself is pure Python and is not present on the GPU box),
and by tests to cross-check the numpy oracle.  It restates src/optimize.py:13-75,152-162 in
design-matrix form and keeps the reference's costs: all K decoders are evaluated densely on all
T points, the decoder parameters keep ``requires_grad=True`` (so autograd also forms the weight
gradients the reference never uses), draws come from ``torch.randint`` per MC sample.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def make_decoder(W):
    """nn.Sequential 2->128->ReLU->128->ReLU->X (src/train.py:80-85) from weight arrays."""
    net = nn.Sequential(nn.Linear(2, 128), nn.ReLU(), nn.Linear(128, 128), nn.ReLU(),
                        nn.Linear(128, W["W3"].shape[0]))
    with torch.no_grad():
        for idx, (w, b) in zip((0, 2, 4), (("W1", "b1"), ("W2", "b2"), ("W3", "b3"))):
            net[idx].weight.copy_(torch.as_tensor(W[w]))
            net[idx].bias.copy_(torch.as_tensor(W[b]))
    return net


def design_matrix(basis, t, n_poly):
    tn = t * n_poly
    seg = torch.clamp(tn.floor().long(), max=n_poly - 1)
    u = tn - seg.to(t.dtype)
    pw = torch.stack([u ** i for i in range(4)], dim=1)
    rows = basis.view(n_poly, 4, -1)[seg]
    return torch.einsum("ti,tik->tk", pw, rows)


class SplineBatch(nn.Module):
    def __init__(self, a, b, basis, omega, n_poly):
        super().__init__()
        self.a, self.b, self.basis, self.n_poly = a, b, basis, n_poly
        self.omega = nn.Parameter(omega.clone())

    def forward(self, t):
        P = design_matrix(self.basis, t, self.n_poly)
        lin = (1 - t[:, None, None]) * self.a[None] + t[:, None, None] * self.b[None]
        return lin + torch.einsum("tk,bkd->tbd", P, self.omega)


def energy_mc(model, decoders, t, M, draws=None):
    T, B = t.shape[0], model.a.shape[0]
    z = model(t)
    X = torch.stack([d(z) for d in decoders], dim=0)
    it = torch.arange(T - 1)[:, None]
    ib = torch.arange(B)[None, :]
    total = torch.zeros(B, dtype=z.dtype)
    for m in range(M):
        if draws is None:
            d1 = torch.randint(0, len(decoders), (T - 1, B))
            d2 = torch.randint(0, len(decoders), (T - 1, B))
        else:
            d1, d2 = draws[m, 0], draws[m, 1]
        x1 = X[d1, it, ib]
        x2 = X[d2, it + 1, ib]
        total = total + ((x2 - x1) ** 2).sum(dim=2).sum(dim=0)
    return total / M


def run_steps(model, decoders, t, steps, M=2, lr=1e-3, draws=None, penalty_w=1000.0, opt=None):
    """`steps` iterations of the loop at src/optimize.py:155-162.  Returns (energies, opt)."""
    opt = opt or torch.optim.Adam([model.omega], lr=lr)
    out = []
    for s in range(steps):
        opt.zero_grad()
        e = energy_mc(model, decoders, t, M, None if draws is None else draws[s])
        pen = ((model(t[-1:]) - model.b[None]) ** 2).sum(dim=(0, 2))
        (e + penalty_w * pen).sum().backward()
        opt.step()
        out.append(e.detach())
    return torch.stack(out), opt
