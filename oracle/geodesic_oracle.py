"""CPU oracle for the geodesic curve-energy hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The shipped path (``vlg_b200``) never routes through it.

It restates, in plain numpy with a hand-derived backward pass, what the reference
computes with PyTorch autograd.  Every function cites the reference file:line it
follows (paths relative to the reference repo root).  The arithmetic itself lives in a
third-party dependency that is not vendored in the reference (PyTorch, pinned
``torch==2.1.0`` in ``configs/requirements.txt:1``); the published algorithms restated
here are ``nn.Linear``/``ReLU``, ``torch.optim.Adam`` (non-amsgrad, no weight decay) and
``torch.randint`` (replaced by explicit draws).

Parity pinning: the reference has no tests or golden vectors of its own.
The oracle is therefore pinned against outputs of the reference itself: the goldens in
``tests/golden/*.npz`` were produced by ``tests/golden/make_golden.py``, which imports
the reference's *own* ``GeodesicSplineBatch`` / ``compute_energy_mc`` /
``torch.optim.Adam`` loop (``src/optimize.py:152-162``) with recorded decoder draws, in
fp32 and fp64.  ``tests/test_oracle.py`` checks this module against them.

Layout conventions (shared with the C ABI in ``include/vlg.h``):
  a, b      [N, 2]            curve end points in latent space
  omega     [N, Kb, 2]        free spline coefficients (Kb = n_poly + 1)
  basis     [4*n_poly, Kb]    null-space basis stored in the spline file
  t         [T]               curve parameter grid (``torch.linspace(0, 1, T)``)
  draws     [S, M, 2, T-1, N] decoder indices, reference order: for each MC sample m the
                              reference draws d1 (role 0) then d2 (role 1)
                              (``src/optimize.py:57-58``)
  decoders  list of dicts W1[H,2] b1[H] W2[H,H] b2[H] W3[X,H] b3[X]
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# spline
# --------------------------------------------------------------------------------------


def segment_coords(t, n_poly):
    """seg = clamp(floor(t*n), max=n-1); u = t*n - seg   (src/optimize.py:27-28)."""
    t = np.asarray(t)
    tn = t * t.dtype.type(n_poly)
    seg = np.minimum(np.floor(tn).astype(np.int64), n_poly - 1)
    u = tn - seg.astype(t.dtype)
    return seg, u


def design_matrix(basis, t, n_poly):
    """P[t,k] = sum_i u_t^i * basis[4*seg_t+i, k]  (SURVEY §8 a-2 restatement of
    src/optimize.py:24-32: poly = P @ omega)."""
    seg, u = segment_coords(t, n_poly)
    dt = basis.dtype
    u = u.astype(dt)
    pw = np.stack([np.ones_like(u), u, u * u, u * u * u], axis=1)  # [T,4]
    rows = basis.reshape(n_poly, 4, -1)[seg]  # [T,4,Kb]
    return np.einsum("ti,tik->tk", pw, rows)


def spline_points(a, b, omega, basis, t, n_poly):
    """z[T,N,2] following the reference operation order (src/optimize.py:22-35):
    coeffs = basis @ omega; poly = sum_i u^i coeffs[seg,i]; z = (1-t) a + t b + poly."""
    dt = omega.dtype
    N, Kb, D = omega.shape
    coeffs = np.einsum("nk,bkd->nbd", basis.astype(dt), omega)  # [4n,N,D]
    coeffs = coeffs.reshape(n_poly, 4, N, D)
    seg, u = segment_coords(np.asarray(t, dtype=dt), n_poly)
    pw = np.stack([np.ones_like(u), u, u * u, u * u * u], axis=1)
    sel = coeffs[seg]  # [T,4,N,D]
    poly = np.einsum("ti,tibd->tbd", pw, sel)
    tt = np.asarray(t, dtype=dt)[:, None, None]
    lin = (1 - tt) * a[None] + tt * b[None]
    return lin + poly


# --------------------------------------------------------------------------------------
# decoders
# --------------------------------------------------------------------------------------


def decoder_hidden(dec, z):
    """Pre-activations of the decoder MLP 2 -> H -> H -> X (src/train.py:80-85)."""
    p1 = z @ dec["W1"].T + dec["b1"]
    h1 = np.maximum(p1, 0)
    p2 = h1 @ dec["W2"].T + dec["b2"]
    h2 = np.maximum(p2, 0)
    x = h2 @ dec["W3"].T + dec["b3"]
    return p1, p2, x


def decoder_mean(dec, z):
    """``GaussianDecoder(z).mean`` (src/train.py:42-46); sigma=5 is irrelevant to .mean.
    For the single VAE (src/single_decoder/vae.py:29-42) pass W3/b3 rows 0:X only."""
    return decoder_hidden(dec, z)[2]


def cast_decoders(decoders, dtype):
    return [{k: np.asarray(v, dtype=dtype) for k, v in d.items()} for d in decoders]


# --------------------------------------------------------------------------------------
# energy + hand-derived gradient
# --------------------------------------------------------------------------------------


def energy_mc(a, b, omega, basis, t, n_poly, decoders, draws):
    """MC ensemble energy E[N] (src/optimize.py:38-75).  draws: [M,2,T-1,N] ints.
    E_b = (1/M) sum_m sum_t || X[d2[m,t,b], t+1, b] - X[d1[m,t,b], t, b] ||^2."""
    return energy_mc_grad(a, b, omega, basis, t, n_poly, decoders, draws, want_grad=False)[0]


def energy_mc_grad(a, b, omega, basis, t, n_poly, decoders, draws, penalty_w=1000.0,
                   want_grad=True):
    """Energy and d(loss)/d(omega) where loss = E + penalty_w * ||z(t=1) - b||^2
    (src/optimize.py:156-161).  Backward is derived by hand: only the
    input gradient of each decoder is formed; the decoder weight gradients that the
    reference also computes (and never uses) are not.

    Returns (E[N], grad[N,Kb,2] or None, z_end_err[N,2])."""
    dt = omega.dtype
    T = len(t)
    N = a.shape[0]
    M = draws.shape[0]
    K = len(decoders)
    z = spline_points(a, b, omega, basis, t, n_poly)  # [T,N,2]
    zf = z.reshape(T * N, 2)
    pre = [decoder_hidden(d, zf) for d in decoders]
    X = np.stack([p[2].reshape(T, N, -1) for p in pre], axis=0)  # [K,T,N,X]
    it = np.arange(T - 1)[:, None]
    ib = np.arange(N)[None, :]
    E = np.zeros(N, dtype=dt)
    G = np.zeros_like(X) if want_grad else None
    for m in range(M):
        d1 = draws[m, 0]
        d2 = draws[m, 1]
        x1 = X[d1, it, ib]
        x2 = X[d2, it + 1, ib]
        diff = x2 - x1
        E += (diff * diff).sum(axis=2).sum(axis=0)
        if want_grad:
            g = (dt.type(2.0) / dt.type(M)) * diff
            np.add.at(G, (d1, it, ib), -g)
            np.add.at(G, (d2, it + 1, ib), g)
    E = E / dt.type(M)
    # end-point penalty: a separate spline evaluation at t[-1:] (src/optimize.py:158-159)
    z_end = spline_points(a, b, omega, basis, np.asarray(t)[-1:], n_poly)[0]
    end_err = z_end - b
    if not want_grad:
        return E, None, end_err
    dz = np.zeros((T * N, 2), dtype=dt)
    for k, dec in enumerate(decoders):
        p1, p2, _ = pre[k]
        g3 = G[k].reshape(T * N, -1)
        dh2 = (g3 @ dec["W3"]) * (p2 > 0)
        dh1 = (dh2 @ dec["W2"]) * (p1 > 0)
        dz += dh1 @ dec["W1"]
    dz = dz.reshape(T, N, 2)
    P = design_matrix(basis.astype(dt), np.asarray(t, dtype=dt), n_poly)  # [T,Kb]
    grad = np.einsum("tk,tbd->bkd", P, dz)
    grad += dt.type(2.0 * penalty_w) * end_err[:, None, :] * P[-1][None, :, None]
    return E, grad, end_err


def energy_single(a, b, omega, basis, t, n_poly, decoder):
    """Deterministic single-decoder energy sum_t ||x_{t+1}-x_t||^2
    (src/single_decoder/optimize_energy_batched.py:51-57)."""
    T, N = len(t), a.shape[0]
    dr = np.zeros((1, 2, T - 1, N), dtype=np.int64)
    return energy_mc(a, b, omega, basis, t, n_poly, [decoder], dr)


def curve_length_single(a, b, omega, basis, t, n_poly, decoder):
    """Polyline length sum_t ||x_{t+1}-x_t||
    (src/single_decoder/optimize_energy_batched.py:42-49)."""
    T, N = len(t), a.shape[0]
    z = spline_points(a, b, omega, basis, t, n_poly)
    x = decoder_mean(decoder, z.reshape(T * N, 2)).reshape(T, N, -1)
    d = x[1:] - x[:-1]
    return np.sqrt((d * d).sum(axis=2)).sum(axis=0)


# --------------------------------------------------------------------------------------
# Adam + driver
# --------------------------------------------------------------------------------------


def adam_update(omega, m, v, g, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor/foreach semantics (torch/optim/adam.py, defaults of
    src/optimize.py:153): step is 1-based.  Scalars are Python doubles exactly as in
    torch; tensor ops stay in omega.dtype."""
    dt = omega.dtype.type
    m = m + (g - m) * dt(1 - beta1)  # lerp_
    v = v * dt(beta2) + (g * g) * dt(1 - beta2)  # mul_().addcmul_()
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / dt(bc2 ** 0.5) + dt(eps)
    omega = omega - dt(step_size) * (m / denom)
    return omega, m, v


def optimize_steps(a, b, omega, basis, t, n_poly, decoders, draws, steps, step0=0,
                   adam_m=None, adam_v=None, lr=1e-3, penalty_w=1000.0):
    """The inner loop of src/optimize.py:155-162 for all N curves (curves are
    independent).  draws: [S,M,2,T-1,N].  Returns dict with omega/m/v after `steps`
    updates, per-step energies [S,N] (energy of the omega *before* that step's update,
    as printed/used at src/optimize.py:164-168)."""
    m = np.zeros_like(omega) if adam_m is None else adam_m.copy()
    v = np.zeros_like(omega) if adam_v is None else adam_v.copy()
    omega = omega.copy()
    energies = []
    grads0 = None
    for s in range(steps):
        E, g, _ = energy_mc_grad(a, b, omega, basis, t, n_poly, decoders, draws[s],
                                 penalty_w=penalty_w)
        if s == 0:
            grads0 = g
        energies.append(E)
        omega, m, v = adam_update(omega, m, v, g, step0 + s + 1, lr=lr)
    return {"omega": omega, "m": m, "v": v, "energy": np.stack(energies), "grad0": grads0}


# --------------------------------------------------------------------------------------
# ensemble disagreement field (init_splines_ensemble.py:47-54)
# --------------------------------------------------------------------------------------


def ensemble_std_norm(grid, decoders):
    """|| std_k f_k(grid) ||_2 with the unbiased (K-1) estimator, before the min-max
    normalisation (src/init_splines_ensemble.py:49-51)."""
    X = np.stack([decoder_mean(d, grid) for d in decoders])  # [K,G,X]
    sd = X.std(axis=0, ddof=1)
    return np.sqrt((sd * sd).sum(axis=1))


# --------------------------------------------------------------------------------------
# counter-based decoder draws (Philox4x32-10), shared definition with the CUDA kernels
# --------------------------------------------------------------------------------------

_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011).  All args uint32 arrays/scalars."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def counter_draws(seed, curve_ids, step, T, M, K):
    """Decoder draws for one optimisation step, keyed on (seed, global curve id, step,
    MC-sample pair, segment) so results do not depend on how curves are sharded.

    Philox counter = (segment t, step, curve id, m // 2), key = (seed lo, seed hi); the
    four output words map to (m even: d1, d2; m odd: d1, d2); a word w becomes the
    decoder index (w * K) >> 32.  Returns int64 [M,2,T-1,N]."""
    curve_ids = np.asarray(curve_ids, dtype=np.uint32)
    N = curve_ids.shape[0]
    t = np.arange(T - 1, dtype=np.uint32)[:, None]
    out = np.zeros((M, 2, T - 1, N), dtype=np.int64)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    for j in range((M + 1) // 2):
        w = philox4x32_10(t, np.uint32(step), curve_ids[None, :], np.uint32(j), k0, k1)
        for q in range(4):
            mm = 2 * j + q // 2
            if mm < M:
                out[mm, q % 2] = (w[q].astype(np.uint64) * np.uint64(K)) >> np.uint64(32)
    return out


# --------------------------------------------------------------------------------------
# least-squares spline fit to a path (init_splines_ensemble.py:172-193)
# --------------------------------------------------------------------------------------


def fit_spline_to_path(target, basis, n_poly):
    """argmin_omega mean((lin + P_L omega - target)^2): the optimum the reference's
    LBFGS(max_iter=50) loop (src/init_splines_ensemble.py:175-192) converges towards.
    target [L,2]; a,b = target[0], target[-1].  Returns omega [Kb,2] (float64 solve)."""
    L = target.shape[0]
    tt = np.linspace(0.0, 1.0, L).astype(np.float32).astype(np.float64)
    P = design_matrix(basis.astype(np.float64), tt, n_poly)
    a, b = target[0].astype(np.float64), target[-1].astype(np.float64)
    lin = (1 - tt)[:, None] * a[None] + tt[:, None] * b[None]
    rhs = target.astype(np.float64) - lin
    omega, *_ = np.linalg.lstsq(P, rhs, rcond=None)
    return omega
