"""CPU: the oracle (oracle/geodesic_oracle.py) against goldens produced by the reference's
own code (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import geodesic_oracle as O
from tests import helpers as Hh

MC_CASES = ["ens_seed12_euclid", "ens_seed12_entropy", "ens_seed12_cov_k3", "synth_np8_T256", "synth_np4_T130"]


@pytest.mark.parametrize("tag", MC_CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_optimize_steps_matches_reference(tag, prec):
    g = Hh.load(tag)
    dt = np.float32 if prec == "f32" else np.float64
    K, T, S = int(g["K"]), int(g["T"]), int(g["steps"])
    decs = Hh.decoder_list(Hh.decoder_arrays(g), K, dt)
    draws = Hh.regen_draws(g)[:S]
    r = O.optimize_steps(g["a"].astype(dt), g["b"].astype(dt), g["omega_init"].astype(dt), g["basis"].astype(dt),
                         Hh.tgrid(T, dt), int(g["n_poly"]), decs, draws, S)
    tolE, tolG = (5e-6, 1e-5) if prec == "f32" else (2e-7, 1e-7)  # golden f64 energy sits in an fp32 accumulator
    assert np.abs(r["energy"] / g[f"energy_{prec}"] - 1).max() < tolE
    assert Hh.relerr(r["grad0"], g[f"grad0_{prec}"]) < tolG
    assert np.abs(r["omega"] - g[f"omega_{prec}"]).max() < (2e-6 if prec == "f32" else 1e-9)
    assert Hh.relerr(r["m"], g[f"m_{prec}"]) < (1e-5 if prec == "f32" else 1e-7)
    assert Hh.relerr(r["v"], g[f"v_{prec}"]) < (2e-5 if prec == "f32" else 1e-6)


def test_spline_points_match_reference():
    g = Hh.load("ens_seed12_euclid")
    z = O.spline_points(g["a"], g["b"], g["omega_init"], g["basis"], Hh.tgrid(2000), int(g["n_poly"]))
    assert np.abs(z[::50] - g["z0_f32"]).max() < 2e-6


def test_design_matrix_equals_spline_form():
    g = Hh.load("synth_np8_T256")
    t = Hh.tgrid(256, np.float64)
    P = O.design_matrix(g["basis"].astype(np.float64), t, 8)
    om = g["omega_init"].astype(np.float64)
    a, b = g["a"].astype(np.float64), g["b"].astype(np.float64)
    z = O.spline_points(a, b, om, g["basis"].astype(np.float64), t, 8)
    z2 = (1 - t)[:, None, None] * a[None] + t[:, None, None] * b[None] + np.einsum("tk,bkd->tbd", P, om)
    assert np.abs(z - z2).max() < 1e-12


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_single_decoder_energy_and_length(prec):
    g = Hh.load("single_seed123")
    dt = np.float32 if prec == "f32" else np.float64
    dec = Hh.decoder_list(Hh.decoder_arrays(g), 1, dt)[0]
    args = (g["a"].astype(dt), g["b"].astype(dt), g["omega_init"].astype(dt), g["basis"].astype(dt), Hh.tgrid(2000, dt),
            int(g["n_poly"]))
    E = O.energy_single(*args, dec)
    # single-decoder energies are sums of tiny differences of large numbers: fp32 noise ~1e-4
    assert np.abs(E / g[f"energy_{prec}"][0] - 1).max() < (5e-4 if prec == "f32" else 1e-9)
    S = int(g["steps"])
    T, N = 2000, g["a"].shape[0]
    draws = np.zeros((S, 1, 2, T - 1, N), dtype=np.int64)
    r = O.optimize_steps(*args, [dec], draws, S)
    L = O.curve_length_single(args[0], args[1], r["omega"], *args[3:], dec)
    assert np.abs(L / g[f"length_{prec}"] - 1).max() < (2e-4 if prec == "f32" else 1e-7)
    if prec == "f64":
        assert np.abs(r["energy"] / g["energy_f64"] - 1).max() < 1e-8
        assert Hh.relerr(r["grad0"], g["grad0_f64"]) < 1e-8


def test_std_field():
    g = Hh.load("std_field_seed12")
    w = Hh.load("evae_seed12_decoders")
    s = O.ensemble_std_norm(g["grid"].astype(np.float64), Hh.decoder_list(w, 10, np.float64))
    assert np.abs(s / g["std_norm_f64"] - 1).max() < 1e-10
    s32 = O.ensemble_std_norm(g["grid"], Hh.decoder_list(w, 10, np.float32))
    assert np.abs(s32 / g["std_norm_f32"] - 1).max() < 2e-5


def test_lstsq_fit_is_the_lbfgs_fixed_point():
    """The reference stops LBFGS after 50 iterations; the least-squares optimum has a loss that is
    never worse and the two coefficient sets agree closely."""
    g = Hh.load("lbfgs_fit")
    basis = g["basis"]
    for i, L in enumerate(g["lens"]):
        target = g["targets"][i, :L]
        om = O.fit_spline_to_path(target, basis, 4)
        tt = np.linspace(0, 1, L).astype(np.float32).astype(np.float64)
        P = O.design_matrix(basis.astype(np.float64), tt, 4)
        lin = (1 - tt)[:, None] * target[0][None] + tt[:, None] * target[-1][None]

        def loss(o):
            return np.mean((lin + P @ o - target) ** 2)

        assert loss(om) <= loss(g["omega_lbfgs"][i].astype(np.float64)) * (1 + 1e-9) + 1e-12
        assert np.abs(om - g["omega_lbfgs"][i]).max() < 5e-3


# Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the Random123 distribution)
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,out", PHILOX_KAT)
def test_philox_known_answers(ctr, key, out):
    w = O.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
    assert tuple(int(x) for x in w) == out


def test_counter_draws_shape_and_range():
    d = O.counter_draws(seed=7, curve_ids=np.arange(5) + 100, step=3, T=64, M=3, K=10)
    assert d.shape == (3, 2, 63, 5) and d.min() >= 0 and d.max() <= 9
    # sharding independence: a curve's stream only depends on its global id
    d2 = O.counter_draws(seed=7, curve_ids=np.array([102]), step=3, T=64, M=3, K=10)
    assert (d2[..., 0] == d[..., 2]).all()
    # roughly uniform
    big = O.counter_draws(seed=1, curve_ids=np.arange(64), step=0, T=2000, M=2, K=10)
    freq = np.bincount(big.ravel(), minlength=10) / big.size
    assert np.abs(freq - 0.1).max() < 0.005


def test_adam_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=(3, 5, 2)).astype(np.float32)
    p = torch.nn.Parameter(torch.tensor(p0))
    opt = torch.optim.Adam([p], lr=1e-3)
    om, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for s in range(1, 6):
        gnp = (rng.normal(size=p0.shape) * 1e4).astype(np.float32)
        p.grad = torch.tensor(gnp)
        opt.step()
        om, m, v = O.adam_update(om, m, v, gnp, s)
    assert np.abs(om - p.detach().numpy()).max() < 1e-7
