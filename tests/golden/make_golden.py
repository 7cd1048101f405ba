"""Generate the golden fixtures in this directory by running the REFERENCE's own code.

Run once, in the build container (where /root/reference is mounted read-only):

    python tests/golden/make_golden.py

It imports the reference's ``GeodesicSplineBatch`` / ``compute_energy_mc``
(src/optimize.py:13-75), the single-decoder ``compute_energy`` /
``compute_geodesic_lengths`` (src/single_decoder/optimize_energy_batched.py:42-57),
``construct_nullspace_basis`` (src/single_decoder/optimize_energy.py:58-102) and drives
them exactly like the loop at src/optimize.py:152-162, with ``torch.randint`` replaced by
*recorded* draws so the MC energy is reproducible.  matplotlib / seaborn are not installed
here and are stubbed (they are only used for plotting).  Nothing from the reference is
copied: only its outputs on committed inputs are stored.

The GPU box has no /root/reference: tests read only the .npz files written here.
"""
import sys
import types
import zlib
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm",
             "seaborn", "mpl_toolkits", "mpl_toolkits.axes_grid1"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, str(REF))

import src.optimize as ref_opt  # noqa: E402
import src.single_decoder.optimize_energy as ref_single  # noqa: E402
import src.single_decoder.optimize_energy_batched as ref_sb  # noqa: E402
from src.single_decoder.vae import VAE  # noqa: E402
from src.train import (EVAE, GaussianDecoder, GaussianEncoder, GaussianPrior,  # noqa: E402
                       make_decoder_net, make_encoder_net)

torch.set_num_threads(8)


def make_draws(seed, S, M, T, N, K):
    """Recorded decoder draws, layout [S,M,2,T-1,N] (order of src/optimize.py:57-58)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, K, (S, M, 2, T - 1, N), generator=g, dtype=torch.int64)


class DrawFeeder:
    """Stands in for torch.randint inside compute_energy_mc: hands out the recorded
    draws in call order (d1 then d2 for each MC sample)."""

    def __init__(self, draws):
        self.draws = draws
        self.step = 0
        self.call = 0

    def __call__(self, low, high, size, device=None, **kw):
        M = self.draws.shape[1]
        m, role = divmod(self.call, 2)
        out = self.draws[self.step, m, role]
        assert tuple(out.shape) == tuple(size), (out.shape, size)
        self.call += 1
        if self.call == 2 * M:
            self.call = 0
            self.step += 1
        return out


def load_evae(path, dtype):
    enc = GaussianEncoder(make_encoder_net(50, 2))
    dec = GaussianDecoder(make_decoder_net(2, 50))
    model = EVAE(GaussianPrior(2), enc, dec, num_decoders=10)
    model.load_state_dict(torch.load(path, map_location="cpu"))
    model.eval()
    return model.to(dtype)


def decoder_arrays(decoders, rows=None):
    out = {}
    for name, idx in (("W1", 0), ("W2", 2), ("W3", 4)):
        w = torch.stack([d.decoder_net[idx].weight.detach() for d in decoders]).float().numpy()
        bb = torch.stack([d.decoder_net[idx].bias.detach() for d in decoders]).float().numpy()
        if name == "W3" and rows is not None:
            w, bb = w[:, :rows], bb[:, :rows]
        out[name] = w
        out["b" + name[1]] = bb
    return out


def run_reference_mc(decoders, a, b, omega0, basis, n_poly, T, draws, steps, dtype,
                     record_after=()):
    """The loop of src/optimize.py:152-162 with the reference's own classes."""
    t_vals = torch.linspace(0, 1, T).to(dtype)
    a, b, basis = a.to(dtype), b.to(dtype), basis.to(dtype)
    model = ref_opt.GeodesicSplineBatch(a, b, basis, omega0.to(dtype).clone(), n_poly)
    opt = torch.optim.Adam([model.omega], lr=1e-3)
    feeder = DrawFeeder(draws)
    real_randint = torch.randint
    ref_opt.torch.randint = feeder
    energies, grad0, snaps, z0 = [], None, {}, None
    try:
        for step in range(steps):
            opt.zero_grad()
            if step == 0:
                with torch.no_grad():
                    z0 = model(t_vals).clone()
            energy = ref_opt.compute_energy_mc(model, decoders, t_vals, M=draws.shape[1])
            endpoint_error = (model(t_vals[-1:]) - b[None]) ** 2
            loss = energy + 1000 * endpoint_error.sum(dim=(0, 2))
            loss.sum().backward()
            if step == 0:
                grad0 = model.omega.grad.detach().clone()
            opt.step()
            energies.append(energy.detach().clone())
            if (step + 1) in record_after:
                st = opt.state[model.omega]
                snaps[step + 1] = (model.omega.detach().clone(), st["exp_avg"].clone(),
                                   st["exp_avg_sq"].clone())
    finally:
        ref_opt.torch.randint = real_randint
    st = opt.state[model.omega]
    return dict(energy=torch.stack(energies), grad0=grad0, omega=model.omega.detach().clone(),
                m=st["exp_avg"].clone(), v=st["exp_avg_sq"].clone(), snaps=snaps, z0=z0)


def np64(x):
    return x.detach().double().numpy()


def np32(x):
    return x.detach().float().numpy()


def ensemble_case(tag, model_path, spline_path, sl, steps, draw_seed, k_active=10, M=2,
                  long_steps=0, zero_omega=False):
    blob = torch.load(spline_path, map_location="cpu", weights_only=False)
    chunk = blob["spline_data"][sl]
    n_poly = chunk[0]["n_poly"]
    basis = chunk[0]["basis"]
    a = torch.stack([d["a"] for d in chunk])
    b = torch.stack([d["b"] for d in chunk])
    om = torch.stack([d["omega_init"] for d in chunk])
    if zero_omega:
        om = torch.zeros_like(om)
    T, N = 2000, len(chunk)
    draws = make_draws(draw_seed, max(steps, long_steps), M, T, N, k_active)
    res = {}
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        model = load_evae(model_path, dtype)
        decs = [model.decoder[i] for i in range(k_active)]
        res[name] = run_reference_mc(decs, a, b, om, basis, n_poly, T, draws, steps, dtype,
                                     record_after=(1,))
        if long_steps:
            res[name + "_long"] = run_reference_mc(decs, a, b, om, basis, n_poly, T, draws,
                                                   long_steps, dtype)
    out = dict(a=np32(a), b=np32(b), omega_init=np32(om), basis=np32(basis),
               n_poly=n_poly, T=T, M=M, K=k_active, steps=steps, draw_seed=draw_seed,
               draws_crc=zlib.crc32(draws.numpy().astype(np.uint8).tobytes()),
               draws_shape=np.array(draws.shape))
    for name in ("f32", "f64"):
        r = res[name]
        cv = np32 if name == "f32" else np64
        out[f"energy_{name}"] = cv(r["energy"])
        out[f"grad0_{name}"] = cv(r["grad0"])
        out[f"omega_{name}"] = cv(r["omega"])
        out[f"m_{name}"] = cv(r["m"])
        out[f"v_{name}"] = cv(r["v"])
        o1, m1, v1 = r["snaps"][1]
        out[f"omega1_{name}"] = cv(o1)
        out[f"m1_{name}"] = cv(m1)
        out[f"v1_{name}"] = cv(v1)
        if long_steps:
            rl = res[name + "_long"]
            out[f"long_energy_{name}"] = cv(rl["energy"])
            out[f"long_omega_{name}"] = cv(rl["omega"])
    out["z0_f32"] = np32(res["f32"]["z0"])[::50]  # every 50th point
    out["long_steps"] = long_steps
    np.savez_compressed(OUT / f"{tag}.npz", **out)
    print(tag, "E0", out["energy_f32"][0][:3], "rel f32-f64",
          np.abs(out["energy_f32"] / out["energy_f64"] - 1).max())


def synthetic_case(tag, n_poly, T, N, K, M, steps, seed):
    """Random-init decoders (default nn.Linear init) / n_poly 8 / short T: the shapes of
    BASELINE config 5, at a size the oracle finishes instantly."""
    torch.manual_seed(seed)
    decs32 = [GaussianDecoder(make_decoder_net(2, 50)) for _ in range(K)]
    basis, _ = ref_single.construct_nullspace_basis(n_poly, "cpu")
    a = (torch.rand(N, 2) * 6 - 3)
    b = (torch.rand(N, 2) * 6 - 3)
    om = 0.1 * torch.randn(N, n_poly + 1, 2)
    draws = make_draws(seed + 1, steps, M, T, N, K)
    out = dict(a=np32(a), b=np32(b), omega_init=np32(om), basis=np32(basis), n_poly=n_poly,
               T=T, M=M, K=K, steps=steps, draw_seed=seed + 1,
               draws_crc=zlib.crc32(draws.numpy().astype(np.uint8).tobytes()),
               draws_shape=np.array(draws.shape), **decoder_arrays(decs32))
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        decs = [d.to(dtype) for d in decs32]
        r = run_reference_mc(decs, a, b, om, basis, n_poly, T, draws, steps, dtype,
                             record_after=(1,))
        cv = np32 if name == "f32" else np64
        for k in ("energy", "grad0", "omega", "m", "v"):
            out[f"{k}_{name}"] = cv(r[k])
        decs32 = [d.float() for d in decs32]
    np.savez_compressed(OUT / f"{tag}.npz", **out)
    print(tag, out["energy_f32"][0][:3])


def single_case(tag, seed, count, steps):
    """Single-decoder VAE (BASELINE config 2): deterministic energy, 500-step committed
    answers from the reference's own artifact."""
    vae = VAE(input_dim=50, latent_dim=2)
    vae.load_state_dict(torch.load(REF / f"src/artifacts/vae_best_seed{seed}.pth", map_location="cpu"))
    vae.eval()
    init = torch.load(REF / f"src/artifacts/spline_batch_seed{seed}.pt", map_location="cpu",
                      weights_only=False)["spline_data"][:count]
    done = torch.load(REF / f"src/artifacts/spline_batch_optimized_batched_seed{seed}.pt",
                      map_location="cpu", weights_only=False)[:count]
    n_poly = init[0]["n_poly"]
    # the batched script recomputes the basis (optimize_energy_batched.py:74); the stored
    # one is what the committed omegas refer to -> keep both and record them.
    basis_file = init[0]["basis"]
    a = torch.stack([d["a"] for d in init])
    b = torch.stack([d["b"] for d in init])
    om = torch.stack([d["omega_init"] for d in init])
    T = 2000
    out = dict(a=np32(a), b=np32(b), omega_init=np32(om), basis=np32(basis_file), n_poly=n_poly,
               T=T, steps=steps,
               committed_length_geodesic=np.array([d["length_geodesic"] for d in done]),
               committed_omega_optimized=np32(torch.stack([d["omega_optimized"] for d in done])),
               committed_basis=np32(done[0]["basis"]))
    dn = vae.decoder.decoder_net
    out.update(W1=np32(dn[0].weight)[None], b1=np32(dn[0].bias)[None],
               W2=np32(dn[2].weight)[None], b2=np32(dn[2].bias)[None],
               W3=np32(dn[4].weight)[None, :50], b3=np32(dn[4].bias)[None, :50])
    for name, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        v = VAE(input_dim=50, latent_dim=2)
        v.load_state_dict(vae.state_dict())
        v = v.to(dtype).eval()
        t_vals = torch.linspace(0, 1, T).to(dtype)
        model = ref_sb.GeodesicSplineBatch(a.to(dtype), b.to(dtype), basis_file.to(dtype),
                                           om.to(dtype).clone(), n_poly)
        opt = torch.optim.Adam([model.omega], lr=1e-3)
        energies = []
        for step in range(steps):
            opt.zero_grad()
            energy = ref_sb.compute_energy(model, v.decoder, t_vals)
            endpoint_error = (model(t_vals[-1:]) - b.to(dtype)[None]) ** 2
            loss = energy + 1000 * endpoint_error.sum(dim=(0, 2))
            loss.sum().backward()
            if step == 0:
                out[f"grad0_{name}"] = (np32 if name == "f32" else np64)(model.omega.grad)
            opt.step()
            energies.append(energy.detach().clone())
        cv = np32 if name == "f32" else np64
        out[f"energy_{name}"] = cv(torch.stack(energies))
        out[f"omega_{name}"] = cv(model.omega.detach())
        out[f"length_{name}"] = cv(ref_sb.compute_geodesic_lengths(model, v.decoder, t_vals))
    np.savez_compressed(OUT / f"{tag}.npz", **out)
    print(tag, out["energy_f32"][0][:3], out["length_f32"][:3], out["committed_length_geodesic"][:3])


def std_field_case(tag, model_path):
    """Ensemble disagreement field (src/init_splines_ensemble.py:47-54)."""
    model = load_evae(model_path, torch.float32)
    gx, gy = torch.meshgrid(torch.linspace(-3.5, 3.5, 48), torch.linspace(-3.5, 3.5, 48), indexing="ij")
    grid = torch.stack([gx, gy], dim=-1).view(-1, 2)
    with torch.no_grad():
        outputs = torch.stack([d(grid).mean for d in model.decoder])
        s32 = outputs.std(dim=0).norm(dim=1)
        m64 = load_evae(model_path, torch.float64)
        o64 = torch.stack([d(grid.double()).mean for d in m64.decoder])
        s64 = o64.std(dim=0).norm(dim=1)
    np.savez_compressed(OUT / f"{tag}.npz", grid=np32(grid), std_norm_f32=np32(s32), std_norm_f64=np64(s64))
    print(tag, s32[:3])


def lbfgs_fit_case(tag):
    """Spline fit to a poly-line (src/init_splines_ensemble.py:172-193) on synthetic
    grid paths: the reference's LBFGS(max_iter=50) from omega=0."""
    basis, _ = ref_single.construct_nullspace_basis(4, "cpu")
    rng = np.random.default_rng(0)
    targets, omegas, lens = [], [], []
    for L in (12, 57, 140, 233):
        steps = rng.integers(-1, 2, size=(L, 2)).astype(np.float32) * 0.035
        steps[:, 0] += 0.02
        target = torch.tensor(np.cumsum(steps, axis=0), dtype=torch.float32)
        a, b = target[0], target[-1]
        spline = ref_sb.GeodesicSplineBatch(a[None], b[None], basis,
                                            omega=torch.zeros((1, basis.shape[1], 2)), n_poly=4)
        t_vals = torch.linspace(0, 1, len(target))
        opt = torch.optim.LBFGS([spline.omega], max_iter=50)

        def closure():
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(spline(t_vals).squeeze(1), target)
            loss.backward()
            return loss

        opt.step(closure)
        pad = np.zeros((233, 2), np.float32)
        pad[:L] = target.numpy()
        targets.append(pad)
        lens.append(L)
        omegas.append(np32(spline.omega.detach()[0]))
    np.savez_compressed(OUT / f"{tag}.npz", targets=np.stack(targets), lens=np.array(lens),
                        omega_lbfgs=np.stack(omegas), basis=np32(basis))
    print(tag, omegas[0][:2])


def basis_case(tag):
    out = {}
    for n in (2, 4, 8):
        bs, C = ref_single.construct_nullspace_basis(n, "cpu")
        out[f"basis_{n}"] = np32(bs)
        out[f"C_{n}"] = np32(C)
    np.savez_compressed(OUT / f"{tag}.npz", **out)


if __name__ == "__main__":
    m12 = REF / "experiment/model_seed12.pt"
    # decoder weights of the committed ensemble checkpoint (real weights for parity)
    np.savez_compressed(OUT / "evae_seed12_decoders.npz",
                        **decoder_arrays(list(load_evae(m12, torch.float32).decoder)))
    # all 45 curves of both init files (inputs only; the oracle recomputes on the fly)
    for kind in ("euclidean", "entropy"):
        blob = torch.load(REF / f"experiment/splines_init_model_seed12/spline_batch_init_{kind}_10.pt",
                          map_location="cpu", weights_only=False)
        sd = blob["spline_data"]
        opt = torch.load(REF / f"experiment/splines_opt_model_seed12/spline_batch_opt_{kind}_10.pt",
                         map_location="cpu", weights_only=False)
        np.savez_compressed(OUT / f"splines_seed12_{kind}_10.npz",
                            a=np32(torch.stack([d["a"] for d in sd])),
                            b=np32(torch.stack([d["b"] for d in sd])),
                            omega_init=np32(torch.stack([d["omega_init"] for d in sd])),
                            basis=np32(sd[0]["basis"]), n_poly=sd[0]["n_poly"],
                            a_index=np.array([d["a_index"] for d in sd]),
                            b_index=np.array([d["b_index"] for d in sd]),
                            committed_geodesic_length=np.array([d["geodesic_length"] for d in opt["spline_data"]]),
                            committed_omega_optimized=np32(torch.stack([d["omega_optimized"] for d in opt["spline_data"]])),
                            committed_steps=opt["metadata"]["steps"])
    ensemble_case("ens_seed12_euclid", m12,
                  REF / "experiment/splines_init_model_seed12/spline_batch_init_euclidean_10.pt",
                  slice(0, 8), steps=6, draw_seed=1234, long_steps=0)
    ensemble_case("ens_seed12_entropy", m12,
                  REF / "experiment/splines_init_model_seed12/spline_batch_init_entropy_10.pt",
                  slice(20, 26), steps=4, draw_seed=77)
    ensemble_case("ens_seed12_cov_k3", m12,
                  REF / "experiment/splines_init_model_seed12/spline_batch_init_euclidean_10.pt",
                  slice(5, 8), steps=3, draw_seed=5, k_active=3, zero_omega=True)
    ensemble_case("ens_seed12_long", m12,
                  REF / "experiment/splines_init_model_seed12/spline_batch_init_euclidean_10.pt",
                  slice(0, 4), steps=2, draw_seed=4321, long_steps=150)
    synthetic_case("synth_np8_T256", n_poly=8, T=256, N=6, K=7, M=3, steps=4, seed=0)
    synthetic_case("synth_np4_T130", n_poly=4, T=130, N=5, K=4, M=1, steps=3, seed=3)
    single_case("single_seed123", 123, count=8, steps=6)
    std_field_case("std_field_seed12", m12)
    lbfgs_fit_case("lbfgs_fit")
    basis_case("nullspace_basis")
