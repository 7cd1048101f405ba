"""CPU: the C-ABI library builds, loads and exports every symbol include/vlg.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "vlg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for need in ["vlg_pack_decoders", "vlg_optimize_steps", "vlg_curve_energy", "vlg_ensemble_std_norm",
                 "vlg_spline_points", "vlg_fit_splines", "vlg_workspace_bytes", "vlg_packed_decoders_bytes"]:
        assert need in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/vlg.h but not exported"


def test_python_binding_covers_the_header(built_lib):
    from vlg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    lib = _lib.load()
    assert lib.vlg_abi_version() == 2
    assert lib.vlg_error_string(0) == b"ok"
    assert b"sm_100" in lib.vlg_error_string(-4)


def test_sizes_and_argument_errors_without_gpu(built_lib):
    from vlg_b200 import _lib
    lib = _lib.load()
    n = lib.vlg_packed_decoders_bytes(10, 128, 50)
    assert n > 10 * (2 * 128 + 128 + 128 * 128 + 128 + 50 * 128 + 50) * 4
    assert lib.vlg_packed_decoders_bytes(10, 64, 50) == 0      # H must be 128
    assert lib.vlg_packed_decoders_bytes(10, 128, 100) == 0    # X too wide
    # fp32 kernel: ReLU-mask scratch per persistent CTA (<= one CTA per curve): whole 2 KB items
    ws = lib.vlg_workspace_bytes(45, 2000, 4, 10, 2, 0)
    assert ws > 256 and (ws - 256) % (45 * 2048) == 0   # 256-byte status header + items
    assert lib.vlg_workspace_bytes(45, 2000, 4, 10, 2, 3) == lib.vlg_workspace_bytes(45, 2000, 4, 10, 2, 1) > 0
    # null pointers are rejected before any CUDA call
    rc = lib.vlg_optimize_steps(None, 10, 50, 10, 4, 2000, 4, 2, 1, 0, None, None, None, None, None, None, None, None,
                                None, 0, 0, 1e-3, 0.9, 0.999, 1e-8, 1000.0, None, None, 0, None, 0, None)
    assert rc == -1
    assert lib.vlg_workspace_status(None, None, None) == -1
    assert lib.vlg_curve_energy(None, 1, 50, 1, 1, 2, 1, 1, None, None, None, None, None, None, None, 0, 0, 0, None, None, 0,
                                None, 0, None) == -1


def test_selftest_kernels_are_not_in_the_product_library(built_lib):
    import subprocess
    import vlg_b200
    syms = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True).stdout
    assert "vlg_selftest" not in syms and "mma_rate" not in syms
    st = vlg_b200.build.build_selftest()
    syms = subprocess.run(["nm", "-D", "--defined-only", str(st)], capture_output=True, text=True).stdout
    for name in vlg_b200._lib.SELFTEST_SIGNATURES:
        assert name in syms


def test_draws_are_range_checked_on_the_host(built_lib):
    import torch
    from vlg_b200 import api
    ok = torch.randint(0, 3, (2, 2, 2, 9, 4))
    assert api._prep_draws(ok, 4, 2, 2, 10, "cpu", 3).shape == (4, 2, 2, 2, 9)
    with pytest.raises(api._lib.VlgError):
        api._prep_draws(ok, 4, 2, 2, 10, "cpu", 2)          # a draw of 2 with two active decoders
    with pytest.raises(api._lib.VlgError):
        api._prep_draws(ok - 1, 4, 2, 2, 10, "cpu", 3)      # negative
    with pytest.raises(api._lib.VlgError):
        api._prep_draws(ok.float(), 4, 2, 2, 10, "cpu", 3)  # not an integer tensor


def test_cpu_tensors_are_rejected_loudly(built_lib):
    import torch
    import vlg_b200
    with pytest.raises(vlg_b200.VlgError):
        vlg_b200.DecoderEnsemble.from_arrays(torch.zeros(1, 128, 2), torch.zeros(1, 128), torch.zeros(1, 128, 128),
                                             torch.zeros(1, 128), torch.zeros(1, 50, 128), torch.zeros(1, 50), "cpu")


def test_nullspace_basis_spans_the_reference_one(built_lib):
    import numpy as np
    import vlg_b200
    from tests import helpers as Hh
    g = Hh.load("nullspace_basis")
    for n in (2, 4, 8):
        basis, C = vlg_b200.construct_nullspace_basis(n)
        basis, C = basis.numpy().astype(np.float64), C.numpy().astype(np.float64)
        ref = g[f"basis_{n}"].astype(np.float64)
        assert basis.shape == ref.shape
        assert np.abs(C - g[f"C_{n}"]).max() == 0
        assert np.abs(C @ basis).max() < 1e-6
        assert np.abs(basis.T @ basis - np.eye(basis.shape[1])).max() < 1e-6
        # same column space: projecting one basis onto the other loses nothing
        assert np.abs(ref @ (ref.T @ basis) - basis).max() < 1e-5
