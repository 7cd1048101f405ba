"""Recipe for oracle/_ref: the UNMODIFIED reference hot-path modules, made importable where the
reference checkout is not mounted (the GPU box), so that bench.py's reference arm and its
`cpu_baseline` / `gpu_eager_baseline` legs time the reference's OWN code
(`cpu_baseline.kind = "reference"`), not a port.

    python oracle/make_ref.py            # needs /root/reference (build container); idempotent

What it does: copies, byte for byte, the modules on the hot path from
/root/reference/src into oracle/_ref/src/ -- `optimize.py` (GeodesicSplineBatch, compute_energy_mc),
`train.py` (EVAE / GaussianDecoder / make_decoder_net), `single_decoder/{optimize_energy,
optimize_energy_batched,vae}.py` -- and writes empty stub modules for the plotting-only imports the
reference makes at module top (matplotlib, seaborn, mpl_toolkits: not installed in this image; never
called on the hot path).  oracle/_ref/ is git-ignored (reference sources are NOT committed to this
repo) but travels to the GPU box with the snapshot, like the built .so files.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.
"""
from __future__ import annotations

import shutil
import sys
import types
from pathlib import Path

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
DST = HERE / "_ref"
MODULES = ["src/optimize.py", "src/train.py", "src/single_decoder/optimize_energy.py",
           "src/single_decoder/optimize_energy_batched.py", "src/single_decoder/vae.py"]
STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm", "seaborn", "mpl_toolkits",
         "mpl_toolkits.axes_grid1"]


def make(verbose: bool = True) -> bool:
    """Populate oracle/_ref from /root/reference.  Returns False (and leaves any existing copy alone)
    when the reference checkout is not present."""
    if not (REF / "src" / "optimize.py").exists():
        if verbose:
            print(f"[make_ref] {REF} not present: keeping {DST} as is ({'exists' if DST.exists() else 'absent'})")
        return DST.exists()
    for m in MODULES:
        dst = DST / m
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / m, dst)
    # the reference has no __init__.py files (namespace packages); empty ones are generated here so that
    # oracle/_ref/src wins over this repo's own drop-in `src` package while the reference is being imported
    for pkg in ("src", "src/single_decoder"):
        (DST / pkg / "__init__.py").write_text("")
    (DST / "SOURCE.txt").write_text(
        "Unmodified copies of johannefranck/vae-latent-geometry modules, made by oracle/make_ref.py from\n"
        f"{REF}; git-ignored; used only by bench.py's reference arm and tests.\n" + "\n".join(MODULES) + "\n")
    if verbose:
        print(f"[make_ref] copied {len(MODULES)} reference modules into {DST}")
    return True


def available() -> bool:
    return all((DST / m).exists() for m in MODULES)


def import_reference():
    """-> (ref_optimize module, ref_train module, ref_single_batched module) of the UNMODIFIED reference,
    loaded from oracle/_ref under a private package name (the repo's own drop-in package is also called
    `src`, so the reference's `src` is imported with sys.path/sys.modules swapped and then restored)."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference is mounted")
    for name in STUBS:   # plotting-only imports at the reference's module tops: stub what is not installed
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(DST))
    try:
        import src.optimize as ref_opt
        import src.single_decoder.optimize_energy_batched as ref_sb
        import src.train as ref_train
        mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    finally:
        sys.path.remove(str(DST))
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    # keep the reference modules reachable under a name that cannot collide with the drop-in `src`
    for k, v in mods.items():
        sys.modules["vlg_reference." + k] = v
    return ref_opt, ref_train, ref_sb


if __name__ == "__main__":
    ok = make()
    sys.exit(0 if ok else 1)
