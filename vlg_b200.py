"""Import shim: the package directory is named ``vae-latent-geometry_b200`` (not a valid
Python identifier), so ``import vlg_b200`` loads it from there under this name."""
import importlib.util
import pathlib
import sys

_pkg = pathlib.Path(__file__).resolve().parent / "vae-latent-geometry_b200"
_spec = importlib.util.spec_from_file_location("vlg_b200", _pkg / "__init__.py",
                                               submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vlg_b200"] = _mod
_spec.loader.exec_module(_mod)
