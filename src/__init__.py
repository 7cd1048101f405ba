"""Drop-in entry points with the reference's flags and file layouts
(`python -m src.optimize`, `python -m src.eval`, `python -m src.init_splines_ensemble`),
running on the B200 engine in ``vlg_b200``."""
