"""CPU oracle -- test infrastructure only (see geodesic_oracle.py)."""
