"""CPU oracle package -- test infrastructure only (see air_oracle.py header)."""
