"""CPU oracle (test infrastructure only -- see regt_oracle.py header)."""
