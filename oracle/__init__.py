"""Test infrastructure: CPU oracle of the Show-and-Tell decoder hot path (see snt_oracle.py header)."""
