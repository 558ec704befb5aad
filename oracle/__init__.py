"""Test infrastructure only: CPU oracle for the MMTG hot path (see mmtg_oracle.py header)."""
