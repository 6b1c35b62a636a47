"""Test infrastructure only (see denseclip_oracle.py header)."""
