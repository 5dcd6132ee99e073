"""Seeded synthetic inputs for tests/ and bench.py (no HG002 data and no network here).
Test and bench infrastructure: nothing under vcfdist_b200/ imports this package."""
