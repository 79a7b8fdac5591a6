"""Synthetic DAB signal generation for tests and benchmarks (not on the product path)."""
