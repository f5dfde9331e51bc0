"""Inference-side subset of the reference data package (see generator.py)."""
