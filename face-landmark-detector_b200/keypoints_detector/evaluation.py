"""Placeholder of the reference module of the same name (empty there too)."""
