"""Heat-map decode wrappers (metrics.py) and drawing helpers (plots.py)."""
