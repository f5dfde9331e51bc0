class KeypointException(Exception):
    """Mirror of reference exceptions/__init__.py:3-4."""
