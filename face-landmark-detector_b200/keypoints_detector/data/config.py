"""Tensor ordering switch of the reference's data pipeline (data/config.py): re-exported from the networks
package so that both modules can never disagree."""
from ..networks.config import IMAGE_ORDERING, IMAGE_ORDERING_CHANNELS_FIRST, IMAGE_ORDERING_CHANNELS_LAST  # noqa: F401
