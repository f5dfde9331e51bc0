"""Mirror of reference data/config.py:2-6 — the one configuration option of the reference."""
IMAGE_ORDERING_CHANNELS_LAST = "channels_last"
IMAGE_ORDERING_CHANNELS_FIRST = "channels_first"

# Default IMAGE_ORDERING = channels_last (the only ordering the CUDA kernels implement; MobileNet in the
# reference already asserts it, mobilenet.py:64-67)
IMAGE_ORDERING = IMAGE_ORDERING_CHANNELS_LAST
