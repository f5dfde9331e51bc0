"""Tensor ordering switch of the reference (networks/config.py).  The CUDA kernels are NHWC-only, so the
one supported value is 'channels_last'; the second constant exists because callers compare against it."""
_ORDERINGS = ("channels_last", "channels_first")
IMAGE_ORDERING_CHANNELS_LAST, IMAGE_ORDERING_CHANNELS_FIRST = _ORDERINGS
IMAGE_ORDERING = _ORDERINGS[0]   # the reference's default as well; its MobileNet builder asserts it (mobilenet.py:64-67)
