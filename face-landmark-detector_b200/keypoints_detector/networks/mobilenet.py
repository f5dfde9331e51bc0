"""MobileNet-v1 encoder (reference networks/mobilenet.py:59-114).  SURVEY §8 row f3 ("next"): the depthwise
3x3 / stride-2 kernels are not built yet; the builder fails loudly instead of silently using another path."""


def get_mobilenet_encoder(input_height=224, input_width=224, pretrained=None, channels=3):
    raise NotImplementedError("fcn_*_mobilenet: depthwise-conv CUDA kernels are a 'next' row (SURVEY §8 f3), not built yet")
