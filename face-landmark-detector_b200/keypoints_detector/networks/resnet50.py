"""ResNet50 encoder (reference networks/resnet50.py:122-182).  SURVEY §8 row f3 ("next"): residual epilogues,
7x7 stride-2 stem and 3x3 stride-2 pool kernels are not built yet; the builder fails loudly."""


def get_resnet50_encoder(input_height=224, input_width=224, pretrained=None, channels=3):
    raise NotImplementedError("fcn_*_resnet50: residual / strided CUDA kernels are a 'next' row (SURVEY §8 f3), not built yet")
