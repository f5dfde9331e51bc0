"""VGG16 encoder — mirror of reference networks/vgg16.py:17-81 (2-2-3-3-3 Conv3x3 'same' + bias + ReLU,
MaxPool 2x2 after each block, no BN).  `pretrained` defaults to None: the reference downloads ImageNet
weights (vgg16.py:76-79), which is impossible offline."""
from .. import _native as N
from .config import IMAGE_ORDERING
from .model import Graph


def get_vgg_encoder(input_height=224, input_width=224, pretrained=None, channels=3):
    assert input_height % 32 == 0
    assert input_width % 32 == 0
    assert IMAGE_ORDERING == "channels_last"
    if pretrained is not None:
        raise ValueError("pretrained ImageNet weights cannot be downloaded in this build; load an .npz with model.load_weights")
    g = Graph(input_height, input_width, channels)
    x = 0
    levels = []
    for b, (n, f) in enumerate(((2, 64), (2, 128), (3, 256), (3, 512), (3, 512)), start=1):
        for i in range(1, n + 1):
            x = g.conv(x, "block%d_conv%d" % (b, i), f, 3, pad="same", act=N.ACT_RELU, pool=2 if i == n else 0)
        levels.append(x)
    return g, levels
