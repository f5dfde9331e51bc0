"""MobileNet-v1 encoder (reference networks/mobilenet.py:16-114), alpha = 1, depth_multiplier = 1.

Layer names are the reference's (conv1 / conv1_bn, conv_dw_%d / conv_dw_%d_bn, conv_pw_%d / conv_pw_%d_bn) so a
converted Keras checkpoint maps by name.  ZeroPadding2D((1,1)) is folded into the following conv's padding, the
BatchNormalization into its weights, relu6 into its epilogue.  In bf16 mode the 1x1 point-wise convs run on the
tensor cores; the 3x3 stem (stride 2, Cin = 3) and the depth-wise convs are HBM-bound CUDA-core kernels.

`pretrained='imagenet'` downloads weights in the reference (mobilenet.py:106-112); there is no network here, so
any value other than None raises.
"""
from .. import _native as N
from .config import IMAGE_ORDERING
from .model import Graph

# (point-wise filters, depth-wise stride) for block_id 1..13 (mobilenet.py:79-103)
_BLOCKS = ((64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1), (512, 1),
           (1024, 2), (1024, 1))
_LEVEL_AFTER = (1, 3, 5, 11, 13)


def get_mobilenet_encoder(input_height=224, input_width=224, pretrained=None, channels=3, graph=None):
    assert IMAGE_ORDERING == "channels_last", "Currently only channels last mode is supported"
    assert input_height % 32 == 0
    assert input_width % 32 == 0
    if pretrained is not None:
        raise ValueError("pretrained=%r needs a download (mobilenet.py:106-112); load converted weights with "
                         "model.load_weights() instead" % (pretrained,))
    g = graph or Graph(input_height, input_width, channels)
    # _conv_block (:16-28)
    x = g.conv(0, "conv1", 32, 3, pad=(1, 1, 1, 1), stride=2, act=N.ACT_RELU6, bias=False, bn=True, bn_name="conv1_bn")
    levels = []
    for i, (f, s) in enumerate(_BLOCKS, start=1):
        # _depthwise_conv_block (:31-56)
        x = g.dwconv(x, "conv_dw_%d" % i, 3, pad=(1, 1, 1, 1), stride=s, act=N.ACT_RELU6, bias=False, bn=True,
                     bn_name="conv_dw_%d_bn" % i)
        x = g.conv(x, "conv_pw_%d" % i, f, 1, act=N.ACT_RELU6, bias=False, bn=True, bn_name="conv_pw_%d_bn" % i)
        if i in _LEVEL_AFTER:
            levels.append(x)
    return g, levels
