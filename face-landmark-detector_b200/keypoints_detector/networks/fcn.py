"""FCN model builders — same names and signatures as reference networks/fcn.py, producing layer lists
for the sm_100a kernels instead of Keras graphs.

Layer semantics follow the reference exactly:
  vanilla_encoder  fcn.py:10-51   5 x (ZeroPadding2D(1) -> Conv2D 3x3 valid + bias -> BN -> ReLU -> MaxPool 2x2)
  crop             fcn.py:55-86   folded into the ADD layer (both operands cropped bottom/right to the smaller)
  fcn_8 / fcn_32   fcn.py:89-150
Dropout layers (fcn.py:99,101) are the identity at inference and are not recorded.
"""
from .. import _native as N
from .config import IMAGE_ORDERING
from .model import Graph
from .utils import get_segmentation_model


def vanilla_encoder(input_height=224, input_width=224, channels=3, graph=None):
    assert IMAGE_ORDERING == "channels_last", "only channels_last is implemented"
    g = graph or Graph(input_height, input_width, channels)
    x = 0
    levels = []
    for i, f in enumerate((64, 128, 256, 256, 256), start=1):
        # ZeroPadding2D((1,1)) + Conv2D(f, 3x3, valid) + BatchNormalization + relu + MaxPooling2D((2,2))
        x = g.conv(x, "conv%d" % i, f, 3, pad=(1, 1, 1, 1), act=N.ACT_RELU, pool=2, bias=True, bn=True, bn_name="bn%d" % i)
        levels.append(x)
    return g, levels


def _head(g, f5):
    # fcn.py:98-103
    o = g.conv(f5, "head7", 4096, 7, pad="same", act=N.ACT_RELU)
    o = g.conv(o, "head1", 4096, 1, act=N.ACT_RELU)
    return o


def fcn_8(n_classes, encoder=vanilla_encoder, input_height=416, input_width=608, channels=3):
    g, levels = encoder(input_height=input_height, input_width=input_width, channels=channels)
    f1, f2, f3, f4, f5 = levels
    o = _head(g, f5)
    o = g.conv(o, "score5", n_classes, 1)                      # :103 (he_normal, bias)
    o = g.deconv(o, "up2a", n_classes, 4, 2)                   # :104
    o2 = g.conv(f4, "score4", n_classes, 1)                    # :107-108
    o = g.add(o, o2, "add4")                                   # :110-112 crop + Add
    o = g.deconv(o, "up2b", n_classes, 4, 2)                   # :114
    o2 = g.conv(f3, "score3", n_classes, 1)                    # :116-117
    o = g.add(o2, o, "seg_feats")                              # :118-119
    o = g.deconv(o, "up8", n_classes, 16, 8)                   # :121
    model = get_segmentation_model(g, o)
    model.model_name = "fcn_8"
    return model


def fcn_32(n_classes, encoder=vanilla_encoder, input_height=416, input_width=608, channels=3):
    g, levels = encoder(input_height=input_height, input_width=input_width, channels=channels)
    f5 = levels[4]
    o = _head(g, f5)
    o = g.conv(o, "score5", n_classes, 1)                      # :142-143 ("seg_feats")
    o = g.deconv(o, "up32", n_classes, 64, 32)                 # :144-145
    model = get_segmentation_model(g, o)
    model.model_name = "fcn_32"
    return model


def _named(builder, encoder_getter, name):
    def f(n_classes, input_height, input_width, channels=3):
        model = builder(n_classes, encoder_getter(), input_height=input_height, input_width=input_width, channels=channels)
        model.model_name = name
        return model
    return f


def _vgg():
    from .vgg16 import get_vgg_encoder
    return get_vgg_encoder


def _mobilenet():
    from .mobilenet import get_mobilenet_encoder
    return get_mobilenet_encoder


def _resnet50():
    from .resnet50 import get_resnet50_encoder
    return get_resnet50_encoder


def fcn_8_vgg(n_classes, input_height=416, input_width=608, channels=3):
    return _named(fcn_8, _vgg, "fcn_8_vgg")(n_classes, input_height, input_width, channels)


def fcn_32_vgg(n_classes, input_height=416, input_width=608, channels=3):
    return _named(fcn_32, _vgg, "fcn_32_vgg")(n_classes, input_height, input_width, channels)


def fcn_8_resnet50(n_classes, input_height=416, input_width=608, channels=3):
    return _named(fcn_8, _resnet50, "fcn_8_resnet50")(n_classes, input_height, input_width, channels)


def fcn_32_resnet50(n_classes, input_height=416, input_width=608, channels=3):
    return _named(fcn_32, _resnet50, "fcn_32_resnet50")(n_classes, input_height, input_width, channels)


def fcn_8_mobilenet(n_classes, input_height=224, input_width=224, channels=3):
    return _named(fcn_8, _mobilenet, "fcn_8_mobilenet")(n_classes, input_height, input_width, channels)


def fcn_32_mobilenet(n_classes, input_height=224, input_width=224, channels=3):
    return _named(fcn_32, _mobilenet, "fcn_32_mobilenet")(n_classes, input_height, input_width, channels)


def fcn_8_vanilla(n_classes, input_height=224, input_width=224, channels=3):
    model = fcn_8(n_classes, vanilla_encoder, input_height=input_height, input_width=input_width, channels=channels)
    model.model_name = "fcn_8_vanilla"
    return model
