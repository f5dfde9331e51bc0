"""ResNet50 encoder (reference networks/resnet50.py:23-182).

Layer names are the reference's: conv1 / bn_conv1, res{stage}{block}_branch{2a,2b,2c,1} / bn{stage}{block}_branch...
Every conv carries a bias (Keras default) and is followed by a BatchNormalization, folded here; the residual
`add -> relu` (:68-69, :117-118) is one ADD layer with a ReLU epilogue.  In bf16 mode every conv runs on the tensor cores:
the 7x7 stride-2 stem through shared-memory im2col (csrc/tc_conv_stem.cu), the stride-2 1x1 convs (first conv of the main
branch and the shortcut, :98,110) through TMA element strides that pick every second pixel (csrc/tc_conv.cu), the stride-1
convs (Cin % 64 == 0) through the halo / per-tap TMA kernels.  The 3x3/2 max-pool and the residual adds are 16-byte
vectorised CUDA-core kernels (csrc/simt_extra.cu).

Levels: the reference returns f1 = conv1 output BEFORE bn_conv1 (:147) and f2 = one_side_pad(stage 2) (:153); no FCN
head of the reference consumes them (fcn_8 uses f3, f4, f5; fcn_32 uses f5), so f1 / f2 here are the tensors after
the fused BN + ReLU and without the one-sided pad.  f3, f4, f5 are exactly the reference's.

`pretrained='imagenet'` downloads weights in the reference (:175-179); any value other than None raises here.
"""
from .. import _native as N
from .config import IMAGE_ORDERING
from .model import Graph


def _block(g, x, filters, stage, block, stride, shortcut):
    f1, f2, f3 = filters
    cb = "res%d%s_branch" % (stage, block)
    bb = "bn%d%s_branch" % (stage, block)
    y = g.conv(x, cb + "2a", f1, 1, stride=stride, act=N.ACT_RELU, bias=True, bn=True, bn_name=bb + "2a")
    y = g.conv(y, cb + "2b", f2, 3, pad="same", act=N.ACT_RELU, bias=True, bn=True, bn_name=bb + "2b")
    y = g.conv(y, cb + "2c", f3, 1, act=N.ACT_NONE, bias=True, bn=True, bn_name=bb + "2c")
    sc = x
    if shortcut:
        sc = g.conv(x, cb + "1", f3, 1, stride=stride, act=N.ACT_NONE, bias=True, bn=True, bn_name=bb + "1")
    return g.add(y, sc, "add%d%s" % (stage, block), act=N.ACT_RELU)


def get_resnet50_encoder(input_height=224, input_width=224, pretrained=None, channels=3, graph=None):
    assert IMAGE_ORDERING == "channels_last", "Currently only channels last mode is supported"
    assert input_height % 32 == 0
    assert input_width % 32 == 0
    if pretrained is not None:
        raise ValueError("pretrained=%r needs a download (resnet50.py:175-179); load converted weights with "
                         "model.load_weights() instead" % (pretrained,))
    g = graph or Graph(input_height, input_width, channels)
    # ZeroPadding2D((3,3)) + Conv2D(64, 7x7, strides 2) + bn_conv1 + relu (:143-148)
    x = g.conv(0, "conv1", 64, 7, pad=(3, 3, 3, 3), stride=2, act=N.ACT_RELU, bias=True, bn=True, bn_name="bn_conv1")
    f1 = x
    x = g.maxpool(x, "pool1", 3, 2)                                                # :149
    levels = [f1]
    for stage, filters, blocks, stride in ((2, (64, 64, 256), "abc", 1), (3, (128, 128, 512), "abcd", 2),
                                           (4, (256, 256, 1024), "abcdef", 2), (5, (512, 512, 2048), "abc", 2)):
        for bi, b in enumerate(blocks):
            x = _block(g, x, filters, stage, b, stride if bi == 0 else 1, bi == 0)
        levels.append(x)
    return g, levels
