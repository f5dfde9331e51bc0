"""68-point landmark regressor (build-defined).

The reference calls an opaque external TF SavedModel through model.signatures["predict"]
(prediction.py:84): uint8 RGB [1,128,128,3] in, >=136 floats normalised to the square face box out.  Its
architecture is not in the reference repository, so this build defines the stand-in from the reference's
own vocabulary: vanilla_encoder (fcn.py:10-51) @128x128x3 -> Flatten (H,W,C) -> Dense(136), with the uint8
input scaled by 1/255 (folded into the first conv's weights).
"""
from .. import _native as N
from .model import Graph, Model


def landmark_regressor(input_height=128, input_width=128, channels=3, n_points=68):
    g = Graph(input_height, input_width, channels)
    x = 0
    for i, f in enumerate((64, 128, 256, 256, 256), start=1):
        x = g.conv(x, "conv%d" % i, f, 3, pad=(1, 1, 1, 1), act=N.ACT_RELU, pool=2, bias=True, bn=True, bn_name="bn%d" % i,
                   in_scale=(1.0 / 255.0) if i == 1 else 1.0)
    g.dense(x, "fc", 2 * n_points)
    m = Model(g, "regression", model_name="landmark_regressor", in_dtype="uint8")
    return m
