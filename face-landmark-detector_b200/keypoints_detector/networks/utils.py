"""get_segmentation_model — mirror of reference networks/utils.py:6-39: append the Reshape + per-pixel
softmax over classes and attach the attributes the prediction/training code reads."""
from .config import IMAGE_ORDERING
from .model import Model


def get_segmentation_model(graph, output):
    assert IMAGE_ORDERING == "channels_last"
    assert output == len(graph.layers), "the segmentation head must be the last recorded layer"
    graph.softmax(output, "softmax")            # utils.py:28-30 (Reshape is a no-op on NHWC memory)
    model = Model(graph, "segmentation", model_name="", in_dtype="float32")
    return model
