"""LANDMARKS_MODELS registry — mirror of reference networks/basic_models.py:59-64.

The reference's 'default' builder (`build_model`, basic_models.py:7-56) cannot be built as written (float
kernel size :32, Reshape/transpose-conv size mismatch :45-50) and lacks the attributes every caller reads
(SURVEY App. A end); the key is kept and mapped to fcn_8 over the vanilla encoder.  'landmark_regressor' and
'fcn_8_vanilla' / 'fcn_32_vanilla' are additions.
"""
from .fcn import (fcn_8, fcn_32, fcn_8_resnet50, fcn_8_mobilenet, fcn_8_vgg, fcn_8_vanilla, fcn_32_vgg, fcn_32_resnet50,
                  fcn_32_mobilenet, vanilla_encoder)
from .regression import landmark_regressor


def build_model(n_classes, input_height=224, input_width=224):
    model = fcn_8(n_classes, vanilla_encoder, input_height=input_height, input_width=input_width)
    model.model_name = "default"
    return model


def fcn_32_vanilla(n_classes, input_height=224, input_width=224, channels=3):
    model = fcn_32(n_classes, vanilla_encoder, input_height=input_height, input_width=input_width, channels=channels)
    model.model_name = "fcn_32_vanilla"
    return model


def landmark_regressor_model(n_classes=136, input_height=128, input_width=128, channels=3):
    """Registry adapter: the regression model's `n_classes` is its output width (2 values per landmark), which is what
    Model.config_dict / save_config record, so save_config -> model_from_checkpoint_path round-trips."""
    return landmark_regressor(input_height, input_width, channels, n_points=int(n_classes) // 2)


LANDMARKS_MODELS = {
    'fcn_8_resnet50': fcn_8_resnet50,
    'fcn_8_mobilenet': fcn_8_mobilenet,
    'fcn_8_vgg': fcn_8_vgg,
    'default': build_model,
    # additions of this build
    'fcn_8_vanilla': fcn_8_vanilla,
    'fcn_32_vanilla': fcn_32_vanilla,
    'fcn_32_vgg': fcn_32_vgg,
    'fcn_32_resnet50': fcn_32_resnet50,
    'fcn_32_mobilenet': fcn_32_mobilenet,
    'landmark_regressor': landmark_regressor_model,
}
