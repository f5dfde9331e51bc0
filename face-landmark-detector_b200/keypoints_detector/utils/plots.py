"""Minimal host-side drawing helpers so _prediction / video_predict keep working (reference
utils/plots.py:6-7,60-72,116-166).  Out of the hot path; cv2 + numpy only, no matplotlib."""
import random

import numpy as np

random.seed(0)
class_colors = [(random.randint(0, 255), random.randint(0, 255), random.randint(0, 255)) for _ in range(5000)]


def get_colored_segmentation_image(seg_arr, n_classes, colors=class_colors):
    """plots.py:60-72 (vectorised): class map (h,w) -> BGR uint8-valued float image."""
    seg_arr = np.asarray(seg_arr)
    lut = np.asarray(colors[:max(int(n_classes), 1)], dtype=np.float64)
    idx = np.clip(seg_arr, 0, len(lut) - 1)
    out = lut[idx]
    out[seg_arr >= n_classes] = 0
    return out


def overlay_seg_image(inp_img, seg_img):
    import cv2
    seg_img = cv2.resize(seg_img, (inp_img.shape[1], inp_img.shape[0]), interpolation=cv2.INTER_NEAREST)
    return (inp_img / 2 + seg_img / 2).astype('uint8')


def visualize_keypoints(kpts_arr, inp_img=None, n_classes=None, colors=class_colors, class_names=None,
                        overlay_img=False, show_legends=False, pred_dim=None):
    """plots.py:116-149; pred_dim=None (the reference's default, on which it crashes at :123) means
    "no final resize"."""
    import cv2
    prediction_width, prediction_height = pred_dim if pred_dim is not None else (None, None)
    if n_classes is None:
        n_classes = np.max(kpts_arr)
    seg_img = get_colored_segmentation_image(kpts_arr, n_classes, colors=colors)
    if inp_img is not None:
        seg_img = cv2.resize(seg_img, (inp_img.shape[1], inp_img.shape[0]), interpolation=cv2.INTER_NEAREST)
    if (prediction_height is not None) and (prediction_width is not None):
        seg_img = cv2.resize(seg_img, (prediction_width, prediction_height), interpolation=cv2.INTER_NEAREST)
        if inp_img is not None:
            inp_img = cv2.resize(inp_img, (prediction_width, prediction_height))
    if overlay_img:
        assert inp_img is not None
        seg_img = overlay_seg_image(inp_img, seg_img)
    if show_legends:
        assert class_names is not None
        legend = np.zeros((len(class_names) * 25 + 25, 125, 3), dtype="uint8") + 255
        for i, (name, color) in enumerate(zip(class_names, colors)):
            cv2.putText(legend, str(name), (5, i * 25 + 17), cv2.FONT_HERSHEY_COMPLEX, 0.5, (0, 0, 0), 1)
            cv2.rectangle(legend, (100, i * 25), (125, i * 25 + 25), tuple(int(c) for c in color), -1)
        h = max(seg_img.shape[0], legend.shape[0])
        out = np.zeros((h, seg_img.shape[1] + legend.shape[1], 3), dtype=seg_img.dtype)
        out[:legend.shape[0], :legend.shape[1]] = legend
        out[:seg_img.shape[0], legend.shape[1]:] = seg_img
        seg_img = out
    return seg_img


def draw_marks(image, marks, color=(0, 255, 0)):
    """plots.py:152-166."""
    import cv2
    for mark in marks:
        cv2.circle(image, (int(mark[0]), int(mark[1])), 2, color, -1, cv2.LINE_AA)
