"""Reference-named entry points of `utils/joints2d_utils.py` (reference lines 5-32)."""
import torch


def undo_keypoint_normalisation(normalised_keypoints, img_wh):
    """[-1, 1] -> pixels.  One FMA per element; fused into ops.orthographic_project(pixel_wh=...) and
    ops.joints2d_loss on the hot path, kept here as a plain tensor op for API parity."""
    return (normalised_keypoints + 1) * (img_wh / 2.0)


def check_joints2d_visibility_torch(joints2d, img_wh):
    """Boolean mask of keypoints inside the image."""
    x, y = joints2d[:, :, 0], joints2d[:, :, 1]
    return ~((x > img_wh) | (y > img_wh) | (x < 0) | (y < 0))
