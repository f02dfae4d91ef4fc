"""Reference-named entry points of `utils/cam_utils.py` (reference lines 5-85) on the B200 kernels."""
import numpy as np
import torch

from . import ops


def orthographic_project_torch(points3D, cam_params):
    """Weak-perspective projection u = s (x + tx), v = s (y + ty); cam_params = [s, tx, ty]."""
    return ops.orthographic_project(points3D, cam_params, 0.0)


def convert_weak_perspective_to_camera_translation_torch(cam_wp, focal_length, resolution):
    """[s, tx, ty] -> [tx, ty, 2 f / (res * s + 1e-9)]  (3 floats per body: plain tensor ops)."""
    tz = 2 * focal_length / (resolution * cam_wp[:, 0] + 1e-9)
    return torch.stack([cam_wp[:, 1], cam_wp[:, 2], tz], dim=-1)


def convert_camera_translation_to_weak_perspective_torch(translation, focal_length, resolution):
    s = 2 * focal_length / (resolution * translation[:, 2] + 1e-9)
    return torch.stack([s, translation[:, 0], translation[:, 1]], dim=-1)


def get_intrinsics_matrix(img_width, img_height, focal_length):
    return np.array([[focal_length, 0.0, img_width / 2.0],
                     [0.0, focal_length, img_height / 2.0],
                     [0.0, 0.0, 1.0]])


def perspective_project_torch(points, rotation, translation, cam_K=None, focal_length=None, img_wh=None):
    """X' = R X + t, divide by depth, apply K (reference lines 54-85): either an explicit `cam_K` (bs,3,3) -- any
    intrinsics matrix -- or the (focal_length, img_wh) form the reference calls (player_recon.py:685-688)."""
    if cam_K is not None:
        return ops.perspective_project_camk(points, rotation, translation, cam_K.to(points.device, torch.float32))
    return ops.perspective_project(points, rotation, translation, focal_length, img_wh)
