"""Reference-named entry point of `utils/rigid_transform_utils.py:27-41`."""
from . import ops


def rot6d_to_rotmat(x):
    """6D rotation representation (Zhou et al.) -> rotation matrices, (B, 6k) -> (B*k, 3, 3)."""
    return ops.rot6d_to_rotmat(x)
