"""Drop-in for the one `smplx.lbs` function the reference's scripts import directly
(`from smplx.lbs import batch_rodrigues`, PlayerReconstruction/predict/predict_3D.py:5; used at
player_recon.py:201,655 and hmr.py:207 to turn axis-angle poses into the rotation matrices passed to
`SMPL(..., pose2rot=False)`).  CUDA float32 through the C-ABI, differentiable; there is no CPU path."""
from .ops import batch_rodrigues  # noqa: F401
