"""Drop-in for the one `smplx.lbs` function the reference's scripts import directly
(`from smplx.lbs import batch_rodrigues`, PlayerReconstruction/predict/predict_3D.py:5; used at
player_recon.py:201,655 and hmr.py:207 to turn axis-angle poses into the rotation matrices passed to
`SMPL(..., pose2rot=False)`).  CUDA float32 through the C-ABI, differentiable; there is no CPU path."""
from .ops import batch_rodrigues  # noqa: F401


def vertices2joints(J_regressor, vertices):
    """smplx.lbs.vertices2joints (imported by the reference at models/smpl_official.py:5 for its three extra
    regressors): einsum('bik,ji->bjk').  Provided for scripts that call it on their own; the SMPL module does not
    use it -- its regressed joints are folded into extra rows of the blend GEMM at pack time, so no vertex is
    re-read.  A plain library contraction, device-agnostic."""
    import torch
    return torch.einsum("bik,ji->bjk", vertices, J_regressor)
