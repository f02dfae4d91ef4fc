#!/usr/bin/env python
"""Derive the SMPL-topology face fixture from the reference tree (run in the build container).

Source (read-only): /root/reference/Python/Soccer/PlayerReconstruction/UVTextureConverter/
config/UV_Processed.mat  -- the DensePose atlas table the reference ships.  SURVEY.md section 4:
`All_vertices[All_Faces-1]-1` is a (13774,3) integer face array over all 6890 SMPL vertices.
The official SMPL `f` has 13776 rows in another order, so this does NOT pin `smpl.faces`; it is
a realistic *topology* (vertex-index locality of the real mesh) used (a) by the synthetic model
generator to place skinning weights / regressors with SMPL-like index locality and (b) by the
bit-exact faces passthrough test.

Output: soccerplayershapepose_b200/data/smpl_topology_faces.npz  (int16, ~60 kB compressed)
"""
import os
import sys
import numpy as np
import scipy.io as sio

REF = "/root/reference/Python/Soccer/PlayerReconstruction/UVTextureConverter/config/UV_Processed.mat"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "soccerplayershapepose_b200", "data", "smpl_topology_faces.npz")


def main():
    if not os.path.exists(REF):
        sys.exit("reference fixture not found (only available in the build container): " + REF)
    m = sio.loadmat(REF)
    all_vertices = m["All_vertices"][0].astype(np.int64)      # (7829,) in [1, 6890]
    all_faces = m["All_Faces"].astype(np.int64)               # (13774, 3) 1-based into All_vertices
    faces = all_vertices[all_faces - 1] - 1                   # 0-based SMPL vertex ids
    assert faces.shape == (13774, 3) and faces.min() == 0 and faces.max() == 6889
    assert len(np.unique(faces)) == 6890
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, faces=faces.astype(np.int16))
    print("wrote", os.path.normpath(OUT), faces.shape)


if __name__ == "__main__":
    main()
