"""SMPL model tensors: loaders for user-supplied files and a seeded synthetic generator.

The licensed SMPL data the reference points at (`PlayerReconstruction/config.py:3-8`:
`additional/smpl/SMPL_NEUTRAL.pkl`, `J_regressor_extra.npy`, `cocoplus_regressor.npy`,
`J_regressor_h36m.npy`) is not in the reference tree, so benches and
tests run on a synthetic model with SMPL's exact shapes / dtypes / sparsity structure
(SURVEY.md Appendix B.3).  This module is *model data*, not the algorithm: it is used by the
product (bench, smoke) and by the oracle alike.

Field names follow what `smplx.SMPL.__init__` registers plus the
three extra regressors of `models/smpl_official.py:17-25`.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, Optional

import numpy as np

NUM_VERTS = 6890
NUM_JOINTS = 24
NUM_BETAS = 10
NUM_POSE_FEATS = 9 * (NUM_JOINTS - 1)  # 207
NUM_FACES = 13776

# SMPL kinematic tree, kintree_table[0] with root -> -1. Bit-exact.
SMPL_PARENTS = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int64)

# smplx.vertex_ids['smplh'] in VertexJointSelector order:
# nose, reye, leye, rear, lear, LBigToe, LSmallToe, LHeel, RBigToe, RSmallToe, RHeel,
# lthumb, lindex, lmiddle, lring, lpinky, rthumb, rindex, rmiddle, rring, rpinky
SMPL_EXTRA_JOINT_VERTEX_IDS = np.array(
    [332, 6260, 2800, 4071, 583,
     3216, 3226, 3387, 6617, 6624, 6787,
     2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133], dtype=np.int64)

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_TOPOLOGY = os.path.join(_DATA_DIR, "smpl_topology_faces.npz")


def load_topology_faces() -> Optional[np.ndarray]:
    """(13774,3) int64 faces over the 6890 SMPL vertices (derived fixture, see
    scripts/make_topology_fixture.py), or None when the fixture is not shipped."""
    if not os.path.exists(_TOPOLOGY):
        return None
    return np.load(_TOPOLOGY)["faces"].astype(np.int64)


def _graph_parts(faces: np.ndarray, nv: int, nparts: int, rng: np.random.Generator):
    """Partition the mesh graph into `nparts` connected regions (multi-source BFS from
    farthest-point seeds).  Gives synthetic skinning weights the vertex-index locality of
    the real SMPL mesh."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import breadth_first_order, shortest_path  # noqa: F401

    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    e = np.concatenate([e, e[:, ::-1]], 0)
    adj = sp.csr_matrix((np.ones(len(e), np.int8), (e[:, 0], e[:, 1])), shape=(nv, nv))
    adj.data[:] = 1

    def bfs_dist(src):
        d = np.full(nv, np.iinfo(np.int32).max, np.int64)
        d[src] = 0
        frontier = np.array([src])
        level = 0
        while len(frontier):
            level += 1
            nb = np.unique(adj[frontier].indices)
            nb = nb[d[nb] > level]
            d[nb] = level
            frontier = nb
        return d

    seeds = [int(rng.integers(nv))]
    dmin = bfs_dist(seeds[0])
    dists = [dmin.copy()]
    while len(seeds) < nparts:
        s = int(np.argmax(np.where(dmin < np.iinfo(np.int32).max, dmin, -1)))
        seeds.append(s)
        d = bfs_dist(s)
        dists.append(d)
        dmin = np.minimum(dmin, d)
    dists = np.stack(dists, 0)                     # (nparts, nv)
    part = np.argmin(dists, 0)
    return part, dists, adj


def make_synthetic_smpl(seed: int = 1234, use_topology: bool = True, statistics: str = "compact") -> Dict[str, np.ndarray]:
    """Seeded synthetic SMPL-shaped model.

    `statistics="wide"` is the harder variant for the packed layouts (the number of virtual joint rows and the
    skinning plan's slot reloads depend on the sparsity structure): regressor rows that draw from many mesh parts
    (H36M 6-10 parts and ~200 non-zeros, COCO-plus 3-6 parts, extra 2-3, J_regressor own part + tree neighbours)
    and skinning rows with 1, 2 or 3 influences next to the usual 4 (real SMPL weights have such rows).

    Magnitudes are body-like so that the 1e-5 m / 1e-4 m tolerances are meaningful:
    template extents ~[0.25, 0.45, 0.12] m, shapedirs sigma 0.03*0.7^l, posedirs sigma 0.002.
    lbs_weights has exactly 4 non-zeros per row (an SMPL property), J_regressor ~32 nnz/row,
    extra / cocoplus rows ~32 nnz, h36m rows ~200 nnz; every regressor row sums to 1.
    """
    rng = np.random.default_rng(seed)
    V, J = NUM_VERTS, NUM_JOINTS
    parents = SMPL_PARENTS.copy()
    faces_topo = load_topology_faces() if use_topology else None

    if faces_topo is not None:
        part, dists, _ = _graph_parts(faces_topo, V, J, rng)
        # relabel parts so that part ids follow a tree-ish order (closest seed pairs adjacent)
        faces = np.concatenate([faces_topo, faces_topo[:2]], 0)            # 13776 rows
    else:
        # index-contiguous parts of uneven size
        cuts = np.sort(rng.choice(np.arange(1, V), J - 1, replace=False))
        part = np.searchsorted(cuts, np.arange(V), side="right")
        dists = None
        faces = rng.integers(0, V, size=(NUM_FACES, 3)).astype(np.int64)

    # tree neighbours of each joint (parent + children)
    nbrs = [[] for _ in range(J)]
    for j in range(1, J):
        nbrs[j].append(int(parents[j]))
        nbrs[int(parents[j])].append(j)

    centers = rng.standard_normal((J, 3)) * np.array([0.25, 0.45, 0.12])
    v_template = centers[part] + rng.standard_normal((V, 3)) * 0.04

    sig = 0.03 * 0.7 ** np.arange(NUM_BETAS)
    shapedirs = rng.standard_normal((V, 3, NUM_BETAS)) * sig
    posedirs = rng.standard_normal((NUM_POSE_FEATS, V * 3)) * 0.002

    # skinning weights, exactly 4 nnz per row.  With the mesh topology: the 4 graph-nearest part
    # seeds with smooth distance-decaying weights (spatially coherent like real SMPL weights, so
    # neighbouring vertex ids mostly share their joint set).  Without: own part + 3 tree neighbours.
    lbs_weights = np.zeros((V, J), np.float64)
    if dists is not None:
        order = np.argsort(dists, axis=0, kind="stable")[:4].T           # (V,4) joint ids
        d4 = np.take_along_axis(dists.T, order, axis=1).astype(np.float64)
        w = np.exp(-(d4 - d4[:, :1]) / 6.0) * (0.9 + 0.2 * rng.random((V, 4)))
        w /= w.sum(1, keepdims=True)
        np.put_along_axis(lbs_weights, order, w, axis=1)
    else:
        for v in range(V):
            j0 = int(part[v])
            cand = list(dict.fromkeys(nbrs[j0] + [k for n in nbrs[j0] for k in nbrs[n] if k != j0]))
            while len(cand) < 3:
                c = int(rng.integers(J))
                if c != j0 and c not in cand:
                    cand.append(c)
            others = rng.choice(cand, 3, replace=False)
            w = rng.dirichlet([4.0, 1.0, 0.6, 0.3])
            lbs_weights[v, j0] = w[0]
            lbs_weights[v, others] = w[1:]

    wide = statistics == "wide"
    if statistics not in ("compact", "wide"):
        raise ValueError("statistics must be 'compact' or 'wide'")
    if wide:
        # rows with fewer than 4 influences: drop the smallest weights of ~half of the vertices and renormalise
        keepn = rng.choice([1, 2, 3, 4], size=V, p=[0.15, 0.15, 0.2, 0.5])
        order_w = np.argsort(-lbs_weights, axis=1)
        for v in range(V):
            lbs_weights[v, order_w[v, keepn[v]:]] = 0.0
        lbs_weights /= lbs_weights.sum(1, keepdims=True)

    def sparse_rows(nrows, nnz, parts_per_row):
        R = np.zeros((nrows, V), np.float64)
        for r in range(nrows):
            npart = parts_per_row if np.isscalar(parts_per_row) else int(rng.integers(parts_per_row[0], parts_per_row[1] + 1))
            ps = rng.choice(J, npart, replace=False)
            pool = np.nonzero(np.isin(part, ps))[0]
            if len(pool) < nnz:
                pool = np.arange(V)
            idx = rng.choice(pool, nnz, replace=False)
            w = rng.random(nnz) + 0.05
            R[r, idx] = w / w.sum()
        return R

    J_regressor = np.zeros((J, V), np.float64)
    for j in range(J):
        pool = np.nonzero(np.isin(part, [j] + nbrs[j]) if wide else part == j)[0]
        if len(pool) < 32:
            pool = np.arange(V)
        idx = rng.choice(pool, 32, replace=False)
        w = rng.random(32) + 0.05
        J_regressor[j, idx] = w / w.sum()

    model = dict(
        v_template=v_template.astype(np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=posedirs.astype(np.float32),
        J_regressor=J_regressor.astype(np.float32),
        lbs_weights=lbs_weights.astype(np.float32),
        parents=parents,
        faces=faces.astype(np.int64),
        extra_joints_idxs=SMPL_EXTRA_JOINT_VERTEX_IDS.copy(),
        J_regressor_extra=sparse_rows(9, 32, (2, 3) if wide else 1).astype(np.float32),
        J_regressor_cocoplus=sparse_rows(19, 64 if wide else 32, (3, 6) if wide else 2).astype(np.float32),
        J_regressor_h36m=sparse_rows(17, 200, (6, 10) if wide else 2).astype(np.float32),
    )
    return model


def _to_np(x):
    """chumpy arrays (official .pkl) expose `.r`; scipy sparse exposes `.toarray()`."""
    if hasattr(x, "toarray"):
        x = x.toarray()
    if hasattr(x, "r") and not isinstance(x, np.ndarray):
        x = x.r
    return np.asarray(x)


class _ChumpyStub:
    """Stands in for `chumpy.ch.Ch` when chumpy is not installed: the official SMPL .pkl stores v_template,
    shapedirs, posedirs, weights, J as chumpy arrays, and unpickling them only needs their state -- the value of
    a plain `Ch` leaf is the ndarray in its `x` attribute, which is what `.r` returns."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"x": state})

    @property
    def r(self):
        if "x" not in self.__dict__:
            raise ValueError("chumpy object without a stored value: install chumpy to load this file")
        return np.asarray(self.__dict__["x"])


class _TolerantUnpickler(pickle.Unpickler):
    """pickle.load(f, encoding='latin1') as smplx does, except that classes of a missing `chumpy` package resolve to
    `_ChumpyStub` instead of raising ModuleNotFoundError."""

    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            if module.split(".")[0] == "chumpy":
                return _ChumpyStub
            raise


def load_smpl_file(path: str) -> Dict[str, np.ndarray]:
    """Load an official SMPL file (`.pkl` as read by smplx, latin1, or its `.npz` conversion).

    Reproduces the buffer construction of `smplx.SMPL.__init__`:
    posedirs -> reshape(6890*3, 207).T ; parents = kintree_table[0] with [0] = -1.
    Only the first 10 shape components are kept (the reference uses num_betas=10).
    """
    if path.endswith(".npz"):
        raw = dict(np.load(path, allow_pickle=True))
    else:
        with open(path, "rb") as f:
            raw = _TolerantUnpickler(f, encoding="latin1").load()
    shapedirs = _to_np(raw["shapedirs"])[:, :, :NUM_BETAS]
    posedirs = _to_np(raw["posedirs"])
    if posedirs.ndim == 3:                       # (6890,3,207) -> (207, 20670)
        posedirs = posedirs.reshape(-1, posedirs.shape[-1]).T
    kin = _to_np(raw["kintree_table"]).astype(np.int64)
    parents = kin[0].copy()
    parents[0] = -1
    return dict(
        v_template=_to_np(raw["v_template"]).astype(np.float32),
        shapedirs=shapedirs.astype(np.float32),
        posedirs=np.ascontiguousarray(posedirs).astype(np.float32),
        J_regressor=_to_np(raw["J_regressor"]).astype(np.float32),
        lbs_weights=_to_np(raw["weights"]).astype(np.float32),
        parents=parents,
        faces=_to_np(raw["f"]).astype(np.int64),
        extra_joints_idxs=SMPL_EXTRA_JOINT_VERTEX_IDS.copy(),
    )


def load_smpl_model(model_path: str, gender: str = "neutral",
                    extra_regressor_paths: Optional[Dict[str, str]] = None) -> Dict[str, np.ndarray]:
    """`model_path` is a file or a directory holding SMPL_{GENDER}.pkl|npz, like
    `smplx.SMPL(model_path=...)` accepts (`models/smpl_official.py:15-16`).  The three extra
    regressors are read from `extra_regressor_paths` (keys `J_regressor_extra`,
    `J_regressor_cocoplus`, `J_regressor_h36m`; defaults = `config.py` paths)."""
    from . import config

    if os.path.isdir(model_path):
        for ext in ("pkl", "npz"):
            cand = os.path.join(model_path, "SMPL_{}.{}".format(gender.upper(), ext))
            if os.path.exists(cand):
                model_path = cand
                break
        else:
            raise FileNotFoundError("no SMPL_{}.pkl|npz under {}".format(gender.upper(), model_path))
    model = load_smpl_file(model_path)
    paths = dict(J_regressor_extra=config.J_REGRESSOR_EXTRA_PATH,
                 J_regressor_cocoplus=config.COCOPLUS_REGRESSOR_PATH,
                 J_regressor_h36m=config.H36M_REGRESSOR_PATH)
    paths.update(extra_regressor_paths or {})
    for k, p in paths.items():
        model[k] = np.load(p).astype(np.float32)
    return model


def validate_model(model: Dict[str, np.ndarray]) -> None:
    V = model["v_template"].shape[0]
    J = model["J_regressor"].shape[0]
    assert model["v_template"].shape == (V, 3)
    assert model["shapedirs"].shape[:2] == (V, 3)
    assert model["posedirs"].shape == (9 * (J - 1), V * 3)
    assert model["lbs_weights"].shape == (V, J)
    p = np.asarray(model["parents"])
    assert p.shape == (J,) and p[0] == -1 and np.all(p[1:] < np.arange(1, J)) and np.all(p[1:] >= 0)
    for k in ("J_regressor_extra", "J_regressor_cocoplus", "J_regressor_h36m"):
        if k in model:
            assert model[k].shape[1] == V
