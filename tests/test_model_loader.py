"""Official-model loader round trip (VERDICT r01 "loader path unexecuted").

The licensed SMPL files are absent, so the synthetic model is written in the OFFICIAL key layout -- a latin1 / protocol-2
pickle with `weights`, `kintree_table` (uint32, root = 2^32-1), `f`, `posedirs (6890,3,207)`, `shapedirs` as an object
exposing `.r` (chumpy), scipy-sparse `J_regressor` -- plus the three `.npy` regressors at the reference's relative paths
(PlayerReconstruction/config.py:3-8).  `SMPL(config.SMPL_MODEL_DIR, batch_size=1)` is then constructed exactly as
the reference does (player_recon.py:147) and compared bit-for-bit with the dict-constructed module.
"""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from soccerplayershapepose_b200 import config
from soccerplayershapepose_b200.model_io import load_smpl_model
from soccerplayershapepose_b200.smpl import SMPL


class FakeCh:
    """Stand-in for chumpy.ch.Ch as found in the official .pkl: the array sits behind `.r`."""

    def __init__(self, a):
        self.r = np.asarray(a)


def write_official_layout(model, root):
    add = os.path.join(root, "PlayerReconstruction", "additional")
    os.makedirs(os.path.join(add, "smpl"))
    V = model["v_template"].shape[0]
    kin = np.zeros((2, 24), np.uint32)
    kin[0] = np.asarray(model["parents"]).astype(np.int64).astype(np.uint32)      # root -> 4294967295
    kin[1] = np.arange(24)
    assert kin[0, 0] == np.uint32(4294967295)
    rng = np.random.default_rng(0)
    sd = np.concatenate([model["shapedirs"].astype(np.float64), rng.standard_normal((V, 3, 2))], 2)   # 12 components
    raw = {
        "v_template": model["v_template"].astype(np.float64),
        "shapedirs": FakeCh(sd),
        "posedirs": model["posedirs"].astype(np.float64).T.reshape(V, 3, 207),
        "J_regressor": sp.csc_matrix(model["J_regressor"].astype(np.float64)),
        "weights": model["lbs_weights"].astype(np.float64),
        "kintree_table": kin,
        "f": model["faces"].astype(np.uint32),
        "bs_style": "lbs", "bs_type": "lrotmin",
    }
    with open(os.path.join(add, "smpl", "SMPL_NEUTRAL.pkl"), "wb") as f:
        pickle.dump(raw, f, protocol=2)
    np.save(os.path.join(add, "J_regressor_extra.npy"), model["J_regressor_extra"])
    np.save(os.path.join(add, "cocoplus_regressor.npy"), model["J_regressor_cocoplus"].astype(np.float64))
    np.save(os.path.join(add, "J_regressor_h36m.npy"), model["J_regressor_h36m"])


KEYS = ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "parents", "faces", "extra_joints_idxs",
        "J_regressor_extra", "J_regressor_cocoplus", "J_regressor_h36m")


def test_loader_reproduces_the_model_bit_for_bit(synthetic_model, tmp_path, monkeypatch):
    write_official_layout(synthetic_model, str(tmp_path))
    monkeypatch.chdir(tmp_path)
    assert config.SMPL_MODEL_DIR == os.path.join("PlayerReconstruction", "additional", "smpl")
    loaded = load_smpl_model(config.SMPL_MODEL_DIR)
    for k in KEYS:
        a, b = np.asarray(loaded[k]), np.asarray(synthetic_model[k])
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(a, b), k
    assert loaded["posedirs"].flags["C_CONTIGUOUS"] and loaded["parents"][0] == -1
    # the reference's construction call (player_recon.py:147) and the attributes its callers read
    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=1)
    ref = SMPL(synthetic_model, batch_size=1)
    sd, rd = smpl.state_dict(), ref.state_dict()
    assert list(sd) == list(rd)
    for k in sd:
        assert torch.equal(sd[k], rd[k]), k
    assert smpl.faces.shape == (13776, 3) and np.array_equal(smpl.faces, synthetic_model["faces"])
    assert smpl.shapedirs.shape == (6890, 3, 10) and smpl.posedirs.shape == (207, 20670)
    assert smpl.betas.shape == (1, 10) and smpl.body_pose.shape == (1, 69)
    with pytest.raises(FileNotFoundError):
        load_smpl_model(config.SMPL_MODEL_DIR, gender="male")


def test_npz_conversion_is_accepted(synthetic_model, tmp_path, monkeypatch):
    write_official_layout(synthetic_model, str(tmp_path))
    monkeypatch.chdir(tmp_path)
    d = os.path.join("PlayerReconstruction", "additional", "smpl")
    raw = pickle.load(open(os.path.join(d, "SMPL_NEUTRAL.pkl"), "rb"), encoding="latin1")
    os.remove(os.path.join(d, "SMPL_NEUTRAL.pkl"))
    np.savez(os.path.join(d, "SMPL_NEUTRAL.npz"), v_template=raw["v_template"], shapedirs=raw["shapedirs"].r,
             posedirs=raw["posedirs"], J_regressor=raw["J_regressor"].toarray(), weights=raw["weights"],
             kintree_table=raw["kintree_table"], f=raw["f"])
    loaded = load_smpl_model(d)
    for k in KEYS:
        assert np.array_equal(np.asarray(loaded[k]), np.asarray(synthetic_model[k])), k


@pytest.mark.gpu
def test_file_constructed_module_computes_the_same(synthetic_model, tmp_path, monkeypatch):
    write_official_layout(synthetic_model, str(tmp_path))
    monkeypatch.chdir(tmp_path)
    dev = torch.device("cuda", 0)
    a = SMPL(config.SMPL_MODEL_DIR, batch_size=1).to(dev)       # player_recon.py:147
    b = SMPL(synthetic_model, batch_size=1).to(dev)
    g = torch.Generator().manual_seed(5)
    betas = torch.randn(3, 10, generator=g).to(dev)
    pose = (torch.randn(3, 72, generator=g) * 0.3).to(dev)
    oa = a(betas=betas, body_pose=pose[:, 3:], global_orient=pose[:, :3])
    ob = b(betas=betas, body_pose=pose[:, 3:], global_orient=pose[:, :3])
    assert torch.equal(oa.vertices, ob.vertices) and torch.equal(oa.joints, ob.joints)
    assert oa.joints.shape == (3, 90, 3)


def test_official_pickle_loads_without_chumpy(synthetic_model, tmp_path, monkeypatch):
    """The official file pickles its arrays as `chumpy.ch.Ch` objects.  With chumpy absent (as in this image) the
    loader substitutes a stub that recovers the stored ndarray instead of failing in `pickle.load`."""
    import sys
    import types
    assert "chumpy" not in sys.modules
    mod, sub = types.ModuleType("chumpy"), types.ModuleType("chumpy.ch")

    class Ch:                                   # what the writer's chumpy would have pickled: state dict with `x`
        def __init__(self, x):
            self.x = np.asarray(x)
            self.dterms = ("x",)

    Ch.__module__, Ch.__qualname__ = "chumpy.ch", "Ch"
    sub.Ch, mod.ch = Ch, sub
    monkeypatch.setitem(sys.modules, "chumpy", mod)
    monkeypatch.setitem(sys.modules, "chumpy.ch", sub)
    write_official_layout(synthetic_model, str(tmp_path))
    path = os.path.join(str(tmp_path), "PlayerReconstruction", "additional", "smpl", "SMPL_NEUTRAL.pkl")
    raw = pickle.load(open(path, "rb"), encoding="latin1")
    raw["shapedirs"] = Ch(raw["shapedirs"].r)
    raw["v_template"] = Ch(raw["v_template"])
    raw["weights"] = Ch(raw["weights"])
    with open(path, "wb") as f:
        pickle.dump(raw, f, protocol=2)
    monkeypatch.delitem(sys.modules, "chumpy")          # ... and the reader has no chumpy
    monkeypatch.delitem(sys.modules, "chumpy.ch")
    with pytest.raises(ModuleNotFoundError):
        pickle.load(open(path, "rb"), encoding="latin1")
    monkeypatch.chdir(tmp_path)
    loaded = load_smpl_model(config.SMPL_MODEL_DIR)
    for k in KEYS:
        assert np.array_equal(np.asarray(loaded[k]), np.asarray(synthetic_model[k])), k
