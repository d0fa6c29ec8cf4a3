"""The oracle's featuriser against the golden vectors produced by the reference's own NumPy
code (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np

from oracle import featurize as fz


def test_senders_bit_exact_all_casp14(casp14):
    total = 0
    for name, e in casp14.items():
        g = fz.featurize(e["pos"], e["gt"], e["exists"], 50)
        assert g["n_node"] == e["n_valid"], name
        assert np.array_equal(g["senders"], e["senders"]), name
        assert np.array_equal(g["receivers"], np.repeat(np.arange(e["n_valid"]), 50))
        total += g["n_node"]
    assert total == 5616 and len(casp14) == 31


def test_edge_features_match_reference_after_fp32_cast(casp14):
    full = 0
    for name, e in casp14.items():
        g = fz.featurize(e["pos"], e["gt"], e["exists"], 50)
        f32 = np.ascontiguousarray(g["edge_features"].astype(np.float32))
        assert hashlib.sha256(f32.tobytes()).hexdigest() == e["feat_sha256"], name
        if e["edge_features"] is not None:
            assert np.array_equal(f32, e["edge_features"]), name
            full += 1
    assert full >= 3


def test_incomplete_residues_are_dropped(casp14):
    # T1029 and T1041 each have one residue without a complete backbone
    for name in ("T1029", "T1041"):
        e = casp14[name]
        assert e["pos"].shape[0] == e["n_valid"] + 1
        assert int(fz.valid_residue_mask(e["gt"]).sum()) == e["n_valid"]


def test_no_rank_ties_on_casp14(casp14):
    e = casp14["T1073"]
    g = fz.featurize(e["pos"], e["gt"], e["exists"], 50)
    assert not fz.has_rank_ties(fz.pairwise_distance(g["centroid"]), 50)


def test_n_equals_k_includes_self():
    rng = np.random.default_rng(0)
    pos = np.zeros((50, 37, 3), np.float32)
    pos[:, :5] = rng.normal(size=(50, 5, 3)).astype(np.float32) + np.arange(50)[:, None, None] * 1.5
    gt = np.zeros((50, 37), bool)
    gt[:, [0, 1, 2, 4]] = True
    g = fz.featurize(pos, gt, gt, 50)
    assert g["k"] == 50 and (g["senders"].reshape(50, 50)[:, 0] == np.arange(50)).all()


def test_frames_are_orthonormal(casp14):
    e = casp14["T1082"]
    g = fz.featurize(e["pos"], e["gt"], e["exists"], 50)
    for a, b in (("u", "u"), ("v", "v"), ("n", "n")):
        assert np.allclose((g[a] * g[b]).sum(-1), 1.0, atol=1e-12)
    assert np.allclose((g["u"] * g["v"]).sum(-1), 0.0, atol=1e-12)
    assert np.allclose(np.cross(g["u"], g["v"]), g["n"], atol=1e-12)
