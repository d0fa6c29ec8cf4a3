import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "protein-structure-tokenizer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI shared library, built in-tree if it is not there yet (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, PKG)
    import build as pst_build

    return pst_build.build()


@pytest.fixture(scope="session")
def casp14():
    """The 31 bundled CASP14 structures as atom37 arrays + the reference's own graph outputs."""
    a = np.load(os.path.join(GOLDEN, "casp14_atom37.npz"))
    g = np.load(os.path.join(GOLDEN, "casp14_graph_ref.npz"))
    names = [str(n) for n in a["names"]]
    lens = a["lengths"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    gt = np.unpackbits(a["atom37_gt_exists"], axis=1)[:, :37].astype(bool)
    ex = np.unpackbits(a["atom37_atom_exists"], axis=1)[:, :37].astype(bool)
    nv = g["n_valid"]
    soffs = np.concatenate([[0], np.cumsum(nv.astype(np.int64) * 50)])
    out = {}
    for i, n in enumerate(names):
        sl = slice(offs[i], offs[i + 1])
        out[n] = {
            "pos": a["atom37_positions"][sl],
            "gt": gt[sl],
            "exists": ex[sl],
            "n_valid": int(nv[i]),
            "senders": g["senders"][soffs[i] : soffs[i + 1]].astype(np.int64),
            "feat_sha256": str(g["edge_features_sha256"][i]),
            "feat_sum": float(g["edge_features_sum"][i]),
            "edge_features": g[f"edge_features_{n}"] if f"edge_features_{n}" in g.files else None,
        }
    return out


def valid_atoms(entry):
    """Host-side residue filtering (data/preprocessing.py:99-117): keep residues with N, CA, C, O."""
    gt = entry["gt"]
    keep = gt[:, 0] & gt[:, 1] & gt[:, 2] & gt[:, 4]
    mask = (entry["gt"] & entry["exists"])[keep]
    return np.ascontiguousarray(entry["pos"][keep], np.float32), np.ascontiguousarray(mask, np.uint8)
