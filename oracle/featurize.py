"""Oracle (TEST INFRASTRUCTURE): NumPy fp64 restatement of the reference's host
featuriser -- atom37 -> backbone frames -> centroid k-NN graph -> 27-d edge
features.  See ``oracle/__init__.py`` for who may import this.

Reference files followed (paths relative to the reference repo root):
  * structure_tokenizer/data/preprocessing.py:69-189      (masking, frames, filtering)
  * structure_tokenizer/model/quat_affine.py:378-522      (backbone frames)
  * structure_tokenizer/utils/protein_utils.py:257-281    (RBF distance features)
  * structure_tokenizer/utils/protein_utils.py:325-438    (k-NN graph, orientation features)
  * structure_tokenizer/data/protein_structure_sample.py:64-70 (missing-backbone mask)
Third-party arithmetic restated here (not vendored in the reference):
  * scipy.spatial.distance.cdist (scipy==1.9.3, euclidean): sqrt((dx*dx+dy*dy)+dz*dz)
  * numpy.argsort (numpy 1.24 default kind): contract = stable order by (distance, index)

Pinned: ``tests/golden/make_golden.py`` runs the reference's own NumPy
functions through import stubs and checks this module bit-for-bit.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

# atom37 slots used by the path (reference: data/residue_constants.py:539-593)
N_INDEX, CA_INDEX, C_INDEX, CB_INDEX, O_INDEX = 0, 1, 2, 3, 4
NUM_RBF = 15
NUM_EDGE_FEATURES = 27


def backbone_frames(n_xyz: np.ndarray, ca_xyz: np.ndarray, c_xyz: np.ndarray):
    """Per-residue frame axes (u, v, n), fp64, same operation order as
    quat_affine.py:406-522 (`make_canonical_transform` followed by the transpose
    in `make_transform_from_reference`) and preprocessing.py:94-97 (u, v, n are
    columns 0, 1, 2 of the returned rotation = rows of Rn.Rc).

    Every 3-term product-sum is evaluated left to right, as the reference's
    `_multiply` / `apply_rot_to_vec` do (quat_affine.py:378-403, 175-185).
    """
    n_xyz = np.asarray(n_xyz, np.float64)
    ca_xyz = np.asarray(ca_xyz, np.float64)
    c_xyz = np.asarray(c_xyz, np.float64)
    t = -ca_xyz
    nv = n_xyz + t
    cv = c_xyz + t
    cx, cy, cz = cv[:, 0], cv[:, 1], cv[:, 2]
    zero = np.zeros_like(cx)
    one = np.ones_like(cx)

    den_xy = np.sqrt(1e-20 + cx * cx + cy * cy)
    s1 = -cy / den_xy
    c1 = cx / den_xy
    r1 = [[c1, -s1, zero], [s1, c1, zero], [zero, zero, one]]

    den_xyz = np.sqrt(1e-20 + cx * cx + cy * cy + cz * cz)
    s2 = cz / den_xyz
    c2 = np.sqrt(cx * cx + cy * cy) / den_xyz
    r2 = [[c2, zero, s2], [zero, one, zero], [-s2, zero, c2]]

    def mat3(a, b):
        return [
            [a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] for j in range(3)]
            for i in range(3)
        ]

    rc = mat3(r2, r1)
    x, y, z = nv[:, 0], nv[:, 1], nv[:, 2]
    ny = rc[1][0] * x + rc[1][1] * y + rc[1][2] * z
    nz = rc[2][0] * x + rc[2][1] * y + rc[2][2] * z
    den_n = np.sqrt(1e-20 + ny * ny + nz * nz)
    sn = -nz / den_n
    cn = ny / den_n
    rn = [[one, zero, zero], [zero, cn, -sn], [zero, sn, cn]]
    m = mat3(rn, rc)  # rows of m are the frame axes
    u = np.stack(m[0], axis=-1)
    v = np.stack(m[1], axis=-1)
    n = np.stack(m[2], axis=-1)
    return u, v, n


def valid_residue_mask(gt_exists: np.ndarray) -> np.ndarray:
    """protein_structure_sample.py:64-70: a residue is kept iff N, CA, C and O
    coordinates were given."""
    g = np.asarray(gt_exists, bool)
    return g[:, CA_INDEX] & g[:, N_INDEX] & g[:, C_INDEX] & g[:, O_INDEX]


def centroids(atom_pos: np.ndarray, atom_mask: np.ndarray) -> np.ndarray:
    """protein_utils.py:373-378: mean over the present atoms of each residue.
    np.mean(axis=0) of an [n_atoms, 3] array is a sequential fp64 sum in atom
    order divided by the count; adding +0.0 for absent slots is exact, so the
    sum can run over all slots (this is also the form the CUDA kernel uses)."""
    pos = np.asarray(atom_pos, np.float64)
    m = np.asarray(atom_mask, bool)
    acc = np.zeros((pos.shape[0], 3), np.float64)
    for a in range(pos.shape[1]):
        acc = acc + np.where(m[:, a, None], pos[:, a, :], 0.0)
    cnt = m.sum(axis=1).astype(np.float64)
    return acc / cnt[:, None]


def pairwise_distance(x: np.ndarray) -> np.ndarray:
    """scipy cdist (euclidean, fp64): sqrt((dx*dx + dy*dy) + dz*dz), no FMA."""
    d = x[:, None, :] - x[None, :, :]
    return np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])


def knn_senders(dist: np.ndarray, k: int) -> np.ndarray:
    """protein_utils.py:385-389.  Stable ascending order by (distance, index);
    ranks 1..k when n > k (rank 0 is the residue itself), all n ranks when
    n == k (self included, :365-367,385-387)."""
    n = dist.shape[0]
    order = np.argsort(dist, axis=-1, kind="stable")
    if k >= n:
        return order[:, :n]
    return order[:, 1 : k + 1]


def rbf_features(d: np.ndarray) -> np.ndarray:
    """protein_utils.py:257-281: exp(-d^2 / 1.5^x), x = 0..14, cast to fp32."""
    scales = [float(1.5**x) for x in range(NUM_RBF)]
    return np.stack([np.exp(-((d - 0.0) ** 2) / s) for s in scales], axis=-1).astype(np.float32)


def edge_features(dist_e, senders, receivers, ca, u, v, n) -> np.ndarray:
    """protein_utils.py:403-434: [15 RBF | p | q | k | t] per edge, fp64
    (the RBF block has already been rounded to fp32 by the reference)."""
    rbf = rbf_features(dist_e)
    basis = np.stack([n, u, v], axis=1)  # [n_res, 3(rows n,u,v), 3]
    b = basis[receivers]  # [E,3,3]
    diff = ca[senders] - ca[receivers]

    def rot(vec):
        return np.stack(
            [b[:, j, 0] * vec[:, 0] + b[:, j, 1] * vec[:, 1] + b[:, j, 2] * vec[:, 2] for j in range(3)],
            axis=-1,
        )

    p = rot(diff)
    q = rot(n[senders])
    kk = rot(u[senders])
    t = rot(v[senders])
    return np.concatenate([rbf.astype(np.float64), p, q, kk, t], axis=1)


def featurize(atom_pos, gt_exists, atom_exists, num_neighbor: int = 50) -> Dict[str, np.ndarray]:
    """preprocessing.py:69-189 for noise_level=0, residue_loc_is_alphac=True and
    no crop.  `atom_pos` is [n, A, 3] with slots 0,1,2 = N, CA, C (A = 37 for the
    atom37 layout; A = 5 works for backbone+CB+O inputs).

    Returns the *unpadded* graph: senders/receivers int64 [n_valid*K],
    edge_features fp64 [n_valid*K, 27], plus the intermediates used by tests.
    """
    pos = np.asarray(atom_pos, np.float64)
    gt = np.asarray(gt_exists, bool)
    mask = gt & np.asarray(atom_exists, bool)  # preprocessing.py:72
    keep = valid_residue_mask(gt)
    u, v, n = backbone_frames(pos[:, N_INDEX], pos[:, CA_INDEX], pos[:, C_INDEX])
    pos, mask, u, v, n = pos[keep], mask[keep], u[keep], v[keep], n[keep]
    n_res = pos.shape[0]
    k = min(num_neighbor, n_res)
    ca = pos[:, CA_INDEX]
    cen = centroids(pos, mask)
    dist = pairwise_distance(cen)
    send = knn_senders(dist, k)
    recv = np.repeat(np.arange(n_res), k)
    dist_e = np.take_along_axis(dist, send, axis=1).reshape(-1)
    send = send.reshape(-1)
    feats = edge_features(dist_e, send, recv, ca, u, v, n)
    return {
        "n_node": n_res,
        "k": k,
        "keep": keep,
        "centroid": cen,
        "u": u,
        "v": v,
        "n": n,
        "dist": dist_e,
        "senders": send.astype(np.int64),
        "receivers": recv.astype(np.int64),
        "edge_features": feats,
    }


def has_rank_ties(dist: np.ndarray, k: int) -> bool:
    """True iff a row has an exact distance tie among ranks 0..k+1, i.e. the
    reference's unstable argsort would be implementation-defined there."""
    srt = np.sort(dist, axis=-1)[:, : k + 2]
    return bool((np.diff(srt, axis=-1) == 0).any())
