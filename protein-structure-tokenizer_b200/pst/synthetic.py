"""Seeded synthetic protein backbones: the inputs for the
BASELINE.json configs that are not the bundled CASP14 files.

Each structure is a compact self-avoiding CA random walk (bond 3.80 A, bond
angle U(85,145) deg, dihedral from a helix-like / strand-like mixture, confined
to a sphere of radius 1.3*2.2*L^0.38 A, >= 3.6 A between non-adjacent CAs) with
N, C, O placed from a fixed local geometry around every CA.  Coordinates are
rounded to 3 decimals and stored as fp32, like PDB files.  Residue type is GLY:
the heavy atoms are exactly N, CA, C, O, so the k-NN anchor (centroid) is the
mean of those four atoms (reference utils/protein_utils.py:373-378).

Layout returned: ``backbone[L, 4, 3]`` fp32 in atom order N, CA, C, O.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

CA_BOND = 3.80
MIN_CA_DIST = 3.6


def _place(p1, p2, p3, bond, angle, dihedral):
    """NeRF: next point from the previous three (batched over leading dim)."""
    bc = p3 - p2
    bc /= np.linalg.norm(bc, axis=-1, keepdims=True)
    nrm = np.cross(p2 - p1, bc)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True) + 1e-12
    m = np.cross(nrm, bc)
    d = np.stack(
        [-bond * np.cos(angle), bond * np.sin(angle) * np.cos(dihedral), bond * np.sin(angle) * np.sin(dihedral)], -1
    )
    return p3 + d[:, 0:1] * bc + d[:, 1:2] * m + d[:, 2:3] * nrm


def _ca_walk(rng: np.random.Generator, batch: int, length: int) -> np.ndarray:
    radius = 1.3 * 2.2 * length**0.38
    ca = np.zeros((batch, length, 3))
    ca[:, 1] = [CA_BOND, 0, 0]
    ca[:, 2] = ca[:, 1] + CA_BOND * np.array([math.cos(math.radians(70)), math.sin(math.radians(70)), 0])
    for i in range(3, length):
        todo = np.ones(batch, bool)
        best = np.zeros((batch, 3))
        best_score = np.full(batch, -np.inf)
        for attempt in range(64):
            idx = np.nonzero(todo)[0]
            if idx.size == 0:
                break
            nb = idx.size
            angle = np.radians(rng.uniform(85.0, 145.0, nb))
            helix = rng.random(nb) < 0.5
            dih = np.where(helix, rng.normal(50.0, 15.0, nb), rng.normal(-170.0, 25.0, nb))
            if attempt >= 8:  # stuck: open the dihedral up
                dih = rng.uniform(-180.0, 180.0, nb)
            cand = _place(ca[idx, i - 3], ca[idx, i - 2], ca[idx, i - 1], CA_BOND, np.pi - angle, np.radians(dih))
            dmin = np.linalg.norm(ca[idx, : i - 1] - cand[:, None, :], axis=-1).min(axis=1)
            rr = np.linalg.norm(cand - ca[idx, :i].mean(axis=1), axis=-1)
            ok = (dmin >= MIN_CA_DIST) & (rr <= radius)
            score = np.minimum(dmin - MIN_CA_DIST, 0) * 10 + np.minimum(radius - rr, 0)
            better = score > best_score[idx]
            best[idx[better]] = cand[better]
            best_score[idx[better]] = score[better]
            best[idx[ok]] = cand[ok]
            todo[idx[ok]] = False
        ca[:, i] = best
    return ca


def _backbone_from_ca(ca: np.ndarray) -> np.ndarray:
    """N, C, O around each CA from the directions to the neighbouring CAs."""
    # chain ends: parallelogram extrapolation (never collinear with the real neighbour)
    nxt = np.concatenate([ca[:, 1:], ca[:, -1:] + (ca[:, -2:-1] - ca[:, -3:-2])], axis=1)
    prv = np.concatenate([ca[:, :1] - (ca[:, 2:3] - ca[:, 1:2]), ca[:, :-1]], axis=1)
    a = nxt - ca
    a /= np.linalg.norm(a, axis=-1, keepdims=True)
    b = prv - ca
    b /= np.linalg.norm(b, axis=-1, keepdims=True)
    w = a - b
    w /= np.linalg.norm(w, axis=-1, keepdims=True)
    m = a + b
    mn = np.linalg.norm(m, axis=-1, keepdims=True)
    z = np.cross(a, b)
    z /= np.linalg.norm(z, axis=-1, keepdims=True) + 1e-12
    m = np.where(mn > 1e-3, m / np.maximum(mn, 1e-12), np.cross(z, w))
    g = math.cos(math.radians(30)) * m + math.sin(math.radians(30)) * z
    g -= (g * w).sum(-1, keepdims=True) * w
    g /= np.linalg.norm(g, axis=-1, keepdims=True)
    half = math.radians(111.0 / 2)
    dir_c = math.cos(half) * g + math.sin(half) * w
    dir_n = math.cos(half) * g - math.sin(half) * w
    c = ca + 1.525 * dir_c
    n = ca + 1.459 * dir_n
    dir_o = 0.55 * dir_c - 0.83 * g + 0.10 * z
    dir_o /= np.linalg.norm(dir_o, axis=-1, keepdims=True)
    o = c + 1.231 * dir_o
    return np.stack([n, ca, c, o], axis=2)


def _has_knn_ties(backbone: np.ndarray, k: int) -> bool:
    cen = backbone.astype(np.float64)
    cen = (((cen[:, 0] + cen[:, 1]) + cen[:, 2]) + cen[:, 3]) / 4.0
    d = cen[:, None, :] - cen[None, :, :]
    d2 = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    srt = np.sort(d2, axis=-1)[:, : k + 2]
    return bool((np.diff(srt, axis=-1) == 0).any())


def make_backbones(seed: int, lengths: List[int], k: int = 50, group: int = 64) -> List[np.ndarray]:
    """One fp32 [L,4,3] backbone per entry of `lengths`.  Structures with an exact
    centroid-distance tie among the first k+2 ranks of any row are re-drawn
    (the k-NN order would be implementation-defined in the reference there)."""
    rng = np.random.default_rng(seed)
    out: List[np.ndarray] = [None] * len(lengths)  # type: ignore
    by_len = {}
    for i, L in enumerate(lengths):
        by_len.setdefault(int(L), []).append(i)
    for L, idxs in sorted(by_len.items()):
        pending = list(idxs)
        while pending:
            cur, pending = pending[:group], pending[group:]
            ca = _ca_walk(rng, len(cur), L)
            ca -= ca.mean(axis=1, keepdims=True)
            bb = np.round(_backbone_from_ca(ca), 3).astype(np.float32)
            for j, i in enumerate(cur):
                if _has_knn_ties(bb[j], min(k, L - 2)) or not np.isfinite(bb[j]).all():
                    pending.append(i)
                else:
                    out[i] = bb[j]
    return out


def pack_backbones(backbones: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate into the ragged device layout: atoms fp32 [sum L, 4, 3], offsets int32 [B+1]."""
    offsets = np.zeros(len(backbones) + 1, np.int32)
    offsets[1:] = np.cumsum([b.shape[0] for b in backbones])
    return np.ascontiguousarray(np.concatenate(backbones, axis=0), dtype=np.float32), offsets


def backbone_to_atom37(backbone: np.ndarray):
    """Expand [L,4,3] (N,CA,C,O) to the parser's atom37 layout for a GLY chain."""
    L = backbone.shape[0]
    pos = np.zeros((L, 37, 3), np.float32)
    gt = np.zeros((L, 37), bool)
    for src, dst in enumerate((0, 1, 2, 4)):
        pos[:, dst] = backbone[:, src]
        gt[:, dst] = True
    return pos, gt, gt.copy()


def bucketed_lengths(seed: int, count: int, lo: int = 64, hi: int = 2048, step: int = 64) -> np.ndarray:
    """cfg5: lengths log-uniform on [lo, hi], snapped to multiples of `step`."""
    rng = np.random.default_rng(seed)
    L = np.exp(rng.uniform(math.log(lo), math.log(hi), count))
    return np.clip((np.round(L / step) * step).astype(np.int64), lo, hi)
