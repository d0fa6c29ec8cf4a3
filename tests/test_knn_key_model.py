"""The packed-key selection rule of `knn_warp_kernel` (csrc/featurize.cu) as a NumPy model, CPU only.

The kernel sorts 64-bit keys = fp64 bit pattern of the (squared) centroid distance with the 11 lowest mantissa bits
replaced by the candidate index, and hands a row to the exact (distance, index) kernel when two adjacent entries
of the sorted head (ranks 0 .. K+2) have truncated values that are equal (distance keys) or equal / adjacent
(squared-distance keys, the default).  The claim checked here: every row the rule does NOT flag comes out in the
stable ascending (distance, index) order of the reference (utils/protein_utils.py:380-399), including the
constructed case the rule exists for: two different squared distances that round to the same distance.
The kernel itself is compared with the oracle on the GPU (tests/test_gpu_featurize.py, tests/test_gpu_fullsize.py).
"""
import numpy as np
import pytest

from oracle import featurize as fz

K = 50


def packed_select(cen: np.ndarray, k: int, mode: str, gap: int):
    """Returns (senders [L, k] by packed key order, flagged [L]).  mode: 'dist' or 'd2'."""
    L = cen.shape[0]
    d = cen[:, None, :] - cen[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    val = np.sqrt(d2) if mode == "dist" else d2
    keys = (np.ascontiguousarray(val).view(np.uint64) & ~np.uint64(0x7FF)) | np.arange(L, dtype=np.uint64)[None, :]
    order = np.argsort(keys, axis=1, kind="stable")  # keys are unique within a row
    head = np.take_along_axis(keys, order[:, : min(L, k + 3)], axis=1) >> np.uint64(11)  # ranks 0 .. K+2
    flagged = ((head[:, 1:] - head[:, :-1]) <= np.uint64(gap)).any(axis=1)
    first = 0 if L == k else 1
    return order[:, first : first + k], flagged


def reference_senders(cen: np.ndarray, k: int) -> np.ndarray:
    return fz.knn_senders(fz.pairwise_distance(cen), k)


def _casp14_centroids(entry):
    keep = fz.valid_residue_mask(entry["gt"])
    mask = (entry["gt"] & entry["exists"])[keep]
    return fz.centroids(entry["pos"][keep].astype(np.float64), mask)


@pytest.mark.parametrize("mode,gap", [("dist", 0), ("d2", 1)])
def test_unflagged_rows_equal_the_reference_on_casp14(casp14, mode, gap):
    rows = flagged_rows = 0
    for name, e in casp14.items():
        cen = _casp14_centroids(e)
        s, flagged = packed_select(cen, K, mode, gap)
        ref = e["senders"].reshape(-1, K)  # the reference's own output (tests/golden/make_golden.py)
        assert np.array_equal(s[~flagged], ref[~flagged]), name
        rows += len(cen)
        flagged_rows += int(flagged.sum())
    assert rows == 5616
    assert flagged_rows <= 2  # the exact kernel is a rare path on real structures


@pytest.mark.parametrize("mode,gap", [("dist", 0), ("d2", 1)])
def test_lattice_ties_are_flagged(mode, gap):
    g = np.arange(4, dtype=np.float64) * 3.8
    cen = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)  # 64 points, many equal distances
    s, flagged = packed_select(cen, K, mode, gap)
    assert flagged.all()
    rng = np.random.default_rng(3)
    cen = cen + rng.normal(0, 0.3, cen.shape)  # generic positions: nothing to flag, order = reference
    s, flagged = packed_select(cen, K, mode, gap)
    assert not flagged.any() and np.array_equal(s, reference_senders(cen, K))


def _same_sqrt_pair(straddle: bool):
    """Two points p_lo, p_hi (z differs by a few ulps) whose squared distances from the origin are different doubles
    that round to the SAME distance; with `straddle` the smaller square ends in eleven 1-bits, so the truncated keys
    are adjacent instead of equal."""
    rng = np.random.default_rng(11 if straddle else 5)
    for _ in range(4000):
        x, y = rng.uniform(3.0, 6.0, 2)
        z0 = rng.uniform(1e-3, 4e-3)  # small z: one ulp of z moves the square by less than one ulp of the sum
        z = z0 * (1.0 + np.arange(1 << 15) * 2.0**-52 * 64)
        d2 = (x * x + y * y) + z * z
        bits = d2.view(np.uint64)
        nz = np.nonzero(np.diff(bits) > 0)[0]  # places where the square steps to the next double(s)
        lo, hi = d2[nz], d2[nz + 1]
        ok = np.sqrt(lo) == np.sqrt(hi)
        low11 = lo.view(np.uint64) & np.uint64(0x7FF)
        ok &= (low11 == np.uint64(0x7FF)) if straddle else (low11 < np.uint64(0x700))
        if ok.any():
            i = nz[np.nonzero(ok)[0][0]]
            return np.array([x, y, z[i]]), np.array([x, y, z[i + 1]])
    raise AssertionError("no same-sqrt pair found")


@pytest.mark.parametrize("straddle", [False, True])
def test_constructed_same_sqrt_pair_needs_the_adjacent_key_rule(straddle):
    p_lo, p_hi = _same_sqrt_pair(straddle)
    rng = np.random.default_rng(1)
    far = rng.uniform(-1.0, 1.0, (K + 7, 3))
    far = far / np.linalg.norm(far, axis=1, keepdims=True) * rng.uniform(9.0, 30.0, (K + 7, 1))
    cen = np.concatenate([np.zeros((1, 3)), far])
    cen[5], cen[9] = p_hi, p_lo  # the larger square has the SMALLER index: the reference puts it first
    ref = reference_senders(cen, K)
    assert list(ref[0, :2]) == [5, 9]
    d = fz.pairwise_distance(cen)
    assert d[0, 5] == d[0, 9]
    # distance keys: the two truncated values are equal -> flagged
    _, flagged = packed_select(cen, K, "dist", 0)
    assert flagged[0]
    # squared keys: equal truncated values fall back to index order (right here by luck), adjacent ones order the pair
    # by its squares, i.e. the other way round; either way the row MUST be flagged, and gap <= 1 does it ...
    s, flagged = packed_select(cen, K, "d2", 1)
    assert list(s[0, :2]) == ([9, 5] if straddle else [5, 9]) and flagged[0]
    # ... while 'equal truncated values only' would let the straddling pair through with the wrong order
    _, flagged0 = packed_select(cen, K, "d2", 0)
    assert bool(flagged0[0]) == (not straddle)


def coarse_preselect(cen: np.ndarray, k: int):
    """Round-2 phase 1 of `knn_warp_kernel` as a NumPy model: 32-bit keys = fp32(d2) with the 11 lowest mantissa bits
    replaced by the index; survivors = the 64 smallest; a row is handed to the exact kernel unless the truncated key of
    coarse rank min(k + 3, 62) is strictly below that of rank 63.  Returns (survivor index sets, flagged)."""
    L = cen.shape[0]
    d = cen[:, None, :] - cen[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    keys = (d2.astype(np.float32).view(np.uint32) & ~np.uint32(0x7FF)) | np.arange(L, dtype=np.uint32)[None, :]
    order = np.argsort(keys, axis=1, kind="stable")
    if L <= 64:
        return order, np.zeros(L, bool)
    srt = np.take_along_axis(keys, order[:, :64], axis=1) >> np.uint32(11)
    rk = min(k + 3, 62)
    return order[:, :64], srt[:, rk] >= srt[:, 63]


def _exact_head(cen: np.ndarray, n: int) -> np.ndarray:
    d = cen[:, None, :] - cen[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    return np.argsort(d2, axis=1, kind="stable")[:, :n]


def test_coarse_survivors_hold_the_exact_head_on_casp14_and_synthetic(casp14):
    """the claim phase 2 relies on: an unflagged row's 64 survivors contain every candidate of exact rank 0 .. K+3
    (what the exact sort and the clash check of the head look at)"""
    from pst import synthetic as syn

    sets = [_casp14_centroids(e) for e in casp14.values()]
    for bb in syn.make_backbones(77, [512, 1024, 2048, 130]):
        c = bb.astype(np.float64)
        sets.append((((c[:, 0] + c[:, 1]) + c[:, 2]) + c[:, 3]) / 4.0)
    rows = flagged_rows = 0
    for cen in sets:
        surv, flagged = coarse_preselect(cen, K)
        head = _exact_head(cen, min(K + 4, cen.shape[0]))
        ok = np.array([set(h.tolist()) <= set(s.tolist()) for h, s in zip(head, surv)])
        assert ok[~flagged].all()
        rows += len(cen)
        flagged_rows += int(flagged.sum())
    assert flagged_rows <= rows // 1000  # the exact kernel stays a rare path


def test_coarse_preselection_flags_what_it_cannot_separate():
    """more than 64 candidates whose squared distances agree to fp32's 13 kept mantissa bits: the coarse keys cannot tell
    which of them belong to the head, the row must be flagged (points on a sphere around the first residue)"""
    rng = np.random.default_rng(5)
    v = rng.standard_normal((200, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    cen = np.concatenate([np.zeros((1, 3)), 10.0 * v * (1.0 + 1e-7 * rng.random((200, 1)))])
    surv, flagged = coarse_preselect(cen, K)
    assert flagged[0]
    head = _exact_head(cen, K + 4)
    ok = np.array([set(h.tolist()) <= set(s.tolist()) for h, s in zip(head, surv)])
    assert ok[~flagged].all()
