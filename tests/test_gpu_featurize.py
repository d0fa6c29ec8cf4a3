"""CUDA featuriser / k-NN (pst_featurize_knn) against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import valid_atoms

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tok(built_lib):
    import torch
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp32")
    return StructureTokenizer(cfg, init_params(cfg, 0, "rich"))


def _run(tok, structs, masks=None):
    import torch

    lens = [s.shape[0] for s in structs]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    atoms = torch.from_numpy(np.concatenate(structs)).cuda()
    mask = None if masks is None else torch.from_numpy(np.concatenate(masks)).cuda()
    s, f = tok.featurize_device(atoms, mask, torch.from_numpy(offs).cuda(), len(structs), int(offs[-1]))
    torch.cuda.synchronize()
    return s.cpu().numpy(), f.cpu().numpy(), offs


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def test_casp14_senders_bit_exact_and_features(tok, casp14):
    names = list(casp14)
    structs, masks = zip(*[valid_atoms(casp14[n]) for n in names])
    s, f, offs = _run(tok, structs, masks)
    assert tok.read_status() == 0
    n_feat = 0
    for i, n in enumerate(names):
        e = casp14[n]
        sl = slice(offs[i] * 50, offs[i + 1] * 50)
        assert np.array_equal(s[sl].astype(np.int64), e["senders"]), n
        assert abs(float(f[sl].astype(np.float64).sum()) - e["feat_sum"]) < 1e-2, n
        if e["edge_features"] is not None:
            u = ulp_diff(f[sl], e["edge_features"])
            # fp64 arithmetic rounded once: <= 1 fp32 ulp from the reference (CUDA vs NumPy exp / sum order)
            tiny = np.abs(e["edge_features"]) < 1e-30
            assert (u[~tiny] <= 1).all(), (n, int(u[~tiny].max()))
            assert (u == 0).mean() > 0.999
            n_feat += 1
    assert n_feat >= 3


def test_synthetic_backbone4_matches_oracle(tok):
    from oracle import featurize as fz
    from pst import synthetic as syn

    bbs = syn.make_backbones(20240517, [64, 127, 200, 512, 50, 333])
    s, f, offs = _run(tok, bbs)
    assert tok.read_status() == 0
    for i, bb in enumerate(bbs):
        pos, gt, ex = syn.backbone_to_atom37(bb)
        g = fz.featurize(pos, gt, ex, 50)
        sl = slice(offs[i] * 50, offs[i + 1] * 50)
        assert np.array_equal(s[sl].astype(np.int64), g["senders"]), i
        ref = g["edge_features"].astype(np.float32)
        u = ulp_diff(f[sl], ref)
        tiny = np.abs(ref) < 1e-30
        assert (u[~tiny] <= 1).all(), (i, int(u[~tiny].max()))


def test_length_equal_k_includes_self(tok):
    from pst import synthetic as syn

    bb = syn.make_backbones(3, [50])[0]
    s, _, _ = _run(tok, [bb])
    assert (s.reshape(50, 50)[:, 0] == np.arange(50)).all()


def test_too_short_structure_raises_status(tok):
    from pst import synthetic as syn

    bb = syn.make_backbones(4, [60])[0][:40]
    _run(tok, [bb])
    assert tok.read_status() == -3
    with pytest.raises(NotImplementedError):
        tok.tokenize([bb])


def test_oversized_batch_is_refused_by_the_abi(tok):
    # include/pst_abi.h PST_MAX_EDGES_PER_CALL: kernels index edge rows with 32-bit integers
    too_many = (1 << 26) // tok.cfg.num_neighbor + 1
    assert tok.lib.pst_workspace_bytes(tok._h, too_many, 1) == 0
    assert tok.lib.pst_workspace_bytes(tok._h, too_many - 1, 1) > 0
    with pytest.raises(ValueError):
        tok._workspace(too_many, 1)


def test_misaligned_workspace_is_refused(tok):
    import torch

    R, B, K = 64, 1, tok.cfg.num_neighbor
    atoms = torch.zeros((R, 4, 3), dtype=torch.float32, device=tok.device)
    offs = torch.tensor([0, R], dtype=torch.int32, device=tok.device)
    senders = torch.empty((R * K,), dtype=torch.int32, device=tok.device)
    ws = tok._workspace(R, B)
    extra = torch.empty(ws.numel() + 256, dtype=torch.uint8, device=tok.device)
    rc = tok.lib.pst_featurize_knn(tok._h, tok._stream(), atoms.data_ptr(), None, 4, offs.data_ptr(), B, R,
                                   senders.data_ptr(), None, extra.data_ptr() + 8, ws.numel())
    assert rc == -1  # PST_ERR_BAD_ARGUMENT: include/pst_abi.h asks for a 256-byte aligned workspace


def test_atom37_and_backbone4_layouts_agree(tok):
    from pst import synthetic as syn

    bb = syn.make_backbones(5, [128])[0]
    pos, gt, _ = syn.backbone_to_atom37(bb)
    s4, f4, _ = _run(tok, [bb])
    s37, f37, _ = _run(tok, [pos], [gt.astype(np.uint8)])
    assert np.array_equal(s4, s37) and np.array_equal(f4, f37)


def test_exact_distance_ties_follow_stable_argsort(tok):
    """Residues on a regular lattice: many exactly equal distances.  The packed-key k-NN kernel must detect them
    and hand those rows to the exact (distance, index) kernel; the result is the stable argsort of the reference
    contract."""
    from oracle import featurize as fz

    g = np.stack(np.meshgrid(np.arange(5), np.arange(5), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    ca = g * 4.0  # 100 residues on a 4 A lattice
    bb = np.stack([ca + np.float32([-1.2, 0.6, 0.1]), ca, ca + np.float32([1.1, 0.9, -0.2]), ca + np.float32([1.6, 2.0, 0.3])], axis=1)
    bb = np.round(bb, 3).astype(np.float32)
    s, f, offs = _run(tok, [bb])
    assert tok.read_status() == 0
    pos = np.zeros((100, 37, 3), np.float32)
    gt = np.zeros((100, 37), bool)
    for src, dst in enumerate((0, 1, 2, 4)):
        pos[:, dst] = bb[:, src]
        gt[:, dst] = True
    o = fz.featurize(pos, gt, gt, 50)
    assert fz.has_rank_ties(fz.pairwise_distance(o["centroid"]), 50)  # the input really has ties
    assert np.array_equal(s.astype(np.int64), o["senders"])
    u = ulp_diff(f, o["edge_features"].astype(np.float32))
    tiny = np.abs(o["edge_features"].astype(np.float32)) < 1e-30
    assert (u[~tiny] <= 1).all()


@pytest.mark.parametrize("straddle", [False, True])
def test_same_sqrt_pair_is_decided_by_the_exact_kernel(tok, straddle):
    """Two candidates whose SQUARED distances to row 0 are different doubles that round to the same distance: the
    reference orders them by index, the scan's squared keys by their squares.  With `straddle` the truncated keys
    are adjacent rather than equal, the case the head check's "differ by at most one" rule exists for
    (tests/test_knn_key_model.py is the CPU statement of that rule).  Centroid = CA here (only CA is marked present)."""
    from oracle import featurize as fz

    rng = np.random.default_rng(11 if straddle else 5)
    pair = None
    for _ in range(20000):
        x, y = rng.uniform(3.0, 6.0, 2).astype(np.float32).astype(np.float64)
        z0 = np.float32(rng.uniform(5e-5, 4e-4))
        z = (z0.view(np.int32) + np.arange(1 << 14, dtype=np.int32)).view(np.float32).astype(np.float64)  # successive fp32 values
        d2 = (x * x + y * y) + z * z
        nz = np.nonzero(np.diff(d2.view(np.uint64)) > 0)[0]
        lo, hi = d2[nz], d2[nz + 1]
        low11 = lo.view(np.uint64) & np.uint64(0x7FF)
        ok = (np.sqrt(lo) == np.sqrt(hi)) & ((low11 == np.uint64(0x7FF)) if straddle else (low11 < np.uint64(0x700)))
        if ok.any():
            i = nz[np.nonzero(ok)[0][0]]
            pair = (np.array([x, y, z[i]]), np.array([x, y, z[i + 1]]))
            break
    assert pair is not None
    L, K = 58, 50
    far = rng.uniform(-1.0, 1.0, (L - 1, 3))
    far = far / np.linalg.norm(far, axis=1, keepdims=True) * rng.uniform(9.0, 30.0, (L - 1, 1))
    ca = np.concatenate([np.zeros((1, 3)), far]).astype(np.float32)
    ca[5], ca[9] = pair[1].astype(np.float32), pair[0].astype(np.float32)  # larger square at the smaller index
    assert np.array_equal(ca[5].astype(np.float64), pair[1]) and np.array_equal(ca[9].astype(np.float64), pair[0])
    atoms = np.zeros((L, 4, 3), np.float32)
    atoms[:, 1] = ca
    atoms[:, 0] = ca + np.float32([-0.5, 1.4, 0.0])
    atoms[:, 2] = ca + np.float32([1.5, 0.0, 0.0])
    atoms[:, 3] = ca + np.float32([2.0, 1.0, 0.5])
    mask = np.zeros((L, 4), np.uint8)
    mask[:, 1] = 1
    s, _, _ = _run(tok, [atoms], [mask])
    assert tok.read_status() == 0
    cen = ca.astype(np.float64)
    d = fz.pairwise_distance(cen)
    assert d[0, 5] == d[0, 9]
    ref = fz.knn_senders(d, K)
    assert list(ref[0, :2]) == [5, 9]
    assert np.array_equal(s.reshape(L, K).astype(np.int64), ref)
