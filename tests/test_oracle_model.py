"""Properties of the oracle's model forward and quantiser (CPU): constants of SURVEY appendix A.9,
padding independence (the reason the CUDA path may run ragged), round trips."""
import numpy as np
import pytest

from oracle import featurize as fz
from oracle import model as om


def _graph(seed, L):
    from pst import synthetic as syn

    bb = syn.make_backbones(seed, [L])[0]
    pos, gt, ex = syn.backbone_to_atom37(bb)
    return fz.featurize(pos, gt, ex, 50)


def _params(cfg, flavour="rich"):
    from pst.config import TokenizerConfig
    from pst.weights import init_params

    tc = TokenizerConfig(seq_max_size=cfg.seq_max_size, max_out_len=cfg.max_out_len, downsampling_ratio=cfg.downsampling_ratio,
                         levels=list(cfg.levels))
    return init_params(tc, 3, flavour)


def test_fsq_constants_table():
    # SURVEY appendix A.9 (NumPy fp32)
    h, o, s, b = om.fsq_constants([8, 8, 8, 5, 5, 5])
    assert np.allclose(h, [3.4965, 3.4965, 3.4965, 1.998, 1.998, 1.998], atol=1e-6)
    assert np.allclose(s[:3], 0.143983, atol=1e-6) and (s[3:] == 0).all()
    assert list(b) == [1, 8, 64, 512, 2560, 12800]
    h, o, s, b = om.fsq_constants([4] * 6)
    assert np.allclose(h, 1.4985, atol=1e-6) and np.allclose(s, 0.346627, atol=1e-6)
    assert list(b) == [1, 4, 16, 64, 256, 1024]


@pytest.mark.parametrize("levels,code", [([4, 4, 3, 3, 3], 218), ([4, 4, 4, 3, 3, 3], 874), ([4] * 6, 2730), ([8, 8, 8, 5, 5, 5], 32036)])
def test_masked_token_code(levels, code):
    z = np.random.default_rng(0).normal(size=(4, len(levels))).astype(np.float32)
    t = om.fsq_tokens(z, levels, n_valid_tokens=2)
    assert (t[2:] == code).all()


@pytest.mark.parametrize("levels", [[4, 4, 3, 3, 3], [4, 4, 4, 3, 3, 3], [4] * 6, [8, 8, 8, 5, 5, 5]])
def test_pack_unpack_round_trip_and_range(levels):
    rng = np.random.default_rng(1)
    z = rng.normal(0, 2.0, (5000, len(levels))).astype(np.float32)
    b = om.fsq_bound(z, levels)
    t = om.fsq_pack(b, levels)
    assert t.max() < np.prod(levels)
    assert np.array_equal(om.indexes_to_codes(t, levels), np.rint(b).astype(np.int64))
    all_codes = om.indexes_to_codes(np.arange(np.prod(levels)), levels)
    assert np.array_equal(om.fsq_pack(all_codes.astype(np.float32), levels), np.arange(np.prod(levels)).astype(np.uint32))


def test_round_half_to_even():
    assert list(om.fsq_pack(np.array([[0.5], [1.5], [-0.5], [-1.5]], np.float32), [8])) == [4, 6, 4, 2]


@pytest.mark.parametrize("df,levels", [(1, [4] * 6), (4, [8, 8, 8, 5, 5, 5])])
def test_padding_independence(df, levels):
    """Reference-style padded forward (N = 512, self-loop padding edges, MaskedLN, -1e9 attention bias, dead ops)
    gives the same tokens as the ragged valid-rows-only forward."""
    cfg = om.OracleConfig(downsampling_ratio=df, max_out_len=512 // df, levels=levels)
    P = _params(cfg)
    g = _graph(5, 96)
    n, K, N = g["n_node"], 50, 512
    z_ragged = om.encode(P, cfg, g["edge_features"], g["senders"], n)
    send = np.concatenate([g["senders"], np.repeat(np.arange(n, N), K)])
    feat = np.concatenate([g["edge_features"], np.zeros(((N - n) * K, 27))])
    z_pad = om.encode(P, cfg, feat, send, N, n_valid=n, dense_attention=True, include_dead_ops=True)
    T = n // df
    assert np.abs(z_pad[:T] - z_ragged).max() < 2e-5
    t_pad = om.fsq_tokens(z_pad, levels, T)
    assert np.array_equal(t_pad[:T], om.fsq_tokens(z_ragged, levels))
    masked = {tuple([4] * 6): 2730, (8, 8, 8, 5, 5, 5): 32036}[tuple(levels)]
    assert (t_pad[T:] == masked).all()


def test_structure_permutation_invariance_of_batching():
    cfg = om.OracleConfig()
    P = _params(cfg)
    ga, gb = _graph(6, 64), _graph(7, 80)
    za = om.encode(P, cfg, ga["edge_features"], ga["senders"], ga["n_node"])
    za2 = om.encode(P, cfg, ga["edge_features"], ga["senders"], ga["n_node"])
    assert np.array_equal(za, za2)  # deterministic
    assert za.shape == (64, 6) and om.encode(P, cfg, gb["edge_features"], gb["senders"], gb["n_node"]).shape == (80, 6)


def test_pe_table_matches_formula():
    t = om.pe_table(np.array([0, 1, 7, -3]), 512)
    assert t.shape == (4, 128)
    assert np.allclose(t[0, 0::2], 1.0) and np.allclose(t[0, 1::2], 0.0)  # k odd -> cos(0), k even -> sin(0)
    k = 4  # even: sin(x*pi / n^(2k/d))
    assert abs(t[2, k - 1] - np.sin(7 * np.pi / 512 ** (2 * k / 128))) < 1e-6
    k = 5  # odd: cos(x*pi / n^(2(k-1)/d))
    assert abs(t[3, k - 1] - np.cos(-3 * np.pi / 512 ** (2 * (k - 1) / 128))) < 1e-6


def test_product_pe_table_is_the_same_data():
    from pst.weights import pe_table

    for n in (128, 512, 1024):
        pos = np.arange(-(n - 1), n)
        assert np.array_equal(pe_table(pos, n), om.pe_table(pos, n))
