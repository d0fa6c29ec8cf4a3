"""The oracle's model forward against fixtures produced by EXECUTING THE REFERENCE'S OWN MODEL SOURCE
(Vq3D.encode_and_quantize over the NumPy stand-ins for jax / haiku of tests/golden/refshim.py; generator:
tests/golden/make_golden_model.py).  This is what pins oracle/model.py: layer order, concat order, masks,
LayerNorm variants, layer_stack parameter slicing, FSQ constants and the Haiku parameter names all come from
the reference's code; only the array library underneath (NumPy fp32 instead of XLA) is not the reference's,
which is what the 2e-5 tolerance on the bounded latents covers."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import featurize as fz
from oracle import model as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(4096, 1), (64000, 4), (432, 1), (4096, 2), (64000, 1), (1728, 1)]
LARGE_CASES = [(64000, 1), (1728, 1)]  # >= 2 000 tokens each: tokens + bounded latents only
TOL_BOUNDED = 2e-5  # fp32 on both sides, different summation order / libm; |bounded| <= 3.5


def load_case(codebook, df):
    from pst.config import TokenizerConfig
    from pst.weights import init_params

    f = np.load(os.path.join(GOLDEN, f"model_ref_{codebook}_df{df}.npz"))
    levels = [int(x) for x in f["levels"]]
    cfg = TokenizerConfig(seq_max_size=int(f["seq_max_size"]), max_out_len=int(f["max_out_len"]), downsampling_ratio=df,
                          levels=levels, precision="fp32")
    params = init_params(cfg, int(f["seed"]), "rich")
    sha = hashlib.sha256(b"".join(np.ascontiguousarray(params[k]).tobytes() for k in sorted(params))).hexdigest()
    assert sha == str(f["param_sha256"]), "init_params no longer reproduces the weights the fixture was made with"
    return f, cfg, params


def oracle_cfg(cfg):
    return om.OracleConfig(seq_max_size=cfg.seq_max_size, graph_max_neighbor=cfg.num_neighbor,
                           downsampling_ratio=cfg.downsampling_ratio, max_out_len=cfg.max_out_len, levels=list(cfg.levels))


@pytest.mark.parametrize("codebook,df", CASES)
def test_oracle_matches_reference_source(casp14, codebook, df):
    f, cfg, params = load_case(codebook, df)
    ocfg = oracle_cfg(cfg)
    for name in (str(n) for n in f["names"]):
        e = casp14[name]
        g = fz.featurize(e["pos"].astype(np.float64), e["gt"], e["exists"], cfg.num_neighbor)
        assert g["n_node"] == int(f[f"{name}/n_valid"])
        z, inter = om.encode(params, ocfg, g["edge_features"], g["senders"], g["n_node"], return_intermediates=True)
        nt = g["n_node"] // df
        ref_b, ref_t = f[f"{name}/bounded"], f[f"{name}/tokens"]
        assert ref_b.shape == (nt, len(cfg.levels)) and ref_t.shape == (cfg.max_out_len,)
        if f"{name}/pre_proj" in f.files:
            assert np.abs(inter["pre_proj"] - f[f"{name}/pre_proj"]).max() < 2e-6
        assert np.abs(om.fsq_bound(z, cfg.levels) - ref_b).max() < TOL_BOUNDED
        t = om.fsq_tokens(z, cfg.levels)
        amb = om.fsq_ambiguous(z, cfg.levels, tol=1e-4)
        assert np.array_equal(t[~amb], ref_t[:nt][~amb])
        assert (t == ref_t[:nt]).mean() > 0.995
        # the reference's own packing of its own bounded values (model/quantize.py:188,209,113-120)
        assert np.array_equal(om.fsq_pack(ref_b, cfg.levels), ref_t[:nt])
        # padded tokens carry the all-zero code (mask -> bounded 0, model/quantize.py:184)
        assert (ref_t[nt:] == om.fsq_pack(np.zeros((1, len(cfg.levels)), np.float32), cfg.levels)[0]).all()


def test_oracle_reference_faithful_form_matches_too(casp14):
    """padded to seq_max_size, dense masked attention, dead ops kept: the form the CPU baseline times"""
    f, cfg, params = load_case(64000, 4)
    ocfg = oracle_cfg(cfg)
    name = "T1046s1"
    e = casp14[name]
    g = fz.featurize(e["pos"].astype(np.float64), e["gt"], e["exists"], cfg.num_neighbor)
    n, K, N = g["n_node"], cfg.num_neighbor, cfg.seq_max_size
    feats = np.zeros((N * K, 27), np.float32)
    feats[: n * K] = g["edge_features"]
    send = np.repeat(np.arange(N), K)  # padded rows: K self loops (data/preprocessing.py:261-271)
    send[: n * K] = g["senders"]
    z = om.encode(params, ocfg, feats, send, N, n_valid=n, dense_attention=True, include_dead_ops=True)
    nt = n // 4
    assert np.abs(om.fsq_bound(z[:nt], cfg.levels) - f[f"{name}/bounded"]).max() < TOL_BOUNDED
    assert np.array_equal(om.fsq_tokens(z, cfg.levels, n_valid_tokens=nt), f[f"{name}/tokens"])


def test_haiku_parameter_names_map_onto_ours():
    """every parameter the reference's `init` creates for the tokenize path is found, exactly once, by the
    suffix matching of pst/weights.py, with the same shape"""
    from pst.config import TokenizerConfig
    from pst.weights import _find, from_haiku, init_params, pack_weights

    with open(os.path.join(GOLDEN, "model_ref_param_names.json")) as fh:
        names = json.load(fh)
    for key, shapes in names.items():
        codebook, df = key.split("_df")
        cfg = TokenizerConfig.named(int(codebook), int(df))
        mine = init_params(cfg, 0, "rich")
        tree = {}
        rng = np.random.default_rng(0)
        for full, shp in shapes.items():
            mod, p = full.rsplit("/", 1)
            tree.setdefault("forward_vq3_d/" + mod, {})[p] = rng.standard_normal(shp).astype(np.float32)
        flat = from_haiku(tree)
        assert len(flat) == len(shapes) == len(mine)
        for k, v in mine.items():
            assert _find(flat, k).shape == v.shape, k
        blob = pack_weights(flat, cfg)
        assert blob.shape == pack_weights(mine, cfg).shape


def test_full_checkpoint_with_decoder_side_names_still_packs():
    """A released checkpoint also holds the decoder: `cross_attn_upsampling/cross_attn_scaler_iteration/...` repeats the
    down-sampler's inner names (model/model.py:70-98, modules.py:513), plus `up_proj`, the structure module, ...  The
    suffix matching must still find every encode-side parameter exactly once and ignore the rest."""
    from pst.config import TokenizerConfig
    from pst.weights import from_haiku, init_params, pack_weights

    cfg = TokenizerConfig.named(4096, 1)
    mine = init_params(cfg, 3, "rich")
    tree = {}
    rng = np.random.default_rng(1)
    for k, v in mine.items():
        mod, p = k.rsplit("/", 1)
        tree.setdefault("forward_vq3_d/vq3_d/~/" + mod, {})[p] = v
        if "cross_attn_downsampling" in mod:  # decoder twin with identical inner names and shapes, different values
            twin = mod.replace("cross_attn_downsampling", "cross_attn_upsampling")
            tree.setdefault("forward_vq3_d/vq3_d/~/" + twin, {})[p] = rng.standard_normal(v.shape).astype(np.float32)
    tree["forward_vq3_d/vq3_d/~decode/up_proj"] = {"w": rng.standard_normal((6, 128)).astype(np.float32), "b": np.zeros(128, np.float32)}
    tree["forward_vq3_d/vq3_d/~/structure_module/fold_iteration/invariant_point_attention/q_scalar"] = {
        "weights": rng.standard_normal((128, 192)).astype(np.float32), "bias": np.zeros(192, np.float32)}
    assert np.array_equal(pack_weights(from_haiku(tree), cfg), pack_weights(mine, cfg))
