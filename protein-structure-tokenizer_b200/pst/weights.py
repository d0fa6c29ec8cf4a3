"""Weights of the encode path: reference-rule random initialisation, the constant
positional-encoding tables, and the packer that turns Haiku-named arrays into the
prepared fp32 blob `pst_model_create` uploads (layout = the order of `pack_weights`,
mirrored by csrc/api.cu::pst_fill_weight_pointers).

Reference behaviour (paths under structure_tokenizer/):
  initialisers     model/utils.py:151-171 -> hk.initializers.VarianceScaling(1.0, "fan_in",
                   "truncated_normal"); biases 0; LayerNorm scale 1 / offset 0; gating_w 0,
                   gating_b 1 (model/modules.py:354-363)
  PE               model/positional_encoding_layer.py:49-66 (tables are constants of the model:
                   node x = i, edge x = sender - receiver, base n = seq_max_size,
                   model/structure_encoder.py:77-87; tokens x = t, base n = max_out_len,
                   model/modules.py:486-500)
  parameter names  Haiku module/param naming of model/{structure_encoder,gnn_layers,modules,model}.py;
                   a checkpoint is matched by name suffix (prefixes such as "vq3_d/~/" vary).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from .config import TokenizerConfig

D = 128
NUM_HEAD = 4
HEAD_DIM = 32
EDGE_FEATS = 27
FEAT_PAD = 32
C8 = 8


def pe_table(positions, n: int, d: int = D) -> np.ndarray:
    """Sinusoidal table, fp32.  k = 1..d: odd k -> cos(x*pi / n**(2(k-1)/d)), even k ->
    sin(x*pi / n**(2k/d)).  The argument follows the reference's fp32 operation order; the
    trig value is the correctly rounded fp32 function of that fp32 argument."""
    x = np.asarray(positions, np.int64).astype(np.float32)[:, None]
    k = np.arange(1, d + 1, dtype=np.int32)[None, :]
    odd = (k % 2) == 1
    expo = np.where(odd, (2 * (k - 1)).astype(np.float32) / np.float32(d), (2 * k).astype(np.float32) / np.float32(d)).astype(np.float32)
    den = np.power(np.float32(n), expo, dtype=np.float32)
    arg = ((x * np.float32(math.pi)).astype(np.float32) / den).astype(np.float32).astype(np.float64)
    return np.where(odd, np.cos(arg), np.sin(arg)).astype(np.float32)


def _trunc_normal(rng: np.random.Generator, shape, fan_in: int) -> np.ndarray:
    """hk.initializers.VarianceScaling(1.0, 'fan_in', 'truncated_normal')."""
    std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2
    return (out * std).astype(np.float32)


def init_params(cfg: TokenizerConfig, seed: int = 0, flavour: str = "ref") -> Dict[str, np.ndarray]:
    """Random weights of the encode path with the reference's initialisers.

    flavour 'ref'    : exactly the reference's init (biases 0, LN 1/0, gating_w 0 / gating_b 1).
            'spread' : 'ref' with down_proj.w x 12 (otherwise 1-7 codes are ever hit).
            'rich'   : 'spread' plus randomised biases, LN scale/offset and gating weights, so
                       that every parameter influences the output (used by parity tests).
    Names are Haiku-style `module/.../param`; scaler-block params carry a leading
    dim of sc_num_block (layer_stack, model/layer_stack.py:132-172)."""
    assert flavour in ("ref", "spread", "rich")
    rng = np.random.default_rng(seed)
    rich = flavour == "rich"
    p: Dict[str, np.ndarray] = {}

    def lin(name, fin, fout, wname="w", bname="b", lead=()):
        p[f"{name}/{wname}"] = _trunc_normal(rng, (*lead, fin, fout), fin)
        b = np.zeros((*lead, fout), np.float32)
        if rich:
            b = (0.1 * rng.standard_normal(b.shape)).astype(np.float32)
        p[f"{name}/{bname}"] = b

    def norm(name, lead=()):
        s = np.ones((*lead, D), np.float32)
        o = np.zeros((*lead, D), np.float32)
        if rich:
            s = (1.0 + 0.1 * rng.standard_normal(s.shape)).astype(np.float32)
            o = (0.1 * rng.standard_normal(o.shape)).astype(np.float32)
        p[f"{name}/scale"], p[f"{name}/offset"] = s, o

    lin("structure_encoder/init_node_embed", D, D)
    lin("structure_encoder/init_edge_embed", D + EDGE_FEATS, D)
    for l in range(cfg.gnn_layers):
        pre = "mpnn_layer" + ("" if l == 0 else f"_{l}")
        for i, (a, b) in enumerate([(3 * D, D), (D, D), (D, D)]):
            lin(f"{pre}/node_mlp_0/linear_{i}", a, b)
        for i, (a, b) in enumerate([(D, 4 * D), (4 * D, D)]):
            lin(f"{pre}/node_mlp_1/linear_{i}", a, b)
        for i, (a, b) in enumerate([(3 * D, D), (D, D), (D, D)]):
            lin(f"{pre}/edge_mlp/linear_{i}", a, b)
        for j in range(3):
            norm(f"{pre}/norm_msg" + ("" if j == 0 else f"_{j}"))
    nb = cfg.num_blocks
    it = "cross_attn_downsampling/cross_attn_scaler_iteration"
    norm(f"{it}/cross_attention/query_norm", (nb,))
    norm(f"{it}/cross_attention/data_norm", (nb,))
    att = f"{it}/cross_attention/attention"
    for w in ("query_w", "key_w", "value_w"):
        p[f"{att}/{w}"] = _trunc_normal(rng, (nb, D, NUM_HEAD, HEAD_DIM), NUM_HEAD * D)
    p[f"{att}/gating_w"] = np.zeros((nb, D, NUM_HEAD, HEAD_DIM), np.float32)
    p[f"{att}/gating_b"] = np.ones((nb, NUM_HEAD, HEAD_DIM), np.float32)
    p[f"{att}/output_w"] = _trunc_normal(rng, (nb, NUM_HEAD, HEAD_DIM, D), NUM_HEAD * HEAD_DIM)
    p[f"{att}/output_b"] = np.zeros((nb, D), np.float32)
    if rich:
        p[f"{att}/gating_w"] = _trunc_normal(rng, (nb, D, NUM_HEAD, HEAD_DIM), D)
        p[f"{att}/gating_b"] = (1.0 + 0.3 * rng.standard_normal((nb, NUM_HEAD, HEAD_DIM))).astype(np.float32)
        p[f"{att}/output_b"] = (0.1 * rng.standard_normal((nb, D))).astype(np.float32)
    for tr in ("resampled_transition", "original_transition"):
        norm(f"{it}/{tr}/input_layer_norm", (nb,))
        lin(f"{it}/{tr}/transition1", D, 2 * D, "weights", "bias", (nb,))
        lin(f"{it}/{tr}/transition2", 2 * D, D, "weights", "bias", (nb,))
    lin("down_proj", D, len(cfg.levels))
    if flavour in ("spread", "rich"):
        p["down_proj/w"] = p["down_proj/w"] * np.float32(12.0)
    return p


def canonical_name(name: str) -> str:
    """Haiku module path -> the name used here: the "~" / "~method" components Haiku inserts for
    modules built in __init__ / in a method other than __call__ are dropped, e.g.
    `vq3_d/~/structure_encoder/~/graph_neural_network/~/mpnn_layer/node_mlp_0/~/linear_0/w` ->
    `vq3_d/structure_encoder/graph_neural_network/mpnn_layer/node_mlp_0/linear_0/w`
    (the full list of the reference's names is pinned in tests/golden/model_ref_param_names.json)."""
    return "/".join(p for p in name.replace("//", "/").split("/") if p and not p.startswith("~"))


def from_haiku(tree) -> Dict[str, np.ndarray]:
    """A Haiku parameter tree ({module_name: {param: array}}, or a flat {"module/param": array} dict,
    e.g. a released checkpoint after scripts/inference_runner.py:236-248) -> flat dict keyed by
    canonical names.  Parameters of the decoder side are carried along untouched; `pack_weights`
    only looks up the encode-path names, by suffix."""
    flat: Dict[str, np.ndarray] = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            for n, a in v.items():
                flat[canonical_name(f"{k}/{n}")] = np.asarray(a)
        else:
            flat[canonical_name(k)] = np.asarray(v)
    return flat


def _find(params: Dict[str, np.ndarray], suffix: str) -> np.ndarray:
    if suffix in params:
        return np.asarray(params[suffix], np.float32)
    hits = [k for k in params if canonical_name(k).endswith("/" + suffix)]
    if len(hits) != 1:
        raise KeyError(f"parameter '{suffix}': {len(hits)} matches")
    return np.asarray(params[hits[0]], np.float32)


def pack_weights(params: Dict[str, np.ndarray], cfg: TokenizerConfig) -> np.ndarray:
    """Haiku-named fp32 arrays -> prepared blob.  The PE tables are folded into the input
    embeddings here, once per model (fp64 accumulate, one rounding to fp32):
        node_table    = PE_node . W_node + b_node                [seq_max_size, 128]
        edge_pe_table = PE_edge . W_edge[0:128] + b_edge         [2*seq_max_size-1, 128]
        edge_feat_w   = W_edge[128:155] zero-padded to 32 rows   [32, 128]
    """
    N = cfg.seq_max_size
    out = []

    def put(a, shape):
        a = np.asarray(a, np.float32)
        assert a.shape == tuple(shape), (a.shape, shape)
        out.append(np.ascontiguousarray(a).reshape(-1))

    wn, bn = _find(params, "structure_encoder/init_node_embed/w"), _find(params, "structure_encoder/init_node_embed/b")
    we, be = _find(params, "structure_encoder/init_edge_embed/w"), _find(params, "structure_encoder/init_edge_embed/b")
    put((pe_table(np.arange(N), N).astype(np.float64) @ wn.astype(np.float64) + bn).astype(np.float32), (N, D))
    pe_edge = pe_table(np.arange(-(N - 1), N), N).astype(np.float64)
    put((pe_edge @ we[:D].astype(np.float64) + be).astype(np.float32), (2 * N - 1, D))
    wf = np.zeros((FEAT_PAD, D), np.float32)
    wf[:EDGE_FEATS] = we[D:]
    put(wf, (FEAT_PAD, D))
    for l in range(cfg.gnn_layers):
        pre = "mpnn_layer" + ("" if l == 0 else f"_{l}")
        for i, shp in enumerate([(3 * D, D), (D, D), (D, D)]):
            put(_find(params, f"{pre}/node_mlp_0/linear_{i}/w"), shp)
            put(_find(params, f"{pre}/node_mlp_0/linear_{i}/b"), (D,))
        put(_find(params, f"{pre}/norm_msg/scale"), (D,))
        put(_find(params, f"{pre}/norm_msg/offset"), (D,))
        put(_find(params, f"{pre}/node_mlp_1/linear_0/w"), (D, 4 * D))
        put(_find(params, f"{pre}/node_mlp_1/linear_0/b"), (4 * D,))
        put(_find(params, f"{pre}/node_mlp_1/linear_1/w"), (4 * D, D))
        put(_find(params, f"{pre}/node_mlp_1/linear_1/b"), (D,))
        put(_find(params, f"{pre}/norm_msg_1/scale"), (D,))
        put(_find(params, f"{pre}/norm_msg_1/offset"), (D,))
        for i, shp in enumerate([(3 * D, D), (D, D), (D, D)]):
            put(_find(params, f"{pre}/edge_mlp/linear_{i}/w"), shp)
            put(_find(params, f"{pre}/edge_mlp/linear_{i}/b"), (D,))
        put(_find(params, f"{pre}/norm_msg_2/scale"), (D,))
        put(_find(params, f"{pre}/norm_msg_2/offset"), (D,))
    put(pe_table(np.arange(cfg.max_out_len), cfg.max_out_len), (cfg.max_out_len, D))
    # the parent scope matters: a full released checkpoint also holds the decoder's
    # `cross_attn_upsampling/cross_attn_scaler_iteration/...` with identical inner names (model/model.py:70-98)
    it = "cross_attn_downsampling/cross_attn_scaler_iteration"
    att = f"{it}/cross_attention/attention"
    for b in range(cfg.num_blocks):
        put(_find(params, f"{it}/cross_attention/query_norm/scale")[b], (D,))
        put(_find(params, f"{it}/cross_attention/query_norm/offset")[b], (D,))
        put(_find(params, f"{it}/cross_attention/data_norm/scale")[b], (D,))
        put(_find(params, f"{it}/cross_attention/data_norm/offset")[b], (D,))
        for w in ("query_w", "key_w", "value_w", "gating_w"):
            put(_find(params, f"{att}/{w}")[b].reshape(D, D), (D, D))
        put(_find(params, f"{att}/gating_b")[b].reshape(D), (D,))
        put(_find(params, f"{att}/output_w")[b].reshape(D, D), (D, D))
        put(_find(params, f"{att}/output_b")[b], (D,))
        for tr in ("resampled_transition", "original_transition"):
            put(_find(params, f"{it}/{tr}/input_layer_norm/scale")[b], (D,))
            put(_find(params, f"{it}/{tr}/input_layer_norm/offset")[b], (D,))
            put(_find(params, f"{it}/{tr}/transition1/weights")[b], (D, 2 * D))
            put(_find(params, f"{it}/{tr}/transition1/bias")[b], (2 * D,))
            put(_find(params, f"{it}/{tr}/transition2/weights")[b], (2 * D, D))
            put(_find(params, f"{it}/{tr}/transition2/bias")[b], (D,))
    C = len(cfg.levels)
    wd = np.zeros((D, C8), np.float32)
    wd[:, :C] = _find(params, "down_proj/w")
    bd = np.zeros((C8,), np.float32)
    bd[:C] = _find(params, "down_proj/b")
    put(wd, (D, C8))
    put(bd, (C8,))
    return np.concatenate(out)


def save_params(path: str, params: Dict[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "|"): v for k, v in params.items()})


def load_params_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as f:
        return {k.replace("|", "/"): f[k] for k in f.files}
