"""Oracle (TEST INFRASTRUCTURE): torch-CPU fp32 restatement of the reference's
device forward `Vq3D.encode_and_quantize` (tokenize path only).  See
``oracle/__init__.py`` for who may import this.  Every function cites the
reference lines it follows (paths relative to structure_tokenizer/).

PINNED against the reference's own model source: tests/golden/make_golden_model.py
executes `Vq3D.encode_and_quantize` (the reference's files, unmodified) over NumPy
stand-ins for jax / haiku (tests/golden/refshim.py) on CASP14 structures for four
released configs; this restatement reproduces those outputs to <= 4e-6 on the bounded
latents with identical tokens (tests/test_golden_model.py).  NOT pinned: XLA's own
fp32 arithmetic (a real JAX run is impossible here: neither jax nor haiku is
installable) - summation order and the last ulp of tanh/exp/sin/cos/pow.

  positional encodings   model/positional_encoding_layer.py:49-150
  input embeddings       model/structure_encoder.py:55-123
  MaskedLayerNorm        model/gnn_layers.py:79-164
  MPNNLayer              model/gnn_layers.py:325-438
  GraphNeuralNetwork     model/modules.py:97-132
  Attention              model/modules.py:281-382
  CrossAttention         model/modules.py:393-424
  Transition             model/modules.py:211-262
  CrossAttentionScaler   model/modules.py:438-534, 545-636
  masks                  model/model.py:264-318, 382-401
  spherical norm, proj   model/model.py:148-174, 414-418
  FSQ                    model/quantize.py:93-120, 141-209
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

D = 128  # encoding_dimension == hidden_dimension == out_emb_size in every released config
NUM_HEAD = 4
HEAD_DIM = 32
EDGE_FEATS = 27

CODEBOOK_LEVELS = {
    432: [4, 4, 3, 3, 3],
    1728: [4, 4, 4, 3, 3, 3],
    4096: [4, 4, 4, 4, 4, 4],
    64000: [8, 8, 8, 5, 5, 5],
}


@dataclass
class OracleConfig:
    """The hyper-parameters the tokenize path reads (config/structure_tokenizer/
    data/ablation_df_*.yaml:15-23, model/gnn/ablation_*_df_*.yaml, model/shared.yaml)."""

    seq_max_size: int = 512
    graph_max_neighbor: int = 50
    downsampling_ratio: int = 1
    max_out_len: int = 512
    levels: List[int] = field(default_factory=lambda: [4, 4, 4, 4, 4, 4])
    gnn_number_layers: int = 3
    sc_num_block: int = 3

    @property
    def num_codes(self) -> int:
        return int(np.prod(self.levels))


# --------------------------------------------------------------------------- PE
def pe_table(positions: np.ndarray, n: int, d: int = D) -> np.ndarray:
    """positional_encoding_layer.py:49-66.  k = 1..d; odd k: cos(x*pi / n^(2(k-1)/d)),
    even k: sin(x*pi / n^(2k/d)).  The argument is formed in fp32 in the reference's
    operation order (float32(x)*float32(pi), divided by float32(n)**float32(exponent));
    the trig function is the correctly rounded fp32 value of that fp32 argument
    (XLA's own sin/cos differ from this by <= a few fp32 ulp)."""
    x = np.asarray(positions, np.int64).astype(np.float32)[:, None]
    k = np.arange(1, d + 1, dtype=np.int32)[None, :]
    odd = (k % 2) == 1
    expo = np.where(odd, (2 * (k - 1)).astype(np.float32) / np.float32(d), (2 * k).astype(np.float32) / np.float32(d))
    expo = expo.astype(np.float32)
    den = np.power(np.float32(n), expo, dtype=np.float32)
    arg = (x * np.float32(math.pi)).astype(np.float32) / den
    arg = arg.astype(np.float32).astype(np.float64)
    return np.where(odd, np.cos(arg), np.sin(arg)).astype(np.float32)


# ---------------------------------------------------------------------- helpers
def _t(x) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x))


def masked_layer_norm(x, mask, scale, offset, eps=1e-5):
    """gnn_layers.py:108-120,162-164 (biased variance over channels, mask on rows)."""
    x = mask * x
    mean = (mask * x).mean(-1, keepdim=True)
    var = (mask * ((x - mean) * (x - mean))).mean(-1, keepdim=True)
    inv = scale * torch.rsqrt(var + eps)
    return inv * (x - mean) + offset


def layer_norm(x, scale, offset, eps=1e-5):
    """hk.LayerNorm(axis=-1) as used at modules.py:227-229,409-415."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) * (x - mean)).mean(-1, keepdim=True)
    return scale * (x - mean) * torch.rsqrt(var + eps) + offset


def gelu(x):
    """jax.nn.gelu default (approximate=True, tanh form); gnn_layers.py:354."""
    return F.gelu(x, approximate="tanh")


class Matmul:
    """fp32 matmul, optionally with operand rounding emulation (used only to
    *predict* the token agreement of reduced-precision GPU modes)."""

    def __init__(self, mode: str = "fp32"):
        self.mode = mode

    def _round(self, x, dt):
        return x.to(dt).to(torch.float32)

    def __call__(self, a, b, level: str = "node"):
        m = self.mode
        if m == "fp32" or (m.endswith("-edge") and level != "edge"):
            return a @ b
        base = m.replace("-edge", "")
        if base in ("bf16", "fp16"):
            dt = torch.bfloat16 if base == "bf16" else torch.float16
            return self._round(a, dt) @ self._round(b, dt)
        if base in ("bf16x3", "fp16x3", "fp16x2"):
            dt = torch.bfloat16 if base.startswith("bf16") else torch.float16
            ah, bh = self._round(a, dt), self._round(b, dt)
            al, bl = self._round(a - ah, dt), self._round(b - bh, dt)
            if base == "fp16x2":  # activations split, weights single
                return ah @ bh + al @ bh
            return ah @ bh + (ah @ bl + al @ bh)
        raise ValueError(m)


# ---------------------------------------------------------------------- forward
@torch.no_grad()
def encode(
    params: Dict[str, np.ndarray],
    cfg: OracleConfig,
    edge_features: np.ndarray,  # [N*K, 27] (fp32 or fp64; cast to fp32 like device_put does)
    senders: np.ndarray,  # [N*K]
    n_rows: int,  # N: number of node rows present (padded 512, or ragged n_valid)
    n_valid: Optional[int] = None,  # rows < n_valid are real residues
    dense_attention: bool = False,  # reference-faithful [T,N] masked softmax instead of the local form
    include_dead_ops: bool = False,  # layer-3 edge update and last original transition
    matmul: Optional[Matmul] = None,
    return_intermediates: bool = False,
):
    """model/model.py:357-420 for one structure.  Returns the pre-quantisation
    latents z [T, C] (T = max_out_len rows when padded, floor(n/df) when ragged)."""
    mm = matmul or Matmul()
    P = {k: _t(v) for k, v in params.items()}
    N = n_rows
    n_valid = N if n_valid is None else n_valid
    K = len(senders) // N
    df = cfg.downsampling_ratio
    s = _t(np.asarray(senders, np.int64))
    r = torch.arange(N).repeat_interleave(K)
    f = _t(np.asarray(edge_features, np.float32))
    mask = (torch.arange(N) < n_valid).to(torch.float32)[:, None]
    inter = {}

    # structure_encoder.py:77-105 -------------------------------------------------
    pe_node = _t(pe_table(np.arange(N), cfg.seq_max_size))
    pe_edge_tab = _t(pe_table(np.arange(-(cfg.seq_max_size - 1), cfg.seq_max_size), cfg.seq_max_size))
    pe_edge = pe_edge_tab[(s - r) + (cfg.seq_max_size - 1)]
    h = mm(pe_node, P["structure_encoder/init_node_embed/w"]) + P["structure_encoder/init_node_embed/b"]
    e = mm(torch.cat([pe_edge, f], -1), P["structure_encoder/init_edge_embed/w"], "edge") + P[
        "structure_encoder/init_edge_embed/b"
    ]
    inter["h0"], inter["e0"] = h, e

    def mlp(x, pre, n, level):
        for i in range(n):
            x = mm(x, P[f"{pre}/linear_{i}/w"], level) + P[f"{pre}/linear_{i}/b"]
            if i < n - 1:
                x = gelu(x)
        return x

    # modules.py:116-132 ; gnn_layers.py:325-438 -------------------------------------
    for l in range(cfg.gnn_number_layers):
        pre = "mpnn_layer" + ("" if l == 0 else f"_{l}")
        m = mlp(torch.cat([h[s], h[r], e], -1), f"{pre}/node_mlp_0", 3, "edge")
        agg = m.reshape(N, K, D).sum(1) / K
        h = masked_layer_norm(h + agg, mask, P[f"{pre}/norm_msg/scale"], P[f"{pre}/norm_msg/offset"])
        h = masked_layer_norm(
            h + mlp(h, f"{pre}/node_mlp_1", 2, "node"), mask, P[f"{pre}/norm_msg_1/scale"], P[f"{pre}/norm_msg_1/offset"]
        )
        if l < cfg.gnn_number_layers - 1 or include_dead_ops:
            em = mlp(torch.cat([h[s], h[r], e], -1), f"{pre}/edge_mlp", 3, "edge")
            e = masked_layer_norm(
                (e + em).reshape(N, K, D), mask[:, None], P[f"{pre}/norm_msg_2/scale"], P[f"{pre}/norm_msg_2/offset"]
            ).reshape(N * K, D)
        inter[f"h{l + 1}"] = h
        inter[f"e{l + 1}"] = e

    # modules.py:486-534 ---------------------------------------------------------
    T = cfg.max_out_len if dense_attention else n_valid // df
    t_valid = n_valid // df
    res = _t(pe_table(np.arange(T), cfg.max_out_len))
    orig = h
    it = "cross_attn_downsampling/cross_attn_scaler_iteration"
    att = f"{it}/cross_attention/attention"
    for b in range(cfg.sc_num_block):
        qn = layer_norm(res, P[f"{it}/cross_attention/query_norm/scale"][b], P[f"{it}/cross_attention/query_norm/offset"][b])
        dn = layer_norm(orig, P[f"{it}/cross_attention/data_norm/scale"][b], P[f"{it}/cross_attention/data_norm/offset"][b])
        q = mm(qn, P[f"{att}/query_w"][b].reshape(D, D)).reshape(-1, NUM_HEAD, HEAD_DIM) * HEAD_DIM ** (-0.5)
        k = mm(dn, P[f"{att}/key_w"][b].reshape(D, D)).reshape(-1, NUM_HEAD, HEAD_DIM)
        v = mm(dn, P[f"{att}/value_w"][b].reshape(D, D)).reshape(-1, NUM_HEAD, HEAD_DIM)
        if dense_attention:
            # model.py:382,264-318: base mask x local mask; modules.py:407 bias
            tt = torch.arange(T)[:, None]
            jj = torch.arange(N)[None, :]
            amask = ((tt < t_valid) & (jj < n_valid) & (jj // df == tt)).to(torch.float32)
            logits = torch.einsum("qhc,khc->hqk", q, k) + 1e9 * (amask - 1.0)[None]
            w = torch.softmax(logits, -1)
            wa = torch.einsum("hqk,khc->qhc", w, v)
        else:
            kk = k[: T * df].reshape(T, df, NUM_HEAD, HEAD_DIM)
            vv = v[: T * df].reshape(T, df, NUM_HEAD, HEAD_DIM)
            logits = torch.einsum("thc,tjhc->thj", q, kk)
            w = torch.softmax(logits, -1)
            wa = torch.einsum("thj,tjhc->thc", w, vv)
        gate = torch.sigmoid(mm(qn, P[f"{att}/gating_w"][b].reshape(D, D)).reshape(-1, NUM_HEAD, HEAD_DIM) + P[f"{att}/gating_b"][b])
        wa = (wa * gate).reshape(-1, D)
        res = res + (mm(wa, P[f"{att}/output_w"][b].reshape(D, D)) + P[f"{att}/output_b"][b])

        def transition(x, name):
            y = layer_norm(x, P[f"{it}/{name}/input_layer_norm/scale"][b], P[f"{it}/{name}/input_layer_norm/offset"][b])
            y = torch.relu(mm(y, P[f"{it}/{name}/transition1/weights"][b]) + P[f"{it}/{name}/transition1/bias"][b])
            return mm(y, P[f"{it}/{name}/transition2/weights"][b]) + P[f"{it}/{name}/transition2/bias"][b]

        res = res + transition(res, "resampled_transition")
        if b < cfg.sc_num_block - 1 or include_dead_ops:
            orig = orig + transition(orig, "original_transition")
    inter["resampled"] = res

    # model.py:169-174, 148-164 ------------------------------------------------
    zn = res / (torch.linalg.norm(res, ord=2, dim=-1, keepdim=True) + 1e-6)
    z = mm(zn, P["down_proj/w"]) + P["down_proj/b"]
    inter["pre_proj"] = zn
    if return_intermediates:
        return z.numpy(), {k: v.numpy() for k, v in inter.items()}
    return z.numpy()


# -------------------------------------------------------------------------- FSQ
def fsq_constants(levels):
    """quantize.py:175-181 in fp32 (levels int32 -> float32 arithmetic)."""
    lv = np.asarray(levels, np.int32)
    half_l = ((lv - 1).astype(np.float32) * np.float32(1 - 1e-3)) / np.float32(2)
    offset = np.where(lv % 2 == 0, np.float32(0.5), np.float32(0.0)).astype(np.float32)
    shift = np.tan(offset / half_l, dtype=np.float32)
    basis = np.concatenate([[1], np.cumprod(lv[:-1])]).astype(np.int64)
    return half_l, offset, shift, basis


def fsq_bound(z: np.ndarray, levels, dtype=np.float32) -> np.ndarray:
    half_l, offset, shift, _ = fsq_constants(levels)
    z = np.asarray(z, dtype)
    return np.tanh(z + shift.astype(dtype)) * half_l.astype(dtype) - offset.astype(dtype)


def fsq_pack(bounded: np.ndarray, levels) -> np.ndarray:
    """quantize.py:188,209,113-120: round half to even, shift by L//2, mixed-radix pack."""
    lv = np.asarray(levels, np.int64)
    _, _, _, basis = fsq_constants(levels)
    q = np.rint(np.asarray(bounded)).astype(np.int64)
    return ((q + lv // 2) * basis).sum(-1).astype(np.uint32)


def fsq_tokens(z: np.ndarray, levels, n_valid_tokens: Optional[int] = None) -> np.ndarray:
    b = fsq_bound(z, levels)
    if n_valid_tokens is not None:  # quantize.py:184 mask -> bounded 0 -> code sum(L//2 * basis)
        b = b.copy()
        b[n_valid_tokens:] = 0
    return fsq_pack(b, levels)


def fsq_ambiguous(z: np.ndarray, levels, tol: float = 1e-5) -> np.ndarray:
    """Tokens whose fp64 bounded value lies within `tol` of a rounding boundary
    in some dimension: there fp32 tanh implementations may legitimately disagree."""
    b = fsq_bound(z, levels, np.float64)
    frac = np.abs(b - np.floor(b) - 0.5)
    return (frac < tol).any(-1)


def indexes_to_codes(tokens: np.ndarray, levels) -> np.ndarray:
    """quantize.py:122-139 with renorm=False: integer codes in [-L//2, ...]."""
    lv = np.asarray(levels, np.int64)
    _, _, _, basis = fsq_constants(levels)
    t = np.asarray(tokens, np.int64)[..., None]
    return ((t // basis) % lv) - lv // 2


def tokenize_structure(params, cfg: OracleConfig, atom_pos, gt_exists, atom_exists, **kw) -> np.ndarray:
    """Full oracle path for one structure: featurise (fp64) -> encode (fp32) -> FSQ."""
    from oracle import featurize as fz

    g = fz.featurize(atom_pos, gt_exists, atom_exists, cfg.graph_max_neighbor)
    z = encode(params, cfg, g["edge_features"], g["senders"], g["n_node"], **kw)
    return fsq_tokens(z, cfg.levels)
