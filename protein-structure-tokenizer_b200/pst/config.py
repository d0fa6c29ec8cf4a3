"""Configuration of the tokenize path.

The reference composes hydra YAML files into an ml_collections.ConfigDict
(structure_tokenizer/utils/utils.py:30-58; files under config/structure_tokenizer/).
hydra is not a dependency here: `load_config` is a small PyYAML composer that honours
the two conventions the reference relies on -- a `defaults: [shared]` list inside a
model file and the `model=...` / `data=...` overrides of scripts/tokenize_pdb.py:102-113
-- and returns attribute-style nested dicts with the same key paths
(`cfg.model.model.codebook.levels`, `cfg.data.data.seq_max_size`, ...).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Any, Dict, List, Optional, Sequence

import yaml

CONFIG_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "configs", "structure_tokenizer")

PRECISIONS = {"fp32": 0, "fp16": 1, "bf16": 2}
CODEBOOK_SURNAME = {432: "0.5k", 1728: "1.7k", 4096: "4k", 64000: "64k"}  # scripts/tokenize_pdb.py:102-104


class AttrDict(dict):
    """dict with attribute access (stands in for ml_collections.ConfigDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return AttrDict({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load_yaml_with_defaults(path: str) -> dict:
    with open(path) as fh:
        doc = yaml.safe_load(fh) or {}
    merged: dict = {}
    for d in doc.pop("defaults", []) or []:
        if isinstance(d, str):
            name = d if d.endswith(".yaml") else d + ".yaml"
            # `shared` sits one directory above the gnn/ files (config/structure_tokenizer/model/shared.yaml)
            for cand in (os.path.join(os.path.dirname(path), name), os.path.join(os.path.dirname(os.path.dirname(path)), name)):
                if os.path.exists(cand):
                    merged = _merge(merged, _load_yaml_with_defaults(cand))
                    break
            else:
                raise FileNotFoundError(f"default '{d}' of {path} not found")
    return _merge(merged, doc)


def load_config(name: str = "vq3d_inference", job_name: str = "tokenize", overrides: Optional[Sequence[str]] = None,
                config_path: Optional[str] = None) -> AttrDict:
    """Same call shape as the reference's `load_config` (utils/utils.py:47-58)."""
    root = config_path or CONFIG_ROOT
    if not os.path.isabs(root):
        root = os.path.abspath(root)
    with open(os.path.join(root, name + ".yaml")) as fh:
        top = yaml.safe_load(fh) or {}
    groups: Dict[str, str] = {}
    for d in top.pop("defaults", []) or []:
        if isinstance(d, dict):
            groups.update({k: str(v) for k, v in d.items()})
    for ov in overrides or []:
        k, v = ov.split("=", 1)
        if k in ("model", "data"):
            groups[k] = v
        else:
            cur = top
            parts = k.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            cur[parts[-1]] = yaml.safe_load(v)
    cfg = dict(top)
    for group, fname in groups.items():
        fname = fname if fname.endswith(".yaml") else fname + ".yaml"
        cfg[group] = _load_yaml_with_defaults(os.path.join(root, group, fname))
    return _wrap(cfg)


@dataclasses.dataclass
class TokenizerConfig:
    """Flat view of the hyper-parameters the hot path reads (maps 1:1 onto pst_config)."""

    seq_max_size: int = 512
    max_out_len: int = 512
    num_neighbor: int = 50
    downsampling_ratio: int = 1
    levels: List[int] = dataclasses.field(default_factory=lambda: [4, 4, 4, 4, 4, 4])
    gnn_layers: int = 3
    num_blocks: int = 3
    precision: str = "fp16"
    max_len: Optional[int] = None  # longest accepted structure; defaults to seq_max_size

    def __post_init__(self):
        if self.max_len is None:
            self.max_len = self.seq_max_size
        if self.precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        if len(self.levels) > 8:
            raise ValueError("at most 8 FSQ levels")

    @property
    def num_codes(self) -> int:
        n = 1
        for l in self.levels:
            n *= l
        return n

    @classmethod
    def from_reference_cfg(cls, cfg: Any, precision: str = "fp16") -> "TokenizerConfig":
        """cfg = load_config(...) result, or any object with the same key paths."""
        data = cfg.data.data
        model = cfg.model.model
        if model.encoder.gnn.gnn_layer.layer_cls != "MPNNLayer":
            raise NotImplementedError("only MPNNLayer encoders are supported (all released configs)")
        if not model.codebook.use_codebook:
            raise NotImplementedError("continuous (no-codebook) variants are not supported")
        if model.codebook.get("renorm", False):
            raise NotImplementedError("codebook.renorm=true is not used by any released config")
        if not model.down_sampler.use_local_attn or model.down_sampler.get("use_global_node", 0):
            raise NotImplementedError("only local-attention down-samplers without a global node are supported")
        if not data.get("graph_residue_loc_is_alphac", True):
            # data/preprocessing.py:147-165: the alternative places the orientation features' positions on the side-chain
            # centroid; the kernels take them from CA (every released data config sets true)
            raise NotImplementedError("graph_residue_loc_is_alphac=false is not supported (all released configs use CA)")
        return cls(
            seq_max_size=int(data.seq_max_size),
            max_out_len=int(model.down_sampler.max_out_len),
            num_neighbor=int(data.graph_max_neighbor),
            downsampling_ratio=int(data.downsampling_ratio),
            levels=[int(x) for x in model.codebook.levels],
            gnn_layers=int(model.encoder.gnn.gnn_number_layers),
            num_blocks=int(model.down_sampler.sc_num_block),
            precision=precision,
        )

    @classmethod
    def named(cls, codebook_size: int = 4096, downsampling: int = 1, seq_max_size: int = 512, precision: str = "fp16") -> "TokenizerConfig":
        """The released ablation configs, by CLI flags (scripts/tokenize_pdb.py:80-98)."""
        levels = {432: [4, 4, 3, 3, 3], 1728: [4, 4, 4, 3, 3, 3], 4096: [4] * 6, 64000: [8, 8, 8, 5, 5, 5]}[codebook_size]
        return cls(seq_max_size=seq_max_size, max_out_len=seq_max_size // downsampling, downsampling_ratio=downsampling,
                   levels=levels, precision=precision)
