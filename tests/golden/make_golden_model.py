#!/usr/bin/env python
"""Generate the model-forward golden fixtures by EXECUTING THE REFERENCE'S OWN MODEL SOURCE.

Runs only in the development container (needs /root/reference).  JAX / Haiku are not installable
here, so `structure_tokenizer.model.model.Vq3D.encode_and_quantize` — the exact callable
`InferenceRunner.prepare_tokenize_fn` wraps (scripts/inference_runner.py:179-191) — is run over
the NumPy stand-ins of tests/golden/refshim.py (read its header for what that does and does not
pin).  Inputs are produced by the reference's own `preprocess_sample` (padded to seq_max_size, as
`make_graph_from_pdb` does, scripts/inference_runner.py:64-72) from bundled CASP14 structures; the
config is composed from the reference's yaml files the way `tokenize_pdb.py` does
(`model=gnn/ablation_<codebook>_df_<df>.yaml data=ablation_df_<df>.yaml`).

Weights: `pst.weights.init_params(cfg, seed, "rich")` (every parameter randomised so each one
influences the output), placed on the Haiku parameter tree that the reference's `init` creates.
That tree's names and shapes ARE part of the fixture (`haiku_param_shapes`): they pin the
checkpoint-name matching of pst/weights.py.

Outputs (tests/golden/):
  model_ref_<codebook>_df<df>.npz   per structure: tokens (uint32, all T rows), bounded latents
                                    `continuous_embedding` [T, C] and `continuous_embedding_pre_proj`
                                    [T, 128] (valid rows), plus names / n_valid / seed / param checksum
  model_ref_param_names.json        Haiku module/param names and shapes of the encode path
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "protein-structure-tokenizer_b200"))
sys.path.insert(0, HERE)

import refshim  # noqa: E402

SURNAME = {432: "0.5k", 1728: "1.7k", 4096: "4k", 64000: "64k"}
CASES = [  # (codebook, df, structures, weight seed)
    (4096, 1, ["T1024", "T1030"], 11),
    (64000, 4, ["T1024", "T1046s1"], 12),
    (432, 1, ["T1046s1"], 13),
    (4096, 2, ["T1030"], 14),
    # larger samples (>= 2 000 tokens each) for the default-mode agreement gate: tokens + bounded latents only
    (64000, 1, ["T1079", "T1037", "T1042", "T1025", "T1041", "T1067", "T1038", "T1090"], 15),
    (1728, 1, ["T1024", "T1030", "T1099", "T1032", "T1056", "T1027", "T1039", "T1043", "T1026", "T1054", "T1046s2", "T1049"], 16),
]
LARGE = {(64000, 1), (1728, 1)}  # no pre-projection embeddings stored (128 floats per token)


def deep_merge(a, b):
    out = dict(a)
    for k, v in b.items():
        out[k] = deep_merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def compose_config(codebook, df):
    """hydra.compose('vq3d_inference', overrides=[model=gnn/ablation_X_df_Y.yaml, data=ablation_df_Y.yaml])
    (scripts/tokenize_pdb.py:106-121, utils/utils.py:47-58): cfg.model = shared.yaml merged under the
    ablation file (its `defaults: [shared]`), cfg.data = the data file."""
    base = os.path.join(REF, "config", "structure_tokenizer")
    load = lambda *p: yaml.safe_load(open(os.path.join(base, *p)))  # noqa: E731
    abl = load("model", "gnn", f"ablation_{SURNAME[codebook]}_df_{df}.yaml")
    assert abl.pop("defaults") == ["shared"]
    model = deep_merge(load("model", "shared.yaml"), abl)
    data = load("data", f"ablation_df_{df}.yaml")
    root = load("vq3d_inference.yaml")
    return refshim.ConfigDict({"model": model, "data": data, "random_seed": root["random_seed"]})


def canonical(name: str) -> str:
    return "/".join(p for p in name.split("/") if not p.startswith("~"))


def main():
    refshim.install(REF)
    import jax
    import haiku as hk
    from structure_tokenizer.data import preprocessing
    from structure_tokenizer.data.protein_structure_sample import ProteinStructureSample
    from structure_tokenizer.model.model import Vq3D

    from pst.config import TokenizerConfig
    from pst.weights import init_params

    ProteinStructureSample.make_protein_features = lambda self: {}  # decoder-loss inputs, unused on this path

    a37 = np.load(os.path.join(HERE, "casp14_atom37.npz"))
    names = [str(n) for n in a37["names"]]
    offs = np.concatenate([[0], np.cumsum(a37["lengths"])])

    def sample(name):
        i = names.index(name)
        sl = slice(offs[i], offs[i + 1])
        n = int(offs[i + 1] - offs[i])
        return ProteinStructureSample(
            chain_id=None, nb_residues=n, aatype=np.eye(21, dtype=np.float32)[np.zeros(n, np.int64)],
            atom37_positions=a37["atom37_positions"][sl].astype(np.float64),
            atom37_gt_exists=np.unpackbits(a37["atom37_gt_exists"][sl], axis=1)[:, :37].astype(bool),
            atom37_atom_exists=np.unpackbits(a37["atom37_atom_exists"][sl], axis=1)[:, :37].astype(bool),
            resolution=0.0, pdb_cluster_size=1)

    only = {tuple(int(v) for v in a.split("_df")) for a in sys.argv[1:]}  # e.g. `64000_df1 1728_df1`: just these cases
    names_path = os.path.join(HERE, "model_ref_param_names.json")
    name_fixture = json.load(open(names_path)) if only and os.path.exists(names_path) else {}
    for codebook, df, structs, seed in CASES:
        if only and (codebook, df) not in only:
            continue
        cfg = compose_config(codebook, df)
        dc = cfg.data.data
        assert dc.downsampling_ratio == df

        # scripts/inference_runner.py:183-190
        def fn(graph, safe_key=None):
            return Vq3D(config=cfg.model, global_config=cfg.data).encode_and_quantize(graph, is_training=False, safe_key=safe_key)

        tokenize = hk.transform(fn)
        graphs = []
        for s in structs:
            np.random.seed(0)
            g = preprocessing.preprocess_sample(
                sample=sample(s), num_neighbor=dc.graph_max_neighbor, downsampling_ratio=dc.downsampling_ratio,
                residue_loc_is_alphac=dc.graph_residue_loc_is_alphac, padding_num_residue=dc.seq_max_size,
                crop_index=dc.seq_max_size, noise_level=0.0).graph
            # batch_collate([1, 1]) + pmap strips the device axis; device_put casts to fp32 / int32
            graphs.append(jax.tree_map(lambda x: refshim._asj(np.asarray(x)[None]), g))
        key = jax.random.PRNGKey(cfg.random_seed)

        t0 = time.time()
        ref_params = tokenize.init(key, graphs[0])
        shapes = {f"{m}/{p}": list(v.shape) for m, d in ref_params.items() for p, v in d.items()}
        print(f"[{codebook} df{df}] init: {len(shapes)} params, {sum(int(np.prod(s)) for s in shapes.values())} floats, {time.time() - t0:.1f}s")
        name_fixture[f"{codebook}_df{df}"] = shapes

        m = cfg.model.model
        tc = TokenizerConfig(seq_max_size=dc.seq_max_size, max_out_len=m.down_sampler.max_out_len, downsampling_ratio=df,
                             levels=list(m.codebook.levels))
        mine = init_params(tc, seed, "rich")
        # place our arrays on the reference's tree: every reference parameter must be covered, exactly once
        used = set()
        params = {}
        for mod, d in ref_params.items():
            params[mod] = {}
            for p, v in d.items():
                c = canonical(f"{mod}/{p}")
                hits = [k for k in mine if c == k or c.endswith("/" + k)]
                assert len(hits) == 1, (mod, p, hits)
                assert mine[hits[0]].shape == tuple(v.shape), (c, mine[hits[0]].shape, v.shape)
                params[mod][p] = refshim._asj(mine[hits[0]])
                used.add(hits[0])
        assert used == set(mine), sorted(set(mine) - used)
        sha = hashlib.sha256(b"".join(np.ascontiguousarray(mine[k]).tobytes() for k in sorted(mine))).hexdigest()

        out = {"names": np.array(structs), "seed": seed, "param_sha256": sha, "levels": np.array(m.codebook.levels),
               "seq_max_size": dc.seq_max_size, "max_out_len": m.down_sampler.max_out_len, "df": df}
        for s, g in zip(structs, graphs):
            t0 = time.time()
            q = tokenize.apply(params, key, g)
            nv = int(np.asarray(g.n_node).reshape(-1)[0])
            nt = int(np.asarray(g.tokens_mask).sum())
            assert nt == nv // df, (nt, nv, df)
            tok = np.asarray(q["tokens"])[0]
            out[f"{s}/tokens"] = tok.astype(np.uint32)
            out[f"{s}/bounded"] = np.asarray(q["continuous_embedding"])[0, :nt].astype(np.float32)
            if (codebook, df) not in LARGE:
                out[f"{s}/pre_proj"] = np.asarray(q["continuous_embedding_pre_proj"])[0, :nt].astype(np.float32)
            out[f"{s}/n_valid"] = nv
            print(f"  {s}: n_valid={nv} tokens={nt} distinct={len(np.unique(tok[:nt]))} "
                  f"perplexity={float(np.asarray(q['perplexity'])):.1f} {time.time() - t0:.1f}s")
        np.savez_compressed(os.path.join(HERE, f"model_ref_{codebook}_df{df}.npz"), **out)
    with open(os.path.join(HERE, "model_ref_param_names.json"), "w") as fh:
        json.dump(name_fixture, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
