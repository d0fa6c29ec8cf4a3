#!/usr/bin/env python
"""Generate the committed golden fixtures from the REFERENCE ITSELF.

Runs only in the development container (needs /root/reference).  The reference
cannot be imported as a whole (no jax / haiku / BioPython in the image), but
its pure-NumPy hot-path functions can, through the stub modules below:

  * structure_tokenizer.data.preprocessing.preprocess_sample      (frames, filtering, k-NN, padding glue)
  * structure_tokenizer.utils.protein_utils.compute_nearest_neighbors_graph
  * structure_tokenizer.model.quat_affine.make_transform_from_reference

What is stubbed and why it does not touch the arithmetic: `jax.tree_map` /
`jax.tree_util.tree_map` (tuple mapping only), `jax.nn.one_hot`, the BioPython
parser (we feed atom37 arrays produced by oracle/pdb_ref.py instead), the
decoder-only `all_atom` module and `make_protein_features` (decoder loss
inputs, discarded by the tokenize path: scripts/inference_runner.py:64-72).

Outputs (tests/golden/):
  casp14_atom37.npz     inputs: the 31 bundled CASP14 structures as atom37 fp32 + masks
  casp14_graph_ref.npz  reference outputs: senders for all 31, edge features for 3,
                        fp64 checksums for the rest
It also asserts that oracle/featurize.py reproduces the reference bit-for-bit.
"""
import dataclasses
import glob
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _install_stubs():
    def mk(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    def tree_map(f, *xs):
        x0 = xs[0]
        if isinstance(x0, dict):
            return {k: tree_map(f, *[x[k] for x in xs]) for k in x0}
        if isinstance(x0, (list, tuple)):
            return type(x0)(tree_map(f, *[x[i] for x in xs]) for i in range(len(x0)))
        return f(*xs)

    jnp = mk("jax.numpy", ndarray=np.ndarray, float32=np.float32)
    tu = mk("jax.tree_util", tree_map=tree_map)
    nn = mk("jax.nn", one_hot=lambda x, num_classes: np.eye(num_classes, dtype=np.float32)[np.asarray(x)])
    mk("jax", numpy=jnp, tree_util=tu, tree_map=tree_map, nn=nn)
    mk("tree", map_structure=tree_map)
    mk("biopandas")
    mk("biopandas.pdb", PandasPdb=object)
    mk("jax_dataclasses", pytree_dataclass=dataclasses.dataclass)
    mk("haiku")
    mk("Bio")
    mk("Bio.PDB", PDBParser=object)
    mk("structure_tokenizer.model.all_atom")
    sys.path.insert(0, REF)


def main():
    _install_stubs()
    from structure_tokenizer.data import preprocessing
    from structure_tokenizer.data.protein_structure_sample import ProteinStructureSample
    from structure_tokenizer.model import quat_affine

    from oracle import featurize as ofz
    from oracle import pdb_ref

    ProteinStructureSample.make_protein_features = lambda self: {}

    files = sorted(glob.glob(os.path.join(REF, "casp14_pdbs", "*.pdb")))
    assert len(files) == 31, len(files)
    K, NPAD = 50, 512
    names, lens, pos_all, gt_all, ex_all = [], [], [], [], []
    nvalid, senders_all, feat_sha, feat_sum = [], [], [], []
    full_feat = {}
    for f in files:
        name = os.path.basename(f)[: -len(".pdb")]
        with open(f) as fh:
            s = pdb_ref.parse_pdb(fh.read())
        n = s["nb_residues"]
        sample = ProteinStructureSample(
            chain_id=None,
            nb_residues=n,
            aatype=np.eye(21, dtype=np.float32)[s["aatype"]],
            atom37_positions=s["atom37_positions"],
            atom37_gt_exists=s["atom37_gt_exists"],
            atom37_atom_exists=s["atom37_atom_exists"],
            resolution=0.0,
            pdb_cluster_size=1,
        )
        np.random.seed(0)
        g = preprocessing.preprocess_sample(
            sample=sample,
            num_neighbor=K,
            downsampling_ratio=1,
            residue_loc_is_alphac=True,
            padding_num_residue=NPAD,
            crop_index=NPAD,
            noise_level=0.0,
        ).graph
        nv = int(g.n_node[0])
        send_ref = np.asarray(g.senders)[: nv * K]
        recv_ref = np.asarray(g.receivers)[: nv * K]
        feat_ref = np.asarray(g.edge_features)[: nv * K]
        # padded tail semantics (preprocessing.py:261-271): K self loops per padded node
        assert (np.asarray(g.senders)[nv * K :] == np.repeat(np.arange(nv, NPAD), K)).all()
        assert (np.asarray(g.edge_features)[nv * K :] == 0).all()
        assert int(np.asarray(g.tokens_mask).sum()) == nv and int(np.asarray(g.nodes_mask).sum()) == nv

        # ---- oracle must reproduce the reference bit-for-bit -------------
        o = ofz.featurize(s["atom37_positions"], s["atom37_gt_exists"], s["atom37_atom_exists"], K)
        assert o["n_node"] == nv
        assert (o["senders"] == send_ref).all(), name
        assert (o["receivers"] == recv_ref).all(), name
        rot, _ = quat_affine.make_transform_from_reference(
            n_xyz=s["atom37_positions"][:, 0], ca_xyz=s["atom37_positions"][:, 1], c_xyz=s["atom37_positions"][:, 2]
        )
        keep = o["keep"]
        for col, key in enumerate(("u", "v", "n")):
            assert np.array_equal(rot[keep][:, :, col], o[key]), (name, key)
        ne = (o["edge_features"] != feat_ref).sum()
        ne32 = (o["edge_features"].astype(np.float32) != feat_ref.astype(np.float32)).sum()
        maxd = np.abs(o["edge_features"] - feat_ref).max()
        print(f"{name}: n={n} valid={nv} fp64-mismatch={ne} fp32-mismatch={ne32} max|d|={maxd:.3g}")
        assert ne32 == 0, name

        names.append(name)
        lens.append(n)
        pos_all.append(s["atom37_positions"].astype(np.float32))
        assert np.array_equal(pos_all[-1].astype(np.float64), s["atom37_positions"])
        gt_all.append(s["atom37_gt_exists"])
        ex_all.append(s["atom37_atom_exists"])
        nvalid.append(nv)
        senders_all.append(send_ref.astype(np.int16))
        f32 = np.ascontiguousarray(feat_ref.astype(np.float32))
        feat_sha.append(hashlib.sha256(f32.tobytes()).hexdigest())
        feat_sum.append(float(f32.astype(np.float64).sum()))
        if nv <= 100:
            full_feat[name] = f32

    np.savez_compressed(
        os.path.join(HERE, "casp14_atom37.npz"),
        names=np.array(names),
        lengths=np.array(lens, np.int32),
        atom37_positions=np.concatenate(pos_all),
        atom37_gt_exists=np.packbits(np.concatenate(gt_all), axis=1),
        atom37_atom_exists=np.packbits(np.concatenate(ex_all), axis=1),
    )
    np.savez_compressed(
        os.path.join(HERE, "casp14_graph_ref.npz"),
        names=np.array(names),
        n_valid=np.array(nvalid, np.int32),
        senders=np.concatenate(senders_all),
        edge_features_sha256=np.array(feat_sha),
        edge_features_sum=np.array(feat_sum),
        **{f"edge_features_{k}": v for k, v in full_feat.items()},
    )
    print("total residues", sum(lens), "valid", sum(nvalid), "full-feature fixtures:", list(full_feat))


if __name__ == "__main__":
    main()
