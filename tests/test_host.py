"""Host-side logic (CPU): config composer, weight packing, synthetic generator."""
import numpy as np
import pytest


def test_config_composer_matches_cli_selection():
    from pst.config import CODEBOOK_SURNAME, TokenizerConfig, load_config

    for size, df in ((4096, 1), (64000, 4), (432, 1), (1728, 1), (4096, 2)):
        cfg = load_config("vq3d_inference", overrides=[f"model=gnn/ablation_{CODEBOOK_SURNAME[size]}_df_{df}.yaml", f"data=ablation_df_{df}.yaml"])
        tc = TokenizerConfig.from_reference_cfg(cfg)
        assert tc == TokenizerConfig.named(size, df)
        assert tc.num_codes == size and cfg.model.model.codebook.num_codes == size
        assert cfg.data.data.graph_max_neighbor == 50 and cfg.data.data.seq_max_size == 512
        assert cfg.model.weight_paths == f"weights/{CODEBOOK_SURNAME[size]}_df_{df}/"
    assert load_config("vq3d_inference").random_seed == 0


def test_unsupported_configs_are_rejected():
    from pst.config import TokenizerConfig, load_config

    cfg = load_config("vq3d_inference")
    cfg.model.model.encoder.gnn.gnn_layer.layer_cls = "GNNLayer"
    with pytest.raises(NotImplementedError):
        TokenizerConfig.from_reference_cfg(cfg)
    with pytest.raises(ValueError):
        TokenizerConfig(precision="int8")


def test_init_params_follow_reference_rules():
    from pst.config import TokenizerConfig
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1)
    p = init_params(cfg, 0, "ref")
    w = p["mpnn_layer_1/edge_mlp/linear_0/w"]
    assert w.shape == (384, 128)
    std = np.sqrt(1 / 384)
    assert abs(w.std() - std) / std < 0.05 and np.abs(w).max() <= 2 * std / 0.8796 + 1e-6
    assert (p["mpnn_layer/node_mlp_0/linear_0/b"] == 0).all()
    att = "cross_attn_downsampling/cross_attn_scaler_iteration/cross_attention/attention"
    assert (p[f"{att}/gating_w"] == 0).all() and (p[f"{att}/gating_b"] == 1).all()
    assert p[f"{att}/query_w"].shape == (3, 128, 4, 32)
    q = p[f"{att}/query_w"]
    assert abs(q.std() - np.sqrt(1 / 512)) / np.sqrt(1 / 512) < 0.05  # fan_in = 4*128 (Haiku conv-style rule)
    s = init_params(cfg, 0, "spread")
    assert np.allclose(s["down_proj/w"], 12 * p["down_proj/w"])
    assert len(p) == 95


def test_pack_weights_layout_and_suffix_matching():
    from pst.config import TokenizerConfig
    from pst.weights import init_params, pack_weights, pe_table

    cfg = TokenizerConfig.named(4096, 1)
    p = init_params(cfg, 1, "rich")
    blob = pack_weights(p, cfg)
    prefixed = {"vq3_d/~/" + k: v for k, v in p.items()}
    assert np.array_equal(pack_weights(prefixed, cfg), blob)
    node = blob[: 512 * 128].reshape(512, 128)
    ref = pe_table(np.arange(512), 512) @ p["structure_encoder/init_node_embed/w"] + p["structure_encoder/init_node_embed/b"]
    assert np.abs(node - ref).max() < 1e-5
    wf = blob[512 * 128 + 1023 * 128 :][: 32 * 128].reshape(32, 128)
    assert np.array_equal(wf[:27], p["structure_encoder/init_edge_embed/w"][128:]) and (wf[27:] == 0).all()
    assert np.array_equal(blob[-8:-2], p["down_proj/b"]) and (blob[-2:] == 0).all()


def test_synthetic_backbones_are_seeded_and_physical():
    from pst import synthetic as syn

    a = syn.make_backbones(123, [64, 96])
    b = syn.make_backbones(123, [64, 96])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    bb = a[1]
    assert bb.dtype == np.float32 and bb.shape == (96, 4, 3)
    assert np.array_equal(np.round(bb.astype(np.float64), 3).astype(np.float32), bb)
    ca = bb[:, 1].astype(np.float64)
    bond = np.linalg.norm(ca[1:] - ca[:-1], axis=-1)
    assert np.abs(bond - 3.8).max() < 0.01
    d = np.linalg.norm(ca[:, None] - ca[None], axis=-1) + 99 * np.eye(96)
    assert d.min() > 3.4
    atoms, offs = syn.pack_backbones(a)
    assert atoms.shape == (160, 4, 3) and list(offs) == [0, 64, 160]
    L = syn.bucketed_lengths(1, 1000)
    assert L.min() >= 64 and L.max() <= 2048 and (L % 64 == 0).all()


def test_load_and_build_batch_pads_and_truncates(tmp_path):
    """Token files -> padded int32 batch (scripts/inference_runner.py:114-133)."""
    from pst.inference_runner import load_and_build_batch

    a = np.arange(5, dtype=np.uint32).reshape(1, -1)
    b = (np.arange(9, dtype=np.uint32) + 100).reshape(1, -1)
    np.save(tmp_path / "a_tokens", a)
    np.save(tmp_path / "b_tokens", b)
    out = load_and_build_batch([str(tmp_path / "a_tokens.npy"), str(tmp_path / "b_tokens.npy")], 7, 4097)
    assert out.dtype == np.int32 and out.shape == (2, 7)
    assert out[0].tolist() == [0, 1, 2, 3, 4, 4097, 4097]
    assert out[1].tolist() == [100, 101, 102, 103, 104, 105, 106]


def test_runner_host_pipeline_with_a_stand_in_callable(built_lib, tmp_path):
    """InferenceRunner.tokenize's host side without a GPU: files are parsed side by side through the C ABI parser,
    batches reach the callable in list order (the last one padded by cycling the list, scripts/inference_runner.py:
    268-275), every structure's tokens land in `<stem>_tokens.npy` as uint32 (1, n), and the first bad file of a
    batch, in list order, raises the reference's error."""
    import types

    import pytest
    from pst import synthetic as syn
    from pst.inference_runner import InferenceRunner
    from test_pdb import _pdb_from_backbone

    lengths = [60, 75, 52, 90, 66]
    bbs = syn.make_backbones(17, lengths)
    files = []
    for i, bb in enumerate(bbs):
        f = tmp_path / f"P{i}.pdb"
        f.write_text(_pdb_from_backbone(bb))
        files.append(str(f))
    seen = []

    def quantize(params, rng, batch):  # stands in for the device call: token t of a structure = its CA count + t
        seen.append([a.shape[0] for a, _ in batch])
        for a, m in batch:
            assert a.shape[1:] == (37, 3) and a.dtype == np.float32 and m.shape == a.shape[:2]
        return {"tokens": [np.arange(a.shape[0], dtype=np.int32) + a.shape[0] for a, _ in batch]}

    data_cfg = types.SimpleNamespace(graph_max_neighbor=50, seq_max_size=512)
    out = tmp_path / "tokens"
    InferenceRunner.tokenize(None, quantize, None, files, str(out), 1, data_cfg, batch_size_per_device=2)
    assert seen == [[60, 75], [52, 90], [66, 60]]  # 5 files, batches of 2: the list is cycled to fill the last batch
    for i, n in enumerate(lengths):
        t = np.load(out / f"P{i}_tokens.npy")
        assert t.dtype == np.uint32 and t.shape == (1, n) and t[0, 0] == n and t[0, -1] == 2 * n - 1
    # a too-short structure and an insertion code in the same batch: the one that comes first in the list raises
    short = tmp_path / "short.pdb"
    short.write_text(_pdb_from_backbone(bbs[0][:20]))
    bad = tmp_path / "bad.pdb"
    bad.write_text(_pdb_from_backbone(bbs[1]).replace(" A   7 ", " A   7B"))
    with pytest.raises(NotImplementedError):
        InferenceRunner.tokenize(None, quantize, None, [str(short), str(bad)], str(tmp_path / "o2"), 1, data_cfg, batch_size_per_device=2)
    with pytest.raises(ValueError):
        InferenceRunner.tokenize(None, quantize, None, [str(bad), str(short)], str(tmp_path / "o3"), 1, data_cfg, batch_size_per_device=2)


def test_preprocessed_sample_files_round_trip_and_feed_the_runner(built_lib, tmp_path):
    """ProteinStructureSample .npy files (structure_tokenizer/data/protein_structure_sample.py:46-62): written and read
    back bit for bit, equal to what the PDB parser produces for the same structure, accepted by the runner next to PDB
    files, and the reference's bundled example parses when the checkout is there."""
    import os
    import types

    from pst import pdb as ppdb
    from pst import synthetic as syn
    from pst.inference_runner import InferenceRunner, load_structures
    from test_pdb import _pdb_from_backbone

    bbs = syn.make_backbones(29, [64, 80])
    texts = [_pdb_from_backbone(bb, resname=rn) for bb, rn in zip(bbs, ("ALA", "XYZ"))]
    paths = []
    for i, t in enumerate(texts):
        s = ppdb.structure_from_pdb_bytes_native(t.encode())
        f = str(tmp_path / f"S{i}.npy")
        ppdb.structure_to_sample_file(s, f, chain_id=f"S{i}")
        r = ppdb.structure_from_sample_file(f)
        assert r.nb_residues == s.nb_residues and np.array_equal(r.aatype, s.aatype)
        assert np.array_equal(r.atom37_positions, s.atom37_positions)
        assert np.array_equal(r.atom37_gt_exists, s.atom37_gt_exists) and np.array_equal(r.atom37_atom_exists, s.atom37_atom_exists)
        d = np.load(f, allow_pickle=True)[()]  # the reference's `cls(**dict_representation)` needs exactly these keys
        assert sorted(d) == sorted(["chain_id", "nb_residues", "aatype", "atom37_positions", "atom37_gt_exists",
                                    "atom37_atom_exists", "resolution", "pdb_cluster_size"])
        assert d["aatype"].shape == (s.nb_residues, 21) and d["aatype"].dtype == np.bool_
        paths.append(f)
    pdb_path = tmp_path / "S1.pdb"
    pdb_path.write_text(texts[1])
    mixed = load_structures([paths[0], str(pdb_path), paths[1]], 50, 512, 2)
    assert np.array_equal(mixed[1][0], mixed[2][0]) and np.array_equal(mixed[1][1], mixed[2][1])
    seen = []

    def quantize(params, rng, batch):
        seen.append([a.shape[0] for a, _ in batch])
        return {"tokens": [np.zeros(a.shape[0], np.int32) for a, _ in batch]}

    out = tmp_path / "tok"
    InferenceRunner.tokenize(None, quantize, None, [paths[0], str(pdb_path)], str(out), 1,
                             types.SimpleNamespace(graph_max_neighbor=50, seq_max_size=512), batch_size_per_device=2)
    assert seen == [[64, 80]] and sorted(os.listdir(out)) == ["S0_tokens.npy", "S1_tokens.npy"]
    with pytest.raises(FileNotFoundError):
        ppdb.structure_from_sample_file(str(tmp_path / "missing.npy"))
    ref = "/root/reference/structure_tokenizer/data/test_data/cif_raw_data.npy"
    if os.path.isfile(ref):
        s = ppdb.structure_from_sample_file(ref)
        assert s.nb_residues == 123 and s.atom37_positions.shape == (123, 37, 3) and s.valid_backbone().sum() >= 50


def _pad_like_the_reference(g, K, N, df):
    """data/preprocessing.py:191-283 on an oracle graph: zero features and K self loops per padded residue."""
    n = g["n_node"]
    feats = np.zeros((N * K, 27), np.float32)
    feats[: n * K] = g["edge_features"]
    send = np.repeat(np.arange(N), K)
    send[: n * K] = g["senders"]
    recv = np.repeat(np.arange(N), K)
    nodes_mask = (np.arange(N) < n)[:, None]
    tokens_mask = (np.arange(N // df) < n // df)[:, None]
    return {"n_node": np.array([n]), "n_edge": np.array([n * K]), "nodes_mask": nodes_mask, "edge_features": feats,
            "tokens_mask": tokens_mask, "senders": send, "receivers": recv}


def test_padded_protein_graph_batches_convert_to_the_ragged_call(casp14):
    """The reference's own batch (ProteinGraph leaves stacked to [Dev, B, ...] by batch_collate,
    scripts/inference_runner.py:77-83; or a BatchDataVQ3D wrapping one, types.py:78-87) -> ragged B1 inputs."""
    import types

    from oracle import featurize as fz
    from pst.inference_runner import is_padded_graph, padded_graph_to_ragged

    K, N = 50, 512
    names = ["T1046s1", "T1031", "T1073", "T1082"]
    gs = []
    for n in names:
        e = casp14[n]
        gs.append(fz.featurize(e["pos"].astype(np.float64), e["gt"], e["exists"], K))
    padded = [_pad_like_the_reference(g, K, N, 1) for g in gs]
    batch = {k: np.stack([p[k] for p in padded]).reshape(2, 2, *padded[0][k].shape) for k in padded[0]}  # [Dev=2, B=2, ...]
    graph = types.SimpleNamespace(**batch)
    wrapped = types.SimpleNamespace(graph=graph, features={})
    assert is_padded_graph(batch) and is_padded_graph(graph) and is_padded_graph(wrapped)
    assert not is_padded_graph([(np.zeros((64, 4, 3), np.float32), None)])
    for b in (batch, graph, wrapped):
        lead, n_valid, feats, send, offsets, n_pad = padded_graph_to_ragged(b, K)
        assert lead == (2, 2) and n_pad == N and n_valid.tolist() == [g["n_node"] for g in gs]
        assert offsets.tolist() == np.concatenate([[0], np.cumsum(n_valid)]).tolist()
        assert np.array_equal(feats, np.concatenate([g["edge_features"] for g in gs]).astype(np.float32))
        assert np.array_equal(send, np.concatenate([g["senders"] for g in gs]))
    with pytest.raises(ValueError):
        padded_graph_to_ragged({**batch, "edge_features": batch["edge_features"][..., :26]}, K)


def test_sample_file_reader_refuses_foreign_pickles(tmp_path):
    """A stray .npy in --pdb_dir must not be able to run code: only NumPy's own reconstruction helpers resolve."""
    import os
    import pickle

    from pst import pdb as ppdb

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > " + str(tmp_path / "pwned"),))

    bad = tmp_path / "evil.npy"
    np.save(bad, {"nb_residues": 1, "x": Evil()}, allow_pickle=True)
    with pytest.raises(pickle.UnpicklingError):
        ppdb.structure_from_sample_file(str(bad))
    assert not (tmp_path / "pwned").exists()
    plain = tmp_path / "plain.npy"
    np.save(plain, np.zeros((4, 3), np.float32))
    with pytest.raises(ValueError):
        ppdb.structure_from_sample_file(str(plain))


def test_load_params_names_the_conversion_step_for_released_checkpoints(tmp_path):
    from pst.config import TokenizerConfig
    from pst.inference_runner import InferenceRunner
    from pst.weights import init_params, save_params

    cfg = TokenizerConfig.named(4096, 1)
    params = init_params(cfg, 0, "spread")
    released = tmp_path / "released"
    released.mkdir()
    np.savez(released / "params.npz", *[params[k] for k in sorted(params)])  # positional leaves arr_0, arr_1, ...
    with pytest.raises(ValueError, match="params_named.npz"):
        InferenceRunner.load_params(str(released))
    save_params(str(released / "params_named.npz"), params)  # what INTEGRATION.md's recipe writes: found first
    got = InferenceRunner.load_params(str(released))
    assert sorted(got) == sorted(params) and all(np.array_equal(got[k], params[k]) for k in params)
    with pytest.raises(FileNotFoundError):
        InferenceRunner.load_params(str(tmp_path / "nowhere"))


def test_alpha_carbon_flag_is_checked():
    from pst.config import TokenizerConfig, load_config

    cfg = load_config(overrides=["model=gnn/ablation_4k_df_1.yaml", "data=ablation_df_1.yaml"])
    cfg.data.data["graph_residue_loc_is_alphac"] = False
    with pytest.raises(NotImplementedError):
        TokenizerConfig.from_reference_cfg(cfg)


def test_data_pipeline_mirror_host_side(tmp_path):
    """filter rule (data/preprocessing.py:29-39), config defaults (data_pipeline.py:47-61) and the saved formats."""
    from pst import pdb as ppdb
    from pst import synthetic as syn
    from pst.data_pipeline import BatchDataVQ3D, DataPipeline, ProteinGraph, filter_out_sample

    bb = syn.make_backbones(4, [64])[0]
    pos, gt, ex = syn.backbone_to_atom37(bb)
    s = ppdb.StructureSample(nb_residues=64, aatype=np.full(64, 7, np.int32), atom37_positions=pos, atom37_gt_exists=gt, atom37_atom_exists=ex)
    assert not filter_out_sample(s, 10, 1000) and filter_out_sample(s, 65, 1000) and filter_out_sample(s, 10, 63)
    pipe = DataPipeline()
    assert pipe.config["num_neighbor"] == 30 and pipe.config["downsampling_ratio"] == 4 and pipe.config["output_format"] == "npy"
    assert pipe.validate_sample(s) and pipe.get_sample_info(s)["valid_residues"] == 64
    g = ProteinGraph(n_node=np.array([3]), n_edge=np.array([6]), nodes_mask=np.ones((4, 1), bool), nodes_original_coordinates=np.zeros((4, 3), np.float32),
                     node_features=np.zeros((4, 3), np.float32), edge_features=np.arange(8 * 27, dtype=np.float32).reshape(8, 27),
                     tokens_mask=np.ones((4, 1), bool), senders=np.arange(8), receivers=np.repeat(np.arange(4), 2))
    for fmt in ("npy", "npz"):
        pipe.config["output_format"] = fmt
        out = str(tmp_path / f"b.{fmt}")
        pipe.save_output(BatchDataVQ3D(g, {}), out)
        back = DataPipeline.load_output(out)
        assert all(np.array_equal(getattr(back.graph, k), getattr(g, k)) for k in g._fields)
