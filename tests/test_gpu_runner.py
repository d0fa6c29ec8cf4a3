"""InferenceRunner mirror end to end on the GPU: PDB files in, `<stem>_tokens.npy` out (reference file format)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tokenize_directory_of_pdbs(built_lib, tmp_path):
    from oracle import featurize as fz
    from oracle import model as om
    from pst import synthetic as syn
    from pst.config import load_config
    from pst.inference_runner import InferenceRunner
    from test_pdb import _pdb_from_backbone, _pdb_to_mmcif

    pdb_dir = tmp_path / "pdbs"
    pdb_dir.mkdir()
    bbs = syn.make_backbones(31, [70, 120, 55])
    for i, bb in enumerate(bbs):
        text = _pdb_from_backbone(bb)
        if i == 1:  # an mmCIF file next to the PDB files: parsed by content, token file named by its stem
            (pdb_dir / f"S{i}.cif").write_text(_pdb_to_mmcif(text))
        else:
            (pdb_dir / f"S{i}.pdb").write_text(text)
    cfg = load_config("vq3d_inference", overrides=["model=gnn/ablation_4k_df_1.yaml", "data=ablation_df_1.yaml"])
    runner = InferenceRunner()
    devices, n = runner.prepare_devices("gpu")
    fn = runner.prepare_tokenize_fn(cfg, devices[:1], precision="fp32")
    params = runner.load_params(str(tmp_path / "nope"), devices, cfg=fn.cfg, allow_random_init=True)
    out_dir = str(tmp_path / "tokens")
    files = sorted(str(p) for p in pdb_dir.iterdir())
    runner.tokenize(None, fn, params, files, out_dir, 1, cfg.data.data, batch_size_per_device=2)
    ocfg = om.OracleConfig(levels=list(fn.cfg.levels))
    for i, bb in enumerate(bbs):
        t = np.load(os.path.join(out_dir, f"S{i}_tokens.npy"))
        assert t.dtype == np.uint32 and t.shape == (1, bb.shape[0])
        pos, gt, ex = syn.backbone_to_atom37(bb)
        g = fz.featurize(pos, gt, ex, 50)
        ref = om.fsq_tokens(om.encode(params, ocfg, g["edge_features"], g["senders"], g["n_node"]), fn.cfg.levels)
        assert (t[0] == ref).mean() > 0.99
    with pytest.raises(FileExistsError):
        runner.tokenize(None, fn, params, files, out_dir, 1, cfg.data.data)
    with pytest.raises(FileNotFoundError):
        runner.load_params(str(tmp_path / "nope"), devices)


def test_casp14_tokens_fp16_agree_with_oracle(built_lib, casp14):
    """BASELINE config 1 inputs (bundled CASP14 structures, 4k codebook, df=1) through the default fp16 mode."""
    from conftest import valid_atoms
    from oracle import featurize as fz
    from oracle import model as om
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp16")
    params = init_params(cfg, 0, "spread")
    tok = StructureTokenizer(cfg, params)
    names = sorted(casp14)[:12]
    structs, masks = zip(*[valid_atoms(casp14[n]) for n in names])
    out = tok.tokenize(structs, masks)
    ocfg = om.OracleConfig()
    agree = total = 0
    for n, t in zip(names, out):
        e = casp14[n]
        g = fz.featurize(e["pos"], e["gt"], e["exists"], 50)
        ref = om.fsq_tokens(om.encode(params, ocfg, g["edge_features"], g["senders"], g["n_node"]), cfg.levels)
        agree += int((t == ref).sum())
        total += len(ref)
    assert agree / total >= 0.995, agree / total


def test_two_devices_concurrently_match_one_device(built_lib):
    """The runner's callable shards a batch over the visible devices (pmap's [Dev, B] reshape,
    scripts/inference_runner.py:299-306) and drives them from one host thread each."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.inference_runner import InferenceRunner
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp16")
    params = init_params(cfg, 0, "spread")
    bbs = syn.make_backbones(5, [90, 64, 130, 77, 101])
    batch = [(bb, None) for bb in bbs]
    one = InferenceRunner.prepare_tokenize_fn(cfg, [0])(params, None, batch)["tokens"]
    two = InferenceRunner.prepare_tokenize_fn(cfg, [0, 1])(params, None, batch)["tokens"]
    assert len(one) == len(two) == len(bbs)
    for a, b in zip(one, two):
        assert np.array_equal(a, b)


def test_chunk_pipeline_matches_single_call(built_lib):
    """tokenize() splits long inputs into chunks of max_rows_per_call residues and pipelines them over two staging
    slots; the tokens must not depend on the chunking (structures are independent), with and without atom masks."""
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp16")
    params = init_params(cfg, 0, "spread")
    lengths = [64, 200, 90, 150, 77, 300, 64, 128, 51, 260, 99]
    bbs = syn.make_backbones(9, lengths)
    whole = StructureTokenizer(cfg, params).tokenize(bbs)
    small = StructureTokenizer(cfg, params, max_rows_per_call=400)
    assert len(small._chunks(lengths)) >= 4
    for rep in range(2):  # second pass reuses the staging slots
        parts = small.tokenize(bbs)
        assert len(parts) == len(whole)
        for a, b, L in zip(whole, parts, lengths):
            assert a.dtype == np.uint32 and a.shape == (L,) and np.array_equal(a, b)
    # atom37 input with masks through the same pipeline
    a37 = [syn.backbone_to_atom37(bb) for bb in bbs]
    structs = [np.ascontiguousarray(p, np.float32) for p, _, _ in a37]
    masks = [(g & e).astype(np.uint8) for _, g, e in a37]
    parts = small.tokenize(structs, masks)
    for a, b in zip(whole, parts):
        assert np.array_equal(a, b)
    assert small.tokenize([]) == []


def test_flat_token_output_equals_the_list_form(built_lib):
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(64000, 4, precision="fp16")
    tok = StructureTokenizer(cfg, init_params(cfg, 0, "spread"), max_rows_per_call=300)
    lengths = [64, 201, 90, 150, 77]
    bbs = syn.make_backbones(12, lengths)
    parts = tok.tokenize(bbs)
    flat, counts = tok.tokenize(bbs, flat=True)
    assert counts.tolist() == [L // 4 for L in lengths] and flat.dtype == np.int32
    assert np.array_equal(flat.view(np.uint32), np.concatenate(parts))


@pytest.mark.parametrize("codebook,df", [(4096, 1), (64000, 4), (4096, 2)])
def test_reference_padded_graph_batch_through_the_callable(built_lib, casp14, codebook, df):
    """The reference's own batch object — ProteinGraph leaves padded to seq_max_size and stacked to [Dev, B, ...]
    (data/preprocessing.py:191-283, scripts/inference_runner.py:77-83) — goes through the runner's callable (boundary
    B1) and comes back as uint32 [Dev, B, T] INCLUDING the padded tail (code 2 730 / 32 036), compared with the tokens
    the reference's model source produced for the same padded graphs (tests/golden/model_ref_*.npz, all T rows)."""
    import types

    from oracle import featurize as fz
    from pst.config import TokenizerConfig
    from pst.inference_runner import InferenceRunner
    from test_golden_model import load_case
    from test_host import _pad_like_the_reference

    f, cfg, params = load_case(codebook, df)
    cfg = TokenizerConfig(seq_max_size=cfg.seq_max_size, max_out_len=cfg.max_out_len, downsampling_ratio=df,
                          levels=list(cfg.levels), precision="fp32")
    names = [str(n) for n in f["names"]]
    padded = []
    for n in names:
        e = casp14[n]
        g = fz.featurize(e["pos"].astype(np.float64), e["gt"], e["exists"], cfg.num_neighbor)
        padded.append(_pad_like_the_reference(g, cfg.num_neighbor, cfg.seq_max_size, df))
    batch = types.SimpleNamespace(**{k: np.stack([p[k] for p in padded])[None] for k in padded[0]})  # [Dev=1, B, ...]
    fn = InferenceRunner.prepare_tokenize_fn(cfg, [0])
    out = fn(params, None, types.SimpleNamespace(graph=batch, features={}))["tokens"]
    T = cfg.seq_max_size // df
    assert out.shape == (1, len(names), T) and out.dtype == np.uint32
    for b, n in enumerate(names):
        ref_t, ref_b = f[f"{n}/tokens"], f[f"{n}/bounded"]
        nt = int(f[f"{n}/n_valid"]) // df
        amb = (np.abs(ref_b - np.floor(ref_b) - 0.5) < 5e-4).any(-1)
        assert np.array_equal(out[0, b, :nt][~amb], ref_t[:nt][~amb]), n
        assert np.array_equal(out[0, b, nt:], ref_t[nt:]), n  # the masked-token code of the padded tail
        # what the reference's save loop keeps (scripts/inference_runner.py:307-317)
        kept = out[0, b, : int(batch.tokens_mask[0, b].sum())]
        assert kept.shape == (nt,)


def test_runner_under_torchrun_shards_files_and_gathers_on_rank0(built_lib, tmp_path):
    """scripts/tokenize_pdb.py under torchrun with 2 ranks: files sharded by LPT, tokens gathered over NCCL, rank 0
    writes every file; the files equal a single-process run."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from conftest import ROOT
    from pst import synthetic as syn
    from test_pdb import _pdb_from_backbone

    pdb_dir = tmp_path / "pdbs"
    pdb_dir.mkdir()
    bbs = syn.make_backbones(41, [70, 120, 55, 200, 64])
    for i, bb in enumerate(bbs):
        (pdb_dir / f"S{i}.pdb").write_text(_pdb_from_backbone(bb))
    cli = os.path.join(ROOT, "scripts", "tokenize_pdb.py")
    common = ["--pdb_dir", str(pdb_dir), "--random_init", "--batch_size_per_device", "2"]
    one = subprocess.run([sys.executable, cli, "--token_save_path", str(tmp_path / "one"), *common], capture_output=True, text=True)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", cli, "--token_save_path", str(tmp_path / "two"), *common], capture_output=True, text=True)
    assert two.returncode == 0, two.stderr[-2000:]
    names = sorted(os.listdir(tmp_path / "one"))
    assert names == sorted(os.listdir(tmp_path / "two")) == [f"S{i}_tokens.npy" for i in range(5)]
    for n in names:
        assert np.array_equal(np.load(tmp_path / "one" / n), np.load(tmp_path / "two" / n))


def test_data_pipeline_mirror_feeds_the_callable(built_lib, casp14, tmp_path):
    """The reference's second caller of the featurisation path (data_pipeline.py:36-332): sample -> padded BatchDataVQ3D
    (GPU featuriser + the reference's padding) -> saved .npy / .npz -> the runner's callable (boundary B1).  The graph
    equals the oracle's padded graph (senders exactly, features to 1 fp32 ulp) and the tokens equal the B2 path's."""
    from conftest import valid_atoms
    from oracle import featurize as fz
    from pst import pdb as ppdb
    from pst.config import TokenizerConfig
    from pst.data_pipeline import DataPipeline
    from pst.inference_runner import InferenceRunner
    from pst.weights import init_params
    from test_host import _pad_like_the_reference

    name = "T1046s1"
    e = casp14[name]
    n_all = e["pos"].shape[0]
    sample = ppdb.StructureSample(nb_residues=n_all, aatype=np.zeros(n_all, np.int32), atom37_positions=e["pos"].astype(np.float32),
                                  atom37_gt_exists=e["gt"], atom37_atom_exists=e["exists"])
    pipe = DataPipeline({"num_neighbor": 50, "downsampling_ratio": 2, "padding_num_residue": 512, "crop_index": 512})
    assert pipe.validate_sample(sample)
    data = pipe.preprocess(sample)
    g = fz.featurize(e["pos"].astype(np.float64), e["gt"], e["exists"], 50)
    ref = _pad_like_the_reference(g, 50, 512, 2)
    assert np.array_equal(data.graph.senders, ref["senders"]) and np.array_equal(data.graph.receivers, ref["receivers"])
    assert np.array_equal(data.graph.n_node, ref["n_node"]) and np.array_equal(data.graph.tokens_mask, ref["tokens_mask"])
    assert np.array_equal(data.graph.nodes_mask, ref["nodes_mask"])
    ulp = np.spacing(np.abs(ref["edge_features"]).astype(np.float32))
    assert (np.abs(data.graph.edge_features - ref["edge_features"]) <= ulp).all()
    # both output formats round-trip
    for fmt in ("npy", "npz"):
        pipe.config["output_format"] = fmt
        out = str(tmp_path / f"g.{fmt}")
        pipe.save_output(data, out)
        back = DataPipeline.load_output(out)
        assert np.array_equal(back.graph.edge_features, data.graph.edge_features) and np.array_equal(back.graph.senders, data.graph.senders)
    # the saved batch through the callable == the atoms through the fused call
    cfg = TokenizerConfig.named(4096, 2, precision="fp32")
    params = init_params(cfg, 0, "spread")
    fn = InferenceRunner.prepare_tokenize_fn(cfg, [0])
    t_graph = fn(params, None, back)["tokens"]
    atoms, mask = valid_atoms(e)
    t_atoms = fn(params, None, [(atoms, mask)])["tokens"][0]
    nt = int(back.graph.tokens_mask.sum())
    assert t_graph.shape == (256,) and nt == t_atoms.shape[0]
    assert (t_graph[:nt] == t_atoms).mean() >= 0.99  # fp32 features (two-call path) vs compact features (fused call)
    with pytest.raises(NotImplementedError):
        DataPipeline({"num_neighbor": 50, "noise_level": 0.1}).preprocess(sample)


def test_ragged_chunk_stream_runs_on_graph_updates(built_lib):
    """A stream of chunks with distinct (B, R, T) through the two staging slots of tokenize(): after the first two
    sightings of a slot's buffers every call is ONE graph launch, retargeted in place to the chunk's sizes
    (cudaGraphExecUpdate), and the tokens equal the kernel-by-kernel path's."""
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp16")
    params = init_params(cfg, 0, "spread")
    lengths = [64, 200, 90, 150, 77, 300, 64, 128, 51, 260, 99, 180, 70, 310, 55, 140]
    bbs = syn.make_backbones(19, lengths)
    eager = StructureTokenizer(cfg, params, max_rows_per_call=420)
    eager.graph_cache_enable(False)
    want = eager.tokenize(bbs)
    tok = StructureTokenizer(cfg, params, max_rows_per_call=420)
    n_chunks = len(tok._chunks(lengths))
    assert n_chunks >= 6
    got = tok.tokenize(bbs)
    st = tok.graph_cache_stats()
    assert all(np.array_equal(a, b) for a, b in zip(want, got))
    # two slots: each slot's first call is eager, its second instantiates, every later call updates (sizes differ) or replays
    assert st["eager"] == 2 and st["instantiate"] == 2 and st["update"] + st["replay"] == n_chunks - 4, st
    assert st["update"] >= 1
    got2 = tok.tokenize(bbs)  # second pass over the same stream
    st2 = tok.graph_cache_stats()
    assert all(np.array_equal(a, b) for a, b in zip(want, got2))
    assert st2["eager"] == 2 and st2["instantiate"] == 2, st2
