"""CUDA encoder + quantiser against the oracle (fp32 CPU restatement of the reference forward)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [(4096, 1), (64000, 1), (64000, 4), (432, 1), (4096, 2)]


def _oracle_cfg(cfg):
    from oracle import model as om

    return om.OracleConfig(seq_max_size=cfg.seq_max_size, graph_max_neighbor=cfg.num_neighbor,
                           downsampling_ratio=cfg.downsampling_ratio, max_out_len=cfg.max_out_len, levels=list(cfg.levels))


def _setup(codebook, df, precision, lengths, seed=11):
    import torch
    from oracle import featurize as fz
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(codebook, df, precision=precision)
    params = init_params(cfg, 1, "rich")
    tok = StructureTokenizer(cfg, params)
    bbs = syn.make_backbones(seed, lengths)
    graphs = []
    for bb in bbs:
        pos, gt, ex = syn.backbone_to_atom37(bb)
        graphs.append(fz.featurize(pos, gt, ex, cfg.num_neighbor))
    return cfg, params, tok, bbs, graphs


@pytest.mark.parametrize("codebook,df", CONFIGS)
def test_fp32_latents_and_tokens_match_oracle(built_lib, codebook, df):
    import torch
    from oracle import model as om

    lengths = [64, 101, 256, 50, 190]
    cfg, params, tok, bbs, graphs = _setup(codebook, df, "fp32", lengths)
    ocfg = _oracle_cfg(cfg)
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    toff = tok.token_offsets(offs)
    feats = np.concatenate([g["edge_features"].astype(np.float32) for g in graphs])
    send = np.concatenate([g["senders"].astype(np.int32) for g in graphs])
    z = tok.encode_graph_device(torch.from_numpy(feats).cuda(), torch.from_numpy(send).cuda(), torch.from_numpy(offs).cuda(),
                                torch.from_numpy(toff).cuda(), len(lengths), int(offs[-1]), int(toff[-1]))
    tokens, bounded = tok.quantize_device(z, want_bounded=True)
    torch.cuda.synchronize()
    z, tokens, bounded = z.cpu().numpy(), tokens.cpu().numpy(), bounded.cpu().numpy()
    C = len(cfg.levels)
    agree = total = 0
    for i, g in enumerate(graphs):
        zr = om.encode(params, ocfg, g["edge_features"], g["senders"], g["n_node"])
        sl = slice(toff[i], toff[i + 1])
        assert zr.shape[0] == toff[i + 1] - toff[i]
        # fp32 on both sides, different summation order: stated tolerance 2e-4 absolute on z (|z| ~ 1)
        assert np.abs(z[sl, :C] - zr).max() < 2e-4, (i, float(np.abs(z[sl, :C] - zr).max()))
        tr = om.fsq_tokens(zr, cfg.levels)
        amb = om.fsq_ambiguous(zr, cfg.levels, tol=5e-4)
        assert np.array_equal(tokens[sl][~amb].astype(np.uint32), tr[~amb]), i
        agree += int((tokens[sl].astype(np.uint32) == tr).sum())
        total += len(tr)
        # quantiser is bit-exact given identical bounded values
        assert np.array_equal(om.fsq_pack(bounded[sl, :C], cfg.levels), tokens[sl].astype(np.uint32))
    assert agree / total > 0.995


def test_fsq_pack_and_inverse_bit_exact(built_lib):
    import torch
    from oracle import model as om
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    rng = np.random.default_rng(0)
    for codebook in (432, 1728, 4096, 64000):
        cfg = TokenizerConfig.named(codebook, 1, precision="fp32")
        tok = StructureTokenizer(cfg, init_params(cfg, 0, "ref"))
        C = len(cfg.levels)
        z = np.zeros((20000, 8), np.float32)
        z[:, :C] = rng.normal(0, 1.5, (20000, C))
        b = np.zeros_like(z)
        b[:, :C] = om.fsq_bound(z[:, :C], cfg.levels)
        b[:64, :C] = np.round(b[:64, :C]) + 0.5  # exact ties: round half to even
        b[:64, :C] = np.clip(b[:64, :C], -(np.array(cfg.levels) // 2) + 0.5, (np.array(cfg.levels) - 1) // 2 - 0.5)
        t_pack = tok.fsq_pack_device(torch.from_numpy(b).cuda()).cpu().numpy()
        assert np.array_equal(t_pack.astype(np.uint32), om.fsq_pack(b[:, :C], cfg.levels))
        t_q = tok.quantize_device(torch.from_numpy(z).cuda()).cpu().numpy().astype(np.uint32)
        ref = om.fsq_tokens(z[:, :C], cfg.levels)
        amb = om.fsq_ambiguous(z[:, :C], cfg.levels, tol=1e-5)
        assert np.array_equal(t_q[~amb], ref[~amb])
        assert t_q.max() < cfg.num_codes
        codes = tok.indexes_to_codes_device(torch.from_numpy(t_pack).cuda()).cpu().numpy()
        assert np.array_equal(codes[:, :C], om.indexes_to_codes(t_pack, cfg.levels).astype(np.float32))
        assert np.array_equal(codes[:, :C], np.rint(b[:, :C]))
        tok.close()


@pytest.mark.parametrize("codebook,df", [(4096, 1), (64000, 4)])
def test_tokenize_end_to_end_matches_oracle(built_lib, codebook, df):
    from oracle import model as om

    lengths = [75, 128, 300, 60]
    cfg, params, tok, bbs, graphs = _setup(codebook, df, "fp32", lengths, seed=23)
    ocfg = _oracle_cfg(cfg)
    out = tok.tokenize(bbs)
    agree = total = 0
    for i, g in enumerate(graphs):
        zr = om.encode(params, ocfg, g["edge_features"], g["senders"], g["n_node"])
        tr = om.fsq_tokens(zr, cfg.levels)
        assert out[i].dtype == np.uint32 and out[i].shape == tr.shape
        agree += int((out[i] == tr).sum())
        total += len(tr)
    assert agree / total > 0.995
    assert tok.launches > 0


@pytest.mark.parametrize("precision,codebook,df,mean_tol,max_tol,min_agree", [
    ("fp16", 4096, 1, 5e-4, 6e-3, 0.995),
    ("fp16", 64000, 4, 5e-4, 6e-3, 0.99),
    ("bf16", 4096, 1, 5e-3, 5e-2, 0.95),
])
def test_tensor_core_modes_match_oracle(built_lib, precision, codebook, df, mean_tol, max_tol, min_agree):
    """tcgen05 edge-level MLPs (16-bit operands, fp32 accumulate) against the fp32 oracle.
    Stated tolerance on the pre-quantisation latents z (|z| ~ 1): fp16 operands mean |dz| <= 5e-4,
    max <= 6e-3; bf16 operands mean <= 5e-3, max <= 5e-2."""
    import torch
    from oracle import model as om

    lengths = [64, 101, 256, 50, 190, 77]
    cfg, params, tok, bbs, graphs = _setup(codebook, df, precision, lengths)
    ocfg = _oracle_cfg(cfg)
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    toff = tok.token_offsets(offs)
    feats = np.concatenate([g["edge_features"].astype(np.float32) for g in graphs])
    send = np.concatenate([g["senders"].astype(np.int32) for g in graphs])
    z = tok.encode_graph_device(torch.from_numpy(feats).cuda(), torch.from_numpy(send).cuda(), torch.from_numpy(offs).cuda(),
                                torch.from_numpy(toff).cuda(), len(lengths), int(offs[-1]), int(toff[-1]))
    tokens = tok.quantize_device(z)
    torch.cuda.synchronize()
    z, tokens = z.cpu().numpy(), tokens.cpu().numpy()
    C = len(cfg.levels)
    errs, agree, total = [], 0, 0
    for i, g in enumerate(graphs):
        zr = om.encode(params, ocfg, g["edge_features"], g["senders"], g["n_node"])
        sl = slice(toff[i], toff[i + 1])
        errs.append(np.abs(z[sl, :C] - zr).ravel())
        tr = om.fsq_tokens(zr, cfg.levels)
        agree += int((tokens[sl].astype(np.uint32) == tr).sum())
        total += len(tr)
    errs = np.concatenate(errs)
    assert np.isfinite(z).all()
    assert errs.mean() <= mean_tol and errs.max() <= max_tol, (float(errs.mean()), float(errs.max()))
    assert agree / total >= min_agree, agree / total


@pytest.mark.parametrize("env,codebook,df", [("PST_MSG_T", 4096, 1), ("PST_FUSED_RESAMPLER", 64000, 4)])
def test_switched_off_kernels_fall_back_to_the_older_gpu_path_with_agreeing_tokens(built_lib, monkeypatch, env, codebook, df):
    """The two environment switches read at model creation (csrc/api.cu): PST_MSG_T=0 runs the message MLPs through
    edge_mlp_tc_kernel<., 0> (the node kernel then writes all four addend tables), PST_FUSED_RESAMPLER=0 runs the
    df > 1 resampler per op.  Both are GPU paths of the same precision mode: latents within the fp16 tolerance of each
    other, tokens agree."""
    import torch

    lengths = [64, 101, 256, 50, 190, 77]
    cfg, params, tok, bbs, graphs = _setup(codebook, df, "fp16", lengths)
    monkeypatch.setenv(env, "0")
    _, _, tok_off, _, _ = _setup(codebook, df, "fp16", lengths)
    monkeypatch.delenv(env)
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    toff = tok.token_offsets(offs)
    feats = torch.from_numpy(np.concatenate([g["edge_features"].astype(np.float32) for g in graphs])).cuda()
    send = torch.from_numpy(np.concatenate([g["senders"].astype(np.int32) for g in graphs])).cuda()
    out = []
    for t in (tok, tok_off):
        z = t.encode_graph_device(feats, send, torch.from_numpy(offs).cuda(), torch.from_numpy(toff).cuda(), len(lengths),
                                  int(offs[-1]), int(toff[-1]))
        n_launch = t.launches
        tokens = t.quantize_device(z)
        torch.cuda.synchronize()
        out.append((z.cpu().numpy(), tokens.cpu().numpy(), n_launch))
    (z0, t0, n0), (z1, t1, n1) = out
    assert n1 > n0 if env == "PST_FUSED_RESAMPLER" else n1 == n0  # the per-op resampler launches ~40 kernels more
    assert np.isfinite(z1).all()
    assert np.abs(z0 - z1).max() <= 6e-3
    assert (t0 == t1).mean() >= 0.99


@pytest.mark.parametrize("codebook,df", [(4096, 1), (64000, 4), (1728, 2)])
def test_fused_fsq_epilogue_is_bit_identical_to_the_quantiser_launch(built_lib, monkeypatch, codebook, df):
    """pst_tokenize in the tensor-core modes: the head of the fused resampler kernel emits the int32 token ids itself
    (csrc/fsq_device.cuh is the one definition of bound / round / pack).  PST_FUSED_FSQ=0 keeps the separate
    fsq_quantize_kernel launch on the same latents: identical ids, one launch more."""
    lengths = [64, 101, 256, 50, 190, 77, 512]
    cfg, params, tok, bbs, graphs = _setup(codebook, df, "fp16", lengths, seed=5)
    monkeypatch.setenv("PST_FUSED_FSQ", "0")
    _, _, tok_off, _, _ = _setup(codebook, df, "fp16", lengths, seed=5)
    monkeypatch.delenv("PST_FUSED_FSQ")
    a = tok.tokenize(bbs)
    n_a = tok.launches
    b = tok_off.tokenize(bbs)
    n_b = tok_off.launches
    assert n_b == n_a + 1
    for x, y, n in zip(a, b, lengths):
        assert x.shape == y.shape == (n // df,)
        assert np.array_equal(x, y)
        assert int(x.max()) < codebook


def test_graph_replay_matches_eager_and_follows_new_data(built_lib):
    """pst_tokenize replays a CUDA graph when its argument set repeats (include/pst_abi.h: pst_graph_cache_enable).
    The replayed call must give the tokens of the eager call, and, because only pointers and sizes are baked into
    the graph, follow new contents written into the same buffers."""
    import torch
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp16")
    tok = StructureTokenizer(cfg, init_params(cfg, 1, "rich"))
    lengths = [80, 130, 64, 200]
    batches = [syn.pack_backbones(syn.make_backbones(seed, lengths)) for seed in (3, 4)]
    offs = batches[0][1]
    toff = tok.token_offsets(offs)
    B, R, T = len(lengths), int(offs[-1]), int(toff[-1])
    o_dev, t_dev = torch.from_numpy(offs).cuda(), torch.from_numpy(toff).cuda()

    # eager results (graph cache off), fresh buffers
    tok.graph_cache_enable(False)
    eager = [tok.tokenize_device(torch.from_numpy(a).cuda(), None, o_dev, t_dev, B, R, T).clone() for a, _ in batches]
    tok.graph_cache_enable(True)

    stage = torch.empty((R, 4, 3), dtype=torch.float32, device="cuda")
    out = torch.empty((T,), dtype=torch.int32, device="cuda")
    for rep in range(3):  # call 1 eager (first sighting), call 2 captures + replays, call 3 replays
        for (a, _), want in zip(batches, eager):
            stage.copy_(torch.from_numpy(a))
            out.zero_()
            tok.tokenize_device(stage, None, o_dev, t_dev, B, R, T, out=out)
            torch.cuda.synchronize()
            assert torch.equal(out, want), rep
    assert tok.read_status() == 0
    # the same on an explicitly named stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for rep in range(3):
            stage.copy_(torch.from_numpy(batches[1][0]), non_blocking=False)
            out.zero_()
            tok.tokenize_device(stage, None, o_dev, t_dev, B, R, T, out=out)
            s.synchronize()
            assert torch.equal(out, eager[1]), rep
    tok.close()


@pytest.mark.parametrize("precision,codebook,df,min_agree", [("bf16", 4096, 1, 0.97), ("fp16", 64000, 4, 0.995), ("fp16", 432, 1, 0.995)])
def test_fused_call_other_modes_agree_with_two_call_path(built_lib, precision, codebook, df, min_agree):
    """The compact feature hand-off of pst_tokenize in the other tensor-core configurations (bf16 operands, df = 4 with
    the per-op resampler, the 5-level codebook) against the two-call path of the same precision mode."""
    import torch
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(codebook, df, precision=precision)
    tok = StructureTokenizer(cfg, init_params(cfg, 1, "rich"))
    lengths = [200, 64, 333, 128]
    atoms, offs = syn.pack_backbones(syn.make_backbones(33, lengths))
    toff = tok.token_offsets(offs)
    B, R, T = len(lengths), int(offs[-1]), int(toff[-1])
    a, o, t = torch.from_numpy(atoms).cuda(), torch.from_numpy(offs).cuda(), torch.from_numpy(toff).cuda()
    fused = tok.tokenize_device(a, None, o, t, B, R, T).clone()
    senders, feats = tok.featurize_device(a, None, o, B, R)
    two_call = tok.quantize_device(tok.encode_graph_device(feats, senders, o, t, B, R, T))
    torch.cuda.synchronize()
    assert tok.read_status() == 0
    assert int(fused.min()) >= 0 and int(fused.max()) < cfg.num_codes
    agree = float((fused == two_call).float().mean())
    assert agree >= min_agree, agree
    tok.close()


def test_fused_call_agrees_with_the_two_call_path(built_lib):
    """pst_tokenize hands the features to the embedding kernel in a compact layout and evaluates the 15 RBFs there in
    fp32; pst_featurize_knn + pst_encode_graph + pst_quantize use the reference's 27 fp32 features (fp64 math).  The
    difference is far below the fp16 rounding of the edge embedding: the token ids must agree (>= 99.9 %)."""
    import torch
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(64000, 1, precision="fp16")
    tok = StructureTokenizer(cfg, init_params(cfg, 1, "rich"))
    lengths = [300, 64, 512, 187, 96, 450]
    atoms, offs = syn.pack_backbones(syn.make_backbones(21, lengths))
    toff = tok.token_offsets(offs)
    B, R, T = len(lengths), int(offs[-1]), int(toff[-1])
    a, o, t = torch.from_numpy(atoms).cuda(), torch.from_numpy(offs).cuda(), torch.from_numpy(toff).cuda()
    fused = tok.tokenize_device(a, None, o, t, B, R, T).clone()
    senders, feats = tok.featurize_device(a, None, o, B, R)
    z = tok.encode_graph_device(feats, senders, o, t, B, R, T)
    two_call = tok.quantize_device(z)
    torch.cuda.synchronize()
    assert tok.read_status() == 0
    agree = float((fused == two_call).float().mean())
    assert agree >= 0.999, agree
    tok.close()


def test_graph_cache_eviction_keeps_results(built_lib):
    """More argument sets than the graph cache holds (8, LRU): entries are evicted and re-captured; every call must
    still give the eager tokens of its own inputs."""
    import torch
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg = TokenizerConfig.named(4096, 1, precision="fp16")
    tok = StructureTokenizer(cfg, init_params(cfg, 1, "rich"))
    cases = []
    for i in range(11):
        lengths = [64 + 8 * i, 90, 70 + i]
        atoms, offs = syn.pack_backbones(syn.make_backbones(100 + i, lengths))
        toff = tok.token_offsets(offs)
        a, o, t = torch.from_numpy(atoms).cuda(), torch.from_numpy(offs).cuda(), torch.from_numpy(toff).cuda()
        out = torch.empty((int(toff[-1]),), dtype=torch.int32, device="cuda")
        cases.append((a, o, t, len(lengths), int(offs[-1]), int(toff[-1]), out))
    tok._workspace(max(c[4] for c in cases), 3)  # one workspace for all: the keys differ by the other arguments
    tok.graph_cache_enable(False)
    want = [tok.tokenize_device(a, None, o, t, B, R, T).clone() for a, o, t, B, R, T, _ in cases]
    tok.graph_cache_enable(True)
    for rep in range(3):
        for (a, o, t, B, R, T, out), w in zip(cases, want):
            out.zero_()
            tok.tokenize_device(a, None, o, t, B, R, T, out=out)
            torch.cuda.synchronize()
            assert torch.equal(out, w), rep
    assert tok.read_status() == 0
    tok.close()


@pytest.mark.parametrize("scale", [8.0, 32.0, 512.0])
def test_scaled_weights_give_agreeing_tokens_or_a_device_status_never_silent_garbage(built_lib, scale):
    """The 16-bit operand modes have a finite range (fp16: 65 504) and the gathered addend tables hold PRE-activation
    values.  With the edge-level weights scaled up, the default mode must either still agree with the all-fp32 mode or
    raise PST_ERR_NON_FINITE through the device status word: never return token ids computed from Inf / NaN."""
    from pst import _lib
    from pst import synthetic as syn
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    cfg16 = TokenizerConfig.named(4096, 1, precision="fp16")
    cfg32 = TokenizerConfig.named(4096, 1, precision="fp32")
    params = dict(init_params(cfg16, 5, "spread"))
    for k in list(params):
        if ("node_mlp_0" in k or "edge_mlp" in k) and k.endswith("/w"):
            params[k] = (params[k] * np.float32(scale)).astype(np.float32)
    bbs = syn.make_backbones(17, [96, 160, 64, 128])
    ref = StructureTokenizer(cfg32, params).tokenize(bbs)
    assert all(np.isfinite(r.astype(np.float64)).all() for r in ref)
    tok = StructureTokenizer(cfg16, params)
    try:
        out = tok.tokenize(bbs)
    except _lib.PstError as e:
        assert e.status == _lib.PST_ERR_NON_FINITE, e
        return
    agree = sum(int((a == b).sum()) for a, b in zip(out, ref)) / sum(len(r) for r in ref)
    # with saturated activations the latents sit on few distinct codes; what must not happen is disagreement from overflow
    assert agree >= 0.97, (scale, agree)
