"""Ragged chunk stream through StructureTokenizer.tokenize (host arrays -> pinned staging -> chunk calls -> host tokens):
kernel-by-kernel launches against graph launches with in-place updates (pst_tokenize's buffer-keyed graph cache).
  python tools/stream_probe.py [n_structures]"""
import sys
import time

sys.path.insert(0, "protein-structure-tokenizer_b200"); sys.path.insert(0, ".")
import numpy as np
import torch

from pst import synthetic as syn
from pst.config import TokenizerConfig
from pst.tokenizer import StructureTokenizer
from pst.weights import init_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lens = [int(v) for v in syn.bucketed_lengths(20240520, 256)]
pool = syn.make_backbones(20240520, lens, group=64)
structs = [pool[i % len(pool)] for i in range(n)]
res = sum(s.shape[0] for s in structs)
cfg = TokenizerConfig.named(64000, 1, seq_max_size=2048, precision="fp16")
params = init_params(cfg, 0, "spread")
out = {}
for mode in ("eager", "graph"):
    tok = StructureTokenizer(cfg, params)
    tok.graph_cache_enable(mode == "graph")
    tok.tokenize(structs[: n // 4])  # warm-up: staging buffers, first captures
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        toks = tok.tokenize(structs)
        ts.append(time.perf_counter() - t0)
    out[mode] = toks
    print(f"{mode:6s} {len(tok._chunks([s.shape[0] for s in structs]))} chunks, {res} residues: best {min(ts) * 1e3:.1f} ms = "
          f"{res / min(ts) / 1e6:.2f} M residues/s   {tok.graph_cache_stats()}")
    tok.close()
assert all(np.array_equal(a, b) for a, b in zip(out["eager"], out["graph"]))
print("tokens identical")
