"""Throughput of the host-level API: StructureTokenizer.tokenize on a long list of structures (chunk pipeline)."""
import sys, time
sys.path.insert(0, "protein-structure-tokenizer_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from pst import synthetic as syn
from pst.config import TokenizerConfig
from pst.tokenizer import StructureTokenizer
from pst.weights import init_params
cfg = TokenizerConfig.named(4096, 1, precision="fp16")
tok = StructureTokenizer(cfg, init_params(cfg, 0, "spread"))
pool = syn.make_backbones(3, [512] * 64, group=64)
structs = [pool[i % 64] for i in range(2048)]  # 1 048 576 residues, 8 chunks of 131 072
tok.tokenize(structs[:512])  # warm-up: staging buffers, workspace
for rep in range(2):
    t0 = time.perf_counter(); out = tok.tokenize(structs); dt = time.perf_counter() - t0
    print(f"tokenize(): {len(structs)} structures, {sum(s.shape[0] for s in structs)} residues in {dt*1e3:.1f} ms = {sum(s.shape[0] for s in structs)/dt/1e6:.2f} M residues/s (host list in, host arrays out)")
