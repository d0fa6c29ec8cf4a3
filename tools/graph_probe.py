"""Does the fused hot call capture into a CUDA graph, and what does replay save?  (experiment)"""
import sys, time
sys.path.insert(0, "protein-structure-tokenizer_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import bench
from pst.config import TokenizerConfig
from pst.tokenizer import StructureTokenizer
from pst.weights import init_params
bbs, atoms, offsets, (codebook, df, seq_max), _ = bench.make_batch("cfg2", 0)
cfg = TokenizerConfig.named(codebook, df, seq_max_size=seq_max, precision="fp16")
tok = StructureTokenizer(cfg, init_params(cfg, 0, "spread"))
toff = tok.token_offsets(offsets)
a = torch.from_numpy(atoms).cuda(); o = torch.from_numpy(offsets).cuda(); t = torch.from_numpy(toff).cuda()
B, R, T = len(bbs), int(offsets[-1]), int(toff[-1])
out = torch.empty((T,), dtype=torch.int32, device="cuda")
def step(): tok.tokenize_device(a, None, o, t, B, R, T, out=out)
for _ in range(3): step()
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timeit(step))
ref = out.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        step()
torch.cuda.synchronize()
out.zero_()
g.replay(); torch.cuda.synchronize()
print("graph tokens equal:", bool((out == ref).all()))
print("graph ms/step", timeit(g.replay))
