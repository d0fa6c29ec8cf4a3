"""Where does the end-to-end pass lose time against the resident pass?  (experiment)"""
import sys
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
B, R, T = len(bbs), int(offsets[-1]), int(toff[-1])
dev = torch.device("cuda", 0)
ap, op, tp = torch.from_numpy(atoms).pin_memory(), torch.from_numpy(offsets).pin_memory(), torch.from_numpy(toff).pin_memory()
def mk():
    return {"a": ap.to(dev), "o": op.to(dev), "t": tp.to(dev), "tok": torch.empty((T,), dtype=torch.int32, device=dev),
            "out": torch.empty((T,), dtype=torch.int32).pin_memory()}
slots = [mk(), mk()]
cs = torch.cuda.Stream()
def timeit(fn, n=20):
    for _ in range(6): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    torch.cuda.current_stream().wait_stream(cs)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n
i = [0]
def resident_one():
    s = slots[0]; tok.tokenize_device(s["a"], None, s["o"], s["t"], B, R, T, out=s["tok"])
def resident_two():
    s = slots[i[0] & 1]; i[0] += 1; tok.tokenize_device(s["a"], None, s["o"], s["t"], B, R, T, out=s["tok"])
def same_stream_copies():
    s = slots[i[0] & 1]; i[0] += 1
    s["a"].copy_(ap, non_blocking=True); s["o"].copy_(op, non_blocking=True); s["t"].copy_(tp, non_blocking=True)
    tok.tokenize_device(s["a"], None, s["o"], s["t"], B, R, T, out=s["tok"])
    s["out"].copy_(s["tok"], non_blocking=True)
def h2d_only_copy_stream():
    s = slots[i[0] & 1]; i[0] += 1
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    with torch.cuda.stream(cs):
        s["a"].copy_(ap, non_blocking=True); s["o"].copy_(op, non_blocking=True); s["t"].copy_(tp, non_blocking=True)
        ev.record(cs)
    main.wait_event(ev)
    tok.tokenize_device(s["a"], None, s["o"], s["t"], B, R, T, out=s["tok"])
for name, fn in (("resident, one slot", resident_one), ("resident, two slots alternating", resident_two),
                 ("copies on the same stream", same_stream_copies), ("H2D on a copy stream (no hazards tracked)", h2d_only_copy_stream),
                 ("resident, one slot (again)", resident_one)):
    print(f"{name:45s} {timeit(fn):.3f} ms/step")
