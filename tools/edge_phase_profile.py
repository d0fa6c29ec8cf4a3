import sys, ctypes as C
sys.path.insert(0, "protein-structure-tokenizer_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import bench
from pst.config import TokenizerConfig
from pst.tokenizer import StructureTokenizer
from pst.weights import init_params
from pst import _lib
bbs, atoms, offsets, (codebook, df, seq_max), _ = bench.make_batch("cfg2", 0)
cfg = TokenizerConfig.named(codebook, df, seq_max_size=seq_max, precision="fp16")
tok = StructureTokenizer(cfg, init_params(cfg, 0, "spread"))
toff = tok.token_offsets(offsets)
a = torch.from_numpy(atoms).cuda(); o = torch.from_numpy(offsets).cuda(); t = torch.from_numpy(toff).cuda()
lib = _lib.load()
out = (C.c_ulonglong * 32)()
for it in range(3):
    tok.tokenize_device(a, None, o, t, len(bbs), int(offsets[-1]), int(toff[-1])); torch.cuda.synchronize()
    lib.pst_debug_edge_profile(out, 1)
v = np.array(list(out), dtype=np.float64).reshape(2, 16)
names = ["top sync", "-", "gather+st+sync", "e wait+issue (t0)", "MMA1 wait", "epi1+sync", "MMA2 wait", "epi2+sync", "MMA3 wait", "reload wait", "LN pass1", "LN pass2+sync", "store issue", "msg epi2+sum+st"]
for mode, launches, label in ((0, 3, "message"), (1, 2, "update")):
    tiles_per_cta = 51200 / 148 / 4
    tot = v[mode].sum()
    print(label, "cycles per tile (group 0):", tot / 148 / launches / tiles_per_cta)
    for i, n in enumerate(names):
        if v[mode][i] > 0: print(f"   {n:16s} {v[mode][i] / 148 / launches / tiles_per_cta:9.0f}  {v[mode][i]/tot*100:5.1f}%")
