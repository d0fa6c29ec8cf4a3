"""Builds libpst_b200.so (sm_100a) in-tree with nvcc.  featurize.cu is compiled with
-fmad=false: its fp64 arithmetic must round like the reference's NumPy host code."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("PST_BUILD_OUT") or os.path.join(HERE, "pst", "libpst_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
SOURCES = {
    "featurize.cu": ["-fmad=false"],
    "encoder_fp32.cu": [],
    # -DPST_T_PROFILE: per-role cycle counters of edge_msg_t_kernel, printed by block 0 (tools/build_variant.sh)
    "edge_mlp_tc.cu": (["-DPST_T_PROFILE"] if os.environ.get("PST_T_PROFILE") else []),
    "linear_tc.cu": [],
    "node_chain_tc.cu": [],
    "quantize.cu": [],
    "api.cu": [],
    "pdb_parse.cc": [],  # host-only C++ (PDB ingest)
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def build(verbose: bool = False, force: bool = False) -> str:
    objdir = os.path.join(HERE, "build_prof" if os.environ.get("PST_T_PROFILE") else "build")
    os.makedirs(objdir, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "pst_abi.h"), __file__]
    newest = max(os.path.getmtime(d) for d in deps)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    objs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc(), *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc(), *ARCH, "-shared", "-o", OUT, *objs]  # cudart is linked statically (nvcc default)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
