"""JAX FFI binding of the C ABI (the reference-side integration BASELINE.json's north star asks for).

Importable only where a JAX >= 0.4.31 with `jax.ffi` is installed: this image has neither jax nor jaxlib, so
nothing here is exercised by the tests; in this repository the caller of the C
ABI is `pst/tokenizer.py` over ctypes with torch-owned buffers.  There is no fallback: without JAX the import
raises.

What it replaces in the reference (scripts/inference_runner.py):
  :179-191  InferenceRunner.prepare_tokenize_fn -> `make_tokenize_fn(tok)` returns a jittable callable
            (atoms, offsets, token_offsets) -> tokens int32[T]   (boundary B2, featurisation included), and
            `make_encode_graph_fn(tok)` a callable on the ProteinGraph leaves (boundary B1).
The handlers (ffi/pst_xla_ffi.cc) only forward XLA's buffers and stream to `pst_tokenize` / `pst_encode_graph`.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import jax  # noqa: F401  (ImportError here is the documented behaviour without JAX)
import jax.numpy as jnp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_SO = os.path.join(HERE, "libpst_xla_ffi.so")


def build(force: bool = False) -> str:
    """g++ ffi/pst_xla_ffi.cc against jax.ffi's headers and libpst_b200.so (in-tree, next to it)."""
    src = os.path.join(ROOT, "ffi", "pst_xla_ffi.cc")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-I", jax.ffi.include_dir(), "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(cuda, "include"), src, "-o", _SO, "-L", HERE, "-lpst_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return _SO


_registered = False


def register() -> None:
    global _registered
    if _registered:
        return
    lib = ctypes.CDLL(build())
    for name in ("PstTokenize", "PstEncodeGraph"):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name)), platform="CUDA")
    _registered = True


def make_tokenize_fn(tok):
    """tok: pst.tokenizer.StructureTokenizer (owns the pst_model handle).  Returns
    f(atoms f32[R, A, 3], offsets i32[B+1], token_offsets i32[B+1], total_tokens: int, atom_mask=None) -> i32[total_tokens].
    `atom_mask` u8[R, A] = gt_exists & atom_exists (data/preprocessing.py:72) is required with atom37 input from real
    files (absent atom slots are zero-filled and must not enter the centroid); None = every slot present."""
    register()
    handle = np.int64(tok._h.value)

    def f(atoms, offsets, token_offsets, total_tokens: int, atom_mask=None):
        R, B = atoms.shape[0], offsets.shape[0] - 1
        ws = int(tok.lib.pst_workspace_bytes(tok._h, int(R), int(B)))
        mask = jnp.zeros((0,), jnp.uint8) if atom_mask is None else jnp.asarray(atom_mask, jnp.uint8)
        out = jax.ffi.ffi_call("PstTokenize", (jax.ShapeDtypeStruct((total_tokens,), jnp.int32),
                                               jax.ShapeDtypeStruct((ws,), jnp.uint8)))(atoms, mask, offsets, token_offsets, model=handle)
        return out[0]

    return f


def make_encode_graph_fn(tok):
    """f(edge_features f32[R*K, 27], senders i32[R*K], offsets, token_offsets, total_tokens) -> (z f32[T, 8], tokens i32[T]):
    the leaves of ProteinGraph the reference's callable reads (types.py:67-75), ragged and concatenated."""
    register()
    handle = np.int64(tok._h.value)
    K = tok.cfg.num_neighbor

    def f(edge_features, senders, offsets, token_offsets, total_tokens: int):
        R, B = senders.shape[0] // K, offsets.shape[0] - 1
        ws = int(tok.lib.pst_workspace_bytes(tok._h, int(R), int(B)))
        z, tokens, _ = jax.ffi.ffi_call("PstEncodeGraph", (jax.ShapeDtypeStruct((total_tokens, 8), jnp.float32),
                                                           jax.ShapeDtypeStruct((total_tokens,), jnp.int32),
                                                           jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
            edge_features, senders, offsets, token_offsets, model=handle, total_residues=np.int64(R))
        return z, tokens

    return f
