"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/pst_abi.h declares.
No compute call is made here (there is no GPU); host-only entry points are exercised."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "pst_abi.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pst_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pst_abi.h but not exported"


def test_header_is_plain_c_and_a_c_program_links_against_the_library(built_lib, tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (what a cgo / JNI / XLA-FFI shim includes) and a C
    program using only host-side entry points must link and run against the shared library without a GPU."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "pst_abi.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    src = tmp_path / "abi_probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "pst_abi.h"\n'
        "int main(void) {\n"
        "  pst_config cfg; pst_model* m = NULL; float blob = 0.f;\n"
        "  memset(&cfg, 0, sizeof cfg);\n"
        "  if (pst_abi_version() != PST_ABI_VERSION) return 1;\n"
        "  if (pst_model_create(&cfg, &blob, 1, 0, &m) != PST_ERR_UNSUPPORTED_CONFIG || m != NULL) return 2;\n"
        "  if (pst_workspace_bytes(NULL, 1, 1) != 0) return 3;\n"
        "  if (strlen(pst_status_string(PST_ERR_WORKSPACE_TOO_SMALL)) == 0) return 4;\n"
        '  puts("ok");\n  return 0;\n}\n')
    exe = tmp_path / "abi_probe"
    libdir = os.path.dirname(built_lib)
    subprocess.run([gcc, "-std=c99", "-I", os.path.dirname(hdr), str(src), "-o", str(exe), "-L", libdir,
                    "-l:" + os.path.basename(built_lib), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def test_python_binding_covers_the_header(built_lib):
    from pst import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.pst_abi_version() == 1
    assert lib.pst_status_string(-3).decode().startswith("structure length")
    assert "Inf / NaN" in lib.pst_status_string(-12).decode()


def test_blob_size_matches_packer(built_lib):
    from pst import _lib
    from pst.config import TokenizerConfig
    from pst.weights import init_params, pack_weights

    lib = _lib.load()
    for codebook, df, seq in ((4096, 1, 512), (64000, 4, 512), (432, 1, 512), (64000, 1, 1024)):
        cfg = TokenizerConfig.named(codebook, df, seq_max_size=seq)
        c = _lib.PstConfig()
        c.abi_version = 1
        c.seq_max_size, c.max_out_len, c.num_neighbor = cfg.seq_max_size, cfg.max_out_len, cfg.num_neighbor
        c.downsampling_ratio, c.num_levels = cfg.downsampling_ratio, len(cfg.levels)
        for i, l in enumerate(cfg.levels):
            c.levels[i] = l
        c.gnn_layers, c.num_blocks, c.precision, c.max_len = 3, 3, 1, cfg.max_len
        blob = pack_weights(init_params(cfg, 0, "ref"), cfg)
        assert lib.pst_weight_blob_floats(C.byref(c)) == blob.size


def test_bad_config_and_no_device_are_reported_not_crashed(built_lib):
    from pst import _lib

    lib = _lib.load()
    c = _lib.PstConfig()  # abi_version 0 -> unsupported
    assert lib.pst_weight_blob_floats(C.byref(c)) == 0
    h = C.c_void_p()
    blob = np.zeros(4, np.float32)
    assert lib.pst_model_create(C.byref(c), blob.ctypes.data, 4, 0, C.byref(h)) == -2
    assert not h.value


def test_product_fails_loudly_without_cuda(built_lib):
    import torch
    from pst.config import TokenizerConfig
    from pst.tokenizer import StructureTokenizer
    from pst.weights import init_params

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cfg = TokenizerConfig.named(4096, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        StructureTokenizer(cfg, init_params(cfg, 0, "ref"))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "protein-structure-tokenizer_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                with open(os.path.join(d, f)) as fh:
                    src = fh.read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(d, f)


def test_jax_ffi_binding_is_gated_on_jax_and_uses_only_abi_symbols():
    """ffi/pst_xla_ffi.cc cannot be built here (no jax headers): check it calls only entry points the header
    declares, and that pst.jax_ffi fails loudly (ImportError) instead of falling back when jax is absent."""
    with open(os.path.join(ROOT, "include", "pst_abi.h")) as fh:
        declared = set(re.findall(r"\b(pst_[a-z0-9_]+)\s*\(", fh.read()))
    with open(os.path.join(ROOT, "ffi", "pst_xla_ffi.cc")) as fh:
        src = fh.read()
    used = set(re.findall(r"\b(pst_[a-z0-9_]+)\s*\(", src))
    assert used and used <= declared, used - declared
    try:
        import jax  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            import pst.jax_ffi  # noqa: F401


def test_jax_ffi_handlers_compile_against_the_header_stand_in():
    """g++ -fsyntax-only of ffi/pst_xla_ffi.cc against tests/ffi_stub (the slice of xla/ffi/api/ffi.h the handlers
    use): the handler parameter lists match their bindings (static_assert in the stand-in), PstTokenize carries the
    atom mask, and every C-ABI call type-checks against include/pst_abi.h."""
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("CUDA headers not found")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-comment", "-Werror", "-I", os.path.join(ROOT, "tests", "ffi_stub"),
                        "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, os.path.join(ROOT, "ffi", "pst_xla_ffi.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(os.path.join(ROOT, "ffi", "pst_xla_ffi.cc")) as fh:
        src = fh.read()
    tok = src[src.index("static ffi::Error TokenizeImpl"):src.index("// B1:")]
    assert "atom_mask" in tok and "/*atom_mask=*/nullptr" not in tok
