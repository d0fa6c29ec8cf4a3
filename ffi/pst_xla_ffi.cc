// XLA FFI custom-call handlers around the C ABI of libpst_b200.so (include/pst_abi.h).
//
// This is the binding BASELINE.json's north star names ("host code stays Python (JAX/Haiku) and calls hand-written
// sm_100a CUDA kernels through a thin C-ABI via JAX FFI custom calls").  It replaces the body of the reference's
// pmapped callable (scripts/inference_runner.py:179-191: Vq3D(...).encode_and_quantize(graph)) at boundary B1
// (`PstEncodeGraph`: the ProteinGraph leaves the callable reads) or, together with make_graph_from_pdb's
// featurisation (:40-74), at boundary B2 (`PstTokenize`: atoms in, token ids out).
//
// A compile check against a stand-in for the XLA header (tests/ffi_stub/xla/ffi/api/ffi.h: the API surface used here,
// nothing more) runs in tests/test_abi.py; the real build is
// NOT POSSIBLE IN THIS IMAGE: it needs the headers of a JAX >= 0.4.31 (`python -c "import jax.ffi;
// print(jax.ffi.include_dir())"`), and neither jax nor jaxlib is installed here (the reference pins jax==0.4.23,
// which predates jax.ffi).  pst/jax_ffi.py compiles and registers it when `import jax` works:
//   g++ -std=c++17 -shared -fPIC -I$(jax.ffi.include_dir()) -I include -I /usr/local/cuda/include \
//       ffi/pst_xla_ffi.cc -o protein-structure-tokenizer_b200/pst/libpst_xla_ffi.so \
//       -L protein-structure-tokenizer_b200/pst -lpst_b200 -Wl,-rpath,'$ORIGIN'
// The handlers only forward XLA's device buffers and stream: no allocation, no synchronisation, scratch is an
// XLA-allocated result buffer (the C ABI's caller-owned workspace).
#include <cuda_runtime_api.h>

#include <cstdint>

#include "pst_abi.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// The model handle is created once from Python (pst_model_create through ctypes, pst/_lib.py) and handed over as
// an integer attribute: XLA FFI attributes are plain scalars.
static inline pst_model* as_model(int64_t handle) { return reinterpret_cast<pst_model*>(static_cast<intptr_t>(handle)); }

static ffi::Error fail(int rc) { return ffi::Error(ffi::ErrorCode::kInternal, pst_status_string(rc)); }

// B2: atoms f32[R, A, 3] (A = 4 or 37), atom_mask u8[R, A] (gt_exists & atom_exists, data/preprocessing.py:72; an EMPTY
// buffer, u8[0], means "every atom slot present": backbone-only input), offsets i32[B+1], token_offsets i32[B+1]
// -> tokens i32[T].  With atom37 input from real PDB files the mask is REQUIRED: the k-NN anchor is the mean of the
// present heavy atoms (utils/protein_utils.py:373-378) and zero-filled absent slots must not enter it.
static ffi::Error TokenizeImpl(cudaStream_t stream, int64_t model, ffi::Buffer<ffi::F32> atoms, ffi::Buffer<ffi::U8> atom_mask,
                               ffi::Buffer<ffi::S32> offsets, ffi::Buffer<ffi::S32> token_offsets,
                               ffi::ResultBuffer<ffi::S32> tokens, ffi::ResultBuffer<ffi::U8> workspace) {
  const int B = static_cast<int>(offsets.dimensions()[0]) - 1;
  const int R = static_cast<int>(atoms.dimensions()[0]);
  const int A = static_cast<int>(atoms.dimensions()[1]);
  const int T = static_cast<int>(tokens->dimensions()[0]);
  const uint8_t* mask = nullptr;
  if (atom_mask.element_count() != 0) {
    if (atom_mask.element_count() != static_cast<size_t>(R) * static_cast<size_t>(A))
      return ffi::Error(ffi::ErrorCode::kInvalidArgument, "atom_mask must be u8[R, A] or empty");
    mask = atom_mask.typed_data();
  }
  const int rc = pst_tokenize(as_model(model), stream, atoms.typed_data(), mask, A, offsets.typed_data(),
                              token_offsets.typed_data(), B, R, T, tokens->typed_data(), workspace->typed_data(),
                              workspace->size_bytes());
  return rc == PST_OK ? ffi::Error::Success() : fail(rc);
}

// B1: edge_features f32[R*K, 27], senders i32[R*K] (local indices), offsets, token_offsets -> z f32[T, 8], tokens i32[T]
static ffi::Error EncodeGraphImpl(cudaStream_t stream, int64_t model, ffi::Buffer<ffi::F32> edge_features,
                                  ffi::Buffer<ffi::S32> senders, ffi::Buffer<ffi::S32> offsets,
                                  ffi::Buffer<ffi::S32> token_offsets, int64_t total_residues, ffi::ResultBuffer<ffi::F32> z,
                                  ffi::ResultBuffer<ffi::S32> tokens, ffi::ResultBuffer<ffi::U8> workspace) {
  const int B = static_cast<int>(offsets.dimensions()[0]) - 1;
  const int R = static_cast<int>(total_residues);
  const int T = static_cast<int>(tokens->dimensions()[0]);
  int rc = pst_encode_graph(as_model(model), stream, edge_features.typed_data(), senders.typed_data(), offsets.typed_data(),
                            token_offsets.typed_data(), B, R, T, z->typed_data(), workspace->typed_data(),
                            workspace->size_bytes());
  if (rc == PST_OK) rc = pst_quantize(as_model(model), stream, z->typed_data(), T, tokens->typed_data(), nullptr);
  return rc == PST_OK ? ffi::Error::Success() : fail(rc);
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(PstTokenize, TokenizeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(PstEncodeGraph, EncodeGraphImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<int64_t>("total_residues")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
