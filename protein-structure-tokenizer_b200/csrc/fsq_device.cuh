// Finite-scalar quantiser, device side: ONE definition of bound / round / pack shared by the stand-alone quantiser
// kernel (quantize.cu) and the heads of the fused resampler kernels (node_chain_tc.cu), so that a token emitted by the
// fused epilogue is bit-identical to pst_quantize applied to the same latent.
// Reference: structure_tokenizer/model/quantize.py:175-181 (bound), :188 (round half to even), :209 -> :113-120 (pack).
#pragma once
#include "pst_internal.h"

struct PstFsqParams {
  float half_l[PST_C8], offset[PST_C8], shift[PST_C8];
  int basis[PST_C8], half_width[PST_C8], levels[PST_C8];
  int C;
};

inline PstFsqParams pst_fsq_params(const pst_model* m) {
  PstFsqParams p{};
  p.C = m->cfg.num_levels;
  for (int c = 0; c < PST_C8; ++c) {
    p.half_l[c] = m->half_l[c]; p.offset[c] = m->fsq_offset[c]; p.shift[c] = m->fsq_shift[c];
    p.basis[c] = m->basis[c]; p.half_width[c] = m->half_width[c];
    p.levels[c] = c < p.C ? m->cfg.levels[c] : 1;
  }
  return p;
}

#ifdef __CUDACC__
// v: the 8 latent slots of a token (slots >= C ignored); on return v holds the bounded values (0 in unused slots).
// Returns the packed token id; `finite` is cleared when a used latent is Inf / NaN.
__device__ __forceinline__ int pst_fsq_token(float (&v)[PST_C8], const PstFsqParams& p, bool& finite) {
  int tok = 0;
#pragma unroll
  for (int c = 0; c < PST_C8; ++c) {
    float bd = 0.f;
    if (c < p.C) {
      finite = finite && isfinite(v[c]);
      bd = tanhf(v[c] + p.shift[c]) * p.half_l[c] - p.offset[c];
      tok += ((int)rintf(bd) + p.half_width[c]) * p.basis[c];
    }
    v[c] = bd;
  }
  return tok;
}
#endif
