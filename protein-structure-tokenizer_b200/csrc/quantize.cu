// Finite-scalar quantiser.  Reference: structure_tokenizer/model/quantize.py
//   bound        :175-181   tanh(z + shift) * half_l - offset, half_l = (L-1)(1-1e-3)/2,
//                           offset = 0.5 for even L, shift = tan(offset / half_l)   [sic: tan]
//   round        :188       jnp.round = round half to even  (rintf in the default rounding mode)
//   pack         :209 -> :113-120 -> :105-107   sum_d (q_d + L_d//2) * prod_{d'<d} L_d'  -> uint32
//   inverse      :122-139   (renorm = False in every released config)
#include "fsq_device.cuh"
#include "pst_internal.h"

namespace {

using FsqParams = PstFsqParams;

__global__ void fsq_quantize_kernel(const float* __restrict__ z, int n, FsqParams p, int32_t* __restrict__ tokens,
                                    float* __restrict__ bounded, int32_t* __restrict__ status) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float4* zp = reinterpret_cast<const float4*>(z + (size_t)t * PST_C8);
  float4 a = zp[0], b = zp[1];
  float v[PST_C8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  bool finite = true;
  tokens[t] = pst_fsq_token(v, p, finite);
  // No silent Inf / NaN: the 16-bit operand modes have a finite range (fp16: 65 504); a latent that arrives here
  // non-finite raises the call's device status word instead of becoming an arbitrary token id.
  if (!finite && status) atomicMin(status, (int)PST_ERR_NON_FINITE);
  if (bounded) {
    float4* bp = reinterpret_cast<float4*>(bounded + (size_t)t * PST_C8);
    bp[0] = make_float4(v[0], v[1], v[2], v[3]);
    bp[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__global__ void fsq_pack_kernel(const float* __restrict__ bounded, int n, FsqParams p, int32_t* __restrict__ tokens) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int tok = 0;
  for (int c = 0; c < p.C; ++c) tok += ((int)rintf(bounded[(size_t)t * PST_C8 + c]) + p.half_width[c]) * p.basis[c];
  tokens[t] = tok;
}

__global__ void fsq_unpack_kernel(const int32_t* __restrict__ tokens, int n, FsqParams p, float* __restrict__ codes) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned tok = (unsigned)tokens[t];
  for (int c = 0; c < PST_C8; ++c) {
    float v = 0.f;
    if (c < p.C) v = (float)((int)((tok / (unsigned)p.basis[c]) % (unsigned)p.levels[c]) - p.half_width[c]);
    codes[(size_t)t * PST_C8 + c] = v;
  }
}

FsqParams make_params(const pst_model* m) { return pst_fsq_params(m); }

}  // namespace

int pst_launch_quantize(const pst_model* m, cudaStream_t st, const float* z, int n, int32_t* tokens, float* bounded,
                        int32_t* status) {
  if (n <= 0) return 0;
  fsq_quantize_kernel<<<(n + 255) / 256, 256, 0, st>>>(z, n, make_params(m), tokens, bounded, status);
  return 1;
}
int pst_launch_fsq_pack(const pst_model* m, cudaStream_t st, const float* bounded, int n, int32_t* tokens) {
  if (n <= 0) return 0;
  fsq_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(bounded, n, make_params(m), tokens);
  return 1;
}
int pst_launch_indexes_to_codes(const pst_model* m, cudaStream_t st, const int32_t* tokens, int n, float* codes) {
  if (n <= 0) return 0;
  fsq_unpack_kernel<<<(n + 255) / 256, 256, 0, st>>>(tokens, n, make_params(m), codes);
  return 1;
}
