// CUDA-core fp32 encoder: every stage of Vq3D.encode as plain, deterministic kernels.
// This is (a) the on-device reference precision mode (PST_PREC_FP32) and (b) the node-level
// and resampler stages of the tensor-core modes (only the edge-level MLPs move to tcgen05,
// see edge_mlp_tc.cu).  Reference behaviour restated (paths relative to the reference repo,
// under structure_tokenizer/):
//   input embeddings   model/structure_encoder.py:77-105 (PE tables folded at weight-pack time)
//   MPNN layer         model/gnn_layers.py:325-438   (concat order sender|receiver|edge, GELU tanh)
//   MaskedLayerNorm    model/gnn_layers.py:108-120,162-164 (biased variance, eps 1e-5)
//   resampler          model/modules.py:281-382 (Attention), 393-424, 211-262 (Transition), 545-636
//   local mask         model/model.py:264-318,382-401 (token t <-> residues [t*df, t*df+df))
//   head               model/model.py:169-174 (x/(||x||+1e-6)), 148-164 (down_proj)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "pst_internal.h"

namespace {

constexpr int D = PST_D;

__device__ __forceinline__ float gelu_tanh(float x) {
  // jax.nn.gelu(approximate=True)
  const float c = 0.7978845608028654f;
  float x3 = x * x * x;
  return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * x3)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int find_segment(const int32_t* __restrict__ offs, int n, int row) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (offs[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// row_base[r] = first row of the structure that owns residue row r
__global__ void row_base_kernel(const int32_t* __restrict__ offsets, int B, int R, int32_t* __restrict__ row_base) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  row_base[r] = offsets[find_segment(offsets, B, r)];
}

// ---------------------------------------------------------------------------------------------
// SGEMM  C[M,N] = epi(A[M,K] . W[K,N]);  128x128x8 tiles, 256 threads, 8x8 micro-tile.
// K % 8 == 0, N % 128 == 0.
struct GemmEpi {
  const float* bias;       // [N] or null
  const float* residual;   // [M,N] or null: C = residual + value
  const float* gather_s;   // [R,N] or null: += gather_s[row_base[row/K] + senders[row]]
  const float* gather_r;   // [R,N]         : += gather_r[row/K]
  const int32_t* senders;  // [M]
  const int32_t* row_base; // [R]
  int knn;                 // K
  float scale;             // value *= scale after bias (1.0 = off)
  int act;                 // 0 none, 1 gelu, 2 relu
};

__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int M, int N,
             int K, GemmEpi ep) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int a_row = tid >> 1, a_col = (tid & 1) * 4;
  const int b_row = tid >> 5, b_col = (tid & 31) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_ok = (m0 + a_row) < M;
  const float* a_ptr = A + (size_t)(m0 + a_row) * K + a_col;
  const float* b_ptr = W + (size_t)b_row * N + n0 + b_col;
  float4 a_reg = a_ok ? *reinterpret_cast<const float4*>(a_ptr) : make_float4(0, 0, 0, 0);
  float4 b_reg = *reinterpret_cast<const float4*>(b_ptr);
  As[0][a_col + 0][a_row] = a_reg.x; As[0][a_col + 1][a_row] = a_reg.y;
  As[0][a_col + 2][a_row] = a_reg.z; As[0][a_col + 3][a_row] = a_reg.w;
  *reinterpret_cast<float4*>(&Bs[0][b_row][b_col]) = b_reg;
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      a_reg = a_ok ? *reinterpret_cast<const float4*>(a_ptr + (size_t)(kt + 1) * BK) : make_float4(0, 0, 0, 0);
      b_reg = *reinterpret_cast<const float4*>(b_ptr + (size_t)(kt + 1) * BK * N);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nxt = cur ^ 1;
      As[nxt][a_col + 0][a_row] = a_reg.x; As[nxt][a_col + 1][a_row] = a_reg.y;
      As[nxt][a_col + 2][a_row] = a_reg.z; As[nxt][a_col + 3][a_row] = a_reg.w;
      *reinterpret_cast<float4*>(&Bs[nxt][b_row][b_col]) = b_reg;
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
    const float* gs = nullptr;
    const float* gr = nullptr;
    if (ep.gather_s) {
      int r = row / ep.knn;
      gs = ep.gather_s + (size_t)(ep.row_base[r] + ep.senders[row]) * N;
      gr = ep.gather_r + (size_t)r * N;
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int col = n0 + (jj == 0 ? tx * 4 : 64 + tx * 4);
      float4 v = make_float4(acc[i][jj * 4 + 0], acc[i][jj * 4 + 1], acc[i][jj * 4 + 2], acc[i][jj * 4 + 3]);
      if (ep.bias) {
        float4 b = *reinterpret_cast<const float4*>(ep.bias + col);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (gs) {
        float4 s = *reinterpret_cast<const float4*>(gs + col);
        float4 r = *reinterpret_cast<const float4*>(gr + col);
        v.x += s.x + r.x; v.y += s.y + r.y; v.z += s.z + r.z; v.w += s.w + r.w;
      }
      if (ep.scale != 1.0f) { v.x *= ep.scale; v.y *= ep.scale; v.z *= ep.scale; v.w *= ep.scale; }
      if (ep.act == 1) { v.x = gelu_tanh(v.x); v.y = gelu_tanh(v.y); v.z = gelu_tanh(v.z); v.w = gelu_tanh(v.w); }
      else if (ep.act == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (ep.residual) {
        float4 r = *reinterpret_cast<const float4*>(ep.residual + (size_t)row * N + col);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      *reinterpret_cast<float4*>(C + (size_t)row * N + col) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-wise kernels: one warp per 128-channel row, lane owns channels [4*lane, 4*lane+4).
__device__ __forceinline__ float4 ln_row(float4 x, const float* __restrict__ scale, const float* __restrict__ offset, int lane) {
  float mean = warp_sum(x.x + x.y + x.z + x.w) * (1.0f / D);
  float dx = x.x - mean, dy = x.y - mean, dz = x.z - mean, dw = x.w - mean;
  float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
  float inv = rsqrtf(var + 1e-5f);
  float4 s = *reinterpret_cast<const float4*>(scale + lane * 4);
  float4 o = *reinterpret_cast<const float4*>(offset + lane * 4);
  return make_float4(s.x * inv * dx + o.x, s.y * inv * dy + o.y, s.z * inv * dz + o.z, s.w * inv * dw + o.w);
}

// out = LN(x + y)   (y may be null)
__global__ void add_ln_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ scale,
                              const float* __restrict__ offset, float* __restrict__ out, int rows) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * D + lane * 4);
  if (y) {
    float4 w = *reinterpret_cast<const float4*>(y + (size_t)row * D + lane * 4);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  *reinterpret_cast<float4*>(out + (size_t)row * D + lane * 4) = ln_row(v, scale, offset, lane);
}

// agg[r] = (sum_k m[r*K+k]) / K      (gnn_layers.py:364-377)
__global__ void segment_mean_kernel(const float* __restrict__ m, int K, float* __restrict__ agg, int rows) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4 s = make_float4(0, 0, 0, 0);
  const float* p = m + (size_t)row * K * D + lane * 4;
  for (int k = 0; k < K; ++k) {
    float4 v = *reinterpret_cast<const float4*>(p + (size_t)k * D);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float kf = (float)K;
  *reinterpret_cast<float4*>(agg + (size_t)row * D + lane * 4) = make_float4(s.x / kf, s.y / kf, s.z / kf, s.w / kf);
}

// h0[r] = node_table[r - row_base[r]]                     (structure_encoder.py:89-92)
__global__ void node_embed_kernel(const float* __restrict__ table, const int32_t* __restrict__ row_base,
                                  float* __restrict__ h, int rows) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  int local = row - row_base[row];
  *reinterpret_cast<float4*>(h + (size_t)row * D + lane * 4) =
      *reinterpret_cast<const float4*>(table + (size_t)local * D + lane * 4);
}

// Tensor-core modes: the same lookup plus the two layer-1 addend tables.  h0 depends only on the position inside the
// structure, so (h0 . W1[0:128]) and (h0 . W1[128:256] + b1) of the first message MLP are position-indexed constant
// tables as well (computed once at model creation by the same tensor-core linear, pst_prepare_layer0_tables).
__global__ void node_embed_tables_kernel(const float* __restrict__ table, const __half* __restrict__ ps0,
                                         const __half* __restrict__ pr0, const int32_t* __restrict__ row_base,
                                         float* __restrict__ h, __half* __restrict__ ps, __half* __restrict__ pr,
                                         __half* __restrict__ h16, int rows) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t local = (size_t)(row - row_base[row]);
  const float4 v = *reinterpret_cast<const float4*>(table + local * D + lane * 4);
  *reinterpret_cast<float4*>(h + (size_t)row * D + lane * 4) = v;
  if (h16) {  // the transposed message kernel gathers fp16(h) for the sender term: no sender table, no separate conversion pass
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(h16 + (size_t)row * D + lane * 4) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  } else {
    *reinterpret_cast<uint2*>(ps + (size_t)row * D + lane * 4) = *reinterpret_cast<const uint2*>(ps0 + local * D + lane * 4);
  }
  *reinterpret_cast<uint2*>(pr + (size_t)row * D + lane * 4) = *reinterpret_cast<const uint2*>(pr0 + local * D + lane * 4);
}

// res0[t] = token_table[t - token_offsets[b]]             (modules.py:486-500)
__global__ void token_embed_kernel(const float* __restrict__ table, const int32_t* __restrict__ token_offsets, int B,
                                   float* __restrict__ res, int T) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  int b = find_segment(token_offsets, B, row);
  int local = row - token_offsets[b];
  *reinterpret_cast<float4*>(res + (size_t)row * D + lane * 4) =
      *reinterpret_cast<const float4*>(table + (size_t)local * D + lane * 4);
}

// e0[e] = edge_pe_table[s - r + (n-1)] + f27[e] . Wf        (structure_encoder.py:94-105)
// One block per receiver row (K edges); thread = output channel, its 27 weights in registers, the
// edge's features broadcast from shared memory as float4.  OutT = float (fp32 mode) or a 16-bit
// type (tensor-core modes keep the edge state in the operand precision of the edge MLPs).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
template <typename OutT>
__device__ __forceinline__ void store_pair(OutT* p, float2 v);
template <> __device__ __forceinline__ void store_pair<float>(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

// 128 threads = 2 receiver rows x 64 channel pairs; each thread keeps the 27 x 2 weights of its channel
// pair in registers and does one packed FFMA2 per feature.
template <typename OutT>
__global__ void __launch_bounds__(128)
edge_embed_kernel(const float* __restrict__ feat, const int32_t* __restrict__ senders,
                  const int32_t* __restrict__ row_base, const float* __restrict__ pe_table,
                  const float* __restrict__ wf, int K, int seq_max, int R, OutT* __restrict__ e) {
  extern __shared__ __align__(16) float s_feat[];  // [2][K][28] then senders [2][K]
  int* s_send = reinterpret_cast<int*>(s_feat + 2 * K * 28);
  const int tid = threadIdx.x;
  const int half = tid >> 6, cp = tid & 63;  // receiver within the block, channel pair
  const int r0 = blockIdx.x * 2;
  for (int t = tid; t < 2 * K * 28; t += 128) {
    const int rr = t / (K * 28), u = t - rr * (K * 28);
    const int k = u / 28, f = u - k * 28;
    const int r = r0 + rr;
    s_feat[t] = (f < PST_EDGE_FEATURES && r < R) ? feat[((size_t)r * K + k) * PST_EDGE_FEATURES + f] : 0.f;
  }
  for (int t = tid; t < 2 * K; t += 128) {
    const int rr = t / K, k = t - rr * K;
    s_send[t] = (r0 + rr < R) ? senders[(size_t)(r0 + rr) * K + k] : 0;
  }
  float2 w[28];
#pragma unroll
  for (int f = 0; f < 28; ++f)
    w[f] = f < PST_EDGE_FEATURES ? *reinterpret_cast<const float2*>(wf + f * D + cp * 2) : make_float2(0.f, 0.f);
  __syncthreads();
  const int r = r0 + half;
  if (r >= R) return;
  const int local_r = r - row_base[r];
  const float* sf = s_feat + half * K * 28;
  const int* ss = s_send + half * K;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const int diff = ss[k] - local_r + (seq_max - 1);
    float2 acc = *reinterpret_cast<const float2*>(pe_table + (size_t)diff * D + cp * 2);
    const float4* fk = reinterpret_cast<const float4*>(sf + k * 28);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float4 x = fk[j];
      acc = ffma2(make_float2(x.x, x.x), w[j * 4 + 0], acc);
      acc = ffma2(make_float2(x.y, x.y), w[j * 4 + 1], acc);
      acc = ffma2(make_float2(x.z, x.z), w[j * 4 + 2], acc);
      acc = ffma2(make_float2(x.w, x.w), w[j * 4 + 3], acc);
    }
    store_pair<OutT>(e + ((size_t)r * K + k) * D + cp * 2, acc);
  }
}

// Local cross attention (modules.py:334-371 with the local mask of model.py:264-318):
// token t of structure b attends residues offsets[b] + t_local*df + [0, df).
// One block (4 warps = 4 heads) per token.
__global__ void __launch_bounds__(128)
local_attention_kernel(const float* __restrict__ q, const float* __restrict__ kx, const float* __restrict__ vx,
                       const float* __restrict__ gate, const int32_t* __restrict__ offsets,
                       const int32_t* __restrict__ token_offsets, int B, int df, float* __restrict__ wa, int T) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const int c = threadIdx.x;  // head = c / 32
  const int b = find_segment(token_offsets, B, t);
  const int row0 = offsets[b] + (t - token_offsets[b]) * df;
  const float qv = q[(size_t)t * D + c];
  float logit[8];
  float mx = -INFINITY;
  for (int i = 0; i < df; ++i) {
    logit[i] = warp_sum(qv * kx[(size_t)(row0 + i) * D + c]);
    mx = fmaxf(mx, logit[i]);
  }
  float den = 0.f;
  for (int i = 0; i < df; ++i) {
    logit[i] = expf(logit[i] - mx);
    den += logit[i];
  }
  float acc = 0.f;
  for (int i = 0; i < df; ++i) acc += (logit[i] / den) * vx[(size_t)(row0 + i) * D + c];
  float g = gate[(size_t)t * D + c];
  g = 1.0f / (1.0f + expf(-g));
  wa[(size_t)t * D + c] = acc * g;
}

// df == 1 shortcut of the local attention: wa = v * sigmoid(gate) (rows of v and tokens coincide)
__global__ void gated_value_kernel(const float* __restrict__ v, const float* __restrict__ gate, float* __restrict__ wa, int n4) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 x = reinterpret_cast<const float4*>(v)[i];
  const float4 g = reinterpret_cast<const float4*>(gate)[i];
  x.x *= 1.0f / (1.0f + expf(-g.x)); x.y *= 1.0f / (1.0f + expf(-g.y));
  x.z *= 1.0f / (1.0f + expf(-g.z)); x.w *= 1.0f / (1.0f + expf(-g.w));
  reinterpret_cast<float4*>(wa)[i] = x;
}

// z[t] = (r / (||r|| + 1e-6)) . Wd + bd      one warp per token; z is [T, 8], unused columns 0
__global__ void head_kernel(const float* __restrict__ res, const float* __restrict__ wd, const float* __restrict__ bd,
                            int C, float* __restrict__ z, int T) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= T) return;
  float4 v = *reinterpret_cast<const float4*>(res + (size_t)row * D + lane * 4);
  float nrm = sqrtf(warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w));
  float d = nrm + 1e-6f;
  v.x /= d; v.y /= d; v.z /= d; v.w /= d;
  float outv = 0.f;
  for (int c = 0; c < PST_C8; ++c) {
    float p = 0.f;
    if (c < C) {
      const float* w = wd + (size_t)(lane * 4) * PST_C8 + c;
      p = v.x * w[0] + v.y * w[PST_C8] + v.z * w[2 * PST_C8] + v.w * w[3 * PST_C8];
    }
    p = warp_sum(p);
    if (lane == c) outv = (c < C) ? p + bd[c] : 0.f;
  }
  if (lane < PST_C8) z[(size_t)row * PST_C8 + lane] = outv;
}

struct Launcher {
  cudaStream_t st;
  const pst_model* model = nullptr;  // set in the tensor-core modes: plain linears go to linear_tc.cu
  int count = 0;
  void gemm(const float* A, const float* W, float* C, int M, int N, int K, GemmEpi ep, int out_half = 0) {
    if (M <= 0) return;
    if (model && !ep.gather_s) {
      int n = pst_launch_linear_tc(model, st, A, W, C, M, N, K, ep.bias, ep.residual, ep.scale, ep.act, out_half);
      if (n > 0) { count += n; return; }
    }
    dim3 grid((M + 127) / 128, N / 128);
    sgemm_kernel<<<grid, 256, 0, st>>>(A, W, C, M, N, K, ep);
    ++count;
  }
  static GemmEpi epi(const float* bias, int act = 0, const float* residual = nullptr, float scale = 1.0f) {
    GemmEpi e{};
    e.bias = bias; e.act = act; e.residual = residual; e.scale = scale;
    return e;
  }
  void add_ln(const float* x, const float* y, const float* s, const float* o, float* out, int rows) {
    if (rows <= 0) return;
    add_ln_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, y, s, o, out, rows);
    ++count;
  }
};

}  // namespace

// Model creation (tensor-core modes, after pst_prepare_linear_tc): the layer-1 addend tables by position,
// [2][seq_max_size][128] fp16, from the node PE table with the tensor-core linear the per-batch path used to run.
int pst_prepare_layer0_tables(pst_model* m) {
  const int n = m->cfg.seq_max_size;
  if (cudaMalloc(&m->layer0_tables, (size_t)2 * n * D * sizeof(uint16_t)) != cudaSuccess) return PST_ERR_CUDA;
  Launcher L{nullptr};
  L.model = m;
  const PstLayerW& w0 = m->w.layer[0];
  float* ps0 = reinterpret_cast<float*>(m->layer0_tables);
  float* pr0 = reinterpret_cast<float*>(m->layer0_tables + (size_t)n * D);
  const int before = L.count;
  L.gemm(m->w.node_table, w0.msg_w1, ps0, n, D, D, Launcher::epi(nullptr), 1);
  L.gemm(m->w.node_table, w0.msg_w1 + D * D, pr0, n, D, D, Launcher::epi(w0.msg_b1), 1);
  (void)before;
  return cudaGetLastError() == cudaSuccess ? PST_OK : PST_ERR_CUDA;
}

int pst_launch_encode_fp32(const pst_model* m, cudaStream_t st, const float* edge_feat,
                           const int32_t* senders, const int32_t* offsets,
                           const int32_t* token_offsets, int B, int R, int T, float* z_out,
                           PstWorkspace& ws, int compact_features, int32_t* fused_tokens, bool* fused_tokens_done) {
  if (fused_tokens_done) *fused_tokens_done = false;
  const pst_config& cfg = m->cfg;
  const int K = cfg.num_neighbor;
  const int E = R * K;
  const bool tc = cfg.precision != PST_PREC_FP32;
  Launcher L{st};
  if (tc) L.model = m;
  int32_t* row_base = ws.row_base;
  {  // input embeddings
  PstSpan embed_span(m, st, 4);
  row_base_kernel<<<(R + 255) / 256, 256, 0, st>>>(offsets, B, R, row_base);
  ++L.count;
  if (tc && m->layer0_tables)
    node_embed_tables_kernel<<<(R + 7) / 8, 256, 0, st>>>(m->w.node_table, reinterpret_cast<const __half*>(m->layer0_tables),
                                                         reinterpret_cast<const __half*>(m->layer0_tables) + (size_t)cfg.seq_max_size * D,
                                                         row_base, ws.h, reinterpret_cast<__half*>(ws.ps),
                                                         reinterpret_cast<__half*>(ws.pr),
                                                         (m->use_msg_t && pst_edge_msg_t_ok(m)) ? reinterpret_cast<__half*>(ws.dn) : nullptr, R);
  else
    node_embed_kernel<<<(R + 7) / 8, 256, 0, st>>>(m->w.node_table, row_base, ws.h, R);
  ++L.count;
  {
    size_t smem = 2 * ((size_t)K * 28 * sizeof(float) + (size_t)K * sizeof(int));
    const int grid = (R + 1) / 2;
    if (compact_features && cfg.precision == PST_PREC_FP32) return PST_ERR_UNSUPPORTED_CONFIG;
    if (cfg.precision == PST_PREC_FP32)
      edge_embed_kernel<float><<<grid, 128, smem, st>>>(edge_feat, senders, row_base, m->w.edge_pe_table, m->w.edge_feat_w, K,
                                                        cfg.seq_max_size, R, ws.e);
    else {
      int n = pst_launch_edge_embed_tc(m, st, edge_feat, senders, row_base, R, reinterpret_cast<uint16_t*>(ws.e), compact_features);
      if (n < 0) return n;
    }
    ++L.count;
  }
  }
  if (tc) {
    // Tensor-core modes: edge-level kernels (edge_mlp_tc.cu) alternate with ONE fused node-level kernel per layer
    // (node_chain_tc.cu).  The gathered addend tables (fp16) ping-pong: (ps, pr) feed the message MLPs,
    // (ps2, pr2) the edge-update MLPs; the node kernel of layer l writes both pairs for what follows it.
    uint16_t* ps = reinterpret_cast<uint16_t*>(ws.ps);
    uint16_t* pr = reinterpret_cast<uint16_t*>(ws.pr);
    uint16_t* ps2 = reinterpret_cast<uint16_t*>(ws.agg);
    uint16_t* pr2 = reinterpret_cast<uint16_t*>(ws.u);
    L.count += pst_launch_abs_senders(m, st, senders, row_base, R, ws.senders_abs);
    if (!m->layer0_tables) {
      const PstLayerW& w0 = m->w.layer[0];
      L.gemm(ws.h, w0.msg_w1, ws.ps, R, D, D, Launcher::epi(nullptr), 1);
      L.gemm(ws.h, w0.msg_w1 + D * D, ws.pr, R, D, D, Launcher::epi(w0.msg_b1), 1);
    }
    bool msg_t = false;
    uint16_t* h16 = reinterpret_cast<uint16_t*>(ws.dn);
    for (int l = 0; l < cfg.gnn_layers; ++l) {
      {
        // message MLP: returns the per-receiver mean of the 2nd hidden layer; the 3rd linear commutes with
        // that mean (no activation follows it) and is applied by the node kernel
        PstSpan span(m, st, 1);
        int n;
        msg_t = m->use_msg_t && pst_edge_msg_t_ok(m);
        if (msg_t) {
          // transposed kernel: the sender term is a product with the gathered rows of fp16(h) (ws.dn is free until the
          // resampler); the node kernel of the previous layer has written it, layer 0 converts the embedding
          if (l == 0 && !m->layer0_tables) L.count += pst_launch_to_half(st, ws.h, h16, (size_t)R * D);  // else written with the embedding
          n = pst_launch_edge_msg_t(m, st, l, reinterpret_cast<const uint16_t*>(ws.e), h16, pr, ws.senders_abs, ws.partial, R);
        } else {
          n = pst_launch_edge_mlp_tc(m, st, l, 0, reinterpret_cast<uint16_t*>(ws.e), ps, pr, ws.senders_abs, row_base, ws.partial, R);  // the node kernel sums the partials itself
        }
        if (n < 0) return n;
        L.count += n;
      }
      {
        PstSpan span(m, st, 3);
        int n = pst_launch_node_update(m, st, l, ws.partial, msg_t ? 6 : 7, ws.h, R, ps2, pr2, ps, pr, msg_t ? h16 : nullptr);
        if (n < 0) return n;
        L.count += n;
      }
      if (l == cfg.gnn_layers - 1) break;  // the last layer's edge update is never read (model.py:385)
      PstSpan span(m, st, 2);
      int n = pst_launch_edge_mlp_tc(m, st, l, 1, reinterpret_cast<uint16_t*>(ws.e), ps2, pr2, ws.senders_abs, row_base, ws.partial, R);
      if (n < 0) return n;
      L.count += n;
    }
  } else
  for (int l = 0; l < cfg.gnn_layers; ++l) {
    const PstLayerW& w = m->w.layer[l];
    // fp32 mode (all CUDA cores): first linear factorised,
    // [h_s|h_r|e].W1 = (h.W1[0:128])[s] + (h.W1[128:256])[r] + e.W1[256:384]
    L.gemm(ws.h, w.msg_w1, ws.ps, R, D, D, Launcher::epi(nullptr));
    L.gemm(ws.h, w.msg_w1 + D * D, ws.pr, R, D, D, Launcher::epi(w.msg_b1));
    {
      PstSpan span(m, st, 1);
      GemmEpi g = Launcher::epi(nullptr, 1);
      g.gather_s = ws.ps; g.gather_r = ws.pr; g.senders = senders; g.row_base = row_base; g.knn = K;
      L.gemm(ws.e, w.msg_w1 + 2 * D * D, ws.t1, E, D, D, g);
      L.gemm(ws.t1, w.msg_w2, ws.t2, E, D, D, Launcher::epi(w.msg_b2, 1));
      L.gemm(ws.t2, w.msg_w3, ws.t1, E, D, D, Launcher::epi(w.msg_b3));
      segment_mean_kernel<<<(R + 7) / 8, 256, 0, st>>>(ws.t1, K, ws.agg, R);
      ++L.count;
    }
    L.add_ln(ws.h, ws.agg, w.ln0_s, w.ln0_o, ws.h, R);
    // feed-forward 128 -> 512 -> 128
    L.gemm(ws.h, w.ffn_w1, ws.u, R, PST_FFN, D, Launcher::epi(w.ffn_b1, 1));
    L.gemm(ws.u, w.ffn_w2, ws.tmp, R, D, PST_FFN, Launcher::epi(w.ffn_b2));
    L.add_ln(ws.h, ws.tmp, w.ln1_s, w.ln1_o, ws.h, R);
    if (l == cfg.gnn_layers - 1) break;  // the last layer's edge update is never read (model.py:385)
    L.gemm(ws.h, w.edge_w1, ws.ps, R, D, D, Launcher::epi(nullptr));
    L.gemm(ws.h, w.edge_w1 + D * D, ws.pr, R, D, D, Launcher::epi(w.edge_b1));
    {
      PstSpan span(m, st, 2);
      GemmEpi g = Launcher::epi(nullptr, 1);
      g.gather_s = ws.ps; g.gather_r = ws.pr; g.senders = senders; g.row_base = row_base; g.knn = K;
      L.gemm(ws.e, w.edge_w1 + 2 * D * D, ws.t1, E, D, D, g);
      L.gemm(ws.t1, w.edge_w2, ws.t2, E, D, D, Launcher::epi(w.edge_b2, 1));
      L.gemm(ws.t2, w.edge_w3, ws.t1, E, D, D, Launcher::epi(w.edge_b3));
      L.add_ln(ws.e, ws.t1, w.ln2_s, w.ln2_o, ws.e, E);
    }
  }

  // ---- resampler (CrossAttentionScaler, 3 blocks) -------------------------------------------
  if (tc && cfg.downsampling_ratio == 1) {
    // df = 1: token t attends residue t alone, the three blocks + head are row-local: one fused kernel
    PstSpan span(m, st, 5);
    int n = pst_launch_resampler_df1(m, st, ws.h, row_base, R, z_out, fused_tokens, fused_tokens ? ws.status : nullptr);
    if (n < 0) return n;
    if (fused_tokens && fused_tokens_done && n > 0) *fused_tokens_done = true;
    return L.count + n;
  }
  if (tc && cfg.downsampling_ratio > 1 && cfg.num_blocks <= 3 && m->use_fused_resampler && T > 0) {
    // df > 1: residue track and token track as two chain kernels (node_chain_tc.cu); k / v of the three blocks live in
    // buffers the GNN no longer needs (ws.u holds four [R,128] arrays)
    PstSpan span(m, st, 5);
    float* kv[6] = {ws.kx, ws.vx, ws.u, ws.u + (size_t)R * D, ws.u + 2 * (size_t)R * D, ws.u + 3 * (size_t)R * D};
    int n = pst_launch_resampler_dfn(m, st, ws.h, offsets, token_offsets, B, R, T, kv, ws.q, z_out, fused_tokens,
                                     fused_tokens ? ws.status : nullptr);
    if (n < 0) return n;
    if (fused_tokens && fused_tokens_done && n > 0) *fused_tokens_done = true;
    return L.count + n;
  }
  token_embed_kernel<<<(T + 7) / 8, 256, 0, st>>>(m->w.token_table, token_offsets, B, ws.res, T);
  ++L.count;
  const float* orig = ws.h;
  const float qscale = 0.17677669529663687f;  // 32 ** -0.5 (modules.py:334)
  for (int b = 0; b < cfg.num_blocks; ++b) {
    const PstBlockW& w = m->w.block[b];
    L.add_ln(ws.res, nullptr, w.qn_s, w.qn_o, ws.qn, T);
    L.add_ln(orig, nullptr, w.dn_s, w.dn_o, ws.dn, R);
    L.gemm(ws.qn, w.wg, ws.g, T, D, D, Launcher::epi(w.bg));
    L.gemm(ws.dn, w.wv, ws.vx, R, D, D, Launcher::epi(nullptr));
    if (cfg.downsampling_ratio == 1) {
      // each token attends exactly one residue (itself): softmax over a single logit is exactly 1.0f, so the
      // query / key projections cannot influence the result and are skipped; out = v * sigmoid(gate).
      if (T > 0) {
        gated_value_kernel<<<(T * (D / 4) + 255) / 256, 256, 0, st>>>(ws.vx, ws.g, ws.wa, T * (D / 4));
        ++L.count;
      }
    } else {
      L.gemm(ws.qn, w.wq, ws.q, T, D, D, Launcher::epi(nullptr, 0, nullptr, qscale));
      L.gemm(ws.dn, w.wk, ws.kx, R, D, D, Launcher::epi(nullptr));
      if (T > 0) {
        local_attention_kernel<<<T, 128, 0, st>>>(ws.q, ws.kx, ws.vx, ws.g, offsets, token_offsets, B,
                                                  cfg.downsampling_ratio, ws.wa, T);
        ++L.count;
      }
    }
    L.gemm(ws.wa, w.wo, ws.res, T, D, D, Launcher::epi(w.bo, 0, ws.res));
    // resampled transition
    L.add_ln(ws.res, nullptr, w.rt_ln_s, w.rt_ln_o, ws.qn, T);
    L.gemm(ws.qn, w.rt_w1, ws.u, T, PST_TRANS, D, Launcher::epi(w.rt_b1, 2));
    L.gemm(ws.u, w.rt_w2, ws.res, T, D, PST_TRANS, Launcher::epi(w.rt_b2, 0, ws.res));
    // original transition (dead in the last block: modules.py:624-629 output unused)
    if (b < cfg.num_blocks - 1) {
      L.add_ln(orig, nullptr, w.ot_ln_s, w.ot_ln_o, ws.dn, R);
      L.gemm(ws.dn, w.ot_w1, ws.u, R, PST_TRANS, D, Launcher::epi(w.ot_b1, 2));
      L.gemm(ws.u, w.ot_w2, ws.orig, R, D, PST_TRANS, Launcher::epi(w.ot_b2, 0, orig));
      orig = ws.orig;
    }
  }
  if (T > 0) {
    head_kernel<<<(T + 7) / 8, 256, 0, st>>>(ws.res, m->w.down_w, m->w.down_b, cfg.num_levels, z_out, T);
    ++L.count;
  }
  return L.count;
}
