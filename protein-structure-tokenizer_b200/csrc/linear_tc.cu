// Node-level linears on the tensor cores at (near) fp32 accuracy: C[M,N] = epi(A[M,K] . W[K,N]).
//
// Used in the tensor-core precision modes for every linear that runs once per residue / token
// (reference: the hk.Linear / hk.nets.MLP / common_modules.Linear calls of
// structure_tokenizer/model/gnn_layers.py:352-361,385-394,410-419 (first-layer factors, FFN),
// model/modules.py:239-251 (Transition), :303-380 (q/k/v/gate/output projections)).
// These feed LayerNorms and the quantiser directly, so they keep fp32-level accuracy: every fp32
// operand is split x = hi + lo into two fp16 values (22 significant bits) and the product is
// evaluated as hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM (three tcgen05.mma per k-step;
// the dropped lo.lo term is below 2^-22 relative).
//
// One CTA (128 threads) computes a 128 x 128 output tile.  Per 64-wide K block: A rows are read
// row-coalesced as fp32, split and written as two K-major SWIZZLE_128B operand images; the matching
// weight images (pre-split, pre-swizzled at model load) are copied in; one thread issues the MMAs
// and commits to an mbarrier.  Several CTAs are resident per SM (64 KB smem, 128 TMEM columns each),
// so one CTA's loads overlap another's MMAs and epilogue.  The epilogue reads the accumulator with one
// thread per row, applies bias / scale / activation, transposes through shared memory and stores
// row-coalesced (adding the residual there).
#include <cuda_fp16.h>

#include <vector>

#include "pst_internal.h"

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr uint32_t kImg = kBM * kBK * 2;      // 16 KB: one [128 x 64] fp16 operand image
constexpr uint32_t kSmem = 4 * kImg + 64;     // A_hi, A_lo, W_hi, W_lo (+ barrier, TMEM slot); reused as fp32 staging

__host__ __device__ __forceinline__ uint32_t swz64(uint32_t row, uint32_t k) {  // element (row, k<64)
  return row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float gelu_tanh(float x) {
  const float c = 0.7978845608028654f;
  return 0.5f * x * (1.0f + tanhf(c * (x + 0.044715f * x * x * x)));
}

struct LinearParams {
  const float* A;
  const uint16_t* w_img;  // [K/64][N/128][2 (hi,lo)][16 KB]
  const float* bias;      // [N] or null
  const float* residual;  // [M,N] or null
  float* C;
  int out_half;           // 1: C is __half [M,N] (the gathered addend tables of the edge MLPs)
  int M, N, K;
  float scale;
  int act;                // 0 none, 1 gelu(tanh), 2 relu
  uint32_t idesc;
};

__global__ void __launch_bounds__(128) linear_tc_kernel(LinearParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAhi = smem;
  uint8_t* sAlo = smem + kImg;
  uint8_t* sWhi = smem + 2 * kImg;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 4 * kImg);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * kBM;
  const int nc = blockIdx.y;
  const int n_chunks = p.N / kBN;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_slot;
  const uint32_t mbar_addr = smem_u32(mbar);
  uint32_t parity = 0;
  const int sub = tid >> 4, c16 = tid & 15;  // 16 threads per row (16 x float4 = 64 floats), 8 rows per pass

  const int n_kb = p.K / kBK;
  for (int kb = 0; kb < n_kb; ++kb) {
    if (kb > 0) {  // the MMAs of the previous block must have read the buffers
      mbar_wait(mbar_addr, parity);
      parity ^= 1;
    }
    // A block: fp32 -> (hi, lo) fp16 images
#pragma unroll 8
    for (int it = 0; it < 16; ++it) {
      const int r = it * 8 + sub;
      float4 x = make_float4(0, 0, 0, 0);
      if (m0 + r < p.M) x = __ldg(reinterpret_cast<const float4*>(p.A + (size_t)(m0 + r) * p.K + kb * kBK) + c16);
      const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      const __half2 l0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y), l1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
      const uint32_t off = swz64(r, c16 * 4);
      *reinterpret_cast<uint2*>(sAhi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      *reinterpret_cast<uint2*>(sAlo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
    // W block (hi and lo images are adjacent): 32 KB straight copy
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.w_img + ((size_t)(kb * n_chunks + nc) * 2) * (kImg / 2));
      uint4* dst = reinterpret_cast<uint4*>(sWhi);
#pragma unroll 8
      for (int i = 0; i < 16; ++i) dst[i * 128 + tid] = __ldg(src + i * 128 + tid);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), w_hi = smem_u32(sWhi), w_lo = w_hi + kImg;
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // 64 = 4 x UMMA_K(16)
        const uint32_t o = j * 32;
        umma(tmem_acc, make_desc(a_hi + o), make_desc(w_hi + o), p.idesc, (kb | j) ? 1u : 0u);
        umma(tmem_acc, make_desc(a_hi + o), make_desc(w_lo + o), p.idesc, 1u);
        umma(tmem_acc, make_desc(a_lo + o), make_desc(w_hi + o), p.idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
    }
  }
  mbar_wait(mbar_addr, parity);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: one thread per row -> staging (fp32 [128][128], float4 slot j stored at j ^ (row & 7)) ----
  float* S = reinterpret_cast<float*>(smem);
  const uint32_t tmem_row = tmem_acc + ((uint32_t)(warp * 32) << 16);
  const int n0 = nc * kBN;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    float v[32];
    tmem_ld32(tmem_row + q * 32, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 o = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
      if (p.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + q * 32) + j);
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      if (p.scale != 1.0f) { o.x *= p.scale; o.y *= p.scale; o.z *= p.scale; o.w *= p.scale; }
      if (p.act == 1) { o.x = gelu_tanh(o.x); o.y = gelu_tanh(o.y); o.z = gelu_tanh(o.z); o.w = gelu_tanh(o.w); }
      else if (p.act == 2) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      const int slot = q * 8 + j;
      *reinterpret_cast<float4*>(S + tid * 128 + ((slot ^ (tid & 7)) << 2)) = o;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {
    const int lane = tid & 31;  // 32 threads per row (32 x float4 = 128 floats), 4 rows per pass
#pragma unroll 8
    for (int it = 0; it < 32; ++it) {
      const int r = it * 4 + warp;
      if (m0 + r < p.M) {
        float4 o = *reinterpret_cast<const float4*>(S + r * 128 + ((lane ^ (r & 7)) << 2));
        const size_t g = (size_t)(m0 + r) * p.N + n0 + lane * 4;
        if (p.residual) {
          const float4 rr = *reinterpret_cast<const float4*>(p.residual + g);
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (p.out_half) {
          const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.C) + g) =
              make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        } else {
          *reinterpret_cast<float4*>(p.C + g) = o;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(128u) : "memory");
  }
}

// w: fp32 [K,N] row-major -> images [K/64][N/128][hi,lo][128 (n) x 64 (k)] swizzled
__global__ void build_split_image_kernel(const float* __restrict__ w, int K, int N, uint16_t* __restrict__ img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int k = idx / N, n = idx - k * N;
  const float x = w[idx];
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  const int kb = k / kBK, kk = k % kBK, nc = n / kBN, nn = n % kBN;
  uint16_t* base = img + ((size_t)(kb * (N / kBN) + nc) * 2) * (kImg / 2);
  base[swz64(nn, kk) >> 1] = *reinterpret_cast<const uint16_t*>(&hi);
  base[(kImg + swz64(nn, kk)) >> 1] = *reinterpret_cast<const uint16_t*>(&lo);
}

struct Entry {
  const float* w;
  int K, N;
  uint16_t* img;
};

}  // namespace

struct PstLinearRegistry {
  std::vector<Entry> entries;
};

static int register_weight(pst_model* m, const float* w, int K, int N) {
  uint16_t* img = nullptr;
  if (cudaMalloc(&img, (size_t)K * N * 2 * sizeof(uint16_t)) != cudaSuccess) return PST_ERR_CUDA;
  build_split_image_kernel<<<(K * N + 255) / 256, 256>>>(w, K, N, img);
  m->linear_tc->entries.push_back({w, K, N, img});
  return PST_OK;
}

int pst_prepare_linear_tc(pst_model* m) {
  m->linear_tc = new PstLinearRegistry();
  const int D = PST_D;
  int rc = PST_OK;
  auto reg = [&](const float* w, int K, int N) { if (rc == PST_OK) rc = register_weight(m, w, K, N); };
  for (int l = 0; l < m->cfg.gnn_layers; ++l) {
    const PstLayerW& L = m->w.layer[l];
    reg(L.msg_w1, D, D); reg(L.msg_w1 + D * D, D, D); reg(L.msg_w3, D, D);
    reg(L.ffn_w1, D, PST_FFN); reg(L.ffn_w2, PST_FFN, D);
    reg(L.edge_w1, D, D); reg(L.edge_w1 + D * D, D, D);
  }
  for (int b = 0; b < m->cfg.num_blocks; ++b) {
    const PstBlockW& B = m->w.block[b];
    reg(B.wq, D, D); reg(B.wk, D, D); reg(B.wv, D, D); reg(B.wg, D, D); reg(B.wo, D, D);
    reg(B.rt_w1, D, PST_TRANS); reg(B.rt_w2, PST_TRANS, D);
    reg(B.ot_w1, D, PST_TRANS); reg(B.ot_w2, PST_TRANS, D);
  }
  if (rc != PST_OK) return rc;
  if (cudaGetLastError() != cudaSuccess) return PST_ERR_CUDA;
  if (cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem) != cudaSuccess) return PST_ERR_CUDA;
  return PST_OK;
}

const uint8_t* pst_linear_tc_image(const pst_model* m, const float* W, int K, int N) {
  if (!m->linear_tc) return nullptr;
  for (const auto& x : m->linear_tc->entries)
    if (x.w == W && x.K == K && x.N == N) return reinterpret_cast<const uint8_t*>(x.img);
  return nullptr;
}

void pst_destroy_linear_tc(pst_model* m) {
  if (!m->linear_tc) return;
  for (auto& e : m->linear_tc->entries) cudaFree(e.img);
  delete m->linear_tc;
  m->linear_tc = nullptr;
}

// Returns 1 if launched, 0 if this weight is not registered (caller falls back to the fp32 SGEMM of the
// fp32 precision mode -- never a CPU path), <0 on error.
int pst_launch_linear_tc(const pst_model* m, cudaStream_t st, const float* A, const float* W, float* C, int M, int N, int K,
                         const float* bias, const float* residual, float scale, int act, int out_half) {
  if (!m->linear_tc || M <= 0) return 0;
  const Entry* e = nullptr;
  for (const auto& x : m->linear_tc->entries)
    if (x.w == W && x.K == K && x.N == N) { e = &x; break; }
  if (!e || (K % kBK) || (N % kBN)) return 0;
  LinearParams p{};
  p.A = A; p.w_img = e->img; p.bias = bias; p.residual = residual; p.C = C;
  p.M = M; p.N = N; p.K = K; p.scale = scale; p.act = act; p.out_half = out_half;
  p.idesc = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);  // fp16 x fp16 -> fp32, M = N = 128
  dim3 grid((M + kBM - 1) / kBM, N / kBN);
  linear_tc_kernel<<<grid, 128, kSmem, st>>>(p);
  return 1;
}
