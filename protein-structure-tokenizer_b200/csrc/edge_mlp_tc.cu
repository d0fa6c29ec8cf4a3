// Edge-level MLPs of the MPNN layers on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Reference op (structure_tokenizer/model/gnn_layers.py:344-361 and :402-436):
//     MLP_{384->128->128->128}([h_s | h_r | e])  with GELU(tanh) between the linears,
//   message mode : agg = mean over the K edges of a receiver                (:364-377)
//   update  mode : e   = MaskedLayerNorm(e + MLP(...))                        (:421-436)
// The first linear is factorised, [h_s|h_r|e].W1 = (h.W1[0:128])[s] + (h.W1[128:256] + b1)[r] + e.W1[256:384]:
// the two node-level products (fp16 tables ps, pr) are computed once per residue by the node-level kernels
// (node_chain_tc.cu / linear_tc.cu) and gathered here straight into the accumulator, so only K = 128 goes
// through the tensor core per edge.  In message mode the third linear commutes with the mean over K (no
// activation follows it), so it is applied to the per-receiver mean of the second hidden layer by the node
// kernel: this kernel returns per-receiver partial row sums (computed by one more MMA, see below).
//
// Structure: persistent CTAs (one per SM, 512 threads).  The three 128x128 16-bit weight matrices of
// the MLP stay resident in shared memory (96 KB, canonical K-major SWIZZLE_128B UMMA layout, image
// built once at model load).  A 128-edge tile lives in one of four SLOTS: a 32 KB A-operand buffer and a
// 128-column fp32 accumulator in TMEM (4 x 128 = all 512 columns).  A PRODUCER warpgroup prepares slots in
// sequence order   [TMA load of e  ||  gather ps[s] + pr[r] -> accumulator]  -> MMA1,   and three WORK GROUPS of
// 4 epilogue warps each pick the prepared tiles up round robin and run   epilogue1 (GELU, fp32 regs -> 16-bit
// smem) -> MMA2 -> epilogue2 -> {MMA3 -> TMA re-read of e -> residual + LayerNorm -> TMA store  |  row-sum MMA ->
// partial sums}:   with four slots for three groups the load, the gather and the first product of the next tile
// overlap the epilogues of the three tiles in flight.  In the 32x32b TMEM load layout each thread owns one
// accumulator row (= one edge), so the LayerNorm and the gathers are thread-local.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>

#include "pst_internal.h"

namespace {

constexpr int kTileM = 128;
constexpr int kD = 128;
constexpr int kGroups = 4;
constexpr int kThreads = kGroups * 128;      // threads of the tile groups (edge MLP: epilogue warps; embedding kernel: all)
// edge MLP kernel: four tile SLOTS (A buffer + 128-column accumulator each), three work groups of epilogue warps and one
// producer warpgroup (warps 12..15, one per TMEM lane quarter)
constexpr int kImagesPerMlp = 4;  // operand images per MLP in pst_model::tc: W1[256:384], W2, W3, W1[0:128]
#ifndef PST_UPD_SLOTS
#define PST_UPD_SLOTS 4
#endif
constexpr int kSlots = PST_UPD_SLOTS;
constexpr int kWorkGroups = kSlots - 1;
constexpr int kMlpThreads = (kWorkGroups + 1) * 128;
// register split (setmaxnreg): 512 threads launch with 128 registers each; the work groups shrink to 120, the producer
// warpgroup grows to 152 (3 * 128 * 120 + 128 * 152 = 65 536)
constexpr int kRegsEpilogue = 120;
constexpr int kRegsGather = 152;
static_assert(kSlots <= kGroups, "the A buffers and the selection-matrix slots are sized by kGroups");
constexpr uint32_t kMatBytes = 128 * 128 * 2;  // one 16-bit 128x128 operand image
constexpr uint32_t kKBlockBytes = 128 * 128;    // 128 rows x 64 elements x 2 B
constexpr uint32_t kSmemW = 3 * kMatBytes;
constexpr uint32_t kSmemA = kGroups * kMatBytes;
constexpr uint32_t kSmemVec = 4 * 128 * sizeof(float) + 64;  // b2, b3, ln scale, ln offset; per-group segment bases
constexpr uint32_t kSmemMisc = 384;  // MMA, TMA, addend-ready and accumulator-free mbarriers, TMEM slot
constexpr uint32_t kSmemTotal = kSmemW + kSmemA + kSmemVec + kSmemMisc;
static_assert(kSmemTotal <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

struct EdgeMlpParams {
  const uint16_t* w_image;  // 3 x 32 KB pre-swizzled operand images (global)
  const float* b2;
  const float* b3;
  const float* ln_s;
  const float* ln_o;
  uint16_t* e;              // [E,128] 16-bit edge state (read; rewritten in update mode)
  const __half* ps;         // [R,128] fp16  (h.W1[0:128])
  const __half* pr;         // [R,128] fp16  (h.W1[128:256] + b1)
  const int32_t* senders;   // [E] ABSOLUTE sender rows (row_base[receiver] + local index, abs_senders_kernel)
  const int32_t* row_base;  // [R]
  float* partial;           // message mode: [num_tiles][4][128] partial row sums of the 2nd hidden layer
  int E;
  int K;
  int num_tiles;
  uint32_t idesc;
  uint32_t idesc_sum;       // message mode: fp16 MN-major A x K-major B, M = 128, N = 16 (row sums)
};

// byte offset of element (row, k) inside a 128x128 16-bit K-major SWIZZLE_128B operand image
__host__ __device__ __forceinline__ uint32_t swz_offset(uint32_t row, uint32_t k) {
  uint32_t kb = k >> 6, kk = k & 63;
  return kb * kKBlockBytes + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
// Every failed poll of an mbarrier is a shared-memory-class operation on the L1 data pipe (ncu: ~800-1000 SYNCS per tile
// showed up as that many data-pipe wavefronts, a sixth of the pipe that bounds the edge MLP kernel), so a waiting warp
// asks the hardware to suspend it (suspend-time hint, ns) and backs off between polls.
#ifndef PST_MBAR_HINT_NS
#define PST_MBAR_HINT_NS 2000
#endif
#ifndef PST_MBAR_SLEEP_NS
#define PST_MBAR_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"((uint32_t)PST_MBAR_HINT_NS)
        : "memory");
    if (PST_MBAR_SLEEP_NS > 0 && !done) __nanosleep(PST_MBAR_SLEEP_NS);
  } while (!done);
}
// ---- TMA (cp.async.bulk.tensor): the 128 x 128 16-bit edge-state tile moves between HBM and the A buffer as two
// [128 rows x 64 columns] SWIZZLE_128B boxes, which is exactly the K-major operand image layout.  The copies run
// on the async proxy: no LDG / STS / LDS / STG wavefronts on the L1 data pipe that bounds this kernel.
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_tile_load(uint32_t sA_addr, const CUtensorMap* map, int row0, uint32_t mbar) {
  mbar_expect_tx(mbar, kMatBytes);
  tma_load_2d(sA_addr, map, 0, row0, mbar);
  tma_load_2d(sA_addr + kKBlockBytes, map, 64, row0, mbar);
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  // K-major, SWIZZLE_128B, 8-row atoms of 1024 B: SBO = 1024, LBO unused, version 1 (Blackwell)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// The same [128 x 128] 16-bit image read as an MN-major operand (M/N = the image's 128 K-elements, K = its rows):
// 64 MN-elements (128 B) contiguous, the next 64 at LBO = 16 KB (the image's second K block), 8 K-rows per 1024 B
// atom (SBO); canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(kKBlockBytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MMA issue: umma_f16 / umma_commit are called by ALL 32 lanes of a converged warp and elect the issuing lane inside
// the asm statement.  Called from `if (lane == 0)` code the compiler cannot know that a single thread is active and
// wraps every UTCHMMA in an ELECT / BRA.U.ANY loop (~10 instructions per MMA); with the election inside a statement
// the whole warp executes the UTCHMMAs are emitted back to back (round 2, found on the node-level kernels).
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar_addr) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(mbar_addr)
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// GELU is the tanh form (jax.nn.gelu default): 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3); see gelu2().
// tanh.approx.f32 (one MUFU op, relative error 2^-11) is below the 16-bit rounding the result gets when it is
// repacked as the next GEMM's operand; measured with the oracle: no change in token agreement.
template <typename T16>
struct Pack;
template <>
struct Pack<__half> {
  static __device__ __forceinline__ uint32_t two(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <>
struct Pack<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t two(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

template <typename T16>
struct Unpack;
template <>
struct Unpack<__half> {
  static __device__ __forceinline__ float2 two(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
};
template <>
struct Unpack<__nv_bfloat16> {
  static __device__ __forceinline__ float2 two(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
};

// Called by the whole first warp of a tile group (converged).  Not inlined on purpose: inlining lets the compiler hoist the ~60 loop-invariant
// descriptor words of the three GEMMs into registers of ALL threads (the epilogue warps run with 80 registers).
__device__ __noinline__ void issue_gemm(uint32_t tmem_acc, uint32_t sA_addr, uint32_t sW_addr, uint32_t idesc,
                                        uint32_t mbar_addr, uint32_t accumulate_first) {
  tc_fence_after();
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // K = 128 = 8 x UMMA_K(16); 4 steps per 64-element swizzle block
    uint32_t off = (j >> 2) * kKBlockBytes + (j & 3) * 32;
    umma_f16(tmem_acc, make_smem_desc(sA_addr + off), make_smem_desc(sW_addr + off), idesc, j > 0 ? 1u : accumulate_first);
  }
  umma_commit(mbar_addr);
}

// ---- packed fp32 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): two elements per instruction -------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
// GELU (tanh form) of two values.  The kernel computes G(x) = 2 gelu(x) = x + x tanh(u), u = x (c0 + c1 x^2): the
// factor 0.5 is folded into the operand that consumes the activation (W2 / W3 images and the row-sum selection
// matrix are scaled by 0.5 at build time: exact in fp16 / bf16), which saves one packed op per pair.
constexpr float kActScale = 0.5f;   // what the consumer of G(x) multiplies by
__device__ __forceinline__ float2 gelu2(float2 x) {
#ifdef PST_ABL_NOGELU  // timing experiment: no activation math
  return x;
#endif
  const float2 x2 = mul2(x, x);
  const float2 p = fma2(x2, make_float2(0.0356774081f, 0.0356774081f), make_float2(0.7978845608f, 0.7978845608f));
  const float2 u = mul2(x, p);  // sqrt(2/pi) (x + 0.044715 x^3)
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  return fma2(x, t, x);
}
// 32-byte global load through the read-only path (LDG.E.256 on sm_100)
__device__ __forceinline__ void ldg256(const uint32_t* ptr, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(ptr));
}
__device__ __forceinline__ void tmem_ld32v(uint32_t taddr, float2 (&v)[16]) { tmem_ld32(taddr, *reinterpret_cast<float (*)[32]>(&v)); }
// The same load in two halves, so that the next chunk's load is in flight while the current chunk is computed: the wait
// names the registers as in/out operands, which keeps every use of them behind it.
__device__ __forceinline__ void tmem_ld32v_issue(uint32_t taddr, float2 (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(&v[0]);
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
}
__device__ __forceinline__ void tmem_ld_wait(float2 (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(&v[0]);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st32v(uint32_t taddr, const float2 (&v)[16]) { tmem_st32(taddr, *reinterpret_cast<const float (*)[32]>(&v)); }

// 32 consecutive K-elements [k0, k0+32) of `row` -> the group's 16-bit A image (4 x 16 B, swizzled)
template <typename T16>
__device__ __forceinline__ void store_a_chunk2(uint8_t* sA, int row, int k0, const float2 (&v)[16]) {
#ifdef PST_ABL_NOSTS  // timing experiment: one of the four 16-byte stores only
#pragma unroll
  for (int c = 0; c < 1; ++c) {
    uint4 pk;
    pk.x = Pack<T16>::two(v[0].x + v[4].x + v[8].x + v[12].x, v[0].y + v[4].y + v[8].y + v[12].y);
    pk.y = Pack<T16>::two(v[1].x + v[5].x + v[9].x + v[13].x, v[1].y + v[5].y + v[9].y + v[13].y);
    pk.z = Pack<T16>::two(v[2].x + v[6].x + v[10].x + v[14].x, v[2].y + v[6].y + v[10].y + v[14].y);
    pk.w = Pack<T16>::two(v[3].x + v[7].x + v[11].x + v[15].x, v[3].y + v[7].y + v[11].y + v[15].y);
    *reinterpret_cast<uint4*>(sA + swz_offset(row, k0 + c * 8)) = pk;
  }
  return;
#endif
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 pk;
    pk.x = Pack<T16>::two(v[c * 4 + 0].x, v[c * 4 + 0].y);
    pk.y = Pack<T16>::two(v[c * 4 + 1].x, v[c * 4 + 1].y);
    pk.z = Pack<T16>::two(v[c * 4 + 2].x, v[c * 4 + 2].y);
    pk.w = Pack<T16>::two(v[c * 4 + 3].x, v[c * 4 + 3].y);
    *reinterpret_cast<uint4*>(sA + swz_offset(row, k0 + c * 8)) = pk;
  }
}

// One GELU epilogue: accumulator row (128 fp32 columns in TMEM) [+ bias] -> G(x) -> 16-bit operand image row.
// (Measured and rejected: software-pipelining the TMEM loads over two register sets (+2 %), tanh.approx.f16x2 with
// the last fma in fp16x2 (+4 % and 0.06 % fewer equal tokens): the kernel is not bound by the XU pipe or TMEM latency.)
template <typename T16, bool BIAS>
__device__ __forceinline__ void gelu_chunk(uint8_t* sA, int gt, int q, const float* sBias, float2 (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (BIAS) {
      const float4 b = *reinterpret_cast<const float4*>(sBias + q * 32 + c * 4);
      v[c * 2] = gelu2(add2(v[c * 2], make_float2(b.x, b.y)));
      v[c * 2 + 1] = gelu2(add2(v[c * 2 + 1], make_float2(b.z, b.w)));
    } else {
      v[c * 2] = gelu2(v[c * 2]);
      v[c * 2 + 1] = gelu2(v[c * 2 + 1]);
    }
  }
  store_a_chunk2<T16>(sA, gt, q * 32, v);
}

template <typename T16, bool BIAS>
__device__ __forceinline__ void gelu_epilogue(uint32_t tmem_row, uint8_t* sA, int gt, const float* sBias) {
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    float2 v[16];
#ifdef PST_ABL_NOLDTM  // timing experiment: one accumulator load per epilogue instead of four
    if (q > 0) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = make_float2((float)(gt + c + q) * 0.001f, (float)(gt - c) * 0.002f);
    } else
#endif
    tmem_ld32v(tmem_row + q * 32, v);
    gelu_chunk<T16, BIAS>(sA, gt, q, sBias, v);
  }
}

// ---- data movement -----------------------------------------------------------------------------------------
// The 16-bit edge-state tile moves by TMA (load, re-read for the residual, store): its two [128 x 64] SWIZZLE_128B
// boxes are the K-major operand image.  The gathered addend rows never touch shared memory: each thread reads its
// own sender / receiver rows with 32-byte loads and writes its accumulator row with tcgen05.st.  Activations are
// written thread-per-row into the operand image (16-byte chunks at chunk ^ (row & 7): conflict-free).

template <typename T16, int MODE>
__global__ void __launch_bounds__(kMlpThreads, 1) edge_mlp_tc_kernel(EdgeMlpParams p, const __grid_constant__ CUtensorMap tmap_e) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sAall = smem + kSmemW;
  float* sVec = reinterpret_cast<float*>(smem + kSmemW + kSmemA);                 // b2, b3, ln_s, ln_o [128] each
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kSmemW + kSmemA + 2048 + 64);
  uint64_t* tbar = mbar + kSlots;  // per slot: completion of the TMA loads of the edge-state tile
  uint64_t* abar = tbar + kSlots;  // per slot: the gather warps have preloaded the accumulator (4 arrivals)
  uint64_t* gbar = abar + kSlots;  // per slot: GEMM 1 (issued by the producer) has completed
  uint64_t* fbar = gbar + kSlots;  // per slot x lane quarter: the epilogue warp has read its accumulator rows for the last time
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fbar + 4 * kSlots);
  // Phase parities.  Every barrier completes a fixed number of times per tile of its slot, and a waiter can never be
  // more than one completion behind: gbar / abar / fbar / (message mode) tbar once, mbar twice (GEMM 2, then GEMM 3 or
  // the row sums; only the tile's own work group waits on it), (update mode) tbar twice (load, re-read).  GEMM 1 has
  // its own barrier because a work group may arrive at a slot while the slot's previous tile, owned by another
  // group, is still between its second and third product: on a shared barrier that phase has the parity it waits for.

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  // ---- one-time setup -------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_image);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    const int n16 = (MODE == 0 ? 2 : 3) * (int)(kMatBytes / 16);
    for (int i = tid; i < n16; i += kMlpThreads) dst[i] = src[i];
    if (MODE == 0) {
      uint4* z = reinterpret_cast<uint4*>(sW + 2 * kMatBytes);
      for (int i = tid; i < (int)(kSlots * 4096 / 16); i += kMlpThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid < 128) {
      sVec[tid] = p.b2[tid];
      sVec[128 + tid] = MODE == 1 ? p.b3[tid] : 0.f;
      sVec[256 + tid] = MODE == 1 ? p.ln_s[tid] : 0.f;
      sVec[384 + tid] = MODE == 1 ? p.ln_o[tid] : 0.f;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(smem_u32(&mbar[i]), 1);
    for (int i = 0; i < kSlots; ++i) mbar_init(smem_u32(&tbar[i]), 1);
    for (int i = 0; i < kSlots; ++i) mbar_init(smem_u32(&abar[i]), 4);
    for (int i = 0; i < kSlots; ++i) mbar_init(smem_u32(&gbar[i]), 1);
    for (int i = 0; i < 4 * kSlots; ++i) mbar_init(smem_u32(&fbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sW_addr = smem_u32(sW);
  // The CTA's tile sequence j = 0, 1, 2, ...: four consecutive tiles per round.  Tile j lives in slot j & 3 (A buffer +
  // accumulator) and is finished by work group j % 3; validity is monotone in j.
  const int round_stride = (int)gridDim.x * kSlots;
  auto tile_of = [&](int j) { return (int)blockIdx.x * kSlots + (j % kSlots) + (j / kSlots) * round_stride; };

  if (warp >= 4 * kWorkGroups) {
    // ================================ producer warpgroup =============================================
    // Warp 12 + q owns TMEM lane quarter q of EVERY slot.  For each tile, in sequence order, each lane reads its
    // edge's sender row of the fp16 table (h.W1a) and the receiver's row of (h.W1b + b1) (256 B each, eight 32-byte
    // loads, issued BEFORE the slot's accumulator is free: their L2 latency -- there is no L1 next to 226 KB of shared
    // memory -- overlaps the slot's previous tile), sums them in fp32 and writes its accumulator row with
    // tcgen05.st.  Warp 12 then issues GEMM 1 (acc += e . W1[256:384]) as soon as the slot's TMA load has landed, so a
    // work group that picks the tile up finds the first product done: the load, the gather and GEMM 1 of tile j + 3
    // overlap the epilogues of tiles j .. j + 2.
    if (kSlots == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsGather));
    const int q = warp - 4 * kWorkGroups;
    const int lane = tid & 31;
    const int last_recv = (p.E - 1) / p.K;
    if (q == 0 && lane == 0) {
#pragma unroll
      for (int j = 0; j < kSlots; ++j)
        if (tile_of(j) < p.num_tiles) tma_tile_load(smem_u32(sAall + j * kMatBytes), &tmap_e, tile_of(j) * kTileM, smem_u32(&tbar[j]));
    }
    int tile = tile_of(0);
    int sender = tile < p.num_tiles ? __ldg(p.senders + min(tile * kTileM + q * 32 + lane, p.E - 1)) : 0;
    for (int j = 0; tile < p.num_tiles; ++j) {
      const int s = j % kSlots;
      const uint32_t par = (uint32_t)(j / kSlots) & 1u;
      const int er = min(tile * kTileM + q * 32 + lane, p.E - 1);
      uint32_t a[64], b[64];
      {
        const uint32_t* psr = reinterpret_cast<const uint32_t*>(p.ps + (size_t)sender * kD);
        const uint32_t* prr = reinterpret_cast<const uint32_t*>(p.pr + (size_t)min(er / p.K, last_recv) * kD);
#if defined(PST_ABL_NOGATHER_PS)  // timing experiment: one 32-byte load of the sender row instead of eight
        ldg256(psr, &a[0]);
#pragma unroll
        for (int i = 8; i < 64; ++i) a[i] = a[i & 7];
#pragma unroll
        for (int i = 0; i < 8; ++i) ldg256(prr + i * 8, &b[i * 8]);
#elif defined(PST_ABL_NOGATHER_PR)  // timing experiment: one 32-byte load of the receiver row instead of eight
#pragma unroll
        for (int i = 0; i < 8; ++i) ldg256(psr + i * 8, &a[i * 8]);
        ldg256(prr, &b[0]);
#pragma unroll
        for (int i = 8; i < 64; ++i) b[i] = b[i & 7];
#elif defined(PST_ABL_NOGATHER)  // timing experiment: one 32-byte load per table instead of eight
        ldg256(psr, &a[0]);
        ldg256(prr, &b[0]);
#pragma unroll
        for (int i = 8; i < 64; ++i) { a[i] = a[i & 7]; b[i] = b[i & 7]; }
#else
#pragma unroll
        for (int i = 0; i < 8; ++i) ldg256(psr + i * 8, &a[i * 8]);
#pragma unroll
        for (int i = 0; i < 8; ++i) ldg256(prr + i * 8, &b[i * 8]);
#endif
      }
      const int ntile = tile_of(j + 1);
      if (ntile < p.num_tiles) sender = __ldg(p.senders + min(ntile * kTileM + q * 32 + lane, p.E - 1));
      mbar_wait(smem_u32(&fbar[s * 4 + q]), par ^ 1u);  // the slot's previous tile has left the accumulator rows
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)(s * 128) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float2 v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = add2(Unpack<__half>::two(a[c4 * 16 + c]), Unpack<__half>::two(b[c4 * 16 + c]));
#ifdef PST_ABL_NOTMEMST  // timing experiment: one of the four accumulator stores only
        if (c4 == 0)
#endif
        tmem_st32v(trow + c4 * 32, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&abar[s])) : "memory");
      if (q == 0) {
        mbar_wait(smem_u32(&abar[s]), par);
        mbar_wait(smem_u32(&tbar[s]), MODE == 0 ? par : 0u);
        issue_gemm(tmem_base + (uint32_t)(s * 128), smem_u32(sAall + s * kMatBytes), sW_addr, p.idesc, smem_u32(&gbar[s]), 1u);
        __syncwarp();
      }
      tile = ntile;
    }
  } else {
    // ================================ work groups (epilogue warps) ========================================
    if (kSlots == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    const int g = warp >> 2;          // work group
    const int gt = tid & 127;         // thread within the group == accumulator row
    const int wq = warp & 3;          // TMEM lane quarter this warp may access
    for (int j = g; tile_of(j) < p.num_tiles; j += kWorkGroups) {
      const int s = j % kSlots;
      const uint32_t par = (uint32_t)(j / kSlots) & 1u;
      const int tile = tile_of(j);
      const int row0 = tile * kTileM;
      const int first_recv = row0 / p.K;
      const int last_row = min(p.E - row0, kTileM);  // valid rows in this tile
      uint8_t* sA = sAall + s * kMatBytes;
      const uint32_t sA_addr = smem_u32(sA);
      const uint32_t tmem_acc = tmem_base + (uint32_t)(s * 128);
      const uint32_t tmem_row = tmem_acc + ((uint32_t)(wq * 32) << 16);
      const uint32_t mbar_addr = smem_u32(&mbar[s]);
      const uint32_t tbar_addr = smem_u32(&tbar[s]);
      const uint32_t fbar_addr = smem_u32(&fbar[s * 4 + wq]);
      // ---- 1. GEMM 1 (issued by the producer warpgroup, normally long done) -> epilogue 1: GELU -> A image -------
      mbar_wait(smem_u32(&gbar[s]), par);
      tc_fence_after();
      gelu_epilogue<T16, false>(tmem_row, sA, gt, sVec);
      fence_proxy_async();
      tc_fence_before();
      group_sync(g);
      // ---- 2. GEMM 2 ----------------------------------------------------------------------------------
      if (gt < 32) issue_gemm(tmem_acc, sA_addr, sW_addr + kMatBytes, p.idesc, mbar_addr, 0u);
      mbar_wait(mbar_addr, 0u);
      tc_fence_after();
      if (MODE == 1) {
        // ---- 3a. epilogue 2: + b2, GELU -> A image; GEMM 3; epilogue 3: residual + LayerNorm -------------
        gelu_epilogue<T16, true>(tmem_row, sA, gt, sVec);
        fence_proxy_async();
        tc_fence_before();
        group_sync(g);
        if (gt < 32) issue_gemm(tmem_acc, sA_addr, sW_addr + 2 * kMatBytes, p.idesc, mbar_addr, 0u);
        mbar_wait(mbar_addr, 1u);
        tc_fence_after();
        // pass 1: x = acc + b3 + e.  The tile is re-read (an L2 hit) by TMA into the free A buffer, in the operand
        // image layout, so each thread finds its own row conflict-free; statistics; x -> TMEM.
        if (gt == 0) tma_tile_load(sA_addr, &tmap_e, row0, tbar_addr);
        mbar_wait(tbar_addr, 1u);
        float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          float2 v[16];
          tmem_ld32v(tmem_row + q * 32, v);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint4 pk = *reinterpret_cast<const uint4*>(sA + swz_offset(gt, q * 32 + jj * 8));
            const float4 ba = *reinterpret_cast<const float4*>(sVec + 128 + q * 32 + jj * 8);
            const float4 bb = *reinterpret_cast<const float4*>(sVec + 128 + q * 32 + jj * 8 + 4);
            v[jj * 4 + 0] = add2(v[jj * 4 + 0], add2(Unpack<T16>::two(pk.x), make_float2(ba.x, ba.y)));
            v[jj * 4 + 1] = add2(v[jj * 4 + 1], add2(Unpack<T16>::two(pk.y), make_float2(ba.z, ba.w)));
            v[jj * 4 + 2] = add2(v[jj * 4 + 2], add2(Unpack<T16>::two(pk.z), make_float2(bb.x, bb.y)));
            v[jj * 4 + 3] = add2(v[jj * 4 + 3], add2(Unpack<T16>::two(pk.w), make_float2(bb.z, bb.w)));
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            sum2 = add2(sum2, v[c]);
            sq2 = fma2(v[c], v[c], sq2);
          }
          tmem_st32v(tmem_row + q * 32, v);
        }
        const float mean = (sum2.x + sum2.y) * (1.0f / kD);
        const float var = fmaxf((sq2.x + sq2.y) * (1.0f / kD) - mean * mean, 0.f);
        const float inv = rsqrtf(var + 1e-5f);
        const float2 inv2 = make_float2(inv, inv), nmi2 = make_float2(-mean * inv, -mean * inv);
        // pass 2: y = ((v - mean) * inv) * scale + offset (two packed fmas per pair) -> 16-bit image of the new edge state
        // (each thread touches only its own row of the buffer: no barrier between the passes)
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          float2 v[16];
          tmem_ld32v(tmem_row + q * 32, v);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 ls = *reinterpret_cast<const float4*>(sVec + 256 + q * 32 + c * 4);
            const float4 lo = *reinterpret_cast<const float4*>(sVec + 384 + q * 32 + c * 4);
            v[c * 2] = fma2(fma2(v[c * 2], inv2, nmi2), make_float2(ls.x, ls.y), make_float2(lo.x, lo.y));
            v[c * 2 + 1] = fma2(fma2(v[c * 2 + 1], inv2, nmi2), make_float2(ls.z, ls.w), make_float2(lo.z, lo.w));
          }
          store_a_chunk2<T16>(sA, gt, q * 32, v);
        }
        // this warp has read its accumulator rows for the last time: the producer may preload the slot's next tile
        tc_fence_before();
        __syncwarp();
        if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fbar_addr) : "memory");
        fence_proxy_async();
        group_sync(g);
        if (gt == 0) {  // TMA store of the new edge state (rows beyond E are clipped), then the slot's next tile comes in
          tma_store_2d(&tmap_e, 0, row0, sA_addr);
          tma_store_2d(&tmap_e, 64, row0, sA_addr + kKBlockBytes);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          const int nt = tile_of(j + kSlots);
          if (nt < p.num_tiles) tma_tile_load(sA_addr, &tmap_e, nt * kTileM, tbar_addr);
        }
      } else {
        // ---- 3b. epilogue 2 (message mode): + b2, GELU, partial sums over the rows of each receiver ------
        // The activated tile is staged as fp16 in the operand image layout (the precision the third GEMM's operand
        // has in update mode).  The per-receiver row sums are one more tensor-core product: the staged tile, read
        // as an MN-major A operand (M = the 128 channels, K = the 128 edge rows: the same bytes, transposed by the
        // descriptor), times a 0/1 selection matrix Sel^T [N = 16 (4 used) x K = 128 rows] built per tile, gives
        // D[channel, receiver] in fp32 in 16 TMEM columns; thread = channel reads its 4 sums.
        uint8_t* sSel = sW + 2 * kMatBytes + s * 4096;  // the W3 slot is not loaded in this mode
        gelu_epilogue<__half, true>(tmem_row, sA, gt, sVec);
        {
          // Sel^T[s][r] for r = gt: K block r >> 6, row s (128 B each), 16-byte chunk ((r & 63) >> 3) ^ s
          uint8_t* sel = sSel + (gt >> 6) * 2048 + ((gt & 7) << 1);
          const int kc = (gt & 63) >> 3;
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            const int lo_u = (first_recv + s4) * p.K - row0;
            const bool in = gt >= lo_u && gt < lo_u + p.K && gt < last_row;
            *reinterpret_cast<uint16_t*>(sel + s4 * 128 + ((kc ^ s4) << 4)) = in ? (kActScale == 0.5f ? (uint16_t)0x3800 : (uint16_t)0x3C00) : (uint16_t)0;  // 0.5 or 1.0
          }
        }
        fence_proxy_async();
        tc_fence_before();
        group_sync(g);
        if (gt < 32) {
          tc_fence_after();
          const uint32_t sel_addr = smem_u32(sSel);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {  // 16 edge rows per step: two 8-row atoms of the image (SBO = 1024)
            umma_f16(tmem_acc, make_smem_desc_mn(sA_addr + jj * 2048), make_smem_desc(sel_addr + (jj >> 2) * 2048 + (jj & 3) * 32),
                     p.idesc_sum, jj > 0 ? 1u : 0u);
          }
          umma_commit(mbar_addr);
        }
        mbar_wait(mbar_addr, 1u);
        tc_fence_after();
        if (gt == 0) {  // the row-sum product has read the A buffer: the slot's next tile comes in
          const int nt = tile_of(j + kSlots);
          if (nt < p.num_tiles) tma_tile_load(sA_addr, &tmap_e, nt * kTileM, tbar_addr);
        }
        {
          float r0, r1, r2, r3;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3) : "r"(tmem_row) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // last TMEM read of the tile: the producer may preload the slot's next tile
          tc_fence_before();
          __syncwarp();
          if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fbar_addr) : "memory");
          float* dst = p.partial + (size_t)tile * 4 * kD + gt;
          dst[0] = r0; dst[kD] = r1; dst[2 * kD] = r2; dst[3 * kD] = r3;
        }
      }
    }
    if (MODE == 1 && gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the TMA stores have landed
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// Message MLP, TRANSPOSED formulation (round 2; fp16 operands, even K >= 43).
//
// The kernel above computes  D[edge, channel] = X . W  with the edges on the accumulator rows.  Its limit is the L1 /
// shared-memory data pipe (profiles/r02_edge/README.md): addend rows gathered through registers, bias / receiver vectors,
// a selection matrix and one more product for the row sums.  Here every product is transposed,
//     D^T[channel, edge] = W^T . X^T        (A = the weight image [n][k], exactly the B image of the kernel above;
//                                            B = the edge-level operand),
// so that a thread owns one CHANNEL (accumulator row = TMEM lane) and walks along the edges (columns):
//   * sender term     (h.W1a)[s]   = one more product: the fp16 `h` rows of the senders are gathered by 16-byte
//                                    cp.async copies into a K-major B operand image and multiplied by W1a^T on the
//                                    tensor core: no register gather, no tcgen05.st;
//   * receiver term   (h.W1b+b1)[r] = a per-thread scalar per receiver (at most 4 receivers per 128-edge tile), added in
//                                    epilogue 1; b2 is a per-thread scalar as well: no bias vectors in shared memory;
//   * the activations of epilogue 1 are stored thread-per-row (row = channel) and read as an MN-major B operand;
//   * the per-receiver sums over the K edges are sums over column ranges of a thread's own row: no selection matrix, no
//     row-sum product, no activation store in epilogue 2.
// Operand traffic per tile: 16 x 16 KB of tensor-core operand reads and ONE 32 KB activation store (before: 2 940 LSU +
// 1 310 operand wavefronts).  Structure: a ring of four 16 KB stages (H k-blocks 0 / 1, E k-blocks 0 / 1) filled by a
// two loader warps, an MMA warp that runs the first two products (sender + edge term) of the next tiles into one of four
// 128-column accumulator slots while four work groups (4 epilogue warps each) finish 64-edge HALF tiles: second product
// with N = 64 on the group's own [128 x 64] activation image, partial sums per half tile.
constexpr int kTWorkGroups = 4;  // 4 warps each; a group finishes one 64-edge HALF tile at a time
constexpr int kTThreads = 640;   // warps 0-15: the work groups; warps 16, 18: loaders; warp 17: MMA issuer; warp 19: idle
constexpr int kTStages = 4;
constexpr int kTAccSlots = 4;    // 128-column accumulators (one per 128-edge tile)
constexpr uint32_t kTSmemW = 3 * kMatBytes;                    // W1a, W1c, W2 images
constexpr uint32_t kTSmemRing = kTStages * kKBlockBytes;       // 4 x [128 rows x 64 k] operand blocks
constexpr uint32_t kTSmemAct = kTWorkGroups * kKBlockBytes;    // one [128 channels x 64 edges] activation image per work group
constexpr uint32_t kTSmemMisc = 256;
constexpr uint32_t kTSmemTotal = kTSmemW + kTSmemRing + kTSmemAct + kTSmemMisc;
static_assert(kTSmemTotal <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

#ifdef PST_T_PROFILE  // per-role cycle accounting, printed by block 0 (debug builds only)
#define TPROF_DECL long long tp_last = clock64(); long long tp_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define TPROF(i) do { long long _t = clock64(); tp_acc[i] += _t - tp_last; tp_last = _t; } while (0)
#define TPROF_PRINT(name) do { if (blockIdx.x == 0 && lane == 0) printf("%s warp %d: %lld %lld %lld %lld %lld %lld %lld %lld\n", name, warp, \
    tp_acc[0], tp_acc[1], tp_acc[2], tp_acc[3], tp_acc[4], tp_acc[5], tp_acc[6], tp_acc[7]); } while (0)
#else
#define TPROF_DECL do {} while (0)
#define TPROF(i) do {} while (0)
#define TPROF_PRINT(name) do {} while (0)
#endif

struct MsgTParams {
  const uint16_t* w1a_image;  // W1[0:128]   image (32 KB)
  const uint16_t* w_image;    // W1[256:384], W2 images (2 x 32 KB; W2 pre-scaled by 0.5, see gelu2)
  const float* b2;
  const __half* pr;           // [R,128] fp16  (h.W1b + b1)
  const __half* h16;          // [R,128] fp16 node state (gathered by sender)
  const int32_t* senders;     // [E] ABSOLUTE sender rows
  float* partial;             // [num_tiles][4][128] per-receiver partial row sums of the 2nd hidden layer
  int E, K, R, num_tiles;
  uint32_t idesc;             // A, B K-major
  uint32_t idesc_bmn;         // A K-major, B MN-major
};

__global__ void __launch_bounds__(kTThreads, 1) edge_msg_t_kernel(MsgTParams p, const __grid_constant__ CUtensorMap tmap_e) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sRing = smem + kTSmemW;
  uint8_t* sAct = smem + kTSmemW + kTSmemRing;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kTSmemW + kTSmemRing + kTSmemAct);  // per stage: operand block landed
  uint64_t* empty = full + kTStages;     // per stage: the products that read it have completed
  uint64_t* gbar = empty + kTStages;     // per accumulator slot: sender + edge products done
  uint64_t* fbar = gbar + kTAccSlots;    // per accumulator slot: both work groups have read it for the last time (8 warps)
  uint64_t* mbar = fbar + kTAccSlots;    // per work group x quarter: second product done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * kTWorkGroups);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  {
    const uint4* src_a = reinterpret_cast<const uint4*>(p.w1a_image);
    const uint4* src = reinterpret_cast<const uint4*>(p.w_image);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int i = tid; i < (int)(kMatBytes / 16); i += kTThreads) dst[i] = src_a[i];
    for (int i = tid; i < (int)(2 * kMatBytes / 16); i += kTThreads) dst[kMatBytes / 16 + i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i < kTStages; ++i) { mbar_init(smem_u32(&full[i]), i < 2 ? 64 : 1); mbar_init(smem_u32(&empty[i]), 1); }
    for (int i = 0; i < kTAccSlots; ++i) { mbar_init(smem_u32(&gbar[i]), 1); mbar_init(smem_u32(&fbar[i]), 8); }
    for (int i = 0; i < 2 * kTWorkGroups; ++i) mbar_init(smem_u32(&mbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sW_addr = smem_u32(sW);
  const uint32_t ring_addr = smem_u32(sRing);
  auto tile_of = [&](int j) { return (int)blockIdx.x + j * (int)gridDim.x; };

  if (warp == 16 || warp == 18) {
    // ================================ loader warps ===================================================
    // per tile: the sender rows of fp16 `h` -> stages 0 / 1 (the two k-blocks of a [128 x 128] K-major SWIZZLE_128B
    // image) by 16-byte cp.async copies, sixteen lanes per 256-byte row (two rows per warp instruction: whole lines on
    // the L2 side, no registers in flight, each lane computes its swizzled destination; TMA tile::gather4 moves one
    // 128-byte row segment per ~16 cycles per SM, 4 000 cycles per tile: measured, profiles/r02_edge), warp 16 the tile's
    // rows 0..63, warp 18 rows 64..127; warp 16 also loads the edge-state tile -> stages 2 / 3 by TMA.
    const int lw = (warp - 16) >> 1;     // 0 / 1: which 64 rows
    const int max_row = p.R - 1;
    auto load_idx = [&](int tile_) {  // sender rows of this warp's rows 2 lane, 2 lane + 1
      int2 v = make_int2(0, 0);
      if (tile_ < p.num_tiles) {
        v = __ldg(reinterpret_cast<const int2*>(p.senders + (size_t)tile_ * kTileM + lw * 64 + lane * 2));  // the workspace pads the array
        v.x = min(max(v.x, 0), max_row); v.y = min(max(v.y, 0), max_row);  // rows beyond E hold anything: any valid row will do
      }
      return v;
    };
    const int half = lane >> 4;          // which of the instruction's two rows
    const int chunk = lane & 15;         // 16-byte chunk of the row
    const uint32_t dst_lane = ring_addr + (uint32_t)(chunk >> 3) * kKBlockBytes;  // k-block = stage
    int tile = tile_of(0);
    int2 idx = load_idx(tile);
    TPROF_DECL;
    for (int j = 0; tile < p.num_tiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      TPROF(0);
      mbar_wait(smem_u32(&empty[0]), ph ^ 1u);
      mbar_wait(smem_u32(&empty[1]), ph ^ 1u);
      TPROF(1);  // waiting for the H stages
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {  // rows 2 i, 2 i + 1 of this warp's 64: their indices sit in lane i
        const int ia = __shfl_sync(0xffffffffu, idx.x, i);
        const int ib = __shfl_sync(0xffffffffu, idx.y, i);
        const int r = lw * 64 + 2 * i + half;
        const __half* src = p.h16 + (size_t)(half ? ib : ia) * kD + chunk * 8;
        const uint32_t dst = dst_lane + (uint32_t)r * 128u + ((uint32_t)((chunk & 7) ^ (r & 7)) << 4);
#ifdef PST_ABL_T_NOGATHER  // timing experiment: one copy in eight
        if ((i & 7) == 0)
#endif
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
      // both stages complete when every lane's copies have landed (the barriers expect 2 x 32 arrivals)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[0])) : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[1])) : "memory");
      TPROF(2);  // issuing the gather
      if (lw == 0) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(smem_u32(&empty[2 + kb]), ph ^ 1u);
          if (lane == 0) {
            mbar_expect_tx(smem_u32(&full[2 + kb]), kKBlockBytes);
            tma_load_2d(ring_addr + (2 + kb) * kKBlockBytes, &tmap_e, kb * 64, tile * kTileM, smem_u32(&full[2 + kb]));
          }
        }
      }
      TPROF(3);  // waiting for the E stages + issuing their loads
      tile = tile_of(j + 1);
      idx = load_idx(tile);
    }
    TPROF_PRINT("loader [-, wait H stages, issue gather, wait + issue E, ...]");
  } else if (warp == 17) {
    // ================================ MMA warp ========================================================
    // D^T = W1a^T . H_s^T + W1c^T . E^T  into the tile's accumulator slot, one operand block (4 k-steps) at a time
    TPROF_DECL;
    for (int j = 0; tile_of(j) < p.num_tiles; ++j) {
      const int slot = j % kTAccSlots;
      const uint32_t acc = tmem_base + (uint32_t)(slot * 128);
      TPROF(0);
      mbar_wait(smem_u32(&fbar[slot]), ((uint32_t)(j / kTAccSlots) & 1u) ^ 1u);  // the slot's previous tile has been read
      TPROF(1);  // waiting for a free accumulator slot
      tc_fence_after();
#pragma unroll
      for (int st = 0; st < kTStages; ++st) {
        mbar_wait(smem_u32(&full[st]), (uint32_t)j & 1u);
        TPROF(2 + st);  // waiting for stage st (+ issuing the previous stage's products)
        if (st < 2) fence_proxy_async();  // the cp.async copies are generic-proxy writes; the tensor core reads through the async proxy
        tc_fence_after();
        {
          const uint32_t wblk = sW_addr + (st < 2 ? 0u : kMatBytes) + (uint32_t)(st & 1) * kKBlockBytes;
          const uint32_t xblk = ring_addr + (uint32_t)st * kKBlockBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(acc, make_smem_desc(wblk + ks * 32), make_smem_desc(xblk + ks * 32), p.idesc, (st | ks) ? 1u : 0u);
          umma_commit(smem_u32(&empty[st]));
          if (st == kTStages - 1) umma_commit(smem_u32(&gbar[slot]));
        }
        __syncwarp();
#ifndef PST_T_NO_PACING
        // Pacing: the tensor pipe runs the products of all issuers in order, and the work groups' second products are on
        // their critical path while these are a tile or two ahead.  Never more than one operand block (4 k-steps,
        // 256 cycles) of look-ahead work is queued in front of them.
        mbar_wait(smem_u32(&empty[st]), (uint32_t)j & 1u);
#endif
      }
    }
    TPROF_PRINT("mma [-, wait slot, wait stage 0, 1, 2, 3]");
  } else if (warp < 16) {
    // ================================ work groups (epilogue warps) ========================================
    // Group g (4 warps, warp w owns TMEM lane quarter w & 3 = 32 output channels) finishes the 64-edge half
    // (g & 1) of the tiles j = g >> 1, (g >> 1) + 2, ...: four independent half tiles in flight per SM, one warp of each on
    // every SM sub-partition, so that the MUFU-bound epilogues of some overlap the tensor-core waits of the others.
    // A thread = one channel x 64 edges.
    const int g = warp >> 2;
    const int wq = warp & 3;
    const int hf = g & 1;
    const int gt = wq * 32 + lane;    // output channel == accumulator row
    uint8_t* act = sAct + g * kKBlockBytes;
    const uint32_t act_addr = smem_u32(act);
    const uint32_t mbar_addr = smem_u32(&mbar[2 * g]);
    const float b2m = __ldg(p.b2 + gt);
    const int last_recv = (p.E - 1) / p.K;
    uint32_t mparity = 0;
    // receiver-term scalars of the half tile (fp16 table; at most 3 receivers in 64 edges), fetched one ahead
    __half prn[3];
    auto load_pr = [&](int tile_, __half (&out)[3]) {
      if (tile_ < p.num_tiles) {
        const int fr = (tile_ * kTileM + hf * 64) / p.K;
#pragma unroll
        for (int i = 0; i < 3; ++i) out[i] = __ldg(p.pr + (size_t)min(fr + i, last_recv) * kD + gt);
      }
    };
    load_pr(tile_of(g >> 1), prn);
    TPROF_DECL;
    for (int j = g >> 1; tile_of(j) < p.num_tiles; j += 2) {
      TPROF(0);
      const int slot = j % kTAccSlots;
      const int tile = tile_of(j);
      const int u0 = tile * kTileM + hf * 64;          // first edge of the half tile
      const int first_recv = u0 / p.K;
      const int last_row = min(p.E - u0, 64);          // valid columns (<= 0: none)
      const uint32_t trow = tmem_base + (uint32_t)(slot * 128 + hf * 64) + ((uint32_t)(wq * 32) << 16);
      // column (relative to the half tile) where receiver slot i starts: b1 < b2; slot 0 starts at or before column 0;
      // both even (K and 64 are)
      const int b1 = (first_recv + 1) * p.K - u0, b2c = b1 + p.K;
      const float pr0 = __half2float(prn[0]), pr1 = __half2float(prn[1]), pr2 = __half2float(prn[2]);
      load_pr(tile_of(j + 2), prn);
      // ---- epilogue 1: x = acc + receiver term -> G(x) -> activation image (row = channel, columns = edges) ----------
      mbar_wait(smem_u32(&gbar[slot]), (uint32_t)(j / kTAccSlots) & 1u);
      TPROF(1);  // waiting for the first two products
      tc_fence_after();
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int c0 = qq * 32;
        float2 v[16];
        tmem_ld32v(trow + c0, v);
        const int i0 = (c0 >= b1) + (c0 >= b2c);
        const float pa = i0 == 0 ? pr0 : i0 == 1 ? pr1 : pr2;
        const float pb = i0 == 0 ? pr1 : pr2;
        const int t = (i0 == 0 ? b1 : i0 == 1 ? b2c : 1 << 20) - c0;  // columns of this chunk before the next receiver
        if (t >= 32) {
          const float2 pp = make_float2(pa, pa);
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = gelu2(add2(v[c], pp));
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float pc = 2 * c < t ? pa : pb;
            v[c] = gelu2(add2(v[c], make_float2(pc, pc)));
          }
        }
        store_a_chunk2<__half>(act, gt, c0, v);
      }
      TPROF(2);  // epilogue 1
      fence_proxy_async();
      tc_fence_before();
      group_sync(g);
      TPROF(3);  // group barrier
      // ---- second product: D^T[:, half] = W2^T . act1 (B = the activation image read MN-major, N = 64) ----------------
      if ((warp & 3) == 0) {
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(slot * 128 + hf * 64);
        const uint32_t w2 = sW_addr + 2 * kMatBytes;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_f16(acc, make_smem_desc(w2 + (ks >> 2) * kKBlockBytes + (ks & 3) * 32), make_smem_desc_mn(act_addr + ks * 2048),
                   p.idesc_bmn, ks > 0 ? 1u : 0u);
        umma_commit(mbar_addr);
      }
      mbar_wait(mbar_addr, mparity);
      mparity ^= 1;
      TPROF(4);  // second product
      tc_fence_after();
      // ---- epilogue 2: + b2, G, sums over the columns of each receiver (thread-local) ------------------------------
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int c0 = qq * 32;
        float2 v[16];
        tmem_ld32v(trow + c0, v);
        const float2 bb = make_float2(b2m, b2m);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = gelu2(add2(v[c], bb));
        if (c0 + 32 > last_row) {  // last tile only: columns beyond E do not exist
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (c0 + 2 * c >= last_row) v[c].x = 0.f;
            if (c0 + 2 * c + 1 >= last_row) v[c].y = 0.f;
          }
        }
        const int i0 = (c0 >= b1) + (c0 >= b2c);
        const int t = (i0 == 0 ? b1 : i0 == 1 ? b2c : 1 << 20) - c0;
        float2 sa2 = make_float2(0.f, 0.f), sb2 = make_float2(0.f, 0.f);
        if (t >= 32) {
          float2 sc2 = make_float2(0.f, 0.f), sd2 = make_float2(0.f, 0.f), se2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 16; c += 4) {  // four independent chains
            sa2 = add2(sa2, v[c]); sc2 = add2(sc2, v[c + 1]); sd2 = add2(sd2, v[c + 2]); se2 = add2(se2, v[c + 3]);
          }
          sa2 = add2(add2(sa2, sc2), add2(sd2, se2));
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            if (2 * c < t) sa2 = add2(sa2, v[c]);
            else sb2 = add2(sb2, v[c]);
          }
        }
        const float sa = sa2.x + sa2.y, sb = sb2.x + sb2.y;
        if (i0 == 0) { s0 += sa; s1 += sb; }
        else if (i0 == 1) { s1 += sa; s2 += sb; }
        else s2 += sa;
      }
      TPROF(5);  // epilogue 2
      // last TMEM read of the half tile: when both halves are through, the MMA warp may reuse the slot
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fbar[slot])) : "memory");
      float* dst = p.partial + ((size_t)tile * 2 + hf) * 4 * kD + gt;  // [64-edge half tile][receiver slot][channel]
      dst[0] = kActScale * s0; dst[kD] = kActScale * s1; dst[2 * kD] = kActScale * s2;
      TPROF(6);
    }
    TPROF_PRINT("work [-, wait products 1+2, epilogue 1, barrier, product 3, epilogue 2, store]");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// Input edge embedding on the tensor cores (reference: structure_tokenizer/model/structure_encoder.py:94-105):
//   e0 = PE_edge(s - r) . W[0:128] + b  (a constant table row, folded at weight-pack time)  +  f27 . W[128:155]
// Per 128-edge tile: the 27 features are read as one contiguous coalesced block, split into fp16 hi/lo
// operand images (K padded to 32), multiplied by the hi/lo images of W[128:155] with three tcgen05.mma per
// k-step (hi.hi + hi.lo + lo.hi: fp32-level accuracy); meanwhile the fp16 table rows are gathered
// row-coalesced into the (then free) A buffer; the epilogue adds them to the accumulator row, packs the
// 16-bit edge state in place and the tile is copied out row-coalesced.  Same 4-group persistent structure
// as the edge MLP kernel.
constexpr uint32_t kEmbSmemW = 2 * 16384;  // W hi / lo images, [128 n x 64 k] (k < 32 used)
// + per group: the tile's 128 table row indices and a 16 KB landing buffer for the first k-block of the gathered rows
constexpr uint32_t kEmbSmemTotal = kEmbSmemW + kSmemA + kGroups * 512 + 64 + kGroups * kKBlockBytes;
static_assert(kEmbSmemTotal <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

struct EmbedParams {
  const uint16_t* w_img;     // hi image (16 KB) then lo image (16 KB)
  const __half* table;       // [2*seq_max-1, 128] fp16
  const float* feat;         // [E, 27], or [E, 16] in the compact layout (see the kernel)
  const int32_t* senders;    // [E]
  const int32_t* row_base;   // [R]
  uint16_t* e;               // [E, 128] out
  int E, K, num_tiles, seq_max;
  uint32_t idesc;
};

__device__ __forceinline__ uint32_t swz_k64(uint32_t row, uint32_t kk) {  // element (row, kk < 64) of a [128 x 64] image
  return row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1));
}

template <typename T16, bool COMPACT>
__global__ void __launch_bounds__(kThreads, 1) edge_embed_tc_kernel(EmbedParams p, const __grid_constant__ CUtensorMap tmap_e) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sAall = smem + kEmbSmemW;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kEmbSmemW + kSmemA + kGroups * 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + kGroups);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = warp >> 2, gt = tid & 127, wq = warp & 3;
  uint8_t* sA = sAall + g * kMatBytes;
  int* sIdx = reinterpret_cast<int*>(smem + kEmbSmemW + kSmemA) + g * 128;  // PE-table row of each edge of the group's tile
  uint8_t* sT = smem + kEmbSmemW + kSmemA + kGroups * 512 + 64 + g * kKBlockBytes;  // landing buffer (k-block 0 of the table rows)
  const uint32_t sT_addr = smem_u32(sT);
  // The fp16 PE-table rows of the tile, one 64-column k-block at a time, by 16-byte cp.async copies into the operand
  // image layout: eight lanes per 128-byte half row, four rows per warp instruction (whole lines; each thread loading
  // its own row with eight 32-byte loads cost one L1 data-pipe wavefront per load, 1 024 per tile, and 0.24 ms of this
  // kernel's 0.59: profiles/r02_edge/README.md).  One commit group per call.
  auto gather_table_block = [&](int kb, uint32_t dst_base) {
    const int chunk = gt & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 16 + (gt >> 3);
#ifdef PST_ABL_EMB_NOTABLE  // timing experiment: one copy in eight
      if (i != 0) continue;
#endif
      const __half* src = p.table + (size_t)sIdx[r] * kD + kb * 64 + chunk * 8;
      const uint32_t dst = dst_base + (uint32_t)r * 128u + ((uint32_t)(chunk ^ (r & 7)) << 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int i = tid; i < (int)(kEmbSmemW / 16); i += kThreads) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) mbar_init(smem_u32(&mbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base + (uint32_t)(g * 128);
  const uint32_t tmem_row = tmem_acc + ((uint32_t)(wq * 32) << 16);
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW), mbar_addr = smem_u32(&mbar[g]);
  uint32_t parity = 0;
  // table row of this thread's edge: sender - receiver (local indices) + seq_max - 1; the two dependent index loads
  // are issued one tile ahead
  auto table_row = [&](int tile_) {
    const int er_ = tile_ * kTileM + gt;
    if (tile_ >= p.num_tiles || er_ >= p.E) return 0;
    const int recv = er_ / p.K;
    return __ldg(p.senders + er_) - (recv - __ldg(p.row_base + recv)) + (p.seq_max - 1);
  };
  int trow_next = table_row(blockIdx.x * kGroups + g);
  bool store_pending = false;  // thread gt == 0: a TMA store of the previous tile may still be reading the buffer

  for (int tile = blockIdx.x * kGroups + g; tile < p.num_tiles; tile += gridDim.x * kGroups) {
    const int row0 = tile * kTileM;
    const int last_row = min(p.E - row0, kTileM);
    const int trow = trow_next;
    // ---- features -> fp16 hi / lo operand images (A_hi at K block 0, A_lo at K block 1 of the buffer) ----
    if (COMPACT) {
      // Fused tokenize path: 16 floats per edge from the k-NN kernel, [d*d, 12 orientation features, 0, 0, 0].  Each
      // thread reads its own edge row (four 16-byte loads, whole sectors), evaluates the 15 RBFs exp(-d*d / 1.5^k)
      // (utils/protein_utils.py:257-281) in fp32 with ex2.approx (relative error ~1e-6 where the value matters: far
      // below the fp16 rounding of the embedding it feeds; the fp64 features of pst_featurize_knn are untouched) and
      // writes its row of both images with eight 16-byte stores.
      const int er = row0 + gt;
      float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0;
      if (gt < last_row) {
        const float4* src = reinterpret_cast<const float4*>(p.feat + (size_t)er * 16);
        r0 = __ldg(src); r1 = __ldg(src + 1); r2 = __ldg(src + 2); r3 = __ldg(src + 3);
      }
      {
        const int nt = tile + gridDim.x * kGroups;
        if (nt < p.num_tiles) {
          const size_t nrow0 = (size_t)nt * kTileM;
          if (gt < 64 && nrow0 + (size_t)gt * 2 < (size_t)p.E)  // 128 rows x 64 B = 64 lines
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.feat + nrow0 * 16) + gt * 128));
          if (gt >= 124 && nrow0 + (size_t)(gt - 124) * 32 < (size_t)p.E)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.senders + nrow0) + (gt - 124) * 128));
        }
      }
      trow_next = table_row(tile + gridDim.x * kGroups);
      float f[32];
      {
        // -log2(e) / 1.5^k, k = 0..14
        constexpr float kC[15] = {-1.44269504f, -0.961796694f, -0.641197796f, -0.427465197f, -0.284976798f,
                                  -0.189984532f, -0.126656355f, -0.0844375698f, -0.0562917132f, -0.0375278088f,
                                  -0.0250185392f, -0.0166790261f, -0.0111193507f, -0.00741290049f, -0.00494193366f};
#pragma unroll
        for (int k = 0; k < 15; ++k) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f[k]) : "f"(r0.x * kC[k]));
        f[15] = r0.y; f[16] = r0.z; f[17] = r0.w;
        f[18] = r1.x; f[19] = r1.y; f[20] = r1.z; f[21] = r1.w;
        f[22] = r2.x; f[23] = r2.y; f[24] = r2.z; f[25] = r2.w;
        f[26] = r3.x;
#pragma unroll
        for (int k = 27; k < 32; ++k) f[k] = 0.f;
        if (gt >= last_row) {
#pragma unroll
          for (int k = 0; k < 15; ++k) f[k] = 0.f;
        }
      }
      // the buffer is free once the previous tile's TMA store has read it: waited for here, behind the loads above
      if (gt == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      sIdx[gt] = trow;
      group_sync(g);
      gather_table_block(0, sT_addr);  // columns 0..63 of the tile's PE-table rows: lands behind the staging and the product
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = f[c * 8 + 2 * j], b = f[c * 8 + 2 * j + 1];
          const __half2 h = __floats2half2_rn(a, b);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
          hi[j] = *reinterpret_cast<const uint32_t*>(&h);
          lo[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const uint32_t off = swz_k64(gt, c * 8);
        *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sA + kKBlockBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else
    {
      const float* f0 = p.feat + (size_t)row0 * PST_EDGE_FEATURES;
      const int n_el = last_row * PST_EDGE_FEATURES;
      // all 27 loads of the thread in flight at once (one HBM round trip), then the next tile's block and sender
      // indices are pulled into L2 (one 128-byte line per thread)
      float x[PST_EDGE_FEATURES];
#pragma unroll
      for (int i = 0; i < PST_EDGE_FEATURES; ++i) {
        const int idx = gt + i * 128;
        x[i] = idx < n_el ? __ldg(f0 + idx) : 0.f;
      }
      {
        const int nt = tile + gridDim.x * kGroups;
        if (nt < p.num_tiles) {
          const size_t nrow0 = (size_t)nt * kTileM;
          const char* nf = reinterpret_cast<const char*>(p.feat + nrow0 * PST_EDGE_FEATURES) + gt * 128;
          if (gt < 108 && nrow0 * PST_EDGE_FEATURES * 4 + (size_t)gt * 128 < (size_t)p.E * PST_EDGE_FEATURES * 4)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nf));
          if (gt >= 124 && nrow0 + (size_t)(gt - 124) * 32 < (size_t)p.E)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.senders + nrow0) + (gt - 124) * 128));
        }
      }
      trow_next = table_row(tile + gridDim.x * kGroups);
      // the buffer is free once the previous tile's TMA store has read it: waited for here, behind the loads above
      if (gt == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      sIdx[gt] = trow;
      group_sync(g);
      gather_table_block(0, sT_addr);  // columns 0..63 of the tile's PE-table rows: lands behind the staging and the product
#pragma unroll
      for (int i = 0; i < PST_EDGE_FEATURES; ++i) {
        const int idx = gt + i * 128;
        const int r = idx / PST_EDGE_FEATURES, kk = idx - r * PST_EDGE_FEATURES;
        const __half hi = __float2half_rn(x[i]);
        const __half lo = __float2half_rn(x[i] - __half2float(hi));
        const uint32_t off = swz_k64(r, kk);
        *reinterpret_cast<__half*>(sA + off) = hi;
        *reinterpret_cast<__half*>(sA + kKBlockBytes + off) = lo;
      }
      // zero the K padding 27..31 of this thread's own row (the matching weight rows are zero, but 0 * garbage
      // must not be NaN)
#pragma unroll
      for (int kk = PST_EDGE_FEATURES; kk < 32; ++kk) {
        const uint32_t off = swz_k64(gt, kk);
        *reinterpret_cast<uint16_t*>(sA + off) = 0;
        *reinterpret_cast<uint16_t*>(sA + kKBlockBytes + off) = 0;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    group_sync(g);
    if (gt < 32) {
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < 2; ++j) {  // K = 32 = 2 x UMMA_K
        const uint32_t o = j * 32;
        umma_f16(tmem_acc, make_smem_desc(sA_addr + o), make_smem_desc(sW_addr + o), p.idesc, j > 0 ? 1u : 0u);
        umma_f16(tmem_acc, make_smem_desc(sA_addr + o), make_smem_desc(sW_addr + 16384 + o), p.idesc, 1u);
        umma_f16(tmem_acc, make_smem_desc(sA_addr + kKBlockBytes + o), make_smem_desc(sW_addr + o), p.idesc, 1u);
      }
      umma_commit(mbar_addr);
    }
    mbar_wait(mbar_addr, parity);
    parity ^= 1;
    tc_fence_after();
    // the product has read the buffer: columns 64..127 of the table rows land in its second k-block, in place, while
    // the first half of the epilogue runs from the landing buffer
    gather_table_block(1, sA_addr + kKBlockBytes);
    // ---- epilogue: acc + table row -> 16-bit edge state in the operand image layout (each thread only touches its
    // own row), which is also the layout of the TMA store boxes ----
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    group_sync(g);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q == 2) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        group_sync(g);
      }
      float2 v[16];
      tmem_ld32v(tmem_row + q * 32, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 pk = q < 2 ? *reinterpret_cast<const uint4*>(sT + swz_k64(gt, q * 32 + j * 8))
                               : *reinterpret_cast<const uint4*>(sA + swz_offset(gt, q * 32 + j * 8));
        v[j * 4 + 0] = add2(v[j * 4 + 0], Unpack<__half>::two(pk.x));
        v[j * 4 + 1] = add2(v[j * 4 + 1], Unpack<__half>::two(pk.y));
        v[j * 4 + 2] = add2(v[j * 4 + 2], Unpack<__half>::two(pk.z));
        v[j * 4 + 3] = add2(v[j * 4 + 3], Unpack<__half>::two(pk.w));
      }
      store_a_chunk2<T16>(sA, gt, q * 32, v);
    }
    fence_proxy_async();
    tc_fence_before();
    group_sync(g);
    if (gt == 0) {  // rows beyond E are clipped by the tensor map
      tma_store_2d(&tmap_e, 0, row0, sA_addr);
      tma_store_2d(&tmap_e, 64, row0, sA_addr + kKBlockBytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      store_pending = true;
    }
  }
  if (gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// W[128:155] (fp32 [32 (27 used), 128]) -> hi / lo fp16 images [128 n x 64 k]; table fp32 -> fp16
__global__ void build_embed_images_kernel(const float* __restrict__ wf, uint16_t* __restrict__ img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128 * 64) return;
  const int n = idx >> 6, kk = idx & 63;
  const float x = kk < PST_EDGE_FEATURES ? wf[kk * kD + n] : 0.f;
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  img[swz_k64(n, kk) >> 1] = *reinterpret_cast<const uint16_t*>(&hi);
  img[(16384 + swz_k64(n, kk)) >> 1] = *reinterpret_cast<const uint16_t*>(&lo);
}
__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

// senders_abs[e] = row_base[receiver(e)] + senders[e]: absolute row of the sender (the k-NN indices are local to
// their structure, gnn_layers.py:344-351 indexes node_inp[senders] per structure)
__global__ void abs_senders_kernel(const int32_t* __restrict__ senders, const int32_t* __restrict__ row_base, int K, int E,
                                   int32_t* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) out[e] = __ldg(row_base + e / K) + __ldg(senders + e);
}

template <typename T16>
__global__ void build_weight_image_kernel(const float* __restrict__ w, uint16_t* __restrict__ image, float scale) {
  // w: fp32 [128 (k), 128 (n)] row-major; image element (n, k) at swz_offset(n, k)
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128 * 128) return;
  int k = idx >> 7, n = idx & 127;
  T16 v = T16(w[idx] * scale);  // scale = 0.5 for the consumers of G(x) = 2 gelu(x): exact
  image[swz_offset(n, k) >> 1] = *reinterpret_cast<uint16_t*>(&v);
}

}  // namespace


int pst_prepare_tc_weights(pst_model* m) {
  const int layers = m->cfg.gnn_layers;
  const size_t per_mlp = kImagesPerMlp * (size_t)kMatBytes;  // W1[256:384], W2, W3, W1[0:128]
  const size_t total = (size_t)layers * 2 * per_mlp;
  if (cudaMalloc(&m->tc_dev, total) != cudaSuccess) return PST_ERR_CUDA;
  m->tc.w = m->tc_dev;
  for (int l = 0; l < layers; ++l) {
    const PstLayerW& L = m->w.layer[l];
    const float* src[2][kImagesPerMlp] = {{L.msg_w1 + 2 * kD * kD, L.msg_w2, L.msg_w3, L.msg_w1},
                                          {L.edge_w1 + 2 * kD * kD, L.edge_w2, L.edge_w3, L.edge_w1}};
    for (int t = 0; t < 2; ++t)
      for (int j = 0; j < kImagesPerMlp; ++j) {
        uint16_t* dst = m->tc_dev + ((size_t)(l * 2 + t) * per_mlp + (size_t)j * kMatBytes) / 2;
        const float scale = (j == 1 || j == 2) ? kActScale : 1.0f;  // the consumers of G(x) = 2 gelu(x)
        if (m->cfg.precision == PST_PREC_FP16)
          build_weight_image_kernel<__half><<<64, 256>>>(src[t][j], dst, scale);
        else
          build_weight_image_kernel<__nv_bfloat16><<<64, 256>>>(src[t][j], dst, scale);
      }
  }
  {
    const int n_table = (2 * m->cfg.seq_max_size - 1) * kD;
    if (cudaMalloc(&m->embed_img_dev, 2 * 16384) != cudaSuccess) return PST_ERR_CUDA;
    if (cudaMalloc(&m->table16_dev, (size_t)n_table * sizeof(uint16_t)) != cudaSuccess) return PST_ERR_CUDA;
    build_embed_images_kernel<<<32, 256>>>(m->w.edge_feat_w, m->embed_img_dev);
    to_half_kernel<<<(n_table + 255) / 256, 256>>>(m->w.edge_pe_table, reinterpret_cast<__half*>(m->table16_dev), n_table);
    if (cudaFuncSetAttribute(edge_embed_tc_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmbSmemTotal) != cudaSuccess ||
        cudaFuncSetAttribute(edge_embed_tc_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmbSmemTotal) != cudaSuccess ||
        cudaFuncSetAttribute(edge_embed_tc_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmbSmemTotal) != cudaSuccess ||
        cudaFuncSetAttribute(edge_embed_tc_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmbSmemTotal) != cudaSuccess)
      return PST_ERR_CUDA;
  }
  if (cudaGetLastError() != cudaSuccess) return PST_ERR_CUDA;
  if (cudaFuncSetAttribute(edge_msg_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmemTotal) != cudaSuccess) return PST_ERR_CUDA;
  cudaError_t e1 = cudaFuncSetAttribute(edge_mlp_tc_kernel<__half, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  cudaError_t e2 = cudaFuncSetAttribute(edge_mlp_tc_kernel<__half, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  cudaError_t e3 = cudaFuncSetAttribute(edge_mlp_tc_kernel<__nv_bfloat16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  cudaError_t e4 = cudaFuncSetAttribute(edge_mlp_tc_kernel<__nv_bfloat16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) return PST_ERR_CUDA;
  return PST_OK;
}

size_t pst_tc_partial_floats(int R, int K) {
  // the transposed message kernel takes its partial sums over 64-edge half tiles, both halves of the last tile included
  size_t tiles = ((size_t)R * K + kTileM - 1) / kTileM;
  return tiles * 2 * 4 * kD;
}

// Tensor map of the 16-bit edge state [E, 128]: boxes of [128 rows x 64 columns], SWIZZLE_128B (= one K block of
// the operand image).  cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime so that
// the library does not link libcuda.
typedef CUresult (*PstEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_edge_state_map(const uint16_t* e, int E, CUtensorMap* out, int box_rows = kTileM) {
  static PstEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return PST_ERR_CUDA;
    fn = reinterpret_cast<PstEncodeTiledFn>(sym);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)E};
  const cuuint64_t strides[1] = {(cuuint64_t)kD * sizeof(uint16_t)};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};  // box_rows = 1: TMA row gather (tile::gather4, tools/probes/gather4_probe.cu)
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<uint16_t*>(e), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? PST_OK : PST_ERR_CUDA;
}

// mode 0: writes per-tile partial row sums of the 2nd hidden layer (partial[num_tiles][4][128]); the node kernel forms the
// per-receiver mean and applies W3, b3.
// mode 1: e <- LN(e + MLP).
int pst_launch_edge_mlp_tc(const pst_model* m, cudaStream_t st, int layer, int mode, uint16_t* e, const uint16_t* ps,
                           const uint16_t* pr, const int32_t* senders, const int32_t* row_base, float* partial, int R) {
  const int K = m->cfg.num_neighbor;
  const PstLayerW& L = m->w.layer[layer];
  EdgeMlpParams p{};
  p.w_image = m->tc.w + ((size_t)(layer * 2 + mode) * kImagesPerMlp * kMatBytes) / 2;
  p.b2 = mode == 0 ? L.msg_b2 : L.edge_b2;
  p.b3 = mode == 0 ? L.msg_b3 : L.edge_b3;
  p.ln_s = L.ln2_s;
  p.ln_o = L.ln2_o;
  p.e = e;
  p.ps = reinterpret_cast<const __half*>(ps);
  p.pr = reinterpret_cast<const __half*>(pr);
  p.senders = senders;
  p.row_base = row_base;
  p.partial = partial;
  p.E = R * K;
  p.K = K;
  p.num_tiles = (p.E + kTileM - 1) / kTileM;
  const uint32_t fmt = m->cfg.precision == PST_PREC_FP16 ? 0u : 1u;  // F16 / BF16
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  // row sums: fp16 x fp16 -> fp32, A MN-major (bit 15), M = 128 channels, N = 16 receivers slots
  p.idesc_sum = (1u << 4) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(kD >> 4) << 24);
  if (p.num_tiles == 0) return 0;
  if (K < 43) return PST_ERR_UNSUPPORTED_CONFIG;  // a 128-row tile must touch at most 4 receivers
  CUtensorMap tmap;
  if (int rc = make_edge_state_map(e, p.E, &tmap)) return rc;
  const bool half = m->cfg.precision == PST_PREC_FP16;
  int grid = m->num_sms;
  const int need = (p.num_tiles + kSlots - 1) / kSlots;
  if (grid > need) grid = need;
  if (mode == 0) {
    if (half) edge_mlp_tc_kernel<__half, 0><<<grid, kMlpThreads, kSmemTotal, st>>>(p, tmap);
    else edge_mlp_tc_kernel<__nv_bfloat16, 0><<<grid, kMlpThreads, kSmemTotal, st>>>(p, tmap);
    return 1;  // the fused node kernel (node_chain_tc.cu) sums the partial rows of each receiver
  }
  if (half) edge_mlp_tc_kernel<__half, 1><<<grid, kMlpThreads, kSmemTotal, st>>>(p, tmap);
  else edge_mlp_tc_kernel<__nv_bfloat16, 1><<<grid, kMlpThreads, kSmemTotal, st>>>(p, tmap);
  return 1;
}

// Transposed message MLP (edge_msg_t_kernel): fp16 operands, even K >= 43.  `h16` = fp16 copy of the node state [R,128]
// (the sender term is a product with its gathered rows), `pr` = fp16 (h.W1b + b1) [R,128].
bool pst_edge_msg_t_ok(const pst_model* m) {
  const int K = m->cfg.num_neighbor;
  return m->cfg.precision == PST_PREC_FP16 && K >= 43 && (K & 1) == 0;
}
int pst_launch_to_half(cudaStream_t st, const float* src, uint16_t* dst, size_t n) {
  if (n == 0) return 0;
  to_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, reinterpret_cast<__half*>(dst), (int)n);
  return 1;
}
int pst_launch_edge_msg_t(const pst_model* m, cudaStream_t st, int layer, const uint16_t* e, const uint16_t* h16, const uint16_t* pr,
                          const int32_t* senders_abs, float* partial, int R) {
  const int K = m->cfg.num_neighbor;
  MsgTParams p{};
  const uint16_t* base = m->tc.w + ((size_t)(layer * 2 + 0) * kImagesPerMlp * kMatBytes) / 2;
  p.w_image = base;                                 // W1[256:384], W2
  p.w1a_image = base + (3 * (size_t)kMatBytes) / 2;  // W1[0:128]
  p.b2 = m->w.layer[layer].msg_b2;
  p.pr = reinterpret_cast<const __half*>(pr);
  p.senders = senders_abs;
  p.partial = partial;
  p.E = R * K;
  p.K = K;
  p.R = R;
  p.num_tiles = (p.E + kTileM - 1) / kTileM;
  p.idesc = (1u << 4) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);  // fp16 x fp16 -> fp32, M = N = 128
  // second product per 64-edge half tile: B operand MN-major, N = 64
  p.idesc_bmn = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
  if (p.num_tiles == 0) return 0;
  p.h16 = reinterpret_cast<const __half*>(h16);
  CUtensorMap tmap_e;
  if (int rc = make_edge_state_map(e, p.E, &tmap_e)) return rc;
  const int grid = m->num_sms < p.num_tiles ? m->num_sms : p.num_tiles;
  edge_msg_t_kernel<<<grid, kTThreads, kTSmemTotal, st>>>(p, tmap_e);
  return 1;
}

int pst_launch_edge_embed_tc(const pst_model* m, cudaStream_t st, const float* feat, const int32_t* senders,
                             const int32_t* row_base, int R, uint16_t* e, int compact) {
  EmbedParams p{};
  p.w_img = m->embed_img_dev;
  p.table = reinterpret_cast<const __half*>(m->table16_dev);
  p.feat = feat;
  p.senders = senders;
  p.row_base = row_base;
  p.e = e;
  p.K = m->cfg.num_neighbor;
  p.E = R * p.K;
  p.num_tiles = (p.E + kTileM - 1) / kTileM;
  p.seq_max = m->cfg.seq_max_size;
  p.idesc = (1u << 4) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);  // fp16 x fp16 -> fp32
  if (p.num_tiles == 0) return 0;
  int grid = m->num_sms;
  const int need = (p.num_tiles + kGroups - 1) / kGroups;
  if (grid > need) grid = need;
  CUtensorMap tmap;
  if (int rc = make_edge_state_map(e, p.E, &tmap)) return rc;
  const bool half = m->cfg.precision == PST_PREC_FP16;
  if (compact) {
    if (half) edge_embed_tc_kernel<__half, true><<<grid, kThreads, kEmbSmemTotal, st>>>(p, tmap);
    else edge_embed_tc_kernel<__nv_bfloat16, true><<<grid, kThreads, kEmbSmemTotal, st>>>(p, tmap);
  } else {
    if (half) edge_embed_tc_kernel<__half, false><<<grid, kThreads, kEmbSmemTotal, st>>>(p, tmap);
    else edge_embed_tc_kernel<__nv_bfloat16, false><<<grid, kThreads, kEmbSmemTotal, st>>>(p, tmap);
  }
  return 1;
}

int pst_launch_abs_senders(const pst_model* m, cudaStream_t st, const int32_t* senders, const int32_t* row_base, int R,
                           int32_t* senders_abs) {
  const int K = m->cfg.num_neighbor, E = R * K;
  if (E <= 0) return 0;
  abs_senders_kernel<<<(E + 255) / 256, 256, 0, st>>>(senders, row_base, K, E, senders_abs);
  return 1;
}
