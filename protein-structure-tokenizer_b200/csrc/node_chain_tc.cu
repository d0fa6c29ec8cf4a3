// Fused node-level chains on the tensor cores (tcgen05 + TMEM), tensor-core precision modes only.
//
// Everything the encoder does once per residue between two edge-level kernels is row-local (every op maps
// row r of [R,128] to row r), so a whole chain runs in ONE persistent kernel per 128-row tile instead of a
// launch per linear / LayerNorm with the [R,128] intermediates bouncing through HBM:
//
//  node_update_kernel  (reference: structure_tokenizer/model/gnn_layers.py:364-419, per MPNN layer)
//      agg = tbar.W3 + b3                       (tbar = per-receiver mean of the message MLP's 2nd hidden layer)
//      h   = MaskedLayerNorm0(h + agg)          (:379-382)
//      h   = MaskedLayerNorm1(h + MLP_{128->512->128}(h))   (:385-399; the 512-wide hidden is chunked 4 x 128
//                                                 and never leaves the SM)
//      out_o = fp16(h.Wout_o + b_o)             the gathered addend tables (3, or 4 for the old message kernel) the next two edge-level
//                                                 kernels need: edge MLP of this layer (:402-419) and message
//                                                 MLP of the next (:344-361), first linear factorised
//
//  resampler_df1_kernel  (reference: model/modules.py:438-636, model/model.py:148-174,414-418,
//                         model/quantize.py:175-209; downsampling_ratio == 1 only)
//      the 3 CrossAttentionScaler blocks + spherical norm + down_proj + FSQ.  With df = 1 token t attends
//      residue t alone (softmax over one logit == 1), so the block is row-local: res += (v * sigmoid(gate)).Wo + bo,
//      res += Transition(res), orig += Transition(orig).
//
// Numerics: as linear_tc.cu - every fp32 operand is split x = hi + lo (two fp16) and each product is evaluated as
// hi.hi + hi.lo + lo.hi with fp32 accumulation in TMEM; LayerNorm / GELU / sigmoid / FSQ in fp32 registers.
//
// Structure (round 2): 1 CTA / SM in clusters of 2; 8 epilogue warps (thread = row x column half: the TMEM 32x32b
// layout), a weight-producer warp and an MMA warp.  A 3-slot ring of 32 KB weight half-units (hi + lo images of one
// 64-row K block of a 128-column weight slice) is streamed from L2 by cp.async.bulk with cluster multicast on full /
// empty mbarriers; tcgen05.commit releases a slot.  The epilogue threads post every product into a command queue, the
// MMA warp issues it.  Activations are A operands IN TMEM wherever a region is free: hidden chunks of the MLPs are
// written in place over their accumulators, node_update_kernel keeps its input / h1 images in a fourth region and has
// no activation in shared memory; the resampler kernels stage their LayerNorm outputs in one or two shared-memory image
// sets (X, U: hi / lo, 2 K blocks, 64 KB each) because all four TMEM regions hold state or accumulators there.  The
// per-row state that must survive a product (res / orig of the resampler) lives in TMEM, not in registers.
#include <cuda_fp16.h>

#include <vector>

#include <cstdio>
#include <cstring>

#include "fsq_device.cuh"
#include "pst_internal.h"

namespace {

constexpr int D = PST_D;
constexpr uint32_t kImgBlk = 16384;        // [128 x 64] fp16 K-major SWIZZLE_128B image
constexpr uint32_t kSlotBytes = 2 * kImgBlk;  // weight half-unit: hi image, lo image of one K block
constexpr uint32_t kSetBytes = 2 * kSlotBytes;  // activation image set: [kb][hi, lo]
constexpr int kSlots = 3;
constexpr int kCluster = 2;  // CTAs per cluster: each weight half-unit is fetched from L2 once per cluster and multicast
constexpr uint32_t kOffX = 0, kOffU = kSetBytes, kOffW = 2 * kSetBytes;
constexpr uint32_t kOffBar = kOffW + kSlots * kSlotBytes;
constexpr uint32_t kBarBytes = 512;  // barriers + TMEM slot (128), command barriers (128), command queue (256)
constexpr uint32_t kSmemBytes = kOffBar + kBarBytes + 2 * 2 * 128 * 4;  // + row-statistics exchange
constexpr int kEpiThreads = 256, kThreads = kEpiThreads + 64;  // + weight producer warp + MMA warp
constexpr int kCmdSlots = 8;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory of sm_100");

__host__ __device__ __forceinline__ uint32_t swz64(uint32_t row, uint32_t k) {
  return row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(mbar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MMA issue.  These four are called by ALL 32 lanes of the (converged) MMA warp and elect the issuing lane inside the
// asm statement.  Round 2 finding: called from `if (tid == X)` code the compiler cannot know that one thread is active
// and wraps every UTCHMMA in an ELECT / BRA.U.ANY loop with its operands re-materialised (~10 instructions, ~115
// cycles per MMA measured: a split 128 x 128 x 128 product took 2 750 cycles for 1 536 cycles of tensor work).  With
// the election inside a statement the whole warp executes, the UTCHMMAs are emitted back to back from uniform registers.
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// A operand from TMEM (K-major only: lane = row, each 32-bit column holds two consecutive K elements), B from shared memory
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of the cluster
__device__ __forceinline__ void umma_commit_mc(uint32_t mbar, uint16_t mask) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(mbar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(mbar)
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

// 32-byte global accesses (LDG / STG .256 on sm_100): a thread that owns a whole row segment touches full sectors
__device__ __forceinline__ void ld256f(const float* ptr, float* r) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(ptr)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void st256f(float* ptr, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]),
               "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
__device__ __forceinline__ void st256u(void* ptr, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// Round 2, measured as a same-box A/B (profiles/r02_edge/README.md is the edge kernel's log; this one: gpurun_out/r2b, r2d):
// the GELU of the 512-wide FFN hidden layer written with ex2 + rcp was 41 % of node_update_kernel's instructions.  One
// MUFU.TANH (relative error ~2^-11, what the edge MLP's GELU uses; the result is split into fp16 hi / lo operand images
// right after) and reciprocal multiplies for the two per-row divisions: node updates 0.95 -> 0.81 ms, resampler + head
// 0.87 -> 0.78 ms per step, token agreement with the fp32 mode unchanged (99.9611 % of 131 072).  -DPST_NODE_EXACT_MATH
// restores the round-1 arithmetic.
#ifndef PST_NODE_EXACT_MATH
#define PST_NODE_GELU_APPROX 1
#define PST_NODE_RECIP_MUL 1
#endif
#ifdef PST_NODE_GELU_APPROX
__device__ __forceinline__ float tanh_fast(float u) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return t;
}
#else
// tanh to ~1e-7 absolute: 1 - 2 / (exp(2u) + 1)   (exp overflow -> +inf -> 1, underflow -> 0 -> -1)
__device__ __forceinline__ float tanh_fast(float u) { return 1.0f - __fdividef(2.0f, __expf(2.0f * u) + 1.0f); }
#endif
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanh_fast(u));
}
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// 32 consecutive K-elements [k0, k0+32) of `row`, fp32 -> hi / lo fp16 images of an activation image set
__device__ __forceinline__ void split_store(uint8_t* set, int row, int k0, const float (&v)[32]) {
  uint8_t* base = set + (k0 >> 6) * kSlotBytes;
  const int kk0 = k0 & 63;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[c * 8 + 2 * j], b = v[c * 8 + 2 * j + 1];
      const __half2 h = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h);
      const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const uint32_t off = swz64(row, kk0 + c * 8);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
#ifndef PST_NODE_NOSPLIT
    *reinterpret_cast<uint4*>(base + kImgBlk + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#endif
  }
}

// state of the weight ring as seen by the MMA-issuing thread / the producer thread
struct Ring {
  uint32_t full, empty, wbase;  // smem addresses: full[kSlots], empty[kSlots] mbarriers, slot 0
  uint32_t n;                   // half-units consumed / produced so far
#ifdef PST_NODE_PROFILE
  long long waited = 0;         // cycles the thread spent waiting on the ring's barriers
#endif
};
// -DPST_NODE_PROFILE: per-phase cycle counters of block 0 (issuing thread, one other epilogue thread, the producer),
// printed at the end of the kernel; the ring's `waited` = cycles the issuing thread waited for weights.
#ifdef PST_NODE_PROFILE
#define NPROF_DECL long long np_last = clock64(); long long np_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define NPROF(i) do { long long _t = clock64(); np_acc[i] += _t - np_last; np_last = _t; } while (0)
#define NPROF_PRINT(name, extra) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 160)) printf("%s thread %d: %lld %lld %lld %lld %lld %lld %lld %lld | ring wait %lld\n", name, \
  (int)threadIdx.x, np_acc[0], np_acc[1], np_acc[2], np_acc[3], np_acc[4], np_acc[5], np_acc[6], np_acc[7], (long long)(extra)); } while (0)
#else
#define NPROF_DECL do {} while (0)
#define NPROF(i) do {} while (0)
#define NPROF_PRINT(name, extra) do {} while (0)
#endif

// one 128 x 128 x 128 product (split precision): A image set x the next two half-units of the ring -> acc.
// flags bit 0: accumulate onto the accumulator's contents; bit 1: the A image set is in TMEM (a_set = TMEM address of a
// 128-column region: K block kb at columns [64 kb, 64 kb + 64): 32 columns of hi pairs, then 32 columns of lo pairs),
// otherwise in shared memory (a_set = address of an activation image set).
constexpr uint32_t kFlagAcc = 1u, kFlagTmemA = 2u;
__device__ __forceinline__ void issue_unit(Ring& r, uint32_t a_set, uint32_t tmem_acc, uint32_t flags, uint32_t idesc,
                                           uint32_t done_bar) {
  const uint32_t accumulate = flags & kFlagAcc;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    const uint32_t slot = r.n % kSlots, par = (r.n / kSlots) & 1;
#ifdef PST_NODE_PROFILE
    const long long t0 = clock64();
#endif
    mbar_wait(r.full + slot * 8, par);
#ifdef PST_NODE_PROFILE
    r.waited += clock64() - t0;
#endif
    tc_after();
    const uint32_t w_hi = r.wbase + slot * kSlotBytes, w_lo = w_hi + kImgBlk;
    if (flags & kFlagTmemA) {
      const uint32_t a_hi = a_set + kb * 64, a_lo = a_hi + 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t o = j * 32;
        umma_ts(tmem_acc, a_hi + j * 8, make_desc(w_hi + o), idesc, (kb | j) ? 1u : accumulate);
#ifndef PST_NODE_NOSPLIT
        umma_ts(tmem_acc, a_hi + j * 8, make_desc(w_lo + o), idesc, 1u);
        umma_ts(tmem_acc, a_lo + j * 8, make_desc(w_hi + o), idesc, 1u);
#endif
      }
    } else {
      const uint32_t a_hi = a_set + kb * kSlotBytes, a_lo = a_hi + kImgBlk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t o = j * 32;
        umma(tmem_acc, make_desc(a_hi + o), make_desc(w_hi + o), idesc, (kb | j) ? 1u : accumulate);
#ifndef PST_NODE_NOSPLIT
        umma(tmem_acc, make_desc(a_hi + o), make_desc(w_lo + o), idesc, 1u);
        umma(tmem_acc, make_desc(a_lo + o), make_desc(w_hi + o), idesc, 1u);
#endif
      }
    }
    umma_commit_mc(r.empty + slot * 8, (uint16_t)((1u << kCluster) - 1));  // the slot is refilled for the whole cluster
    ++r.n;
  }
  umma_commit(done_bar);
}

// Weight streaming, cluster-cooperative: every CTA of the cluster runs the same schedule in lock step.  For each
// half-unit (32 KB) CTA `rank` fetches its 1/kCluster slice from L2 and multicasts it into the same ring slot of
// every CTA; each CTA's full barrier therefore collects kCluster slices.  A slot is refilled only when every CTA has
// released it: the empty barriers count kCluster arrivals (multicast tcgen05.commit).  Without this, all 148 SMs
// pull the same 64 KB per 128 x 128 product from a few L2 slices, which bounds the kernel.
__device__ __forceinline__ void producer_loop(Ring r, const uint8_t* const* sched, int n_sched, int my_tiles) {
  const uint32_t rank = cluster_ctarank();
#ifdef PST_NODE_NOSPLIT  // experiment: hi images only (plain fp16 products)
  constexpr uint32_t kFetch = kImgBlk;
#else
  constexpr uint32_t kFetch = kSlotBytes;
#endif
  constexpr uint32_t kSlice = kFetch / kCluster;
  for (int t = 0; t < my_tiles; ++t)
    for (int i = 0; i < n_sched; ++i) {
      const uint32_t slot = r.n % kSlots;
#ifdef PST_NODE_PROFILE
      const long long t0 = clock64();
#endif
      if (r.n >= kSlots) mbar_wait(r.empty + slot * 8, ((r.n / kSlots) - 1) & 1);
#ifdef PST_NODE_PROFILE
      r.waited += clock64() - t0;
#endif
      mbar_expect_tx(r.full + slot * 8, kFetch);
      bulk_g2s_mc(r.wbase + slot * kSlotBytes + rank * kSlice, sched[i] + rank * kSlice, kSlice, r.full + slot * 8,
                  (uint16_t)((1u << kCluster) - 1));
      ++r.n;
    }
#ifdef PST_NODE_PROFILE
  if (blockIdx.x == 0) printf("producer: %u half-units, waited %lld cycles for free slots\n", r.n, r.waited);
#endif
}

// ---- epilogue thread layout ----------------------------------------------------------------------------------
// 256 epilogue threads: row = tid & 127 (TMEM lane; the lane quarter a warp may touch is warp & 3), half = tid >> 7
// selects the 64 columns [64*half, 64*half + 64) of the row (= K block `half` of an activation image set).  Two
// warps per SM sub-partition hide each other's latencies; row statistics are exchanged through shared memory.
struct Epi {
  int tid, row, half;
  uint32_t lane_off;  // TMEM lane field of this thread's warp
  float* red;         // [2 stages][2 halves][128 rows]
};

// LayerNorm over the full 128-wide row (biased variance of the centred row, eps 1e-5: gnn_layers.py:108-120,162-164;
// hk.LayerNorm is the same formula); x = this thread's 64 columns
// Per-channel parameter vectors (biases, LayerNorm scale / offset) are passed BY VALUE inside the kernel parameter
// struct and read from the constant bank (`cv`, offsets in floats): next to 226 KB of shared memory there is no L1 to
// speak of, so the same vectors read through the global path cost an L2 round trip in every epilogue
// (measured upper bound: 0.19 ms of the 1.9 ms the two node-level kernels take per step).
template <int N>
__device__ __forceinline__ void layer_norm_row(const Epi& e, float (&x)[2][32], const float (&cv)[N], int scale, int offset) {
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 32; ++j) s += x[q][j];
  e.red[e.half * 128 + e.row] = s;
  epi_sync();
  const float mean = (e.red[e.row] + e.red[128 + e.row]) * (1.0f / D);
  float v = 0.f;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = x[q][j] - mean;
      v = fmaf(d, d, v);
    }
  e.red[256 + e.half * 128 + e.row] = v;
  epi_sync();
  const float inv = rsqrtf((e.red[256 + e.row] + e.red[256 + 128 + e.row]) * (1.0f / D) + 1e-5f);
  const int sc = scale + e.half * 64, of = offset + e.half * 64;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 32; ++j) x[q][j] = (cv[sc + q * 32 + j] * inv) * (x[q][j] - mean) + cv[of + q * 32 + j];
}

__device__ __forceinline__ void tmem_ld_half(const Epi& e, uint32_t region, float (&x)[2][32]) {
  tmem_ld32(region + e.lane_off + e.half * 64, x[0]);
  tmem_ld32(region + e.lane_off + e.half * 64 + 32, x[1]);
}
__device__ __forceinline__ void tmem_st_half(const Epi& e, uint32_t region, const float (&x)[2][32]) {
  tmem_st32(region + e.lane_off + e.half * 64, x[0]);
  tmem_st32(region + e.lane_off + e.half * 64 + 32, x[1]);
}
__device__ __forceinline__ void split_store_half(const Epi& e, uint8_t* set, const float (&x)[2][32]) {
  split_store(set, e.row, e.half * 64, x[0]);
  split_store(set, e.row, e.half * 64 + 32, x[1]);
}
// The thread's 64 values (K block `half` of its row) as hi / lo fp16 pairs into ITS OWN 64 columns of a TMEM region:
// an accumulator the thread has just read becomes, in place, the A operand of the next product (no shared memory, no
// swizzle, no proxy fence; the MMA warp reads hi at columns [64 half, +32), lo at [64 half + 32, +32)).
__device__ __forceinline__ void split_store_tmem_half(const Epi& e, uint32_t region, const float (&x)[2][32]) {
  float pk[32];
  uint32_t* u = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const __half2 h = __floats2half2_rn(x[c >> 4][(c & 15) * 2], x[c >> 4][(c & 15) * 2 + 1]);
    u[c] = *reinterpret_cast<const uint32_t*>(&h);
  }
  tmem_st32(region + e.lane_off + e.half * 64, pk);
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float a = x[c >> 4][(c & 15) * 2], b = x[c >> 4][(c & 15) * 2 + 1];
    const float2 f = __half22float2(__floats2half2_rn(a, b));
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    u[c] = *reinterpret_cast<const uint32_t*>(&l);
  }
  tmem_st32(region + e.lane_off + e.half * 64 + 32, pk);
}
// the inverse: the thread's 64 values back from the hi / lo images it wrote (x = hi + lo: the split keeps 22 bits)
__device__ __forceinline__ void load_image_tmem_half(const Epi& e, uint32_t region, float (&x)[2][32]) {
  float hi[32], lo[32];
  tmem_ld32(region + e.lane_off + e.half * 64, hi);
  tmem_ld32(region + e.lane_off + e.half * 64 + 32, lo);
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi[c]));
    const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo[c]));
    x[c >> 4][(c & 15) * 2] = h.x + l.x;
    x[c >> 4][(c & 15) * 2 + 1] = h.y + l.y;
  }
}
__device__ __forceinline__ void load_row_half(const Epi& e, const float* row_ptr, bool valid, float (&x)[2][32]) {
  const float* src = row_ptr + e.half * 64;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (valid) ld256f(src + q * 32 + j * 8, &x[q][j * 8]);
      else {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[q][j * 8 + k] = 0.f;
      }
    }
}
template <int N>
__device__ __forceinline__ void add_bias_half(const Epi& e, float (&x)[2][32], const float (&cv)[N], int bias) {
  const int b = bias + e.half * 64;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 32; ++j) x[q][j] += cv[b + q * 32 + j];
}
// x += r + bias
template <int N>
__device__ __forceinline__ void add_residual_bias_half(const Epi& e, float (&x)[2][32], const float (&r)[2][32],
                                                       const float (&cv)[N], int bias) {
  const int b = bias + e.half * 64;
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int j = 0; j < 32; ++j) x[q][j] += r[q][j] + cv[b + q * 32 + j];
}

struct Setup {
  uint8_t* smem;
  uint32_t tmem_base;
  uint32_t bar_done[4];
  uint32_t cmd_bar, cmd;  // smem addresses: kCmdSlots mbarriers, kCmdSlots 16-byte commands
  Ring ring;
  float* red;
};

__device__ __forceinline__ Setup chain_setup(uint8_t* smem, int tid, int warp) {
  Setup s;
  s.smem = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);  // full[3], empty[3], done[4], tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  if (tid == 0) {
    for (int i = 0; i < 10; ++i) mbar_init(smem_u32(&bars[i]), (i >= 3 && i < 6) ? kCluster : 1);
    for (int i = 0; i < kCmdSlots; ++i) mbar_init(smem_u32(&bars[16 + i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_before();
  __syncthreads();
  cluster_sync();  // every CTA's barriers exist before a peer multicasts into them
  tc_after();
  s.tmem_base = *tmem_slot;
  s.ring.full = smem_u32(&bars[0]);
  s.ring.empty = smem_u32(&bars[3]);
  s.ring.wbase = smem_u32(smem + kOffW);
  s.ring.n = 0;
  for (int i = 0; i < 4; ++i) s.bar_done[i] = smem_u32(&bars[6 + i]);
  s.cmd_bar = smem_u32(&bars[16]);
  s.cmd = smem_u32(smem + kOffBar + 256);
  s.red = reinterpret_cast<float*>(smem + kOffBar + kBarBytes);
  return s;
}

__device__ __forceinline__ void chain_teardown(const Setup& s, int warp) {
  tc_before();
  __syncthreads();
  cluster_sync();  // no CTA leaves while a peer may still multicast into its shared memory / barriers
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s.tmem_base), "r"(512u) : "memory");
}

// MMA groups: every epilogue thread counts them; thread 0 POSTS each group (operand set, accumulator, accumulate flag,
// completion barrier) into a small command queue in shared memory and the MMA warp issues it.  Round 2: with thread 0
// issuing the 24 products of a group itself (descriptor arithmetic + waiting for the weight slots) every other
// epilogue thread waited for it at the next barrier, about 1 000 cycles per group on the critical path of a chain
// that is latency-bound anyway (per-phase counters, -DPST_NODE_PROFILE).  Group i commits to done[i & 3] and the
// threads wait for the groups in issue order, so each barrier has at most one unobserved completion as long as no
// more than four groups are in flight; that also keeps the kCmdSlots = 8 deep queue from wrapping onto an unread entry.
struct Groups {
  uint32_t cmd_bar, cmd;
  uint32_t bar_done[4];
  uint32_t n_issued = 0, n_waited = 0;
  int tid;
#ifdef PST_NODE_PROFILE
  long long pf[6] = {0, 0, 0, 0, 0, 0}, pf_last = 0;  // chunked_mlp: waits, load + activation, image store, publish + post
  __device__ __forceinline__ void prof(int i) { const long long t = clock64(); pf[i] += t - pf_last; pf_last = t; }
#define GPROF(i) G.prof(i)
#else
#define GPROF(i) do {} while (0)
#endif
  __device__ __forceinline__ void issue(uint32_t a_set, uint32_t acc, uint32_t flags) {  // flags: kFlagAcc | kFlagTmemA
    if (tid == 0) {
      const uint32_t slot = n_issued % kCmdSlots;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cmd + slot * 16), "r"(a_set), "r"(acc), "r"(flags),
                   "r"(bar_done[n_issued & 3])
                   : "memory");
      asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(cmd_bar + slot * 8) : "memory");
    }
    ++n_issued;
  }
  __device__ __forceinline__ void wait_next() {
    mbar_wait(bar_done[n_waited & 3], (n_waited >> 2) & 1);
    ++n_waited;
    tc_after();
  }
};
// the MMA warp's loop (all 32 lanes, converged; the issuing lane is elected inside umma / umma_commit): `total` groups,
// in the order the epilogue threads post them
__device__ __forceinline__ void mma_loop(Ring ring, uint32_t cmd_bar, uint32_t cmd, uint32_t idesc, int total) {
  for (int i = 0; i < total; ++i) {
    const uint32_t slot = (uint32_t)i % kCmdSlots;
    mbar_wait(cmd_bar + slot * 8, ((uint32_t)i / kCmdSlots) & 1);
    uint32_t a_set, acc, flags, done;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a_set), "=r"(acc), "=r"(flags), "=r"(done) : "r"(cmd + slot * 16) : "memory");
    tc_after();
    issue_unit(ring, a_set, acc, flags, idesc, done);
  }
#ifdef PST_NODE_PROFILE
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) printf("mma warp: waited %lld cycles for weights\n", ring.waited);
#endif
}
// images written by the epilogue threads become visible to the tensor core; accumulators read by them may be reused
__device__ __forceinline__ void publish() {
  fence_async();
  tc_before();
  epi_sync();
}
// the same for images written into TMEM (tcgen05.st has been waited for by the storing thread)
__device__ __forceinline__ void publish_tmem() {
  tc_before();
  epi_sync();
}

// y (+)= act(X . W1[:, c] + b1[c]) . W2[c, :] accumulated over `chunks` 128-wide chunks of the hidden layer.
// X: image set of the input in shared memory (already published); accH[2]: two hidden accumulators.  The activated
// chunk is written IN PLACE over its accumulator as hi / lo fp16 images and the second-layer product takes its A
// operand from TMEM (round 2; before, one shared-memory image set U held the chunk and the store into it had to wait
// for the previous chunk's second-layer product).  Issue order F1(0) F1(1) F2(0) F1(2) F2(1) F1(3) ... == wait order;
// the tensor pipe executes one thread's products in issue order, so F1(c + 2) overwrites accH[c & 1] only after F2(c)
// has read the image in it.  acc_flags = kFlagAcc: the result is added to what acc_out holds (the residual stream),
// 0: acc_out is overwritten.  acc_out must not be one of the hidden accumulators.  Result complete on return.
// NOTE the schedule of weight half-units must list the units in this issue order.
template <int ACT, int N>  // 1 gelu(tanh), 2 relu
__device__ __forceinline__ void chunked_mlp(Groups& G, const Epi& e, uint32_t X_addr, uint32_t x_flags, uint32_t accH0, uint32_t accH1,
                                            uint32_t acc_out, uint32_t acc_flags, int chunks, const float (&cv)[N], int b1) {
  G.issue(X_addr, accH0, x_flags);  // x_flags: kFlagTmemA if the input image set is in TMEM, else 0
  if (chunks > 1) G.issue(X_addr, accH1, x_flags);
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    const int b = b1 + c * 128 + e.half * 64;
    const uint32_t accH = (c & 1) ? accH1 : accH0;
    GPROF(0);
    if (c >= 2) G.wait_next();  // u_{c-2} . W2[c-2, :] (long done: it precedes F1(c) in the pipe)
    G.wait_next();              // X . W1[:, c]
    GPROF(1);
    float v[2][32];
    tmem_ld_half(e, accH, v);
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = v[q][j] + cv[b + q * 32 + j];
        v[q][j] = ACT == 1 ? gelu_tanh(t) : fmaxf(t, 0.f);
      }
    GPROF(2);
    split_store_tmem_half(e, accH, v);
    GPROF(3);
    publish_tmem();
    G.issue(accH, acc_out, kFlagTmemA | (c > 0 ? kFlagAcc : acc_flags));
    if (c + 2 < chunks) G.issue(X_addr, accH, x_flags);
    GPROF(4);
  }
  for (int c = chunks > 2 ? chunks - 2 : 0; c < chunks; ++c) G.wait_next();  // the last second-layer products
  GPROF(5);
}

// ------------------------------------------------------------------------------------------------------------
struct NodeUpdateParams {
  const float* partial;  // [num_edge_tiles][4][128]: per-receiver partial row sums written by the message-mode edge kernel
  int tile_shift;        // log2 of the edge tile the partial sums are taken over (7: edge_mlp_tc_kernel, 6: edge_msg_t_kernel)
  int K;
  float* h;           // [R,128] in / out
  __half* h16;        // [R,128] out (optional): fp16 copy of the new node state, gathered by the transposed message kernel
  const uint8_t* const* sched;
  int n_sched, n_out;
  __half* out[4];
  // parameter vectors by value (constant bank), offsets in floats
  enum { kB3 = 0, kLn0S = 128, kLn0O = 256, kFfnB1 = 384, kFfnB2 = 896, kLn1S = 1024, kLn1O = 1152, kOutBias = 1280, kCv = 1792 };
  float cv[kCv];      // out_bias[o] at kOutBias + 128 o (zeros where the table has no bias)
  int R, num_tiles;
  uint32_t idesc;
};

__global__ void __launch_bounds__(kThreads, 1) node_update_kernel(const __grid_constant__ NodeUpdateParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  Setup S = chain_setup(smem, tid, warp);
  const int my_tiles = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;  // the same for every CTA (lock-step weight ring)
  if (warp >= kEpiThreads / 32) {
    if (tid == kEpiThreads) producer_loop(S.ring, p.sched, p.n_sched, my_tiles);
    if (warp == kEpiThreads / 32 + 1) mma_loop(S.ring, S.cmd_bar, S.cmd, p.idesc, my_tiles * (p.n_sched / 2));
    __syncwarp();
    chain_teardown(S, warp);
    return;
  }
  Epi e{tid, tid & 127, tid >> 7, (uint32_t)((warp & 3) * 32) << 16, S.red};
  const uint32_t t_acc0 = S.tmem_base + 0, t_x = S.tmem_base + 128, t_acc1 = S.tmem_base + 256, t_acc2 = S.tmem_base + 384;
  // t_x: the hi / lo images of the current activation (A operand of every first-layer product, read from TMEM: with
  // both operands in shared memory a split product moves 192 KB of operands + 64 KB of incoming weights through the
  // 128 B/clock shared-memory port = 2 000+ cycles for 1 536 cycles of tensor work); it doubles as the row's h1.
  Groups G{S.cmd_bar, S.cmd, {S.bar_done[0], S.bar_done[1], S.bar_done[2], S.bar_done[3]}, 0, 0, tid};

  NPROF_DECL;
  for (int ti = 0, tile = blockIdx.x; ti < my_tiles; ++ti, tile += gridDim.x) {  // tiles past the end: all rows invalid
    const int row = tile * 128 + e.row;
    const bool valid = row < p.R;
    float x[2][32];
    NPROF(0);
    // ---- 1. agg = tbar . W3 -------------------------------------------------------------------------------
    // tbar[row] = (sum of the partial row sums of the one or two 128-edge tiles that hold the row's K edges) / K
    {
      const int e0 = row * p.K, t0 = e0 >> p.tile_shift, t1 = (e0 + p.K - 1) >> p.tile_shift;
      load_row_half(e, p.partial + ((size_t)t0 * 4 + (row - (t0 << p.tile_shift) / p.K)) * D, valid, x);
      if (valid && t1 != t0) {
        float y[2][32];
        load_row_half(e, p.partial + ((size_t)t1 * 4 + (row - (t1 << p.tile_shift) / p.K)) * D, true, y);
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 32; ++j) x[q][j] += y[q][j];
      }
      const float kf = (float)p.K;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
#ifdef PST_NODE_RECIP_MUL  // reciprocal multiply instead of IEEE division (see the note at tanh_fast)
        for (int j = 0; j < 32; ++j) x[q][j] = x[q][j] * (1.0f / kf);
#else
        for (int j = 0; j < 32; ++j) x[q][j] = x[q][j] / kf;
#endif
    }
    NPROF(1);  // partial sums loaded + scaled
    split_store_tmem_half(e, t_x, x);
    NPROF(7);  // (profile only) image store of the partial sums
    publish_tmem();
    NPROF(0);  // (profile only) publish
    G.issue(t_x, t_acc0, kFlagTmemA);
    // ---- 2. h1 = LN0(h + agg + b3) -> images in t_x ------------------------------------------------------------
    {
      float hh[2][32];
      load_row_half(e, p.h + (size_t)row * D, valid, hh);  // in flight while the product runs
      G.wait_next();
      NPROF(2);  // product 1 (issue + wait)
      tmem_ld_half(e, t_acc0, x);
      add_residual_bias_half(e, x, hh, p.cv, NodeUpdateParams::kB3);
    }
    layer_norm_row(e, x, p.cv, NodeUpdateParams::kLn0S, NodeUpdateParams::kLn0O);
    split_store_tmem_half(e, t_x, x);
    publish_tmem();
    NPROF(3);  // LN0
    // the next tile's partial sums and node state -> L2 while the tensor pipe works on the FFN: every CTA is in the
    // same phase at the same time, so without this the whole grid waits on HBM at the top of each tile
    if (ti + 1 < my_tiles) {
      const int nrow = row + (int)gridDim.x * 128;
      if (nrow < p.R) {
        const int e0 = nrow * p.K, t0 = e0 >> p.tile_shift, t1 = (e0 + p.K - 1) >> p.tile_shift;
        const float* a0 = p.partial + ((size_t)t0 * 4 + (nrow - (t0 << p.tile_shift) / p.K)) * D + e.half * 64;
        const float* a1 = p.partial + ((size_t)t1 * 4 + (nrow - (t1 << p.tile_shift) / p.K)) * D + e.half * 64;
        const float* a2 = p.h + (size_t)nrow * D + e.half * 64;
        prefetch_l2(a0); prefetch_l2(a0 + 32);
        if (t1 != t0) { prefetch_l2(a1); prefetch_l2(a1 + 32); }
        prefetch_l2(a2); prefetch_l2(a2 + 32);
      }
    }
    // ---- 3. FFN 128 -> 512 -> 128 (gnn_layers.py:385-394), hidden chunked 4 x 128 ---------------------------------
    chunked_mlp<1>(G, e, t_x, kFlagTmemA, t_acc1, t_acc0, t_acc2, 0u, 4, p.cv, NodeUpdateParams::kFfnB1);
    NPROF(4);  // FFN (8 products)
    // ---- 4. h2 = LN1(h1 + ffn + b2) -> global h, X images ------------------------------------------------------
    {
      float h1[2][32];
      tmem_ld_half(e, t_acc2, x);
      load_image_tmem_half(e, t_x, h1);
      add_residual_bias_half(e, x, h1, p.cv, NodeUpdateParams::kFfnB2);
    }
    layer_norm_row(e, x, p.cv, NodeUpdateParams::kLn1S, NodeUpdateParams::kLn1O);
    if (valid) {
      float* hdst = p.h + (size_t)row * D + e.half * 64;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) st256f(hdst + q * 32 + j * 8, &x[q][j * 8]);
      if (p.h16) {
        __half* h16dst = p.h16 + (size_t)row * D + e.half * 64;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t pk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const __half2 v2 = __floats2half2_rn(x[q][j * 16 + 2 * k], x[q][j * 16 + 2 * k + 1]);
              pk[k] = *reinterpret_cast<const uint32_t*>(&v2);
            }
            st256u(h16dst + q * 32 + j * 16, pk);
          }
      }
    }
    NPROF(5);  // LN1 + stores of h
    if (p.n_out > 0) {
      split_store_tmem_half(e, t_x, x);
      publish_tmem();
      // ---- 5. the gathered addend tables of the next edge-level kernels: fp16(h2 . Wout_o + b_o) ---------------
      // three accumulators are free here (the FFN's): three products are posted up front and run back to back on the
      // tensor pipe while the earlier ones are converted and stored; the fourth reuses the first accumulator.
      const uint32_t t_out[4] = {t_acc0, t_acc1, t_acc2, t_acc0};
#pragma unroll
      for (int o = 0; o < 3; ++o)
        if (o < p.n_out) G.issue(t_x, t_out[o], kFlagTmemA);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        if (o >= p.n_out) break;
        G.wait_next();
        NPROF(6);  // tables: waiting for the product
        tmem_ld_half(e, t_out[o], x);
        add_bias_half(e, x, p.cv, NodeUpdateParams::kOutBias + o * 128);
        if (valid) {
          __half* dst = p.out[o] + (size_t)row * D + e.half * 64;
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint32_t pk[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const __half2 v2 = __floats2half2_rn(x[q][j * 16 + 2 * k], x[q][j * 16 + 2 * k + 1]);
                pk[k] = *reinterpret_cast<const uint32_t*>(&v2);
              }
              st256u(dst + q * 32 + j * 16, pk);
            }
        }
        if (o == 0 && p.n_out > 3) {  // every thread has read the first accumulator: it takes the fourth product
          tc_before();
          epi_sync();
          G.issue(t_x, t_out[3], kFlagTmemA);
        }
        NPROF(7);  // tables: epilogue
      }
      tc_before();
      epi_sync();
    } else {
      tc_before();
      epi_sync();
    }
    NPROF(6);
  }
  NPROF_PRINT("node_update [-, partials, product 1, LN0, FFN, LN1, tables: wait, tables: epilogue]", 0);
#ifdef PST_NODE_PROFILE
  if (blockIdx.x == 0 && (tid == 0 || tid == 160))
    printf("  FFN thread %d [-, waits, load + GELU, image store, publish + post, tail wait]: %lld %lld %lld %lld %lld %lld\n", tid, G.pf[0], G.pf[1],
           G.pf[2], G.pf[3], G.pf[4], G.pf[5]);
#endif
  chain_teardown(S, warp);
}

// ------------------------------------------------------------------------------------------------------------
// offsets (floats) of one block's vectors inside ResamplerParams::cv; block b starts at b * kBlk
struct RB {
  enum { kQnS = 0, kQnO = 128, kDnS = 256, kDnO = 384, kBg = 512, kBo = 640, kRtLnS = 768, kRtLnO = 896, kRtB1 = 1024,
         kRtB2 = 1280, kOtLnS = 1408, kOtLnO = 1536, kOtB1 = 1664, kOtB2 = 1920, kBlk = 2048 };
};
struct ResamplerParams {
  const float* h;            // [R,128] node features after the GNN ("original" track input)
  const float* token_table;  // [max_out_len,128] PE of the token index (modules.py:486-500)
  const int32_t* row_base;   // [R] first row of the structure a row belongs to (df = 1: token index == row)
  float* z;                  // [R,8] pre-quantisation latents, unused columns 0
  int32_t* tokens;           // optional [R]: the FSQ epilogue (bound, round, pack) of the fused tokenize call
  int32_t* status;           // the call's status word (PST_ERR_NON_FINITE), with `tokens`
  PstFsqParams fsq;
  int num_blocks;
  const float* down_w;       // [128,8]
  const float* down_b;       // [8]
  int C;
  const uint8_t* const* sched;
  int n_sched;
  int R, num_tiles;
  uint32_t idesc;
  enum { kMaxBlocks = 3 };  // every released config has 3 (config/structure_tokenizer/model/shared.yaml); 4 would not fit the 32 KB parameter space
  float cv[kMaxBlocks * RB::kBlk];  // parameter vectors by value (constant bank)
};
static_assert(sizeof(ResamplerParams) <= 32764, "kernel parameter space");

__global__ void __launch_bounds__(kThreads, 1) resampler_df1_kernel(const __grid_constant__ ResamplerParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  Setup S = chain_setup(smem, tid, warp);
  const int my_tiles = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;  // the same for every CTA (lock-step weight ring)
  if (warp >= kEpiThreads / 32) {
    if (tid == kEpiThreads) producer_loop(S.ring, p.sched, p.n_sched, my_tiles);
    if (warp == kEpiThreads / 32 + 1) mma_loop(S.ring, S.cmd_bar, S.cmd, p.idesc, my_tiles * (p.n_sched / 2));
    __syncwarp();
    chain_teardown(S, warp);
    return;
  }
  uint8_t* X = smem + kOffX;
  uint8_t* U = smem + kOffU;
  const uint32_t X_addr = smem_u32(X), U_addr = smem_u32(U);
  Epi e{tid, tid & 127, tid >> 7, (uint32_t)((warp & 3) * 32) << 16, S.red};
  const uint32_t t_res = S.tmem_base + 0, t_orig = S.tmem_base + 128, t_accA = S.tmem_base + 256, t_accB = S.tmem_base + 384;
  Groups G{S.cmd_bar, S.cmd, {S.bar_done[0], S.bar_done[1], S.bar_done[2], S.bar_done[3]}, 0, 0, tid};

  for (int ti = 0, tile = blockIdx.x; ti < my_tiles; ++ti, tile += gridDim.x) {  // tiles past the end: all rows invalid
    const int row = tile * 128 + e.row;
    const bool valid = row < p.R;
    float x[2][32];
    {
      const int local = valid ? row - __ldg(p.row_base + row) : 0;
      load_row_half(e, p.token_table + (size_t)local * D, valid, x);
      tmem_st_half(e, t_res, x);
      float y[2][32];
      load_row_half(e, p.h + (size_t)row * D, valid, y);
      tmem_st_half(e, t_orig, y);
    }
#pragma unroll 1
    for (int b = 0; b < p.num_blocks; ++b) {
      const int w = b * RB::kBlk;
      // ---- cross attention, df = 1 (modules.py:303-380,407-424): res += (v * sigmoid(gate)) . Wo + bo ------------
      tmem_ld_half(e, t_res, x);
      layer_norm_row(e, x, p.cv, w + RB::kQnS, w + RB::kQnO);
      split_store_half(e, X, x);
      tmem_ld_half(e, t_orig, x);
      layer_norm_row(e, x, p.cv, w + RB::kDnS, w + RB::kDnO);
      split_store_half(e, U, x);
      publish();
      G.issue(X_addr, t_accA, 0u);  // gate = LNq(res) . Wg
      G.issue(U_addr, t_accB, 0u);  // v    = LNd(orig) . Wv
      G.wait_next();
      G.wait_next();
      {
        float g[2][32];
        tmem_ld_half(e, t_accA, g);
        tmem_ld_half(e, t_accB, x);
        const int bg = w + RB::kBg + e.half * 64;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 32; ++j) x[q][j] *= sigmoid_f(g[q][j] + p.cv[bg + q * 32 + j]);
      }
      split_store_half(e, X, x);
      publish();
      G.issue(X_addr, t_accA, 0u);
      G.wait_next();
      {
        float r[2][32];
        tmem_ld_half(e, t_accA, x);
        tmem_ld_half(e, t_res, r);
        add_residual_bias_half(e, x, r, p.cv, w + RB::kBo);
      }
      tmem_st_half(e, t_res, x);
      // ---- resampled transition (modules.py:227-252): res += W2 . relu(W1 . LN(res) + b1) + b2 -----------------------
      layer_norm_row(e, x, p.cv, w + RB::kRtLnS, w + RB::kRtLnO);
      split_store_half(e, X, x);
      publish();
      chunked_mlp<2>(G, e, X_addr, 0u, t_accA, t_accB, t_res, kFlagAcc, 2, p.cv, w + RB::kRtB1);  // accumulated onto res in TMEM
      tmem_ld_half(e, t_res, x);
      add_bias_half(e, x, p.cv, w + RB::kRtB2);
      tmem_st_half(e, t_res, x);
      // ---- original transition; its result is never read after the last block (modules.py:624-629) ---------------
      if (b < p.num_blocks - 1) {
        tmem_ld_half(e, t_orig, x);
        layer_norm_row(e, x, p.cv, w + RB::kOtLnS, w + RB::kOtLnO);
        split_store_half(e, X, x);
        publish();
        chunked_mlp<2>(G, e, X_addr, 0u, t_accA, t_accB, t_orig, kFlagAcc, 2, p.cv, w + RB::kOtB1);
        tmem_ld_half(e, t_orig, x);
        add_bias_half(e, x, p.cv, w + RB::kOtB2);
        tmem_st_half(e, t_orig, x);
      }
    }
    // ---- head (model.py:169-174,148-164): z = (res / (||res|| + 1e-6)) . Wd + bd -------------------------------------
    {
      tmem_ld_half(e, t_res, x);
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 32; ++j) ss = fmaf(x[q][j], x[q][j], ss);
      e.red[e.half * 128 + e.row] = ss;
      epi_sync();
      const float d = sqrtf(e.red[e.row] + e.red[128 + e.row]) + 1e-6f;
      float zc[PST_C8];
#pragma unroll
      for (int c = 0; c < PST_C8; ++c) zc[c] = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
#ifdef PST_NODE_RECIP_MUL
          const float r = x[q][j] * (1.0f / d);
#else
          const float r = x[q][j] / d;
#endif
          const float4* wr = reinterpret_cast<const float4*>(p.down_w + (size_t)(e.half * 64 + q * 32 + j) * PST_C8);
          const float4 w0 = __ldg(wr), w1 = __ldg(wr + 1);
          zc[0] = fmaf(r, w0.x, zc[0]); zc[1] = fmaf(r, w0.y, zc[1]); zc[2] = fmaf(r, w0.z, zc[2]); zc[3] = fmaf(r, w0.w, zc[3]);
          zc[4] = fmaf(r, w1.x, zc[4]); zc[5] = fmaf(r, w1.y, zc[5]); zc[6] = fmaf(r, w1.z, zc[6]); zc[7] = fmaf(r, w1.w, zc[7]);
        }
      float* zs = reinterpret_cast<float*>(U);  // U is free: [128 rows][8]
      if (e.half == 1) {
        *reinterpret_cast<float4*>(zs + e.row * 8) = make_float4(zc[0], zc[1], zc[2], zc[3]);
        *reinterpret_cast<float4*>(zs + e.row * 8 + 4) = make_float4(zc[4], zc[5], zc[6], zc[7]);
      }
      epi_sync();
      if (e.half == 0 && valid) {
        const float4 a = *reinterpret_cast<const float4*>(zs + e.row * 8), b = *reinterpret_cast<const float4*>(zs + e.row * 8 + 4);
        float o[PST_C8] = {zc[0] + a.x, zc[1] + a.y, zc[2] + a.z, zc[3] + a.w, zc[4] + b.x, zc[5] + b.y, zc[6] + b.z, zc[7] + b.w};
#pragma unroll
        for (int c = 0; c < PST_C8; ++c) o[c] = c < p.C ? o[c] + __ldg(p.down_b + c) : 0.f;
        float4* zd = reinterpret_cast<float4*>(p.z + (size_t)row * PST_C8);
        zd[0] = make_float4(o[0], o[1], o[2], o[3]);
        zd[1] = make_float4(o[4], o[5], o[6], o[7]);
        if (p.tokens) {  // quantiser as the epilogue of the chain (model/quantize.py:175-209): int32 token ids
          bool finite = true;
          p.tokens[row] = pst_fsq_token(o, p.fsq, finite);
          if (!finite) atomicMin(p.status, (int)PST_ERR_NON_FINITE);
        }
      }
      tc_before();
      epi_sync();  // zs (in U) and the TMEM state are rewritten by the next tile
    }
  }
  chain_teardown(S, warp);
}

// ------------------------------------------------------------------------------------------------------------
// Resampler for downsampling_ratio > 1 (round 2; reference: model/modules.py:438-636, masks model/model.py:264-318):
// token t of a structure attends its residues t*df .. t*df+df-1.  The residue ("original") track does not depend on
// the token track, so the three blocks split into TWO chain kernels instead of ~50 launches with [R,128] / [T,128]
// intermediates bouncing through HBM:
//   resampler_orig_kernel   per 128-residue tile, for every block b:  k_b = LN_d(orig).Wk, v_b = LN_d(orig).Wv -> HBM,
//                           orig += Transition(orig)  (not after the last block: modules.py:624-629 result unused)
//   resampler_token_kernel  per 128-token tile, for every block b:  q = LN_q(res).Wq * 32^-1/2, gate = LN_q(res).Wg + bg,
//                           a = softmax_j(q . k_b[row0 + j]) . v_b[row0 + j]  per head (thread = token row x 2 heads:
//                           thread-local), res += (a * sigmoid(gate)).Wo + bo, res += Transition(res);  then the head
//                           (spherical norm + down_proj) as in the df = 1 kernel.
struct OB {  // per-block parameter vectors of the residue track
  enum { kDnS = 0, kDnO = 128, kOtLnS = 256, kOtLnO = 384, kOtB1 = 512, kOtB2 = 768, kBlk = 896 };
};
struct ResamplerOrigParams {
  const float* h;            // [R,128] node features after the GNN
  float* k[3];               // per block [R,128]
  float* v[3];
  int num_blocks;
  const uint8_t* const* sched;
  int n_sched;
  int R, num_tiles;
  uint32_t idesc;
  float cv[3 * OB::kBlk];
};

__global__ void __launch_bounds__(kThreads, 1) resampler_orig_kernel(const __grid_constant__ ResamplerOrigParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  Setup S = chain_setup(smem, tid, warp);
  const int my_tiles = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;  // the same for every CTA (lock-step weight ring)
  if (warp >= kEpiThreads / 32) {
    if (tid == kEpiThreads) producer_loop(S.ring, p.sched, p.n_sched, my_tiles);
    if (warp == kEpiThreads / 32 + 1) mma_loop(S.ring, S.cmd_bar, S.cmd, p.idesc, my_tiles * (p.n_sched / 2));
    __syncwarp();
    chain_teardown(S, warp);
    return;
  }
  uint8_t* X = smem + kOffX;
  uint8_t* U = smem + kOffU;
  const uint32_t X_addr = smem_u32(X), U_addr = smem_u32(U);
  Epi e{tid, tid & 127, tid >> 7, (uint32_t)((warp & 3) * 32) << 16, S.red};
  const uint32_t t_orig = S.tmem_base + 0, t_accA = S.tmem_base + 128, t_accB = S.tmem_base + 256, t_accC = S.tmem_base + 384;
  Groups G{S.cmd_bar, S.cmd, {S.bar_done[0], S.bar_done[1], S.bar_done[2], S.bar_done[3]}, 0, 0, tid};

  for (int ti = 0, tile = blockIdx.x; ti < my_tiles; ++ti, tile += gridDim.x) {  // tiles past the end: all rows invalid
    const int row = tile * 128 + e.row;
    const bool valid = row < p.R;
    float x[2][32];
    load_row_half(e, p.h + (size_t)row * D, valid, x);
    tmem_st_half(e, t_orig, x);
#pragma unroll 1
    for (int b = 0; b < p.num_blocks; ++b) {
      const int w = b * OB::kBlk;
      if (b > 0) tmem_ld_half(e, t_orig, x);
      layer_norm_row(e, x, p.cv, w + OB::kDnS, w + OB::kDnO);
      split_store_half(e, X, x);
      publish();
      G.issue(X_addr, t_accA, 0u);  // k = LNd(orig) . Wk
      G.issue(X_addr, t_accB, 0u);  // v = LNd(orig) . Wv
#pragma unroll 1
      for (int kv = 0; kv < 2; ++kv) {
        G.wait_next();
        tmem_ld_half(e, kv ? t_accB : t_accA, x);
        if (valid) {
          float* dst = (kv ? p.v[b] : p.k[b]) + (size_t)row * D + e.half * 64;
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) st256f(dst + q * 32 + j * 8, &x[q][j * 8]);
        }
      }
      if (b < p.num_blocks - 1) {  // original transition (modules.py:227-252)
        tmem_ld_half(e, t_orig, x);
        layer_norm_row(e, x, p.cv, w + OB::kOtLnS, w + OB::kOtLnO);
        split_store_half(e, X, x);
        publish();
        chunked_mlp<2>(G, e, X_addr, 0u, t_accA, t_accB, t_accC, 0u, 2, p.cv, w + OB::kOtB1);
        float r[2][32];
        tmem_ld_half(e, t_accC, x);
        tmem_ld_half(e, t_orig, r);
        add_residual_bias_half(e, x, r, p.cv, w + OB::kOtB2);
        tmem_st_half(e, t_orig, x);
      }
    }
    tc_before();
    epi_sync();  // the TMEM state is rewritten by the next tile
  }
  chain_teardown(S, warp);
}

struct TB {  // per-block parameter vectors of the token track
  enum { kQnS = 0, kQnO = 128, kBg = 256, kBo = 384, kRtLnS = 512, kRtLnO = 640, kRtB1 = 768, kRtB2 = 1024, kBlk = 1152 };
};
struct ResamplerTokenParams {
  const float* token_table;  // [max_out_len,128] PE of the token index (modules.py:486-500)
  const int2* info;          // per token: (index inside its structure, first residue row it attends)
  const float* k[3];         // per block [R,128]
  const float* v[3];
  float* z;                  // [T,8]
  int32_t* tokens;           // optional [T]: the FSQ epilogue of the fused tokenize call
  int32_t* status;
  PstFsqParams fsq;
  int num_blocks, df;
  const float* down_w;
  const float* down_b;
  int C;
  const uint8_t* const* sched;
  int n_sched;
  int T, num_tiles;
  uint32_t idesc;
  float cv[3 * TB::kBlk];
};
static_assert(sizeof(ResamplerTokenParams) <= 32764, "kernel parameter space");

__global__ void __launch_bounds__(kThreads, 1) resampler_token_kernel(const __grid_constant__ ResamplerTokenParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  Setup S = chain_setup(smem, tid, warp);
  const int my_tiles = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp >= kEpiThreads / 32) {
    if (tid == kEpiThreads) producer_loop(S.ring, p.sched, p.n_sched, my_tiles);
    if (warp == kEpiThreads / 32 + 1) mma_loop(S.ring, S.cmd_bar, S.cmd, p.idesc, my_tiles * (p.n_sched / 2));
    __syncwarp();
    chain_teardown(S, warp);
    return;
  }
  uint8_t* X = smem + kOffX;
  uint8_t* U = smem + kOffU;
  const uint32_t X_addr = smem_u32(X), U_addr = smem_u32(U);
  Epi e{tid, tid & 127, tid >> 7, (uint32_t)((warp & 3) * 32) << 16, S.red};
  const uint32_t t_res = S.tmem_base + 0, t_accA = S.tmem_base + 128, t_accB = S.tmem_base + 256, t_accC = S.tmem_base + 384;
  Groups G{S.cmd_bar, S.cmd, {S.bar_done[0], S.bar_done[1], S.bar_done[2], S.bar_done[3]}, 0, 0, tid};
  const float qscale = 0.17677669529663687f;  // 32 ** -0.5 (modules.py:334)

  for (int ti = 0, tile = blockIdx.x; ti < my_tiles; ++ti, tile += gridDim.x) {
    const int t = tile * 128 + e.row;
    const bool valid = t < p.T;
    const int2 inf = valid ? __ldg(p.info + t) : make_int2(0, 0);
    float x[2][32];
    load_row_half(e, p.token_table + (size_t)inf.x * D, valid, x);
    tmem_st_half(e, t_res, x);
#pragma unroll 1
    for (int b = 0; b < p.num_blocks; ++b) {
      const int w = b * TB::kBlk;
      if (b > 0) tmem_ld_half(e, t_res, x);
      layer_norm_row(e, x, p.cv, w + TB::kQnS, w + TB::kQnO);
      split_store_half(e, X, x);
      publish();
      G.issue(X_addr, t_accA, 0u);  // q    = LNq(res) . Wq
      G.issue(X_addr, t_accB, 0u);  // gate = LNq(res) . Wg
      G.wait_next();
      G.wait_next();
      // ---- local attention (modules.py:303-380; model/model.py:264-318: token t sees its own df residues): this thread
      // holds heads 2 half, 2 half + 1 of its token (32 dims each); softmax over df logits per head, all in registers
      {
        tmem_ld_half(e, t_accA, x);
        const float* kb = p.k[b] + (size_t)inf.y * D + e.half * 64;
        const float* vb = p.v[b] + (size_t)inf.y * D + e.half * 64;
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          // df <= 8: the loops are unrolled over 8 with a predicate so that the logits stay in registers
          float logit[8];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            logit[j] = -INFINITY;
            if (j < p.df) {
              float acc = 0.f;
              if (valid) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  float kr[8];
                  ld256f(kb + (size_t)j * D + hq * 32 + c * 8, kr);
#pragma unroll
                  for (int i = 0; i < 8; ++i) acc = fmaf(x[hq][c * 8 + i] * qscale, kr[i], acc);
                }
              }
              logit[j] = acc;
              mx = fmaxf(mx, acc);
            }
          }
          float den = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            logit[j] = j < p.df ? expf(logit[j] - mx) : 0.f;
            den += logit[j];
          }
          float a[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < p.df && valid) {
              const float pj = logit[j] / den;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float vr[8];
                ld256f(vb + (size_t)j * D + hq * 32 + c * 8, vr);
#pragma unroll
                for (int i = 0; i < 8; ++i) a[c * 8 + i] = fmaf(pj, vr[i], a[c * 8 + i]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) x[hq][i] = a[i];
        }
        float g[2][32];
        tmem_ld_half(e, t_accB, g);
        const int bg = w + TB::kBg + e.half * 64;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int j = 0; j < 32; ++j) x[q][j] *= sigmoid_f(g[q][j] + p.cv[bg + q * 32 + j]);
      }
      split_store_half(e, X, x);
      publish();
      G.issue(X_addr, t_accA, 0u);  // . Wo
      G.wait_next();
      {
        float r[2][32];
        tmem_ld_half(e, t_accA, x);
        tmem_ld_half(e, t_res, r);
        add_residual_bias_half(e, x, r, p.cv, w + TB::kBo);
      }
      tmem_st_half(e, t_res, x);
      // ---- resampled transition (modules.py:227-252) ------------------------------------------------------------
      layer_norm_row(e, x, p.cv, w + TB::kRtLnS, w + TB::kRtLnO);
      split_store_half(e, X, x);
      publish();
      chunked_mlp<2>(G, e, X_addr, 0u, t_accA, t_accB, t_accC, 0u, 2, p.cv, w + TB::kRtB1);
      {
        float r[2][32];
        tmem_ld_half(e, t_accC, x);
        tmem_ld_half(e, t_res, r);
        add_residual_bias_half(e, x, r, p.cv, w + TB::kRtB2);
      }
      tmem_st_half(e, t_res, x);
    }
    // ---- head (model.py:169-174,148-164): z = (res / (||res|| + 1e-6)) . Wd + bd -------------------------------------
    {
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 32; ++j) ss = fmaf(x[q][j], x[q][j], ss);
      e.red[e.half * 128 + e.row] = ss;
      epi_sync();
      const float d = sqrtf(e.red[e.row] + e.red[128 + e.row]) + 1e-6f;
      float zc[PST_C8];
#pragma unroll
      for (int c = 0; c < PST_C8; ++c) zc[c] = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
#ifdef PST_NODE_RECIP_MUL
          const float r = x[q][j] * (1.0f / d);
#else
          const float r = x[q][j] / d;
#endif
          const float4* wr = reinterpret_cast<const float4*>(p.down_w + (size_t)(e.half * 64 + q * 32 + j) * PST_C8);
          const float4 w0 = __ldg(wr), w1 = __ldg(wr + 1);
          zc[0] = fmaf(r, w0.x, zc[0]); zc[1] = fmaf(r, w0.y, zc[1]); zc[2] = fmaf(r, w0.z, zc[2]); zc[3] = fmaf(r, w0.w, zc[3]);
          zc[4] = fmaf(r, w1.x, zc[4]); zc[5] = fmaf(r, w1.y, zc[5]); zc[6] = fmaf(r, w1.z, zc[6]); zc[7] = fmaf(r, w1.w, zc[7]);
        }
      float* zs = reinterpret_cast<float*>(U);  // U is free: [128 rows][8]
      if (e.half == 1) {
        *reinterpret_cast<float4*>(zs + e.row * 8) = make_float4(zc[0], zc[1], zc[2], zc[3]);
        *reinterpret_cast<float4*>(zs + e.row * 8 + 4) = make_float4(zc[4], zc[5], zc[6], zc[7]);
      }
      epi_sync();
      if (e.half == 0 && valid) {
        const float4 a = *reinterpret_cast<const float4*>(zs + e.row * 8), bq = *reinterpret_cast<const float4*>(zs + e.row * 8 + 4);
        float o[PST_C8] = {zc[0] + a.x, zc[1] + a.y, zc[2] + a.z, zc[3] + a.w, zc[4] + bq.x, zc[5] + bq.y, zc[6] + bq.z, zc[7] + bq.w};
#pragma unroll
        for (int c = 0; c < PST_C8; ++c) o[c] = c < p.C ? o[c] + __ldg(p.down_b + c) : 0.f;
        float4* zd = reinterpret_cast<float4*>(p.z + (size_t)t * PST_C8);
        zd[0] = make_float4(o[0], o[1], o[2], o[3]);
        zd[1] = make_float4(o[4], o[5], o[6], o[7]);
        if (p.tokens) {  // quantiser as the epilogue of the chain
          bool finite = true;
          p.tokens[t] = pst_fsq_token(o, p.fsq, finite);
          if (!finite) atomicMin(p.status, (int)PST_ERR_NON_FINITE);
        }
      }
      tc_before();
      epi_sync();  // zs (in U) and the TMEM state are rewritten by the next tile
    }
  }
  chain_teardown(S, warp);
}

// per token: its index inside its structure and the first residue row it attends (offsets[b] + local * df)
__global__ void token_info_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ token_offsets, int B, int df, int T,
                                  int2* __restrict__ info) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int lo = 0, hi = B;  // largest b with token_offsets[b] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (token_offsets[mid] <= t) lo = mid; else hi = mid;
  }
  const int local = t - token_offsets[lo];
  info[t] = make_int2(local, offsets[lo] + local * df);
}

struct Entry {
  const float* w;
  int K, N;
  const uint8_t* img;
};

}  // namespace

// linear_tc.cu: image of a registered weight ([K/64][N/128][hi,lo][16 KB]) or null
const uint8_t* pst_linear_tc_image(const pst_model* m, const float* W, int K, int N);

struct PstNodeChain {
  const uint8_t** sched_dev = nullptr;  // all schedules, concatenated
  int layer_off[PST_MAX_LAYERS], layer_n[PST_MAX_LAYERS], layer_nout[PST_MAX_LAYERS];
  int resampler_off = 0, resampler_n = 0;
  int orig_off = 0, orig_n = 0, token_off = 0, token_n = 0;  // df > 1: residue-track / token-track schedules
};

// persistent grid of whole clusters: one CTA per SM, at most one per tile
template <typename P>
static int launch_clustered(void (*kernel)(P), const pst_model* m, cudaStream_t st, int num_tiles, const P& p) {
  int grid = m->num_sms < num_tiles ? m->num_sms : num_tiles;
  grid = (grid + kCluster - 1) / kCluster * kCluster;
  if (grid > m->num_sms) grid = m->num_sms / kCluster * kCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p) == cudaSuccess ? 1 : PST_ERR_CUDA;
}

int pst_prepare_node_chain(pst_model* m) {
  m->node_chain = new PstNodeChain();
  PstNodeChain& C = *m->node_chain;
  std::vector<const uint8_t*> all;
  auto half_unit = [&](const float* W, int K, int N, int kb, int nc) -> const uint8_t* {
    const uint8_t* img = pst_linear_tc_image(m, W, K, N);
    if (!img) return nullptr;
    return img + ((size_t)(kb * (N / 128) + nc) * 2) * kImgBlk;
  };
  bool ok = true;
  auto push_unit = [&](const float* W, int K, int N, int kchunk, int nc) {  // 128(K) x 128(N) unit = 2 half-units
    for (int kb = 0; kb < 2; ++kb) {
      const uint8_t* p = half_unit(W, K, N, kchunk * 2 + kb, nc);
      ok = ok && p;
      all.push_back(p);
    }
  };
  // units of a chunked MLP in the issue order of chunked_mlp(): F1(0) F1(1) F2(0) F1(2) F2(1) ... F2(n-1)
  auto push_chunked = [&](const float* W1, const float* W2, int hidden, int chunks) {
    push_unit(W1, D, hidden, 0, 0);
    if (chunks > 1) push_unit(W1, D, hidden, 0, 1);
    for (int c = 0; c < chunks; ++c) {
      push_unit(W2, hidden, D, c, 0);
      if (c + 2 < chunks) push_unit(W1, D, hidden, 0, c + 2);
    }
  };
  const int L = m->cfg.gnn_layers;
  for (int l = 0; l < L; ++l) {
    const PstLayerW& w = m->w.layer[l];
    C.layer_off[l] = (int)all.size();
    push_unit(w.msg_w3, D, D, 0, 0);
    push_chunked(w.ffn_w1, w.ffn_w2, PST_FFN, 4);
    C.layer_nout[l] = 0;
    if (l < L - 1) {
      const PstLayerW& nx = m->w.layer[l + 1];
      push_unit(w.edge_w1, D, D, 0, 0);
      push_unit(w.edge_w1 + D * D, D, D, 0, 0);
      push_unit(nx.msg_w1 + D * D, D, D, 0, 0);
      push_unit(nx.msg_w1, D, D, 0, 0);  // last: not computed when the next message MLP gathers fp16(h) itself (edge_msg_t_kernel)
      C.layer_nout[l] = 4;
    }
    C.layer_n[l] = (int)all.size() - C.layer_off[l];
  }
  // resampler (df = 1): per block gate, value, output, resampled transition (2 chunks), original transition
  C.resampler_off = (int)all.size();
  for (int b = 0; b < m->cfg.num_blocks; ++b) {
    const PstBlockW& w = m->w.block[b];
    push_unit(w.wg, D, D, 0, 0);
    push_unit(w.wv, D, D, 0, 0);
    push_unit(w.wo, D, D, 0, 0);
    push_chunked(w.rt_w1, w.rt_w2, PST_TRANS, 2);
    if (b < m->cfg.num_blocks - 1) push_chunked(w.ot_w1, w.ot_w2, PST_TRANS, 2);
  }
  C.resampler_n = (int)all.size() - C.resampler_off;
  // resampler (df > 1): residue track (per block key, value, original transition) and token track (per block query,
  // gate, output, resampled transition), in the issue order of resampler_orig_kernel / resampler_token_kernel
  C.orig_off = (int)all.size();
  for (int b = 0; b < m->cfg.num_blocks; ++b) {
    const PstBlockW& w = m->w.block[b];
    push_unit(w.wk, D, D, 0, 0);
    push_unit(w.wv, D, D, 0, 0);
    if (b < m->cfg.num_blocks - 1) push_chunked(w.ot_w1, w.ot_w2, PST_TRANS, 2);
  }
  C.orig_n = (int)all.size() - C.orig_off;
  C.token_off = (int)all.size();
  for (int b = 0; b < m->cfg.num_blocks; ++b) {
    const PstBlockW& w = m->w.block[b];
    push_unit(w.wq, D, D, 0, 0);
    push_unit(w.wg, D, D, 0, 0);
    push_unit(w.wo, D, D, 0, 0);
    push_chunked(w.rt_w1, w.rt_w2, PST_TRANS, 2);
  }
  C.token_n = (int)all.size() - C.token_off;
  if (!ok) return PST_ERR_BAD_ARGUMENT;
  if (cudaMalloc(&C.sched_dev, all.size() * sizeof(void*)) != cudaSuccess) return PST_ERR_CUDA;
  if (cudaMemcpy(C.sched_dev, all.data(), all.size() * sizeof(void*), cudaMemcpyHostToDevice) != cudaSuccess) return PST_ERR_CUDA;
  if (cudaFuncSetAttribute(node_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(resampler_df1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(resampler_orig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(resampler_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes) != cudaSuccess)
    return PST_ERR_CUDA;
  return PST_OK;
}

void pst_destroy_node_chain(pst_model* m) {
  if (!m->node_chain) return;
  if (m->node_chain->sched_dev) cudaFree(m->node_chain->sched_dev);
  delete m->node_chain;
  m->node_chain = nullptr;
}

// h <- node update of MPNN layer `layer` (see the header); for layer < last also the four fp16 addend tables:
// out_edge_s/r for this layer's edge MLP, out_msg_s/r for the next layer's message MLP.
int pst_launch_node_update(const pst_model* m, cudaStream_t st, int layer, const float* partial, int partial_tile_shift, float* h, int R,
                           uint16_t* out_edge_s, uint16_t* out_edge_r, uint16_t* out_msg_s, uint16_t* out_msg_r, uint16_t* h16) {
  if (!m->node_chain || R <= 0) return 0;
  const PstNodeChain& C = *m->node_chain;
  const PstLayerW& w = m->w.layer[layer];
  NodeUpdateParams p{};
  p.partial = partial; p.tile_shift = partial_tile_shift; p.K = m->cfg.num_neighbor; p.h = h;
  p.h16 = reinterpret_cast<__half*>(h16);
  auto put = [&](int off, const float* dev, int n) {  // device pointer into the weight blob -> its host copy
    if (dev) memcpy(p.cv + off, m->blob_host + (dev - m->blob_dev), (size_t)n * sizeof(float));
  };
  put(NodeUpdateParams::kB3, w.msg_b3, 128); put(NodeUpdateParams::kLn0S, w.ln0_s, 128); put(NodeUpdateParams::kLn0O, w.ln0_o, 128);
  put(NodeUpdateParams::kFfnB1, w.ffn_b1, 512); put(NodeUpdateParams::kFfnB2, w.ffn_b2, 128);
  put(NodeUpdateParams::kLn1S, w.ln1_s, 128); put(NodeUpdateParams::kLn1O, w.ln1_o, 128);
  p.sched = C.sched_dev + C.layer_off[layer];
  p.n_sched = C.layer_n[layer];
  p.n_out = C.layer_nout[layer];
  if (p.n_out) {
    const PstLayerW& nx = m->w.layer[layer + 1];
    p.out[0] = reinterpret_cast<__half*>(out_edge_s);
    p.out[1] = reinterpret_cast<__half*>(out_edge_r); put(NodeUpdateParams::kOutBias + 128, w.edge_b1, 128);
    p.out[2] = reinterpret_cast<__half*>(out_msg_r);  put(NodeUpdateParams::kOutBias + 256, nx.msg_b1, 128);
    p.out[3] = reinterpret_cast<__half*>(out_msg_s);
    if (h16) {  // the transposed message kernel takes the sender term from h16: its table (the last one) is not needed
      p.n_out = 3;
      p.n_sched -= 2;
    }
  }
  p.R = R;
  p.num_tiles = (R + 127) / 128;
  p.idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  return launch_clustered(node_update_kernel, m, st, p.num_tiles, p);
}

// z <- the whole resampler + head for downsampling_ratio == 1 (token t == residue t); h = node features after the GNN
int pst_launch_resampler_df1(const pst_model* m, cudaStream_t st, const float* h, const int32_t* row_base, int R, float* z,
                             int32_t* tokens, int32_t* status) {
  if (!m->node_chain || R <= 0 || m->cfg.downsampling_ratio != 1) return 0;
  if (m->cfg.num_blocks > ResamplerParams::kMaxBlocks) return PST_ERR_UNSUPPORTED_CONFIG;
  const PstNodeChain& C = *m->node_chain;
  ResamplerParams p{};
  p.h = h; p.token_table = m->w.token_table; p.row_base = row_base; p.z = z;
  p.tokens = status ? tokens : nullptr; p.status = status; p.fsq = pst_fsq_params(m);
  p.num_blocks = m->cfg.num_blocks;
  for (int b = 0; b < p.num_blocks; ++b) {
    const PstBlockW& w = m->w.block[b];
    auto put = [&](int off, const float* dev, int n) {
      memcpy(p.cv + b * RB::kBlk + off, m->blob_host + (dev - m->blob_dev), (size_t)n * sizeof(float));
    };
    put(RB::kQnS, w.qn_s, 128); put(RB::kQnO, w.qn_o, 128); put(RB::kDnS, w.dn_s, 128); put(RB::kDnO, w.dn_o, 128);
    put(RB::kBg, w.bg, 128); put(RB::kBo, w.bo, 128);
    put(RB::kRtLnS, w.rt_ln_s, 128); put(RB::kRtLnO, w.rt_ln_o, 128); put(RB::kRtB1, w.rt_b1, 256); put(RB::kRtB2, w.rt_b2, 128);
    put(RB::kOtLnS, w.ot_ln_s, 128); put(RB::kOtLnO, w.ot_ln_o, 128); put(RB::kOtB1, w.ot_b1, 256); put(RB::kOtB2, w.ot_b2, 128);
  }
  p.down_w = m->w.down_w; p.down_b = m->w.down_b; p.C = m->cfg.num_levels;
  p.sched = C.sched_dev + C.resampler_off;
  p.n_sched = C.resampler_n;
  p.R = R;
  p.num_tiles = (R + 127) / 128;
  p.idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  return launch_clustered(resampler_df1_kernel, m, st, p.num_tiles, p);
}

// z <- the whole resampler + head for downsampling_ratio > 1: token-info lookup, residue track, token track.
// kv: six [R,128] fp32 buffers (k and v of the three blocks); info: int2 [T].
int pst_launch_resampler_dfn(const pst_model* m, cudaStream_t st, const float* h, const int32_t* offsets, const int32_t* token_offsets,
                             int B, int R, int T, float* const* kv, void* info, float* z, int32_t* tokens, int32_t* status) {
  if (!m->node_chain || R <= 0 || T <= 0 || m->cfg.downsampling_ratio <= 1) return 0;
  if (m->cfg.num_blocks > 3 || m->cfg.downsampling_ratio > 8) return PST_ERR_UNSUPPORTED_CONFIG;
  const PstNodeChain& C = *m->node_chain;
  const int nb = m->cfg.num_blocks;
  token_info_kernel<<<(T + 255) / 256, 256, 0, st>>>(offsets, token_offsets, B, m->cfg.downsampling_ratio, T, static_cast<int2*>(info));
  const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  auto host = [&](const float* dev) { return m->blob_host + (dev - m->blob_dev); };
  {
    ResamplerOrigParams p{};
    p.h = h; p.num_blocks = nb;
    for (int b = 0; b < nb; ++b) {
      const PstBlockW& w = m->w.block[b];
      p.k[b] = kv[2 * b]; p.v[b] = kv[2 * b + 1];
      float* c = p.cv + b * OB::kBlk;
      memcpy(c + OB::kDnS, host(w.dn_s), 512); memcpy(c + OB::kDnO, host(w.dn_o), 512);
      memcpy(c + OB::kOtLnS, host(w.ot_ln_s), 512); memcpy(c + OB::kOtLnO, host(w.ot_ln_o), 512);
      memcpy(c + OB::kOtB1, host(w.ot_b1), 1024); memcpy(c + OB::kOtB2, host(w.ot_b2), 512);
    }
    p.sched = C.sched_dev + C.orig_off; p.n_sched = C.orig_n;
    p.R = R; p.num_tiles = (R + 127) / 128; p.idesc = idesc;
    if (int rc = launch_clustered(resampler_orig_kernel, m, st, p.num_tiles, p); rc < 0) return rc;
  }
  {
    ResamplerTokenParams p{};
    p.token_table = m->w.token_table; p.info = static_cast<const int2*>(info); p.z = z;
    p.tokens = status ? tokens : nullptr; p.status = status; p.fsq = pst_fsq_params(m);
    p.num_blocks = nb; p.df = m->cfg.downsampling_ratio;
    for (int b = 0; b < nb; ++b) {
      const PstBlockW& w = m->w.block[b];
      p.k[b] = kv[2 * b]; p.v[b] = kv[2 * b + 1];
      float* c = p.cv + b * TB::kBlk;
      memcpy(c + TB::kQnS, host(w.qn_s), 512); memcpy(c + TB::kQnO, host(w.qn_o), 512);
      memcpy(c + TB::kBg, host(w.bg), 512); memcpy(c + TB::kBo, host(w.bo), 512);
      memcpy(c + TB::kRtLnS, host(w.rt_ln_s), 512); memcpy(c + TB::kRtLnO, host(w.rt_ln_o), 512);
      memcpy(c + TB::kRtB1, host(w.rt_b1), 1024); memcpy(c + TB::kRtB2, host(w.rt_b2), 512);
    }
    p.down_w = m->w.down_w; p.down_b = m->w.down_b; p.C = m->cfg.num_levels;
    p.sched = C.sched_dev + C.token_off; p.n_sched = C.token_n;
    p.T = T; p.num_tiles = (T + 127) / 128; p.idesc = idesc;
    if (int rc = launch_clustered(resampler_token_kernel, m, st, p.num_tiles, p); rc < 0) return rc;
  }
  return 3;
}
