// Featurisation + k-NN graph.  COMPILED WITH -fmad=false: every fp64
// operation below must round exactly like the reference's NumPy/SciPy host code
// (no FMA contraction), because the k-NN order is a bit-exact contract.
//
// Reference behaviour restated (paths relative to the reference repo):
//   frames     structure_tokenizer/model/quat_affine.py:406-522  (make_canonical_transform +
//              transpose), axes u,v,n = columns 0,1,2 (data/preprocessing.py:94-97)
//   centroid   structure_tokenizer/utils/protein_utils.py:373-378 (sequential fp64 mean of the
//              present atoms; mask = gt_exists & atom_exists, data/preprocessing.py:72)
//   distances  scipy cdist as called at protein_utils.py:380-383: sqrt((dx*dx+dy*dy)+dz*dz)
//   k-NN       protein_utils.py:385-399: ascending (distance, index); ranks 1..K, or 0..K-1
//              when L == K
//   features   protein_utils.py:257-281 (15 RBF) and :403-434 (p,q,k,t in the receiver frame,
//              basis rows [n,u,v], positions = CA), fp64 then one rounding to fp32
#include "pst_internal.h"

namespace {

// prep record: [0:3] centroid, [3:6] CA, [6:9] n, [9:12] u, [12:15] v
__global__ void prep_kernel(const float* __restrict__ atoms, const uint8_t* __restrict__ mask,
                            int apr, int R, double* __restrict__ prep, double* __restrict__ cen) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const float* a = atoms + (size_t)i * apr * 3;
  const uint8_t* m = mask ? mask + (size_t)i * apr : nullptr;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  int cnt = 0;
  for (int s = 0; s < apr; ++s) {
    bool present = m ? (m[s] != 0) : true;
    if (present) {
      sx = sx + (double)a[s * 3 + 0];
      sy = sy + (double)a[s * 3 + 1];
      sz = sz + (double)a[s * 3 + 2];
      ++cnt;
    }
  }
  double c = (double)cnt;
  double* o = prep + (size_t)i * PST_PREP_STRIDE;
  o[0] = sx / c;
  o[1] = sy / c;
  o[2] = sz / c;
  // structure-of-arrays copy (x[R], y[R], z[R]) for the k-NN scan: 32 consecutive candidates of one coordinate are
  // two 128-byte lines per warp load (an array of 32-byte records, two candidates per lane, cost 16 L1 tag requests
  // per load instruction and made the scan the largest consumer of the kernel's saturated L1 data pipe)
  cen[i] = o[0];
  cen[(size_t)R + i] = o[1];
  cen[2 * (size_t)R + i] = o[2];
  double nx = (double)a[0], ny_ = (double)a[1], nz_ = (double)a[2];
  double cax = (double)a[3], cay = (double)a[4], caz = (double)a[5];
  double ccx = (double)a[6], ccy = (double)a[7], ccz = (double)a[8];
  o[3] = cax;
  o[4] = cay;
  o[5] = caz;
  // translate so that CA is the origin
  double x = nx + (-cax), y = ny_ + (-cay), z = nz_ + (-caz);
  double cx = ccx + (-cax), cy = ccy + (-cay), cz = ccz + (-caz);
  double den_xy = sqrt(1e-20 + cx * cx + cy * cy);
  double s1 = -cy / den_xy, c1 = cx / den_xy;
  double den_xyz = sqrt(1e-20 + cx * cx + cy * cy + cz * cz);
  double s2 = cz / den_xyz;
  double c2 = sqrt(cx * cx + cy * cy) / den_xyz;
  // Rc = R2 . R1 (zero terms dropped: adding +-0 is exact)
  double rc00 = c2 * c1, rc01 = c2 * (-s1), rc02 = s2;
  double rc10 = s1, rc11 = c1, rc12 = 0.0;
  double rc20 = (-s2) * c1, rc21 = (-s2) * (-s1), rc22 = c2;
  double ry = rc10 * x + rc11 * y + rc12 * z;
  double rz = rc20 * x + rc21 * y + rc22 * z;
  double den_n = sqrt(1e-20 + ry * ry + rz * rz);
  double sn = -rz / den_n, cn = ry / den_n;
  // rows of Rn . Rc: u = row 0, v = row 1, n = row 2
  double u0 = rc00, u1 = rc01, u2 = rc02;
  double v0 = cn * rc10 + (-sn) * rc20, v1 = cn * rc11 + (-sn) * rc21, v2 = cn * rc12 + (-sn) * rc22;
  double w0 = sn * rc10 + cn * rc20, w1 = sn * rc11 + cn * rc21, w2 = sn * rc12 + cn * rc22;
  o[6] = w0;  o[7] = w1;  o[8] = w2;     // n
  o[9] = u0;  o[10] = u1; o[11] = u2;    // u
  o[12] = v0; o[13] = v1; o[14] = v2;    // v
  o[15] = 0.0;
}

__device__ __forceinline__ bool key_greater(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return (ka > kb) || (ka == kb && ia > ib);
}

// One block per residue row.  Distances to every residue of the same structure are
// sorted (bitonic network in shared memory, keys = fp64 bit patterns, which order
// like the values because distances are >= +0) and the first K+1 ranks are used.
// COMPACT (the fused tokenize path of the tensor-core modes): instead of the 27 fp32 features the kernels write
// 16 floats per edge, [d*d, p, q, k, t (12 orientation features), 0, 0, 0]; the 15 RBFs exp(-d*d / 1.5^k) are then
// evaluated in fp32 by the edge-embedding kernel that consumes them (edge_mlp_tc.cu), see pst_launch_featurize.
template <int kThreads, bool COMPACT>
__device__ __forceinline__ void knn_feature_row(const int row, const double* __restrict__ prep, const int32_t* __restrict__ offsets,
                                                int B, int R, int K, int max_len, int32_t* __restrict__ senders,
                                                float* __restrict__ feat, int32_t* __restrict__ status) {
  extern __shared__ unsigned long long smem_keys[];
  __shared__ int s_struct[2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    int lo = 0, hi = B;  // largest b with offsets[b] <= row
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (offsets[mid] <= row) lo = mid; else hi = mid;
    }
    s_struct[0] = offsets[lo];
    s_struct[1] = offsets[lo + 1] - offsets[lo];
  }
  __syncthreads();
  const int base = s_struct[0];
  const int L = s_struct[1];
  if (L < K || L > max_len) {
    if (tid == 0) atomicMin(status, (int)PST_ERR_LENGTH_OUT_OF_RANGE);
    for (int e = tid; e < K; e += kThreads) senders[(size_t)row * K + e] = 0;
    constexpr int kStride = COMPACT ? 16 : PST_EDGE_FEATURES;
    if (feat)
      for (int t = tid; t < K * kStride; t += kThreads) feat[(size_t)row * K * kStride + t] = 0.f;
    return;
  }
  int n_pad = 64;
  while (n_pad < L) n_pad <<= 1;
  int* s_idx = reinterpret_cast<int*>(smem_keys + n_pad);

  const double* pi = prep + (size_t)row * PST_PREP_STRIDE;
  const double xi = pi[0], yi = pi[1], zi = pi[2];
  for (int j = tid; j < n_pad; j += kThreads) {
    unsigned long long key = 0xFFFFFFFFFFFFFFFFull;
    if (j < L) {
      const double* pj = prep + (size_t)(base + j) * PST_PREP_STRIDE;
      double dx = xi - pj[0], dy = yi - pj[1], dz = zi - pj[2];
      double d = sqrt((dx * dx + dy * dy) + dz * dz);
      key = (unsigned long long)__double_as_longlong(d);
    }
    smem_keys[j] = key;
    s_idx[j] = j;
  }
  __syncthreads();
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n_pad >> 1); t += kThreads) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo | j;
        bool asc = (lo & k) == 0;
        unsigned long long ka = smem_keys[lo], kb = smem_keys[hi];
        int ia = s_idx[lo], ib = s_idx[hi];
        if (key_greater(ka, ia, kb, ib) == asc) {
          smem_keys[lo] = kb; smem_keys[hi] = ka;
          s_idx[lo] = ib; s_idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  const int first = (L == K) ? 0 : 1;  // protein_utils.py:385-389
  for (int e = tid; e < K; e += kThreads) senders[(size_t)row * K + e] = s_idx[first + e];
  if (!feat) return;

  const double ca_x = pi[3], ca_y = pi[4], ca_z = pi[5];
  if (COMPACT) {
    float* out = feat + (size_t)row * K * 16;
    for (int t = tid; t < K * 16; t += kThreads) {
      const int e = t >> 4, c = t & 15;
      const int j = s_idx[first + e];
      double val = 0.0;
      if (c == 0) {
        const double d = __longlong_as_double((long long)smem_keys[first + e]);
        val = d * d;
      } else if (c <= 12) {
        const int g = (c - 1) / 3, r = (c - 1) - 3 * g;
        const double* pj = prep + (size_t)(base + j) * PST_PREP_STRIDE;
        double vx, vy, vz;
        if (g == 0) {
          vx = pj[3] - ca_x; vy = pj[4] - ca_y; vz = pj[5] - ca_z;
        } else {
          const double* sv = pj + 3 + 3 * g;
          vx = sv[0]; vy = sv[1]; vz = sv[2];
        }
        const double* b = pi + 6 + 3 * r;
        val = b[0] * vx + b[1] * vy + b[2] * vz;
      }
      out[t] = (float)val;
    }
    return;
  }
  float* out = feat + (size_t)row * K * PST_EDGE_FEATURES;
  for (int t = tid; t < K * PST_EDGE_FEATURES; t += kThreads) {
    int e = t / PST_EDGE_FEATURES;
    int f = t - e * PST_EDGE_FEATURES;
    int j = s_idx[first + e];
    double val;
    if (f < 15) {
      double d = __longlong_as_double((long long)smem_keys[first + e]);
      double scale = 1.0;
      for (int q = 0; q < f; ++q) scale = scale * 1.5;  // 1.5**f, exact in fp64 for f <= 14
      val = exp(-(d * d) / scale);
    } else {
      int g = (f - 15) / 3;       // 0:p 1:q 2:k 3:t
      int r = (f - 15) - 3 * g;   // basis row: 0:n 1:u 2:v
      const double* pj = prep + (size_t)(base + j) * PST_PREP_STRIDE;
      double vx, vy, vz;
      if (g == 0) {
        vx = pj[3] - ca_x; vy = pj[4] - ca_y; vz = pj[5] - ca_z;
      } else {
        const double* s = pj + 3 + 3 * g;  // g=1: n (6), g=2: u (9), g=3: v (12)
        vx = s[0]; vy = s[1]; vz = s[2];
      }
      const double* b = pi + 6 + 3 * r;
      val = b[0] * vx + b[1] * vy + b[2] * vz;
    }
    out[t] = (float)val;
  }
}

// every row (configurations the warp kernel does not cover)
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
knn_feature_kernel(const double* __restrict__ prep, const int32_t* __restrict__ offsets, int B, int R, int K, int max_len,
                   int32_t* __restrict__ senders, float* __restrict__ feat, int32_t* __restrict__ status) {
  knn_feature_row<kThreads, false>(blockIdx.x, prep, offsets, B, R, K, max_len, senders, feat, status);
}

// only the rows the warp kernel flagged: redo[0] = how many, redo[1..] = which (normally none: a small fixed grid
// reads the count and leaves, instead of one block per row finding its flag clear)
template <int kThreads, bool COMPACT>
__global__ void __launch_bounds__(kThreads)
knn_feature_redo_kernel(const double* __restrict__ prep, const int32_t* __restrict__ offsets, int B, int R, int K, int max_len,
                        int32_t* __restrict__ senders, float* __restrict__ feat, int32_t* __restrict__ status,
                        const int32_t* __restrict__ redo) {
  const int n = redo[0];
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    knn_feature_row<kThreads, COMPACT>(redo[1 + i], prep, offsets, B, R, K, max_len, senders, feat, status);
    __syncthreads();  // the shared key / index arrays are reused by the next row
  }
}

// -------------------------------------------------------------------------------------------------
// Warp-per-row k-NN: the L distances of a row are streamed in chunks of 64 (2 per lane); each chunk is
// sorted with a register/shuffle bitonic network and merged into the running 64 smallest (min against
// the reversed chunk gives a bitonic sequence holding the 64 smallest of the union, then a 6-stage
// bitonic merge).  No block barriers.
//
// Keys are single 64-bit words: the fp64 distance bit pattern with its 11 lowest mantissa bits replaced
// by the candidate index (structures handled here have L <= 2048).  Two candidates whose distances agree
// in the upper 53 bits would be ordered by index instead of by their last 11 bits, so the kernel CHECKS
// the sorted head: if two adjacent entries among ranks 0..K+2 share the truncated distance (this also
// covers exact ties), the row is flagged and recomputed by the exact (distance, index) block kernel
// below.  The emitted order is therefore always the stable ascending argsort of the reference contract.
//
// The scan orders candidates by the SQUARED distance (no fp64 sqrt per candidate; build with -DPST_KNN_DIST_KEYS for
// keys made of the rounded distance itself).  sqrt is monotone, so the two orders differ only where two different
// squares round to the same distance, which the reference then orders by index.  Such squares lie within 4 ulps of
// each other (sqrt(y) - sqrt(x) <= ulp(r) gives (y - x) / x <= 2^-51), i.e. their truncated keys are equal or
// ADJACENT: with squared keys the head check therefore flags every adjacent pair whose truncated values differ by
// at most one (a relative gap below 2^-41: never seen outside constructed cases), and the exact kernel, which does
// take the square roots, decides those rows.  tests/test_knn_key_model.py states the rule as a NumPy model and
// checks it on CASP14, on a lattice and on constructed same-sqrt pairs.
#ifdef PST_KNN_DIST_KEYS
constexpr bool kKeyIsD2 = false;
#else
constexpr bool kKeyIsD2 = true;
#endif
__device__ __forceinline__ void ce_remote(unsigned long long& a, int lane_mask, bool keep_min) {
  const unsigned long long o = __shfl_xor_sync(0xffffffffu, a, lane_mask);
  if ((o < a) == keep_min) a = o;
}
__device__ __forceinline__ void ce_local(unsigned long long& a0, unsigned long long& a1, bool asc) {
  if ((a1 < a0) == asc) {
    const unsigned long long t = a0;
    a0 = a1;
    a1 = t;
  }
}
// stages j = k/2 .. 1 of a bitonic network on 64 elements (element e = 2*lane + r); ascending iff (e & k) == 0
__device__ __forceinline__ void bitonic_stages(unsigned long long& a0, unsigned long long& a1, int lane, int k) {
  const int e0 = 2 * lane;
  const bool asc = (e0 & k) == 0;
#pragma unroll
  for (int j = 32; j >= 2; j >>= 1) {
    if (j < k) {
      const bool keep_min = (((e0 & j) == 0) == asc);
      ce_remote(a0, j >> 1, keep_min);
      ce_remote(a1, j >> 1, keep_min);
    }
  }
  ce_local(a0, a1, asc);
}
__device__ __forceinline__ void bitonic_sort64(unsigned long long& a0, unsigned long long& a1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) bitonic_stages(a0, a1, lane, k);
}

// The same network on 32-bit keys (the coarse pre-selection below): one shuffle and one min / max per exchange instead of
// two shuffles, a two-word compare and two selects.
__device__ __forceinline__ void ce_remote32(unsigned& a, int lane_mask, bool keep_min) {
  const unsigned o = __shfl_xor_sync(0xffffffffu, a, lane_mask);
  const unsigned lo = min(a, o), hi = max(a, o);
  a = keep_min ? lo : hi;
}
__device__ __forceinline__ void bitonic_stages32(unsigned& a0, unsigned& a1, int lane, int k) {
  const int e0 = 2 * lane;
  const bool asc = (e0 & k) == 0;
#pragma unroll
  for (int j = 32; j >= 2; j >>= 1) {
    if (j < k) {
      const bool keep_min = (((e0 & j) == 0) == asc);
      ce_remote32(a0, j >> 1, keep_min);
      ce_remote32(a1, j >> 1, keep_min);
    }
  }
  const unsigned lo = min(a0, a1), hi = max(a0, a1);
  a0 = asc ? lo : hi;
  a1 = asc ? hi : lo;
}
__device__ __forceinline__ void bitonic_sort64_32(unsigned& a0, unsigned& a1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) bitonic_stages32(a0, a1, lane, k);
}

// 1 / 1.5**k, k = 0..14: the RBF argument -d*d / 1.5**k (utils/protein_utils.py:266-270) is formed as a product; the one-ulp
// (fp64) difference from the reference's division is 1e-14 relative on the result, far below the fp32 rounding.
__constant__ double c_rbf_inv_scale[15] = {1.0 / 1.0, 1.0 / 1.5, 1.0 / 2.25, 1.0 / 3.375, 1.0 / 5.0625, 1.0 / 7.59375,
                                          1.0 / 11.390625, 1.0 / 17.0859375, 1.0 / 25.62890625, 1.0 / 38.443359375,
                                          1.0 / 57.6650390625, 1.0 / 86.49755859375, 1.0 / 129.746337890625,
                                          1.0 / 194.6195068359375, 1.0 / 291.92926025390625};

// exp(x) for x <= 0 to ~1e-11 relative, as a double that is then rounded to fp32 (the features are fp32: SURVEY A.4
// gate <= 1 fp32 ulp from float32(exp_fp64)).  n = rint(x log2 e), r = x - n ln2 (two-term Cody-Waite), degree-9
// Taylor polynomial on |r| <= 0.347 (remainder 7e-12), scaled by 2^n through the exponent field.  About a quarter
// of the instructions of the library exp().  Below e^-104 = 2^-150.04 the fp32 result is 0.
__device__ __forceinline__ double exp_neg_for_fp32(double x) {
  if (x < -104.0) return 0.0;
  const double n = rint(x * 1.4426950408889634);
  double r = fma(-n, 6.93147180369123816490e-01, x);
  r = fma(-n, 1.90821492927058770002e-10, r);
  double p = 2.7557319223985893e-06;                 // 1/9!
  p = fma(p, r, 2.4801587301587302e-05);             // 1/8!
  p = fma(p, r, 1.9841269841269841e-04);             // 1/7!
  p = fma(p, r, 1.3888888888888889e-03);             // 1/6!
  p = fma(p, r, 8.3333333333333332e-03);             // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);             // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);             // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const long long bits = ((long long)(int)n + 1023) << 52;  // n >= -151: a normal double
  return p * __longlong_as_double(bits);
}

constexpr int kKnnWarps = 8;
constexpr unsigned long long kIdxMask = 0x7FFull;  // 11 bits: L <= 2048

__device__ __forceinline__ unsigned long long packed_key(const double3& ci, const double* __restrict__ cx, const double* __restrict__ cy,
                                                         const double* __restrict__ cz, int base, int j, int L) {
  if (j >= L) return 0xFFFFFFFFFFFFFFFFull;
  const double dx = ci.x - cx[base + j], dy = ci.y - cy[base + j], dz = ci.z - cz[base + j];
  const double d2 = (dx * dx + dy * dy) + dz * dz;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(kKeyIsD2 ? d2 : sqrt(d2));
  return (bits & ~kIdxMask) | (unsigned long long)j;
}

// Coarse 32-bit key of the pre-selection: the squared distance rounded to fp32 (monotone in the fp64 value), its 11
// lowest mantissa bits replaced by the candidate index.  Monotone: d2_a < d2_b  =>  key_a >> 11 <= key_b >> 11.
__device__ __forceinline__ unsigned coarse_key(const double3& ci, const double* __restrict__ cx, const double* __restrict__ cy,
                                               const double* __restrict__ cz, int base, int j, int L) {
  if (j >= L) return 0xFFFFFFFFu;
  const double dx = ci.x - cx[base + j], dy = ci.y - cy[base + j], dz = ci.z - cz[base + j];
  const float d2 = (float)((dx * dx + dy * dy) + dz * dz);
  return (__float_as_uint(d2) & ~(unsigned)kIdxMask) | (unsigned)j;
}

template <bool COMPACT>
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_warp_kernel(const double* __restrict__ prep, const double* __restrict__ cen, const int32_t* __restrict__ offsets,
                int B, int R, int K, int max_len, int32_t* __restrict__ senders, float* __restrict__ feat,
                int32_t* __restrict__ status, int32_t* __restrict__ redo) {
  __shared__ int s_j[kKnnWarps][64];
  __shared__ double s_d2[kKnnWarps][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kKnnWarps + warp;
  if (row >= R) return;
  int base = 0, L = 0;
  if (lane == 0) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (offsets[mid] <= row) lo = mid; else hi = mid;
    }
    base = offsets[lo];
    L = offsets[lo + 1] - base;
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  L = __shfl_sync(0xffffffffu, L, 0);
  if (L < K || L > max_len) {
    if (lane == 0) atomicMin(status, (int)PST_ERR_LENGTH_OUT_OF_RANGE);
    for (int e = lane; e < K; e += 32) senders[(size_t)row * K + e] = 0;
    constexpr int kStride = COMPACT ? 16 : PST_EDGE_FEATURES;
    if (feat)
      for (int t = lane; t < K * kStride; t += 32) feat[(size_t)row * K * kStride + t] = 0.f;
    return;
  }
  const double* pi = prep + (size_t)row * PST_PREP_STRIDE;
  const double3 ci = make_double3(pi[0], pi[1], pi[2]);
  const double* cx = cen;
  const double* cy = cen + (size_t)R;
  const double* cz = cen + 2 * (size_t)R;
  // ---- phase 1 (round 2): COARSE pre-selection of 64 candidates on 32-bit keys (fp32 squared distance | index): the
  // compare-exchange network was half of the kernel's instructions with 64-bit keys.  Phase 2 recomputes the exact
  // fp64 keys of the 64 survivors and sorts those once.  The coarse key is monotone in the exact squared distance, so
  // every candidate outside the coarse top 64 is at least as far as the 64-th one; the survivors therefore contain the
  // exact top K + 4 whenever the truncated key of coarse rank K + 3 is strictly below that of rank 63 (checked; a
  // row that fails goes to the exact kernel like a row with a clash in its head).
  unsigned c0 = 0, c1 = 0;  // running 64 smallest coarse keys, ascending, element e = 2*lane + r
  const int n_chunks = (L + 63) >> 6;
  // Chunks are visited outwards from the one that holds the row itself; a chunk none of whose keys is below
  // the current 64-th smallest cannot change the result and is skipped after the distance evaluation.
  const int c_home = (row - base) >> 6;
  unsigned bound = 0xFFFFFFFFu;
  // Only keys below the current bound can enter the result.  They are few per chunk once the home chunk has set the
  // bound (about 100 per row over all other chunks at L = 512), so they are COMPACTED into a 64-entry buffer (ballot +
  // popc, order irrelevant: the keys are unique) and one sort-and-merge round is spent per full buffer instead of
  // per chunk (3.4 instead of 5.1 rounds per row at L = 512, 3.9 instead of 6.2 at L = 2 048).
  unsigned* buf = reinterpret_cast<unsigned*>(s_d2[warp]);  // free until the feature phase
  int cnt = 0;
  auto merge_sorted = [&](unsigned a0, unsigned a1) {
    // a ascending; reversed: element e <- element 63 - e (lane 31 - lane, registers swapped); min against c gives a
    // bitonic sequence holding the 64 smallest of the union
    const unsigned r0 = __shfl_xor_sync(0xffffffffu, a1, 31), r1 = __shfl_xor_sync(0xffffffffu, a0, 31);
    c0 = min(r0, c0);
    c1 = min(r1, c1);
    bitonic_stages32(c0, c1, lane, 64);  // bitonic -> ascending
    bound = __shfl_sync(0xffffffffu, c1, 31);  // largest of the current 64 smallest
  };
  auto flush = [&]() {
    __syncwarp();
    unsigned x0 = 2 * lane < cnt ? buf[2 * lane] : 0xFFFFFFFFu;
    unsigned x1 = 2 * lane + 1 < cnt ? buf[2 * lane + 1] : 0xFFFFFFFFu;
    __syncwarp();
    bitonic_sort64_32(x0, x1, lane);
    merge_sorted(x0, x1);
    cnt = 0;
  };
  const unsigned lt = (1u << lane) - 1u;
  for (int step = 0; step < 2 * n_chunks; ++step) {
    const int delta = (step + 1) >> 1;
    const int c = (step & 1) ? c_home + delta : c_home - delta;
    if (step > 0 && (c < 0 || c >= n_chunks)) continue;
    const int j0 = c * 64 + lane;  // lanes on consecutive candidates (which lane holds which key does not matter below)
    const unsigned a0 = coarse_key(ci, cx, cy, cz, base, j0, L);
    const unsigned a1 = coarse_key(ci, cx, cy, cz, base, j0 + 32, L);
    if (step == 0) {  // the home chunk sets the first bound
      c0 = a0;
      c1 = a1;
      bitonic_sort64_32(c0, c1, lane);
      bound = __shfl_sync(0xffffffffu, c1, 31);
      continue;
    }
    unsigned m0 = __ballot_sync(0xffffffffu, a0 < bound), m1 = __ballot_sync(0xffffffffu, a1 < bound);
    int n = __popc(m0) + __popc(m1);
    if (n == 0) continue;
    if (cnt + n > 64) {
      flush();  // tightens the bound: filter this chunk again
      m0 = __ballot_sync(0xffffffffu, a0 < bound);
      m1 = __ballot_sync(0xffffffffu, a1 < bound);
      n = __popc(m0) + __popc(m1);
      if (n == 0) continue;
    }
    if (a0 < bound) buf[cnt + __popc(m0 & lt)] = a0;
    if (a1 < bound) buf[cnt + __popc(m0) + __popc(m1 & lt)] = a1;
    cnt += n;
  }
  if (cnt > 0) flush();
  // the survivors contain the exact head only if the coarse keys separate rank K + 3 from rank 63
  {
    const int rk = min(K + 3, 62);
    const unsigned krk = __shfl_sync(0xffffffffu, (rk & 1) ? c1 : c0, rk >> 1);
    const unsigned k63 = __shfl_sync(0xffffffffu, c1, 31);
    if (L > 64 && (krk >> 11) >= (k63 >> 11)) {  // L <= 64: every candidate is a survivor
      if (lane == 0) redo[1 + atomicAdd(redo, 1)] = row;
      return;
    }
  }
  // ---- phase 2: exact fp64 keys of the 64 survivors, one 64-bit sort
  unsigned long long b0 = c0 == 0xFFFFFFFFu ? 0xFFFFFFFFFFFFFFFFull : packed_key(ci, cx, cy, cz, base, (int)(c0 & (unsigned)kIdxMask), L);
  unsigned long long b1 = c1 == 0xFFFFFFFFu ? 0xFFFFFFFFFFFFFFFFull : packed_key(ci, cx, cy, cz, base, (int)(c1 & (unsigned)kIdxMask), L);
  bitonic_sort64(b0, b1, lane);
  // exactness check on the sorted head (ranks 0 .. K+2): equal (squared keys: equal or adjacent) truncated values ->
  // exact recompute
  {
    constexpr unsigned long long kGap = kKeyIsD2 ? 1ull : 0ull;
    const unsigned long long nxt = __shfl_down_sync(0xffffffffu, b0, 1);  // element 2*lane + 2
    bool clash = (2 * lane + 1 <= K + 2) && ((b1 >> 11) - (b0 >> 11) <= kGap) && (b1 != 0xFFFFFFFFFFFFFFFFull);
    clash = clash || ((2 * lane + 2 <= K + 2) && lane < 31 && ((nxt >> 11) - (b1 >> 11) <= kGap) && (nxt != 0xFFFFFFFFFFFFFFFFull));
    if (__any_sync(0xffffffffu, clash)) {
      if (lane == 0) redo[1 + atomicAdd(redo, 1)] = row;
      return;
    }
  }
  s_j[warp][2 * lane] = (int)(b0 & kIdxMask);
  s_j[warp][2 * lane + 1] = (int)(b1 & kIdxMask);
  __syncwarp();
  const int first = (L == K) ? 0 : 1;  // protein_utils.py:385-389
  for (int e = lane; e < K; e += 32) senders[(size_t)row * K + e] = s_j[warp][first + e];
  if (!feat) return;

  // ---- features: squared distances once per edge, then the 750 RBF items and the 600 orientation items as two
  // divergence-free loops (item -> lane mapping keeps consecutive lanes on consecutive output floats)
  for (int e = lane; e < K; e += 32) {
    const double* cj = prep + (size_t)(base + s_j[warp][first + e]) * PST_PREP_STRIDE;  // [0:3] = centroid (128-byte records)
    const double2 cxy = *reinterpret_cast<const double2*>(cj);
    const double dx = ci.x - cxy.x, dy = ci.y - cxy.y, dz = ci.z - cj[2];
    const double d = sqrt((dx * dx + dy * dy) + dz * dz);  // the reference squares the rounded distance again
    s_d2[warp][e] = d * d;
  }
  __syncwarp();
  if (COMPACT) {
    // 16 floats per edge: [d*d, 12 orientation features, 0, 0, 0]; item -> lane keeps consecutive lanes on consecutive floats.
    // t = lane + 32 * it, so the slot c = t & 15 is a per-lane constant: the receiver's basis row and (for p) its CA stay in
    // registers and every iteration loads only the three doubles of the sender's vector (the kernel is bound by L1
    // data-pipe wavefronts, profiles/r01_kernels_ncu_full.md: the loop used to issue nine scattered loads per item).
    float* out = feat + (size_t)row * K * 16;
    const int c = lane & 15;
    const bool orient = c >= 1 && c <= 12;
    const int g = orient ? (c - 1) / 3 : 0, r = orient ? (c - 1) - 3 * g : 0;  // g: 0:p 1:q 2:k 3:t ; r: basis row 0:n 1:u 2:v
    const double b0 = pi[6 + 3 * r], b1 = pi[7 + 3 * r], b2 = pi[8 + 3 * r];
    // p = B.(CA_j - CA_i); q, k, t = B.axis_j: subtracting +0.0 leaves every double (and its sign) unchanged
    const double ox = g == 0 ? pi[3] : 0.0, oy = g == 0 ? pi[4] : 0.0, oz = g == 0 ? pi[5] : 0.0;
    const int sv_off = 3 + 3 * g;
    for (int t = lane; t < K * 16; t += 32) {
      const int e = t >> 4;
      double val = 0.0;
      if (c == 0) {
        val = s_d2[warp][e];
      } else if (orient) {
        const double* sv = prep + (size_t)(base + s_j[warp][first + e]) * PST_PREP_STRIDE + sv_off;
        const double vx = sv[0] - ox, vy = sv[1] - oy, vz = sv[2] - oz;
        val = b0 * vx + b1 * vy + b2 * vz;
      }
      out[t] = (float)val;
    }
    return;
  }
  float* out = feat + (size_t)row * K * PST_EDGE_FEATURES;
  for (int t = lane; t < K * 15; t += 32) {
    const int e = t / 15, f = t - e * 15;
    out[e * PST_EDGE_FEATURES + f] = (float)exp_neg_for_fp32(-s_d2[warp][e] * c_rbf_inv_scale[f]);
  }
  const double ca_x = pi[3], ca_y = pi[4], ca_z = pi[5];
  for (int t = lane; t < K * 12; t += 32) {
    const int e = t / 12, c = t - e * 12;
    const int g = c / 3, r = c - 3 * g;  // g: 0:p 1:q 2:k 3:t ; r: basis row 0:n 1:u 2:v
    const double* pj = prep + (size_t)(base + s_j[warp][first + e]) * PST_PREP_STRIDE;
    double vx, vy, vz;
    if (g == 0) {
      vx = pj[3] - ca_x; vy = pj[4] - ca_y; vz = pj[5] - ca_z;
    } else {
      const double* sv = pj + 3 + 3 * g;
      vx = sv[0]; vy = sv[1]; vz = sv[2];
    }
    const double* b = pi + 6 + 3 * r;
    out[e * PST_EDGE_FEATURES + 15 + c] = (float)(b[0] * vx + b[1] * vy + b[2] * vz);
  }
}

}  // namespace

// the compact feature layout exists on the warp-kernel path only
bool pst_featurize_compact_ok(const pst_model* m) { return m->cfg.num_neighbor <= 60 && m->cfg.max_len <= 2048; }

int pst_launch_featurize(const pst_model* m, cudaStream_t st, const float* atoms,
                         const uint8_t* mask, int apr, const int32_t* offsets, int B, int R,
                         int32_t* senders, float* edge_feat, double* prep, double* cen, int32_t* status, int32_t* redo,
                         int compact) {
  if (R <= 0) return 0;
  prep_kernel<<<(R + 127) / 128, 128, 0, st>>>(atoms, mask, apr, R, prep, cen);
  int n_pad = 64;
  while (n_pad < m->cfg.max_len) n_pad <<= 1;
  const size_t smem = (size_t)n_pad * (sizeof(unsigned long long) + sizeof(int));
  constexpr int kThreads = 256;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(knn_feature_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (m->cfg.num_neighbor <= 60 && m->cfg.max_len <= 2048) {
    cudaMemsetAsync(redo, 0, sizeof(int32_t), st);  // the counter
    const dim3 wg((R + kKnnWarps - 1) / kKnnWarps), rg(min(R, 2 * m->num_sms));
    if (compact) {
      knn_warp_kernel<true><<<wg, kKnnWarps * 32, 0, st>>>(prep, cen, offsets, B, R,
                                                            m->cfg.num_neighbor, m->cfg.max_len, senders, edge_feat, status, redo);
      // exact recompute of the (normally zero) rows flagged above
      cudaFuncSetAttribute(knn_feature_redo_kernel<kThreads, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      knn_feature_redo_kernel<kThreads, true><<<rg, kThreads, smem, st>>>(prep, offsets, B, R, m->cfg.num_neighbor, m->cfg.max_len,
                                                                           senders, edge_feat, status, redo);
    } else {
      knn_warp_kernel<false><<<wg, kKnnWarps * 32, 0, st>>>(prep, cen, offsets, B, R,
                                                             m->cfg.num_neighbor, m->cfg.max_len, senders, edge_feat, status, redo);
      cudaFuncSetAttribute(knn_feature_redo_kernel<kThreads, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      knn_feature_redo_kernel<kThreads, false><<<rg, kThreads, smem, st>>>(prep, offsets, B, R, m->cfg.num_neighbor, m->cfg.max_len,
                                                                            senders, edge_feat, status, redo);
    }
    return 3;
  }
  if (compact) return PST_ERR_UNSUPPORTED_CONFIG;  // callers ask pst_featurize_compact_ok() first
  knn_feature_kernel<kThreads><<<R, kThreads, smem, st>>>(prep, offsets, B, R, m->cfg.num_neighbor, m->cfg.max_len,
                                                         senders, edge_feat, status);
  return 2;
}
