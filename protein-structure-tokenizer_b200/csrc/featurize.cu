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
                            int apr, int R, double* __restrict__ prep, double4* __restrict__ cen4) {
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
  cen4[i] = make_double4(o[0], o[1], o[2], 0.0);  // compact copy: the k-NN scan streams 32 B per candidate
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
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
knn_feature_kernel(const double* __restrict__ prep, const int32_t* __restrict__ offsets, int B, int R,
                   int K, int max_len, int32_t* __restrict__ senders, float* __restrict__ feat,
                   int32_t* __restrict__ status) {
  extern __shared__ unsigned long long smem_keys[];
  __shared__ int s_struct[2];
  const int row = blockIdx.x;
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
    if (feat)
      for (int t = tid; t < K * PST_EDGE_FEATURES; t += kThreads) feat[(size_t)row * K * PST_EDGE_FEATURES + t] = 0.f;
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


// -------------------------------------------------------------------------------------------------
// Warp-per-row k-NN: the L distances of a row are streamed in chunks of 64 (2 per lane); each chunk is
// sorted with a register/shuffle bitonic network and merged into the running 64 smallest
// (min against the reversed chunk gives a bitonic sequence holding the 64 smallest of the union, then
// a 6-stage bitonic merge).  Ordering is lexicographic on (fp64 distance bits, index): exactly the
// stable ascending argsort the reference's contract is defined by.  No block barriers.
struct Cand {
  unsigned long long key;
  int idx;
};
__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) {
  return (a.key < b.key) || (a.key == b.key && a.idx < b.idx);
}
__device__ __forceinline__ Cand cand_shfl_xor(const Cand& a, int mask) {
  Cand o;
  o.key = __shfl_xor_sync(0xffffffffu, a.key, mask);
  o.idx = __shfl_xor_sync(0xffffffffu, a.idx, mask);
  return o;
}
// compare-exchange across lanes: element index e = 2*lane + r, partner e ^ j (j >= 2), ascending iff asc
__device__ __forceinline__ void ce_remote(Cand& a, int lane_mask, bool keep_min) {
  const Cand o = cand_shfl_xor(a, lane_mask);
  const bool o_less = cand_less(o, a);
  if (o_less == keep_min) a = o;
}
__device__ __forceinline__ void ce_local(Cand& a0, Cand& a1, bool asc) {
  if (cand_less(a1, a0) == asc) {
    const Cand t = a0;
    a0 = a1;
    a1 = t;
  }
}
// stages j = k/2 .. 1 of a bitonic network on 64 elements (2 per lane); dir(e) ascending iff (e & k) == 0
__device__ __forceinline__ void bitonic_stages(Cand& a0, Cand& a1, int lane, int k) {
  const int e0 = 2 * lane;
  const bool asc = (e0 & k) == 0;  // bit k of e0 and e0+1 agree for k >= 2
#pragma unroll
  for (int j = 32; j >= 2; j >>= 1) {
    if (j < k) {
      const bool lower = (e0 & j) == 0;
      const bool keep_min = (lower == asc);
      ce_remote(a0, j >> 1, keep_min);
      ce_remote(a1, j >> 1, keep_min);
    }
  }
  ce_local(a0, a1, asc);
}
__device__ __forceinline__ void bitonic_sort64(Cand& a0, Cand& a1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) bitonic_stages(a0, a1, lane, k);
}

constexpr int kKnnWarps = 8;

__global__ void __launch_bounds__(kKnnWarps * 32)
knn_warp_kernel(const double* __restrict__ prep, const double4* __restrict__ cen4, const int32_t* __restrict__ offsets,
                int B, int R, int K, int max_len, int32_t* __restrict__ senders, float* __restrict__ feat,
                int32_t* __restrict__ status) {
  __shared__ double s_d[kKnnWarps][64];
  __shared__ int s_j[kKnnWarps][64];
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
    if (feat)
      for (int t = lane; t < K * PST_EDGE_FEATURES; t += 32) feat[(size_t)row * K * PST_EDGE_FEATURES + t] = 0.f;
    return;
  }
  const double4 ci = cen4[row];
  Cand b0, b1;  // running 64 smallest, ascending, element e = 2*lane + r
  const int n_chunks = (L + 63) >> 6;
  // Chunks are visited outwards from the one that holds the row itself (sequence neighbours are spatial
  // neighbours, so the running 64-th smallest distance tightens early); a chunk none of whose distances is
  // <= that bound cannot change the result and is skipped after the distance evaluation (warp vote).
  const int c_home = (row - base) >> 6;
  unsigned long long bound = 0xFFFFFFFFFFFFFFFFull;
  for (int step = 0; step < 2 * n_chunks; ++step) {
    // step 0 -> home, then +1, -1, +2, -2, ...
    const int delta = (step + 1) >> 1;
    const int c = (step & 1) ? c_home + delta : c_home - delta;
    if (step == 0 ? false : (c < 0 || c >= n_chunks)) continue;
    Cand a0, a1;
    {
      const int j0 = c * 64 + 2 * lane;
      a0.idx = j0;
      a1.idx = j0 + 1;
      a0.key = a1.key = 0xFFFFFFFFFFFFFFFFull;
      if (j0 < L) {
        const double4 cj = cen4[base + j0];
        const double dx = ci.x - cj.x, dy = ci.y - cj.y, dz = ci.z - cj.z;
        a0.key = (unsigned long long)__double_as_longlong(sqrt((dx * dx + dy * dy) + dz * dz));
      }
      if (j0 + 1 < L) {
        const double4 cj = cen4[base + j0 + 1];
        const double dx = ci.x - cj.x, dy = ci.y - cj.y, dz = ci.z - cj.z;
        a1.key = (unsigned long long)__double_as_longlong(sqrt((dx * dx + dy * dy) + dz * dz));
      }
    }
    if (step > 0 && !__any_sync(0xffffffffu, a0.key <= bound || a1.key <= bound)) continue;
    bitonic_sort64(a0, a1, lane);
    if (step == 0) {
      b0 = a0;
      b1 = a1;
    } else {
      // reversed chunk: element e <- element 63 - e  (lane 31 - lane, registers swapped)
      const Cand r0 = cand_shfl_xor(a1, 31), r1 = cand_shfl_xor(a0, 31);
      if (cand_less(r0, b0)) b0 = r0;
      if (cand_less(r1, b1)) b1 = r1;
      bitonic_stages(b0, b1, lane, 64);  // bitonic -> ascending
    }
    bound = __shfl_sync(0xffffffffu, b1.key, 31);  // largest of the current 64 smallest
  }
  s_d[warp][2 * lane] = __longlong_as_double((long long)b0.key);
  s_d[warp][2 * lane + 1] = __longlong_as_double((long long)b1.key);
  s_j[warp][2 * lane] = b0.idx;
  s_j[warp][2 * lane + 1] = b1.idx;
  __syncwarp();
  const int first = (L == K) ? 0 : 1;  // protein_utils.py:385-389
  for (int e = lane; e < K; e += 32) senders[(size_t)row * K + e] = s_j[warp][first + e];
  if (!feat) return;

  const double* pi = prep + (size_t)row * PST_PREP_STRIDE;
  const double ca_x = pi[3], ca_y = pi[4], ca_z = pi[5];
  float* out = feat + (size_t)row * K * PST_EDGE_FEATURES;
  for (int t = lane; t < K * PST_EDGE_FEATURES; t += 32) {
    const int e = t / PST_EDGE_FEATURES;
    const int f = t - e * PST_EDGE_FEATURES;
    double val;
    if (f < 15) {
      const double d = s_d[warp][first + e];
      double scale = 1.0;
      for (int q = 0; q < f; ++q) scale = scale * 1.5;  // 1.5**f, exact in fp64
      val = exp(-(d * d) / scale);
    } else {
      const int g = (f - 15) / 3;      // 0:p 1:q 2:k 3:t
      const int r = (f - 15) - 3 * g;  // basis row: 0:n 1:u 2:v
      const double* pj = prep + (size_t)(base + s_j[warp][first + e]) * PST_PREP_STRIDE;
      double vx, vy, vz;
      if (g == 0) {
        vx = pj[3] - ca_x; vy = pj[4] - ca_y; vz = pj[5] - ca_z;
      } else {
        const double* s = pj + 3 + 3 * g;
        vx = s[0]; vy = s[1]; vz = s[2];
      }
      const double* b = pi + 6 + 3 * r;
      val = b[0] * vx + b[1] * vy + b[2] * vz;
    }
    out[t] = (float)val;
  }
}

}  // namespace

int pst_launch_featurize(const pst_model* m, cudaStream_t st, const float* atoms,
                         const uint8_t* mask, int apr, const int32_t* offsets, int B, int R,
                         int32_t* senders, float* edge_feat, double* prep, double* cen4, int32_t* status) {
  if (R <= 0) return 0;
  prep_kernel<<<(R + 127) / 128, 128, 0, st>>>(atoms, mask, apr, R, prep, reinterpret_cast<double4*>(cen4));
  if (m->cfg.num_neighbor <= 62) {
    knn_warp_kernel<<<(R + kKnnWarps - 1) / kKnnWarps, kKnnWarps * 32, 0, st>>>(
        prep, reinterpret_cast<const double4*>(cen4), offsets, B, R, m->cfg.num_neighbor, m->cfg.max_len, senders,
        edge_feat, status);
    return 2;
  }
  // K = 63, 64: the warp kernel keeps only 64 candidates (K + 1 needed); use the block-sort kernel
  int n_pad = 64;
  while (n_pad < m->cfg.max_len) n_pad <<= 1;
  size_t smem = (size_t)n_pad * (sizeof(unsigned long long) + sizeof(int));
  constexpr int kThreads = 256;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(knn_feature_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  knn_feature_kernel<kThreads><<<R, kThreads, smem, st>>>(prep, offsets, B, R, m->cfg.num_neighbor,
                                                         m->cfg.max_len, senders, edge_feat, status);
  return 2;
}
