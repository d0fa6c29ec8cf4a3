// Probe: TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) of four arbitrary rows of a [rows x 128] fp16 table
// into a SWIZZLE_128B K-major operand image.  Prints whether the bytes land where the UMMA operand layout expects them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gather4_probe gather4_probe.cu && ./gather4_probe [box_rows]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, const int* rows, uint16_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xFFFFFFFFu;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    // 128 rows x 2 K-blocks of 64 columns: 32 gather4 per K block
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32768u) : "memory");
    for (int kb = 0; kb < 2; ++kb)
      for (int g = 0; g < 32; ++g) {
        const uint32_t dst = smem_u32(smem) + kb * 16384 + g * 4 * 128;
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(dst), "l"(&map), "r"(kb * 64), "r"(rows[g * 4 + 0]), "r"(rows[g * 4 + 1]), "r"(rows[g * 4 + 2]),
              "r"(rows[g * 4 + 3]), "r"(smem_u32(&bar))
            : "memory");
      }
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  }
  for (int i = tid; i < 32768 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

static uint32_t swz_offset(uint32_t row, uint32_t k) {  // the operand image layout of edge_mlp_tc.cu
  uint32_t kb = k >> 6, kk = k & 63;
  return kb * 16384 + row * 128 + ((((kk >> 3) ^ (row & 7)) << 4) | ((kk & 7) << 1));
}

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int R = 1024, D = 128;
  std::vector<uint16_t> table((size_t)R * D);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < D; ++c) table[(size_t)r * D + c] = (uint16_t)((r * 37 + c) & 0xFFFF);
  std::vector<int> rows(128);
  for (int i = 0; i < 128; ++i) rows[i] = (i * 389 + 17) % R;
  uint16_t *d_table, *d_out; int* d_rows;
  cudaMalloc(&d_table, table.size() * 2); cudaMalloc(&d_out, 32768); cudaMalloc(&d_rows, 128 * 4);
  cudaMemcpy(d_table, table.data(), table.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_rows, rows.data(), 128 * 4, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) { printf("no encoder\n"); return 1; }
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)R};
  const cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult rc = reinterpret_cast<EncodeTiledFn>(sym)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d_table, dims, strides, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box rows %d: encode rc=%d\n", box_rows, (int)rc);
  if (rc != CUDA_SUCCESS) return 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(map, d_rows, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> out(16384);
  cudaMemcpy(out.data(), d_out, 32768, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 128 && bad < 8; ++i)
    for (int c = 0; c < D; ++c) {
      const uint16_t want = table[(size_t)rows[i] * D + c], got = out[swz_offset(i, c) / 2];
      if (want != got) { if (bad < 8) printf("row %d col %d: want %04x got %04x\n", i, c, want, got); ++bad; break; }
    }
  printf(bad ? "MISMATCH\n" : "gather4 lands in the SWIZZLE_128B operand image layout: OK\n");
  return bad ? 2 : 0;
}
