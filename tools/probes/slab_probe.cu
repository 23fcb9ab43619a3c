// Probe: can a UMMA K-major SWIZZLE_128B A-descriptor start at a 128-byte line that is NOT a multiple of 8 lines
// (1024 B), with a stride byte offset other than 1024?  That is what addressing the 9 taps of a 3x3 conv as shifted
// views of ONE shared-memory activation slab needs.
//   smem: 256 lines x 128 B loaded by one TMA box (swizzled by absolute address), interpreted as 16 slab rows at a
//   16-line pitch.  MMA M=128: atom a (8 rows) = lines a*16 + s .. a*16 + s + 7, s in {0,1,2}; SBO = 2048 B.
//   B = 64x64 identity, so D[m][n] must equal A_global[line(m)][n].
// Variants: base_offset field = 0 or = s.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../vae_decode_hdr_b200/csrc/ptx.cuh"



__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                int shift, int use_base_off, float* out /*[128][64]*/) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 256 lines x 128 B = 32 KB
  uint8_t* sb = smem + 32768;         // 64 lines x 128 B = 8 KB
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_full, 1);
    ptx::mbar_init(&bar_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<64>(&tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_full, 32768 + 8192);
    ptx::tma_load_2d(sa, &tmA, &bar_full, 0, 0);
    ptx::tma_load_2d(sb, &tmB, &bar_full, 0, 0);
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after_sync();
    const uint32_t idesc = ptx::make_idesc(0u, 128, 64);
    const uint32_t a0 = ptx::smem_u32(sa) + shift * 128;
    const uint64_t da = make_desc(a0, 2048, use_base_off ? shift : 0);
    const uint64_t db = make_desc(ptx::smem_u32(sb), 1024, 0);
    for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, k ? 1u : 0u);
    ptx::umma_commit(&bar_done);
  }
  __syncthreads();
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait(v);
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<64>(tmem);
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int LA = 256;
  __half* hA = (__half*)malloc(LA * 64 * 2);
  __half* hB = (__half*)malloc(64 * 64 * 2);
  for (int l = 0; l < LA; ++l)
    for (int c = 0; c < 64; ++c) hA[l * 64 + c] = __float2half((float)(l * 4 + (c % 4)) * 0.25f + (c / 4) * 256.f * 0.0f);
  // value encodes the line exactly: line*1 + (c%4)*0.25 ... keep it simple: A[l][c] = l + c/64
  for (int l = 0; l < LA; ++l)
    for (int c = 0; c < 64; ++c) hA[l * 64 + c] = __float2half((float)l + (float)c / 64.f);
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2half(n == k ? 1.f : 0.f);
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, LA * 64 * 2); cudaMalloc(&dB, 64 * 64 * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA, LA * 64 * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, 64 * 64 * 2, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_enc enc = (PFN_enc)fn;
  CUtensorMap mA, mB;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)LA}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {64, 64}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  float* hO = (float*)malloc(128 * 64 * 4);
  for (int shift = 0; shift < 8; ++shift)
    for (int bo = 0; bo < 2; ++bo) {
      cudaMemset(dO, 0, 128 * 64 * 4);
      probe<<<1, 128, 42 * 1024>>>(mA, mB, shift, bo, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_off %d: CUDA error %s\n", shift, bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO, dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first_bad = -1;
      for (int m = 0; m < 128; ++m) {
        const int line = (m / 8) * 16 + (m % 8) + shift;
        for (int n = 0; n < 64; ++n) {
          const float want = __half2float(__float2half((float)line + (float)n / 64.f));
          if (hO[m * 64 + n] != want) { if (first_bad < 0) first_bad = m * 64 + n; ++bad; }
        }
      }
      printf("shift %d base_offset_field %d: %s (%d mismatches", shift, bo ? shift : 0, bad ? "WRONG" : "exact", bad);
      if (bad) printf("; first at m=%d n=%d got %.4f", first_bad / 64, first_bad % 64, hO[first_bad]);
      printf(")\n");
    }
  return 0;
}
