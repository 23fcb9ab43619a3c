// GroupNorm(32 groups, eps 1e-6, affine) + SiLU on NHWC bf16 — HBM-streaming kernels.
// Replaces norm1/norm2/norm_out + swish of ComfyUI's Decoder (driven through vae.decode,
// reference hdr_vae_decode.py:859,:1022).
//
//   gn_stats    : per (image, pixel chunk) partial (sum, sum of squares) of every group, fp32,
//                 16-byte loads, smem reduction; partials are written (not atomically added) so the
//                 final reduction order is fixed -> results do not depend on scheduling (and, under
//                 row tiling, not on the number of GPUs once partials are all-reduced);
//   gn_finalize : fixed-order double reduction of the partials -> per (image, channel) scale/shift;
//   gn_apply    : y = silu(x * a[c] + b[c]), 16-byte loads/stores.
// Algorithmic traffic: 2 B read (stats) + 2 B read + 2 B write (apply) per element.
#include "common.cuh"

namespace hdrvae {

constexpr int kGroups = 32;
constexpr int kGnThreads = 256;

__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, int HW, int C, int px_per_block) {
  const int img = blockIdx.y;
  const int chunk = blockIdx.x;
  const int vpp = C >> 3;                       // 16-byte vectors per pixel
  const int cpg = C / kGroups;
  const int vi = threadIdx.x % vpp;
  const int p_off = threadIdx.x / vpp;
  const int p_step = kGnThreads / vpp;
  const int p0 = chunk * px_per_block;
  const int p1 = min(HW, p0 + px_per_block);
  const uint4* base = reinterpret_cast<const uint4*>(x + (long long)img * HW * C);
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (int p = p0 + p_off; p < p1; p += p_step) {
    const uint4 v = __ldg(base + (long long)p * vpp + vi);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    const float2 c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
    s0 += (a.x + a.y) + (b.x + b.y);
    q0 += (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y);
    s1 += (c.x + c.y) + (d.x + d.y);
    q1 += (c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y);
  }
  // Fixed-order block reduction (no float atomics): every thread parks its 4 partials, then 64
  // threads (group, stat) each add the 16 entries that belong to their group in a fixed order.
  __shared__ float parked[kGnThreads][4];
  parked[threadIdx.x][0] = s0;
  parked[threadIdx.x][1] = q0;
  parked[threadIdx.x][2] = s1;
  parked[threadIdx.x][3] = q1;
  __syncthreads();
  if (threadIdx.x < kGroups * 2) {
    const int g = threadIdx.x >> 1, stat = threadIdx.x & 1;
    const int halves_per_group = cpg >> 2;              // 4-channel halves of a 16-byte vector
    float t = 0.f;
    for (int hh = g * halves_per_group; hh < (g + 1) * halves_per_group; ++hh) {
      const int v = hh >> 1, half = hh & 1;
      for (int po = 0; po < p_step; ++po) t += parked[po * vpp + v][half * 2 + stat];
    }
    partial[((long long)img * gridDim.x + chunk) * (kGroups * 2) + threadIdx.x] = t;
  }
}

// one block per image; thread (g, lane8): 32 groups x 8 lanes
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partial, int n_chunks, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift, int C,
                   double count, float eps) {
  const int img = blockIdx.x;
  const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
  double s = 0.0, q = 0.0;
  for (int c = l; c < n_chunks; c += 8) {
    const float* pp = partial + ((long long)img * n_chunks + c) * (kGroups * 2) + g * 2;
    s += (double)pp[0];
    q += (double)pp[1];
  }
  for (int o = 4; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  __shared__ float mean_s[kGroups], rstd_s[kGroups];
  if (l == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_s[g] = (float)mean;
    rstd_s[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int cpg = C / kGroups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int gg = c / cpg;
    const float a = gamma[c] * rstd_s[gg];
    scale[(long long)img * C + c] = a;
    shift[(long long)img * C + c] = beta[c] - mean_s[gg] * a;
  }
}

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.f + __expf(-v)); }

template <bool kSilu>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                const float* __restrict__ shift, int HW, int C, int px_per_block) {
  extern __shared__ float tab[];   // [2][C]
  const int img = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += kGnThreads) {
    tab[c] = scale[(long long)img * C + c];
    tab[C + c] = shift[(long long)img * C + c];
  }
  __syncthreads();
  const int vpp = C >> 3;
  const int vi = threadIdx.x % vpp;
  const int p_off = threadIdx.x / vpp;
  const int p_step = kGnThreads / vpp;
  const int p0 = blockIdx.x * px_per_block;
  const int p1 = min(HW, p0 + px_per_block);
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = tab[vi * 8 + j]; b[j] = tab[C + vi * 8 + j]; }
  const uint4* xin = reinterpret_cast<const uint4*>(x + (long long)img * HW * C);
  uint4* yout = reinterpret_cast<uint4*>(y + (long long)img * HW * C);
  for (int p = p0 + p_off; p < p1; p += p_step) {
    const uint4 v = __ldg(xin + (long long)p * vpp + vi);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      float r0 = fmaf(f.x, a[2 * e], b[2 * e]);
      float r1 = fmaf(f.y, a[2 * e + 1], b[2 * e + 1]);
      if (kSilu) { r0 = silu_f(r0); r1 = silu_f(r1); }
      __nv_bfloat162 pk = __floats2bfloat162_rn(r0, r1);
      ow[e] = *reinterpret_cast<uint32_t*>(&pk);
    }
    yout[(long long)p * vpp + vi] = o;
  }
}

// scratch layout: [partials: B * chunks * 64 floats][scale: B*C][shift: B*C]
size_t gn_scratch_bytes(int B, int C) {
  return ((size_t)B * 4096 * kGroups * 2 + (size_t)2 * B * C) * sizeof(float);
}

static void gn_chunking(int B, int HW, int C, int* chunks, int* px_per_block) {
  const int px_per_iter = kGnThreads / (C >> 3);
  int want = (148 * 8 + B - 1) / B;                 // ~8 CTAs per SM over the whole batch
  int maxc = (HW + px_per_iter * 4 - 1) / (px_per_iter * 4);   // at least 4 iterations per thread
  if (maxc < 1) maxc = 1;
  int c = want < maxc ? want : maxc;
  if (c > 4096) c = 4096;
  int ppb = (HW + c - 1) / c;
  ppb = (ppb + px_per_iter - 1) / px_per_iter * px_per_iter;
  *px_per_block = ppb;
  *chunks = (HW + ppb - 1) / ppb;
}

int launch_groupnorm(const __nv_bfloat16* x, __nv_bfloat16* y, int B, int HW, int C, const float* gamma,
                     const float* beta, bool silu, void* scratch, cudaStream_t s) {
  HDRVAE_REQUIRE(C % 32 == 0 && C >= 128 && C <= 2048 && (kGnThreads % (C >> 3)) == 0,
                 "groupnorm: unsupported channel count %d", C);
  int chunks, ppb;
  gn_chunking(B, HW, C, &chunks, &ppb);
  float* partial = reinterpret_cast<float*>(scratch);
  float* scale = partial + (size_t)B * 4096 * kGroups * 2;
  float* shift = scale + (size_t)B * C;
  gn_stats_kernel<<<dim3(chunks, B), kGnThreads, 0, s>>>(x, partial, HW, C, ppb);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  gn_finalize_kernel<<<B, 256, 0, s>>>(partial, chunks, gamma, beta, scale, shift, C,
                                       (double)HW * (double)(C / kGroups), 1e-6f);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  if (silu)
    gn_apply_kernel<true><<<dim3(chunks, B), kGnThreads, 2 * C * sizeof(float), s>>>(x, y, scale, shift, HW, C, ppb);
  else
    gn_apply_kernel<false><<<dim3(chunks, B), kGnThreads, 2 * C * sizeof(float), s>>>(x, y, scale, shift, HW, C, ppb);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace hdrvae
