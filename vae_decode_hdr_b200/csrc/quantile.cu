// Exact quantiles by MSD radix select (build-side extension named in BASELINE.json north_star: "quantiles by
// radix select"; the reference computes no quantile, SURVEY.md 0.6, so parity is pinned against
// torch.kthvalue on the same tensor: the k-th order statistic is exact, hence bit-exact).
//
// Floats are mapped to order-preserving uint32 keys; three histogram passes (11 + 11 + 10 bits) narrow the
// key of the element of rank k: pass p counts, among the elements whose already-fixed prefix matches, the
// next digit; a single-block scan picks the digit bucket containing rank k.  Every quantile is an
// independent (prefix, rank) pair, all quantiles share each pass over the data (one HBM read per pass).
#include "common.cuh"

namespace hdrvae {

constexpr int kMaxQ = 8;
constexpr int kBins = 2048;

struct SelectState {
  uint32_t prefix[kMaxQ];        // fixed high bits of each quantile's key
  unsigned long long rank[kMaxQ];// remaining rank inside the prefix bucket
};

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending float order == ascending key order
}
__device__ __forceinline__ float key_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// pass 0: shift 21 (11 bits), pass 1: shift 10 (11 bits), pass 2: shift 0 (10 bits)
__global__ void __launch_bounds__(256)
radix_hist_kernel(const float* __restrict__ x, long long n, const SelectState* __restrict__ st, int nq, int pass,
                  unsigned int* __restrict__ hist /*[nq][kBins]*/) {
  extern __shared__ unsigned int sh[];                       // [nq][kBins]
  for (int i = threadIdx.x; i < nq * kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
  const uint32_t dmask = pass == 2 ? 0x3ffu : 0x7ffu;
  const uint32_t pmask = pass == 0 ? 0u : (pass == 1 ? 0xffe00000u : 0xfffffc00u);
  uint32_t pre[kMaxQ];
  for (int q = 0; q < nq; ++q) pre[q] = st->prefix[q];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t k = float_key(x[i]);
    for (int q = 0; q < nq; ++q)
      if ((k & pmask) == pre[q]) atomicAdd(&sh[q * kBins + ((k >> shift) & dmask)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nq * kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one block per quantile: find the bucket that holds the remaining rank, extend the prefix
__global__ void __launch_bounds__(256)
radix_pick_kernel(unsigned int* __restrict__ hist, SelectState* __restrict__ st, int pass, float* __restrict__ out) {
  const int q = blockIdx.x;
  __shared__ unsigned long long part[256];
  const int bins = pass == 2 ? 1024 : 2048;
  const int per = bins / 256;
  unsigned long long s = 0;
  for (int j = 0; j < per; ++j) s += hist[q * kBins + threadIdx.x * per + j];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long r = st->rank[q], acc = 0;
    int t = 0;
    while (t < 255 && acc + part[t] <= r) { acc += part[t]; ++t; }
    int b = t * per;
    while (b < bins - 1 && acc + hist[q * kBins + b] <= r) { acc += hist[q * kBins + b]; ++b; }
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    st->prefix[q] |= (uint32_t)b << shift;
    st->rank[q] = r - acc;
    if (pass == 2) out[q] = key_float(st->prefix[q]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += 256) hist[q * kBins + i] = 0;     // ready for the next pass
}

__global__ void radix_init_kernel(SelectState* st, const unsigned long long* ranks, int nq, unsigned int* hist) {
  for (int i = threadIdx.x; i < nq; i += blockDim.x) { st->prefix[i] = 0; st->rank[i] = ranks[i]; }
  for (int i = threadIdx.x; i < nq * kBins; i += blockDim.x) hist[i] = 0;
}

size_t quantile_scratch_bytes() { return sizeof(SelectState) + kMaxQ * sizeof(unsigned long long) + kMaxQ * kBins * sizeof(unsigned int) + 256; }

// x: device fp32 [n]; ranks_host[q] = 0-based rank of the order statistic wanted; out: device float [nq]
int launch_quantiles(const float* x, long long n, const unsigned long long* ranks_host, int nq, float* out, void* scratch,
                     cudaStream_t s) {
  HDRVAE_REQUIRE(nq >= 1 && nq <= kMaxQ && n >= 1, "quantiles: need 1..%d quantiles of a non-empty tensor", kMaxQ);
  uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
  SelectState* st = reinterpret_cast<SelectState*>(p); p += (sizeof(SelectState) + 127) / 128 * 128;
  unsigned long long* ranks = reinterpret_cast<unsigned long long*>(p); p += kMaxQ * sizeof(unsigned long long);
  unsigned int* hist = reinterpret_cast<unsigned int*>(p);
  HDRVAE_CUDA_OK(cudaMemcpyAsync(ranks, ranks_host, nq * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
  radix_init_kernel<<<1, 256, 0, s>>>(st, ranks, nq, hist);
  HDRVAE_LAUNCHED();
  int grid = ceil_div(n, 256 * 16);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  static PerDeviceOnce once;
  if (once.first()) HDRVAE_CUDA_OK(cudaFuncSetAttribute(radix_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxQ * kBins * 4));
  for (int pass = 0; pass < 3; ++pass) {
    radix_hist_kernel<<<grid, 256, nq * kBins * sizeof(unsigned int), s>>>(x, n, st, nq, pass, hist);
    HDRVAE_LAUNCHED();
    radix_pick_kernel<<<nq, 256, 0, s>>>(hist, st, pass, out);
    HDRVAE_LAUNCHED();
  }
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace hdrvae
