// Fused HDR epilogue (reference hdr_vae_decode.py:837-925 analyze_conv_out, :1009-1161
// intelligent_hdr_decode, :941-1007 mode helpers, :1163-1203 srgb_to_linear, :62-195 orchestration).
//
// The reference needs ~25 full-tensor passes + host syncs, a second decoder run and a third
// conv_out.  Here:
//   phase A  (one pass over the 128-channel activations `pre`):
//            conv_out 3x3 (fp32 CUDA cores, N = 3 is not a tensor-core shape) -> post = clamp((c+1)/2,0,1),
//            128->3 channel MAX-pool (+ first-max index), min/max/sum/sumsq of pre, post, conv, pre3 and the
//            highlight count, as per-block partials;
//   reduce   fixed-order reduction of the partials -> hdrvae_raw_stats (all-reducible across GPUs);
//   scalars  batch-global scalars: SIGMOID/TANH detection, has_hdr, recovered min/max (monotone map of
//            post min/max), range, mean alignment, adaptive compression factor;
//   phase B  elementwise: srgb->linear, logit/atanh recovery, min-max renormalise, mode formula, multiplier,
//            output statistics (HDR / negative pixel counts, min, max).
// Algorithmic traffic: (2 or 4) * 128 B/px read + 24 B/px written (A) + 24 B/px read + 12 B/px written (B).
//
// Arithmetic follows the reference's fp32 op order; __fmul_rn/__fadd_rn keep nvcc from contracting
// mul+add pairs into FMAs the reference does not perform.
#include "../../include/hdrvae.h"
#include "common.cuh"

namespace hdrvae {

// Phase A tile: 64 x 16 output pixels per 256-thread block, 4 horizontally adjacent pixels per thread (register
// blocking: 15 shared-memory loads per 108 FMAs; the one-pixel-per-thread version was LSU-wavefront bound, ncu 92 %).
constexpr int kTileW = 64, kTileH = 16;
constexpr int kPxPerThread = 4;
constexpr int kHaloW = kTileW + 2, kHaloH = kTileH + 2;   // 66 x 18
constexpr int kRowPitch = 68;                        // floats; multiple of 4 so the 16-byte / 8-byte reads are aligned
constexpr int kTilePx = kHaloW * kHaloH;             // 1188 staged pixels
constexpr int kChanStride = kRowPitch * kHaloH;      // 1224 floats (multiple of 4: aligned vector reads)
constexpr int kChunk = 16;                           // channels staged per pass
constexpr int kC = 128;
constexpr int kEpiThreads = (kTileW / kPxPerThread) * kTileH;   // 256

struct PartialA {
  float vmin[4];   // pre, post, conv, pre3
  float vmax[4];
  double vsum[8];  // pre Sx, pre Sxx, post Sx, post Sxx, conv Sx, n_pre, n_post, highlight_count
};

struct PartialB {
  float omin, omax, imax, pad;
  unsigned long long hdr, neg, ihdr, pad2;
};

struct HdrScalars {
  int norm_function, has_hdr;
  float pre_min, pre_max, pre_mean, range;     // 128-channel stats as fp32 (python floats in the reference)
  float rec_min, rec_span;                     // min(recovered), max - min
  float rec_max, aligned_max, cf;
  int cf_is_tensor;                            // adaptive: compression factor computed (else python 1.0)
};

template <typename T> struct Load4;
template <> struct Load4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
};
template <> struct Load4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

template <> struct Load4<__half> {
  static __device__ __forceinline__ float4 ld(const __half* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

// raw 8-byte piece (4 x 16-bit) -> 4 floats
template <typename T> struct Cvt4 {
  static __device__ __forceinline__ float4 cvt(const uint2&) { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct Cvt4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 cvt(const uint2& r) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};
template <> struct Cvt4<__half> {
  static __device__ __forceinline__ float4 cvt(const uint2& r) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

__device__ __forceinline__ float warp_min(float v) { for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(~0u, v, o)); return v; }
__device__ __forceinline__ float warp_max(float v) { for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(~0u, v, o)); return v; }
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(~0u, v, o); return v; }

// ------------------------------------------------------------------------------------ phase A
template <bool kArgmax>
__device__ __forceinline__ void maxpool_update(float v, int c, float& m, int& idx) {
  if (kArgmax) { if (v > m) { m = v; idx = c; } }
  else m = fmaxf(m, v);
}

template <typename T, bool kArgmax>
__global__ void __launch_bounds__(kEpiThreads, 2)
hdr_phase_a_kernel(const T* __restrict__ pre, const float* __restrict__ conv_w /*OIHW [3][128][3][3]*/,
                   const float* __restrict__ conv_b, int H, int W, float* __restrict__ post3,
                   float* __restrict__ pre3, int* __restrict__ argmax3, PartialA* __restrict__ partials, int y_pad,
                   long long img_stride) {
  extern __shared__ float sm[];
  float4* wsm = reinterpret_cast<float4*>(sm);                 // [9][128] (w_r, w_g, w_b, 0)
  float* tile = sm + 9 * kC * 4;                               // [kChunk][kChanStride]: [channel][row][col]
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int tid = threadIdx.x;

  for (int i = tid; i < 9 * kC; i += kEpiThreads) {
    const int tap = i / kC, c = i % kC;
    wsm[i] = make_float4(conv_w[(0 * kC + c) * 9 + tap], conv_w[(1 * kC + c) * 9 + tap],
                         conv_w[(2 * kC + c) * 9 + tap], 0.f);
  }

  const int tcx = tid % (kTileW / kPxPerThread), py = tid / (kTileW / kPxPerThread);
  const int gx0 = x0 + tcx * kPxPerThread, gy = y0 + py;      // first of this thread's 4 pixels
  float acc[kPxPerThread][3];
  float mx[kPxPerThread][3];
  int am[kPxPerThread][3];
#pragma unroll
  for (int q = 0; q < kPxPerThread; ++q) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc[q][k] = 0.f; mx[q][k] = -INFINITY; am[q][k] = 42 * k; }
  }
  float smin = INFINITY, smax = -INFINITY, ssum = 0.f, ssq = 0.f;      // pre stats (this thread's loads)

  // `pre` points at interior row 0; y_pad halo rows above/below hold the neighbour ranks' rows (row tiling)
  const T* img_base = pre + (long long)img * img_stride;
  // Staging loads are software-pipelined for the 16-bit input types (the product path): the raw 8-byte pieces of
  // chunk k+1 are fetched into registers while chunk k is being computed (ncu: the convert-after-load was 28 % of
  // all stall samples); fp32 input (parity entry) loads in place.
  constexpr bool kPrefetch = sizeof(T) == 2;
  constexpr int kPieces = (kTilePx * (kChunk / 4) + kEpiThreads - 1) / kEpiThreads;     // 19
  uint2 pf[kPrefetch ? kPieces : 1];
  auto piece_src = [&](int i, int ch0, bool* ok) -> const T* {
    const int quad = i / kTilePx;
    const int tp = i - quad * kTilePx;
    const int tx = tp % kHaloW, ty = tp / kHaloW;
    const int sx = x0 + tx - 1, sy = y0 + ty - 1;
    *ok = i < kTilePx * (kChunk / 4) && sx >= 0 && sx < W && sy >= -y_pad && sy < H + y_pad;
    return img_base + ((long long)sy * W + sx) * kC + ch0 + quad * 4;
  };
  if (kPrefetch) {
#pragma unroll
    for (int u = 0; u < kPieces; ++u) {
      bool ok;
      const T* src = piece_src(tid + u * kEpiThreads, 0, &ok);
      pf[u] = ok ? __ldg(reinterpret_cast<const uint2*>(src)) : make_uint2(0u, 0u);
    }
  }
  for (int ch0 = 0; ch0 < kC; ch0 += kChunk) {
    __syncthreads();
    // stage [kHaloH x kHaloW] pixels x 16 channels, transposed to [channel][row][col]
    // (pixel-major lane order: a warp stores 32 consecutive pixels of one channel -> conflict-free)
#pragma unroll
    for (int u = 0; u < kPieces; ++u) {
      const int i = tid + u * kEpiThreads;
      if (i >= kTilePx * (kChunk / 4)) break;
      const int quad = i / kTilePx;
      const int tp = i - quad * kTilePx;
      const int tx = tp % kHaloW, ty = tp / kHaloW;
      const int sx = x0 + tx - 1, sy = y0 + ty - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sx >= 0 && sx < W && sy >= -y_pad && sy < H + y_pad) {
        if (kPrefetch) v = Cvt4<T>::cvt(pf[u]);
        else v = Load4<T>::ld(img_base + ((long long)sy * W + sx) * kC + ch0 + quad * 4);
        if (tx >= 1 && tx <= kTileW && ty >= 1 && ty <= kTileH && sy >= 0 && sy < H) {   // pixel owned by this block
          smin = fminf(smin, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
          smax = fmaxf(smax, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
          ssum += (v.x + v.y) + (v.z + v.w);
          ssq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
      }
      float* d = tile + (quad * 4) * kChanStride + ty * kRowPitch + tx;
      d[0] = v.x; d[kChanStride] = v.y; d[2 * kChanStride] = v.z; d[3 * kChanStride] = v.w;
    }
    __syncthreads();
    if (kPrefetch && ch0 + kChunk < kC) {
#pragma unroll
      for (int u = 0; u < kPieces; ++u) {
        bool ok;
        const T* src = piece_src(tid + u * kEpiThreads, ch0 + kChunk, &ok);
        pf[u] = ok ? __ldg(reinterpret_cast<const uint2*>(src)) : make_uint2(0u, 0u);
      }
    }
    const float* tc = tile + py * kRowPitch + tcx * kPxPerThread;    // halo row py = image row gy-1, col = gx0-1
#pragma unroll 1
    for (int c = 0; c < kChunk; ++c) {
      const int cg = ch0 + c;
      // 3 rows x 6 columns of this channel around the thread's 4 pixels
      float a[3][6];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* rp = tc + c * kChanStride + r * kRowPitch;
        const float4 lo = *reinterpret_cast<const float4*>(rp);
        const float2 hi = *reinterpret_cast<const float2*>(rp + 4);
        a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = lo.z; a[r][3] = lo.w; a[r][4] = hi.x; a[r][5] = hi.y;
      }
      // channel MAX-pool 0-41 / 42-83 / 84-125 (hdr_vae_decode.py:1044-1051); strict > keeps the first max.
      // The group is uniform per channel: one branch, then one FMNMX per pixel (compare+select with the index).
      if (cg < 42) {
#pragma unroll
        for (int q = 0; q < kPxPerThread; ++q) maxpool_update<kArgmax>(a[1][q + 1], cg, mx[q][0], am[q][0]);
      } else if (cg < 84) {
#pragma unroll
        for (int q = 0; q < kPxPerThread; ++q) maxpool_update<kArgmax>(a[1][q + 1], cg, mx[q][1], am[q][1]);
      } else if (cg < 126) {
#pragma unroll
        for (int q = 0; q < kPxPerThread; ++q) maxpool_update<kArgmax>(a[1][q + 1], cg, mx[q][2], am[q][2]);
      }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4 w = wsm[(dy * 3 + dx) * kC + cg];
#pragma unroll
          for (int q = 0; q < kPxPerThread; ++q) {
            const float v = a[dy][q + dx];
            acc[q][0] = fmaf(v, w.x, acc[q][0]);
            acc[q][1] = fmaf(v, w.y, acc[q][1]);
            acc[q][2] = fmaf(v, w.z, acc[q][2]);
          }
        }
    }
  }

  // per-pixel results
  float cmin = INFINITY, cmax = -INFINITY, csum = 0.f;
  float pmin = INFINITY, pmax = -INFINITY, psum = 0.f, psq = 0.f;
  float p3min = INFINITY, p3max = -INFINITY;
  float hl = 0.f, nlive = 0.f;
  const float cb[3] = {conv_b[0], conv_b[1], conv_b[2]};
#pragma unroll
  for (int q = 0; q < kPxPerThread; ++q) {
    const int gx = gx0 + q;
    if (gx < W && gy < H) {
      const long long o = (((long long)img * H + gy) * W + gx) * 3;
      nlive += 3.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float cv = acc[q][k] + cb[k];
        // comfy.sd.VAE.decode process_output: clamp((x + 1) / 2, 0, 1)
        const float sv = fminf(fmaxf(__fdiv_rn(__fadd_rn(cv, 1.0f), 2.0f), 0.f), 1.f);
        post3[o + k] = sv;
        pre3[o + k] = mx[q][k];
        if (kArgmax) argmax3[o + k] = am[q][k];
        cmin = fminf(cmin, cv); cmax = fmaxf(cmax, cv); csum += cv;
        pmin = fminf(pmin, sv); pmax = fmaxf(pmax, sv); psum += sv; psq += sv * sv;
        p3min = fminf(p3min, mx[q][k]); p3max = fmaxf(p3max, mx[q][k]);
        hl += mx[q][k] > 1.0f ? 1.f : 0.f;
      }
    }
  }

  // block reduction -> one PartialA per block (fixed order: warp shuffles, then warp 0 over 8 warps)
  __shared__ float rmin[4][8], rmax[4][8];
  __shared__ double rsum[8][8];
  const int warp = tid >> 5, lane = tid & 31;
  const float mins[4] = {warp_min(smin), warp_min(pmin), warp_min(cmin), warp_min(p3min)};
  const float maxs[4] = {warp_max(smax), warp_max(pmax), warp_max(cmax), warp_max(p3max)};
  const double sums[8] = {warp_sum((double)ssum), warp_sum((double)ssq), warp_sum((double)psum), warp_sum((double)psq),
                          warp_sum((double)csum), 0.0, warp_sum((double)nlive), warp_sum((double)hl)};
  if (lane == 0) {
    for (int k = 0; k < 4; ++k) { rmin[k][warp] = mins[k]; rmax[k][warp] = maxs[k]; }
    for (int k = 0; k < 8; ++k) rsum[k][warp] = sums[k];
  }
  __syncthreads();
  if (tid == 0) {
    PartialA pa;
    for (int k = 0; k < 4; ++k) {
      float a = rmin[k][0], b = rmax[k][0];
      for (int w2 = 1; w2 < 8; ++w2) { a = fminf(a, rmin[k][w2]); b = fmaxf(b, rmax[k][w2]); }
      pa.vmin[k] = a; pa.vmax[k] = b;
    }
    for (int k = 0; k < 8; ++k) {
      double a = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) a += rsum[k][w2];
      pa.vsum[k] = a;
    }
    const int ix = min(kTileW, W - x0), iy = min(kTileH, H - y0);
    pa.vsum[5] = (double)ix * iy * kC;        // number of pre elements owned by this block
    partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = pa;
  }
}

// ------------------------------------------------------------------------------------ phase A, conv_out precomputed
// Product path for 16-bit features: conv_out (128 -> 3, 3x3) has already run on the tensor cores (gemm_tc.cu slab
// variant) with the fp32 weights split into fp16 hi + lo rows, so conv8[px] = {hi r,g,b,0, lo r,g,b,0} and
// conv = hi + lo + bias reproduces the fp32 convolution of the 16-bit features to ~1e-7.  What is left is one
// streaming pass over `pre`: the 128 -> 3 channel MAX-pool with first-max index and the statistics.  Same 64 x 16
// pixel block per CTA and the same PartialA layout as hdr_phase_a_kernel; 8 lanes share a pixel (16 channels each:
// a warp load covers 4 pixels x 128 contiguous bytes).
template <typename T, bool kArgmax>
__global__ void __launch_bounds__(kEpiThreads)
hdr_phase_a_pre_kernel(const T* __restrict__ pre, const float4* __restrict__ conv8, const float* __restrict__ conv_b, int H,
                       int W, float* __restrict__ post3, float* __restrict__ pre3, int* __restrict__ argmax3,
                       PartialA* __restrict__ partials, long long img_stride) {
  const int img = blockIdx.z;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane & 7;                 // which 16 channels of the pixel
  const int pq = lane >> 3;                 // which of the warp's 4 pixels
  const T* img_base = pre + (long long)img * img_stride;
  const float cb[3] = {conv_b[0], conv_b[1], conv_b[2]};
  float smin = INFINITY, smax = -INFINITY, ssum = 0.f, ssq = 0.f;
  float cmin = INFINITY, cmax = -INFINITY, csum = 0.f;
  float pmin = INFINITY, pmax = -INFINITY, psum = 0.f, psq = 0.f;
  float p3min = INFINITY, p3max = -INFINITY;
  float hl = 0.f, nlive = 0.f;
  // block pixels in row-major order, 32 per iteration (8 warps x 4).  Lane `sub` holds channels 8 sub .. 8 sub + 7 and
  // 64 + 8 sub .. 64 + 8 sub + 7: each of the two 16-byte loads of a warp covers 4 pixels x 128 contiguous bytes.  The
  // loads of iteration it+1 are issued before iteration it is processed (the pass is latency bound otherwise).
  constexpr int kIters = kTileW * kTileH / 32;
  auto px_of = [&](int it, int* gx, int* gy) {
    const int bp = it * 32 + warp * 4 + pq;
    *gx = x0 + (bp % kTileW); *gy = y0 + (bp / kTileW);
    return *gx < W && *gy < H;
  };
  uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
  {
    int gx, gy;
    if (px_of(0, &gx, &gy)) {
      const uint4* src = reinterpret_cast<const uint4*>(img_base + ((long long)gy * W + gx) * kC + sub * 8);
      n0 = __ldg(src); n1 = __ldg(src + 8);
    }
  }
  for (int it = 0; it < kIters; ++it) {
    int gx, gy;
    const bool live = px_of(it, &gx, &gy);
    const uint4 r0 = n0, r1 = n1;
    if (it + 1 < kIters) {
      int nx, ny;
      if (px_of(it + 1, &nx, &ny)) {
        const uint4* src = reinterpret_cast<const uint4*>(img_base + ((long long)ny * W + nx) * kC + sub * 8);
        n0 = __ldg(src); n1 = __ldg(src + 8);
      }
    }
    float v[16];
    {
      const float4 a = Cvt4<T>::cvt(make_uint2(r0.x, r0.y)), b = Cvt4<T>::cvt(make_uint2(r0.z, r0.w));
      const float4 c = Cvt4<T>::cvt(make_uint2(r1.x, r1.y)), d = Cvt4<T>::cvt(make_uint2(r1.z, r1.w));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
    }
    float mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int am[3] = {0, 42, 84};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = (j < 8 ? 0 : 64) + sub * 8 + (j & 7);
      if (live) {
        smin = fminf(smin, v[j]); smax = fmaxf(smax, v[j]); ssum += v[j]; ssq = fmaf(v[j], v[j], ssq);
      }
      // channel MAX-pool 0-41 / 42-83 / 84-125 (hdr_vae_decode.py:1044-1051); strict > keeps the first max.
      // Predicated, not branched: the 8 lanes of a pixel sit in different groups.
      const bool in0 = c < 42, in1 = c >= 42 && c < 84, in2 = c >= 84 && c < 126;
      if (kArgmax) {
        if (in0 && v[j] > mx[0]) { mx[0] = v[j]; am[0] = c; }
        if (in1 && v[j] > mx[1]) { mx[1] = v[j]; am[1] = c; }
        if (in2 && v[j] > mx[2]) { mx[2] = v[j]; am[2] = c; }
      } else {
        mx[0] = in0 ? fmaxf(mx[0], v[j]) : mx[0];
        mx[1] = in1 ? fmaxf(mx[1], v[j]) : mx[1];
        mx[2] = in2 ? fmaxf(mx[2], v[j]) : mx[2];
      }
    }
    // combine the 8 lanes of the pixel: larger value wins, equal values keep the smaller channel index
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float ov = __shfl_xor_sync(~0u, mx[k], o);
        const int oi = __shfl_xor_sync(~0u, am[k], o);
        if (ov > mx[k] || (kArgmax && ov == mx[k] && oi < am[k])) { mx[k] = ov; am[k] = oi; }
      }
    }
    if (live && sub == 0) {
      const long long px = ((long long)img * H + gy) * W + gx;
      const float4 hi = conv8[px * 2], lo = conv8[px * 2 + 1];
      const float cvv[3] = {(hi.x + lo.x) + cb[0], (hi.y + lo.y) + cb[1], (hi.z + lo.z) + cb[2]};
      nlive += 3.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float cv = cvv[k];
        // comfy.sd.VAE.decode process_output: clamp((x + 1) / 2, 0, 1)
        const float sv = fminf(fmaxf(__fdiv_rn(__fadd_rn(cv, 1.0f), 2.0f), 0.f), 1.f);
        post3[px * 3 + k] = sv;
        pre3[px * 3 + k] = mx[k];
        if (kArgmax) argmax3[px * 3 + k] = am[k];
        cmin = fminf(cmin, cv); cmax = fmaxf(cmax, cv); csum += cv;
        pmin = fminf(pmin, sv); pmax = fmaxf(pmax, sv); psum += sv; psq += sv * sv;
        p3min = fminf(p3min, mx[k]); p3max = fmaxf(p3max, mx[k]);
        hl += mx[k] > 1.0f ? 1.f : 0.f;
      }
    }
  }
  __shared__ float rmin[4][8], rmax[4][8];
  __shared__ double rsum[8][8];
  const float mins[4] = {warp_min(smin), warp_min(pmin), warp_min(cmin), warp_min(p3min)};
  const float maxs[4] = {warp_max(smax), warp_max(pmax), warp_max(cmax), warp_max(p3max)};
  const double sums[8] = {warp_sum((double)ssum), warp_sum((double)ssq), warp_sum((double)psum), warp_sum((double)psq),
                          warp_sum((double)csum), 0.0, warp_sum((double)nlive), warp_sum((double)hl)};
  if (lane == 0) {
    for (int k = 0; k < 4; ++k) { rmin[k][warp] = mins[k]; rmax[k][warp] = maxs[k]; }
    for (int k = 0; k < 8; ++k) rsum[k][warp] = sums[k];
  }
  __syncthreads();
  if (tid == 0) {
    PartialA pa;
    for (int k = 0; k < 4; ++k) {
      float a = rmin[k][0], b = rmax[k][0];
      for (int w2 = 1; w2 < 8; ++w2) { a = fminf(a, rmin[k][w2]); b = fmaxf(b, rmax[k][w2]); }
      pa.vmin[k] = a; pa.vmax[k] = b;
    }
    for (int k = 0; k < 8; ++k) {
      double a = 0.0;
      for (int w2 = 0; w2 < 8; ++w2) a += rsum[k][w2];
      pa.vsum[k] = a;
    }
    const int ix = min(kTileW, W - x0), iy = min(kTileH, H - y0);
    pa.vsum[5] = (double)ix * iy * kC;        // number of pre elements owned by this block
    partials[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = pa;
  }
}

// ------------------------------------------------------------------------------------ reduce partials
__global__ void __launch_bounds__(256)
hdr_reduce_a_kernel(const PartialA* __restrict__ partials, int n, hdrvae_raw_stats* __restrict__ raw) {
  __shared__ float rmin[4][256], rmax[4][256];
  __shared__ double rsum[8][256];
  float mn[4], mx[4];
  double sm[8];
  for (int k = 0; k < 4; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; }
  for (int k = 0; k < 8; ++k) sm[k] = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const PartialA pa = partials[i];
    for (int k = 0; k < 4; ++k) { mn[k] = fminf(mn[k], pa.vmin[k]); mx[k] = fmaxf(mx[k], pa.vmax[k]); }
    for (int k = 0; k < 8; ++k) sm[k] += pa.vsum[k];
  }
  for (int k = 0; k < 4; ++k) { rmin[k][threadIdx.x] = mn[k]; rmax[k][threadIdx.x] = mx[k]; }
  for (int k = 0; k < 8; ++k) rsum[k][threadIdx.x] = sm[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      for (int k = 0; k < 4; ++k) {
        rmin[k][threadIdx.x] = fminf(rmin[k][threadIdx.x], rmin[k][threadIdx.x + s]);
        rmax[k][threadIdx.x] = fmaxf(rmax[k][threadIdx.x], rmax[k][threadIdx.x + s]);
      }
      for (int k = 0; k < 8; ++k) rsum[k][threadIdx.x] += rsum[k][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) { raw->vmin[k] = rmin[k][0]; raw->vmax[k] = rmax[k][0]; }
    for (int k = 0; k < 8; ++k) raw->vsum[k] = rsum[k][0];
  }
}

// ------------------------------------------------------------------------------------ device math shared by scalars / phase B
__device__ __forceinline__ float inverse_sigmoid_f(float x) {          // hdr_vae_decode.py:927-932
  const float c = fminf(fmaxf(x, 1e-7f), (float)(1.0 - 1e-7));
  return logf(__fdiv_rn(c, __fsub_rn(1.0f, c)));                        // torch.logit = log(x / (1 - x))
}
__device__ __forceinline__ float inverse_tanh_f(float x) {             // hdr_vae_decode.py:934-939
  const float c = fminf(fmaxf(x, (float)(-1.0 + 1e-6)), (float)(1.0 - 1e-6));
  return atanhf(c);
}
__device__ __forceinline__ float recover_f(float s, int norm) {        // hdr_vae_decode.py:1085-1093
  return norm == HDRVAE_NORM_TANH ? inverse_tanh_f(s) : (norm == HDRVAE_NORM_SIGMOID ? inverse_sigmoid_f(s) : s);
}
__device__ __forceinline__ float srgb_to_linear_f(float s) {           // hdr_vae_decode.py:1163-1203
  const float a = fabsf(s);
  const float lin = a <= 0.04045f ? __fdiv_rn(a, 12.92f) : powf(__fdiv_rn(__fadd_rn(a, 0.055f), 1.055f), 2.4f);
  const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
  return __fmul_rn(sg, lin);
}
__device__ __forceinline__ float map_recovered_f(float rec, const HdrScalars& sc) {   // :1098-1099
  const float rn = __fdiv_rn(__fsub_rn(rec, sc.rec_min), sc.rec_span);
  return __fadd_rn(__fmul_rn(rn, sc.range), sc.pre_min);
}
__device__ __forceinline__ float aligned_f(float map, const HdrScalars& sc) {         // :1102
  return __fadd_rn(__fsub_rn(map, sc.pre_mean), 1.0f);
}
__device__ __forceinline__ float ev_multiplier_f(float v) {                            // 2^(log2(max(v, 0.001)))
  return exp2f(log2f(fmaxf(v, 0.001f)));
}

__global__ void hdr_scalars_kernel(const hdrvae_raw_stats* __restrict__ raw, HdrScalars* __restrict__ sc,
                                   hdrvae_stats* __restrict__ st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double n_pre = raw->vsum[5], n_post = raw->vsum[6];
  hdrvae_stats s;
  s.pre_min = raw->vmin[0]; s.pre_max = raw->vmax[0];
  s.pre_mean = (double)(float)(raw->vsum[0] / n_pre);                   // torch.mean returns fp32
  {
    const double m = raw->vsum[0] / n_pre;
    double var = (raw->vsum[1] - n_pre * m * m) / (n_pre - 1.0);       // torch.std: unbiased (:865)
    s.pre_std = var > 0.0 ? sqrt(var) : 0.0;
  }
  s.post_min = raw->vmin[1]; s.post_max = raw->vmax[1];
  s.post_mean = (double)(float)(raw->vsum[2] / n_post);
  {
    const double m = raw->vsum[2] / n_post;
    double var = (raw->vsum[3] - n_post * m * m) / (n_post - 1.0);
    s.post_std = var > 0.0 ? sqrt(var) : 0.0;
  }
  s.conv_min = raw->vmin[2]; s.conv_max = raw->vmax[2];
  s.conv_mean = (double)(float)(raw->vsum[4] / n_post);
  s.pre3_min = raw->vmin[3]; s.pre3_max = raw->vmax[3];
  s.highlight_count = (long long)(raw->vsum[7] + 0.5);
  // analyze_conv_out :890-897 (python doubles of fp32 values)
  int norm = HDRVAE_NORM_NONE;
  if (fabs(s.post_max - 1.0) < 1e-3 && fabs(s.post_min - 0.0) < 1e-3) norm = HDRVAE_NORM_SIGMOID;
  else if (fabs(s.post_max - 1.0) < 1e-3 && fabs(s.post_min + 1.0) < 1e-3) norm = HDRVAE_NORM_TANH;
  s.norm_function = norm;
  s.has_hdr = s.pre3_max > (1.0 + 1e-3) ? 1 : 0;                        // :1076-1078
  HdrScalars h;
  h.norm_function = norm; h.has_hdr = s.has_hdr;
  h.pre_min = (float)s.pre_min; h.pre_max = (float)s.pre_max; h.pre_mean = (float)s.pre_mean;
  h.range = (float)(s.pre_max - s.pre_min);                             // :1097 python double -> fp32 scalar
  // recovered = f(post) with f monotone non-decreasing => min/max(recovered) = f(min/max(post))  (:1098)
  h.rec_min = recover_f(raw->vmin[1], norm);
  h.rec_max = recover_f(raw->vmax[1], norm);
  h.rec_span = __fsub_rn(h.rec_max, h.rec_min);
  s.rec_min = h.rec_min; s.rec_max = h.rec_max;
  // adaptive (:1116-1131): aligned is a monotone map of recovered => its max sits at rec_max
  h.aligned_max = aligned_f(map_recovered_f(h.rec_max, h), h);
  h.cf = 1.0f; h.cf_is_tensor = 0;
  if (h.aligned_max > 1.0f && (double)h.aligned_max > s.pre_max) {
    h.cf = __fdiv_rn((float)(s.pre_max - 1.0), __fsub_rn(h.aligned_max, 1.0f));
    h.cf_is_tensor = 1;
  }
  s.aligned_max = s.has_hdr ? (double)h.aligned_max : __longlong_as_double(0x7ff8000000000000LL);
  if (!s.has_hdr) { s.rec_min = s.rec_max = __longlong_as_double(0x7ff8000000000000LL); }
  s.out_min = s.out_max = s.intelligent_max = 0.0;
  s.hdr_pixels = s.negative_pixels = s.intelligent_hdr_pixels = 0;
  s.accepted = 0; s.reserved = 0;
  *sc = h;
  *st = s;
}

// ------------------------------------------------------------------------------------ phase B
__global__ void __launch_bounds__(256)
hdr_phase_b_kernel(const float* __restrict__ post3, const float* __restrict__ pre3, const HdrScalars* __restrict__ scp,
                   long long n, int mode, float factor, float ev, float* __restrict__ out,
                   PartialB* __restrict__ partials) {
  const HdrScalars sc = *scp;
  float omin = INFINITY, omax = -INFINITY, imax = -INFINITY;
  unsigned long long hdr = 0, neg = 0, ihdr = 0;
  // one element: the mode formula, the multiplier, the output statistics
  auto one = [&](float s, float p3) -> float {
    const float ldr = srgb_to_linear_f(s);                                // :1074
    float map = p3, aligned = 1.0f;                                       // :1080-1081
    if (sc.has_hdr) {                                                     // :1082-1102
      map = map_recovered_f(recover_f(s, sc.norm_function), sc);
      aligned = aligned_f(map, sc);
    }
    float r;
    if (mode == HDRVAE_MODE_CONSERVATIVE) {                               // :941-980
      r = p3 > 1.0f ? __fadd_rn(ldr, __fmul_rn(__fmul_rn(__fsub_rn(p3, 1.0f), factor), ldr)) : ldr;
    } else if (mode == HDRVAE_MODE_EXPOSURE) {                            // :982-1007 (un-aligned map)
      r = __fmul_rn(ldr, ev_multiplier_f(map));
    } else if (mode == HDRVAE_MODE_ADAPTIVE_RECOVERY) {                   // :1114-1147
      if (!sc.has_hdr) r = ldr;                                           // deterministic no-HDR rule
      else {
        const float hm = aligned > 1.0f ? 1.0f : 0.0f;
        const float comp = __fadd_rn(__fmul_rn(__fsub_rn(aligned, 1.0f), sc.cf), 1.0f);
        const float mc = __fadd_rn(__fmul_rn(aligned, __fsub_rn(1.0f, hm)), __fmul_rn(comp, hm));
        r = __fmul_rn(ldr, ev_multiplier_f(mc));
      }
    } else {                                                              // :1149-1159
      r = sc.has_hdr ? __fmul_rn(ldr, ev_multiplier_f(aligned)) : ldr;
    }
    ihdr += r > 1.0f; imax = fmaxf(imax, r);                              // :100-102 (before the multiplier)
    if (ev != 1.0f) r = __fmul_rn(r, ev);                                 // :180-182
    omin = fminf(omin, r); omax = fmaxf(omax, r);                         // :188-191
    hdr += r > 1.0f; neg += r < 0.0f;
    return r;
  };
  // 16-byte accesses over the bulk, scalars over the (< 4 element) tail
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(post3) | reinterpret_cast<uintptr_t>(pre3) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const long long n4 = vec_ok ? (n >> 2) : 0;
  const float4* post4 = reinterpret_cast<const float4*>(post3);
  const float4* pre4 = reinterpret_cast<const float4*>(pre3);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 s4 = post4[i], p4 = pre4[i];
    float4 r4;
    r4.x = one(s4.x, p4.x); r4.y = one(s4.y, p4.y); r4.z = one(s4.z, p4.z); r4.w = one(s4.w, p4.w);
    out4[i] = r4;
  }
  for (long long i = (n4 << 2) + blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
    out[i] = one(post3[i], pre3[i]);
  __shared__ float fm[3][8];
  __shared__ unsigned long long cn[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  omin = warp_min(omin); omax = warp_max(omax); imax = warp_max(imax);
  for (int o = 16; o; o >>= 1) {
    hdr += __shfl_xor_sync(~0u, hdr, o); neg += __shfl_xor_sync(~0u, neg, o); ihdr += __shfl_xor_sync(~0u, ihdr, o);
  }
  if (lane == 0) { fm[0][warp] = omin; fm[1][warp] = omax; fm[2][warp] = imax; cn[0][warp] = hdr; cn[1][warp] = neg; cn[2][warp] = ihdr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    PartialB pb;
    pb.omin = fm[0][0]; pb.omax = fm[1][0]; pb.imax = fm[2][0]; pb.hdr = cn[0][0]; pb.neg = cn[1][0]; pb.ihdr = cn[2][0];
    for (int w2 = 1; w2 < 8; ++w2) {
      pb.omin = fminf(pb.omin, fm[0][w2]); pb.omax = fmaxf(pb.omax, fm[1][w2]); pb.imax = fmaxf(pb.imax, fm[2][w2]);
      pb.hdr += cn[0][w2]; pb.neg += cn[1][w2]; pb.ihdr += cn[2][w2];
    }
    pb.pad = 0.f; pb.pad2 = 0;
    partials[blockIdx.x] = pb;
  }
}

// one warp: lanes stride over the partials (min / max / integer sums: order-independent, so still deterministic)
__global__ void hdr_reduce_b_kernel(const PartialB* __restrict__ partials, int n, hdrvae_stats* __restrict__ st) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  float omin = INFINITY, omax = -INFINITY, imax = -INFINITY;
  unsigned long long hdr = 0, neg = 0, ihdr = 0;
  for (int i = threadIdx.x; i < n; i += 32) {
    const PartialB pb = partials[i];
    omin = fminf(omin, pb.omin); omax = fmaxf(omax, pb.omax); imax = fmaxf(imax, pb.imax);
    hdr += pb.hdr; neg += pb.neg; ihdr += pb.ihdr;
  }
  omin = warp_min(omin); omax = warp_max(omax); imax = warp_max(imax);
  for (int o = 16; o; o >>= 1) {
    hdr += __shfl_xor_sync(~0u, hdr, o); neg += __shfl_xor_sync(~0u, neg, o); ihdr += __shfl_xor_sync(~0u, ihdr, o);
  }
  if (threadIdx.x != 0) return;
  st->out_min = omin; st->out_max = omax; st->intelligent_max = imax;
  st->hdr_pixels = (long long)hdr; st->negative_pixels = (long long)neg; st->intelligent_hdr_pixels = (long long)ihdr;
  st->accepted = (ihdr > 0 || imax > 1.1f) ? 1 : 0;                      // :106
}

// ------------------------------------------------------------------------------------ host launchers
constexpr int kPhaseBBlocks = 148 * 8;

struct EpilogueScratch {
  float* post3; float* pre3; PartialA* pa; PartialB* pb; hdrvae_raw_stats* raw; HdrScalars* sc; hdrvae_stats* st;
};

static int n_blocks_a(int B, int H, int W) { return B * ceil_div(H, kTileH) * ceil_div(W, kTileW); }

size_t epilogue_scratch_bytes(int B, int H, int W) {
  size_t b = 0;
  b += align_up((size_t)B * H * W * 3 * sizeof(float), 256) * 2;
  b += align_up((size_t)n_blocks_a(B, H, W) * sizeof(PartialA), 256);
  b += align_up((size_t)kPhaseBBlocks * sizeof(PartialB), 256);
  b += 256 * 3;
  return b;
}

static EpilogueScratch carve(void* scratch, int B, int H, int W) {
  EpilogueScratch e;
  uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
  const size_t img = align_up((size_t)B * H * W * 3 * sizeof(float), 256);
  e.post3 = reinterpret_cast<float*>(p); p += img;
  e.pre3 = reinterpret_cast<float*>(p); p += img;
  e.pa = reinterpret_cast<PartialA*>(p); p += align_up((size_t)n_blocks_a(B, H, W) * sizeof(PartialA), 256);
  e.pb = reinterpret_cast<PartialB*>(p); p += align_up((size_t)kPhaseBBlocks * sizeof(PartialB), 256);
  e.raw = reinterpret_cast<hdrvae_raw_stats*>(p); p += 256;
  e.sc = reinterpret_cast<HdrScalars*>(p); p += 256;
  e.st = reinterpret_cast<hdrvae_stats*>(p);
  return e;
}

void* epilogue_raw_stats_ptr(void* scratch, int B, int H, int W) { return carve(scratch, B, H, W).raw; }
hdrvae_stats* epilogue_stats_dev_ptr(void* scratch, int B, int H, int W) { return carve(scratch, B, H, W).st; }
float* epilogue_post3_ptr(void* scratch, int B, int H, int W) { return carve(scratch, B, H, W).post3; }
float* epilogue_pre3_ptr(void* scratch, int B, int H, int W) { return carve(scratch, B, H, W).pre3; }

template <typename T>
static void launch_phase_a(const void* pre, const float* conv_w, const float* conv_b, int H, int W, float* post3, float* pre3,
                           int* argmax3, PartialA* pa, dim3 grid, size_t smem, int y_pad, long long img_stride, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once.first()) {
    cudaFuncSetAttribute(hdr_phase_a_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(hdr_phase_a_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (argmax3 != nullptr)
    hdr_phase_a_kernel<T, true><<<grid, kEpiThreads, smem, s>>>(reinterpret_cast<const T*>(pre), conv_w, conv_b, H, W, post3, pre3, argmax3, pa, y_pad, img_stride);
  else
    hdr_phase_a_kernel<T, false><<<grid, kEpiThreads, smem, s>>>(reinterpret_cast<const T*>(pre), conv_w, conv_b, H, W, post3, pre3, nullptr, pa, y_pad, img_stride);
}

int launch_epilogue_phase_a(const void* pre, int dtype, int B, int H, int W, const float* conv_w, const float* conv_b,
                            int* argmax3, void* scratch, cudaStream_t s, int y_pad, long long img_stride) {
  if (img_stride == 0) img_stride = (long long)H * W * kC;
  EpilogueScratch e = carve(scratch, B, H, W);
  const dim3 grid(ceil_div(W, kTileW), ceil_div(H, kTileH), B);
  HDRVAE_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "epilogue: image too large for the launch grid");
  const size_t smem = (9 * kC * 4 + kChunk * kChanStride) * sizeof(float);
  if (dtype == HDRVAE_F32) {
    launch_phase_a<float>(pre, conv_w, conv_b, H, W, e.post3, e.pre3, argmax3, e.pa, grid, smem, y_pad, img_stride, s);
  } else if (dtype == HDRVAE_BF16) {
    launch_phase_a<__nv_bfloat16>(pre, conv_w, conv_b, H, W, e.post3, e.pre3, argmax3, e.pa, grid, smem, y_pad, img_stride, s);
  } else if (dtype == HDRVAE_F16) {
    launch_phase_a<__half>(pre, conv_w, conv_b, H, W, e.post3, e.pre3, argmax3, e.pa, grid, smem, y_pad, img_stride, s);
  } else {
    HDRVAE_REQUIRE(false, "epilogue: unsupported activation dtype %d", dtype);
  }
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  hdr_reduce_a_kernel<<<1, 256, 0, s>>>(e.pa, n_blocks_a(B, H, W), e.raw);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

// Phase A with conv_out precomputed on the tensor cores (16-bit features only; see hdr_phase_a_pre_kernel).
// `pre` points at interior row 0 of each image (img_stride elements apart), conv8 is dense [B][H][W][8] fp32.
int launch_epilogue_phase_a_pre(const void* pre, int dtype, int B, int H, int W, const float* conv8, const float* conv_b,
                                int* argmax3, void* scratch, cudaStream_t s, long long img_stride) {
  HDRVAE_REQUIRE(dtype == HDRVAE_F16 || dtype == HDRVAE_BF16, "epilogue (precomputed conv): 16-bit features only");
  if (img_stride == 0) img_stride = (long long)H * W * kC;
  EpilogueScratch e = carve(scratch, B, H, W);
  const dim3 grid(ceil_div(W, kTileW), ceil_div(H, kTileH), B);
  HDRVAE_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "epilogue: image too large for the launch grid");
  const float4* c8 = reinterpret_cast<const float4*>(conv8);
#define HDRVAE_PA_PRE(T, AM) hdr_phase_a_pre_kernel<T, AM><<<grid, kEpiThreads, 0, s>>>(reinterpret_cast<const T*>(pre), c8, conv_b, \
                                                                                       H, W, e.post3, e.pre3, argmax3, e.pa, img_stride)
  if (dtype == HDRVAE_F16) { if (argmax3) HDRVAE_PA_PRE(__half, true); else HDRVAE_PA_PRE(__half, false); }
  else { if (argmax3) HDRVAE_PA_PRE(__nv_bfloat16, true); else HDRVAE_PA_PRE(__nv_bfloat16, false); }
#undef HDRVAE_PA_PRE
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  hdr_reduce_a_kernel<<<1, 256, 0, s>>>(e.pa, n_blocks_a(B, H, W), e.raw);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_epilogue_phase_b(int B, int H, int W, int mode, float factor, float ev, float* out, hdrvae_stats* host_stats,
                            void* scratch, cudaStream_t s) {
  EpilogueScratch e = carve(scratch, B, H, W);
  HDRVAE_REQUIRE(mode >= 0 && mode <= 3, "epilogue: bad mode %d", mode);
  hdr_scalars_kernel<<<1, 32, 0, s>>>(e.raw, e.sc, e.st);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  const long long n = (long long)B * H * W * 3;
  int blocks = ceil_div(n, 256);
  if (blocks > kPhaseBBlocks) blocks = kPhaseBBlocks;
  hdr_phase_b_kernel<<<blocks, 256, 0, s>>>(e.post3, e.pre3, e.sc, n, mode, factor, ev, out, e.pb);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  hdr_reduce_b_kernel<<<1, 32, 0, s>>>(e.pb, blocks, e.st);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  if (host_stats != nullptr) {
    HDRVAE_CUDA_OK(cudaMemcpyAsync(host_stats, e.st, sizeof(hdrvae_stats), cudaMemcpyDeviceToHost, s));
    HDRVAE_CUDA_OK(cudaStreamSynchronize(s));
  }
  return 0;
}

// Multi-GPU batch sharding: merge n gathered raw-statistics blocks (rank order) into one, MIN / MAX / SUM in a fixed
// order.  One thread per scalar; dst may be one of the inputs' storage only if it is not among `blocks`.
__global__ void raw_stats_merge_kernel(const hdrvae_raw_stats* __restrict__ blocks, int n, hdrvae_raw_stats* __restrict__ dst) {
  const int t = threadIdx.x;
  if (t < HDRVAE_RAW_NMIN) {
    float v = blocks[0].vmin[t];
    for (int i = 1; i < n; ++i) v = fminf(v, blocks[i].vmin[t]);
    dst->vmin[t] = v;
  } else if (t < HDRVAE_RAW_NMIN + HDRVAE_RAW_NMAX) {
    const int k = t - HDRVAE_RAW_NMIN;
    float v = blocks[0].vmax[k];
    for (int i = 1; i < n; ++i) v = fmaxf(v, blocks[i].vmax[k]);
    dst->vmax[k] = v;
  } else if (t < HDRVAE_RAW_NMIN + HDRVAE_RAW_NMAX + HDRVAE_RAW_NSUM) {
    const int k = t - HDRVAE_RAW_NMIN - HDRVAE_RAW_NMAX;
    double v = blocks[0].vsum[k];
    for (int i = 1; i < n; ++i) v += blocks[i].vsum[k];
    dst->vsum[k] = v;
  }
}

int launch_raw_stats_merge(const hdrvae_raw_stats* blocks, int n, hdrvae_raw_stats* dst, cudaStream_t s) {
  raw_stats_merge_kernel<<<1, 32, 0, s>>>(blocks, n, dst);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace hdrvae
