// HDR upscaler path: replaces HDRUpscaleWithModel.upscale of the reference
// (/root/reference/hdr_upscale_with_model.py:148-263): two passes of an ESRGAN / RRDBNet 4x model through
// overlapping 512-pixel tiles with the atanh / logit reversal hook on every tile output (:79-107, :110-146),
// feather blending (restated comfy.utils.tiled_scale), then the luma / chroma recombination in YCbCr with a 3x3
// median on Y (:187-224) and the optional local hot-spot fix (:229-258).
//
// The RRDB convs run on the tcgen05 implicit-GEMM kernel of gemm_tc.cu with narrow (32 / 64 column) tiles:
//   * a dense block keeps one 192-channel fp16 buffer per pixel [x | g1 | g2 | g3 | g4]; conv k reads the channel
//     prefix 64 + 32(k-1) through a TMA map whose channel extent is exactly that prefix (the tail of the last
//     128-byte K block is zero-filled by TMA, the packed weights hold zeros there) and writes its LeakyReLU'd 32
//     channels into the next slice: the concat never exists as a copy;
//   * conv5 adds the fp32 residual stream (0.2 conv + x; at the end of an RRDB 0.04 conv + 0.2 x2 + x0 with two
//     residual operands) and writes, besides the fp32 stream, the fp16 operand copy into the next block's buffer;
//   * the two nearest-2x upsamples are folded into the loads of conv_up1 / conv_up2 (4 phase convs each);
//   * conv_last (64 -> 3) computes 32 columns and stores 4.
// Everything after the model (hook, feather blend, YCbCr, median, local fix) is elementwise / stencil work in
// fp32 with the reference's operation order (no FMA contraction) so that it is bit-comparable to the oracle.
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "engine.cuh"

namespace hdrvae {

// ------------------------------------------------------------------------------------------------ kernels
// image fp32 BHWC (3 channels) -> fp16 tile batch [n][h][w][8] (channels 3..7 zero), optional clamp to [-1, 1]
__global__ void up_extract_tiles_kernel(const float* __restrict__ img, int H, int W, const int4* __restrict__ tiles /*b,y0,x0,_*/,
                                        int n, int h, int w, int clamp1, uint4* __restrict__ out) {
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int t = (int)(i / ((long long)w * h));
    const int4 tl = tiles[t];                     // image, y0, x0, clamp flag
    const float* src = img + (((long long)tl.x * H + tl.y + y) * W + tl.z + x) * 3;
    float r = src[0], g = src[1], b = src[2];
    if (clamp1 || tl.w) { r = fminf(fmaxf(r, -1.f), 1.f); g = fminf(fmaxf(g, -1.f), 1.f); b = fminf(fmaxf(b, -1.f), 1.f); }
    const __half2 rg = __floats2half2_rn(r, g), b0 = __floats2half2_rn(b, 0.f);
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&rg);
    o.y = *reinterpret_cast<const uint32_t*>(&b0);
    o.z = 0u; o.w = 0u;
    out[i] = o;
  }
}

__global__ void up_scale_f32_kernel(float* p, int n, float s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] *= s;
}

__device__ __forceinline__ float up_reversal(float v, int kind) {
  if (kind == 1) {                                    // atanh(clamp(v, -1 + 1e-6, 1 - 1e-6))   (:101-105)
    v = fminf(fmaxf(v, -1.f + 1e-6f), 1.f - 1e-6f);
    return atanhf(v);
  }
  if (kind == 2) {                                    // logit(clamp(v, 1e-7, 1 - 1e-7))          (:93-98)
    v = fminf(fmaxf(v, 1e-7f), 1.f - 1e-7f);
    return logf(__fdiv_rn(v, 1.f - v));
  }
  return v;
}

// model output [n][4h][4w][4] fp32 -> reversal hook -> [n][4h][4w][3] (kernel-level parity entry)
__global__ void up_finish_tiles_kernel(const float4* __restrict__ in, long long n_px, int kind, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_px; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    out[i * 3 + 0] = up_reversal(v.x, kind);
    out[i * 3 + 1] = up_reversal(v.y, kind);
    out[i * 3 + 2] = up_reversal(v.z, kind);
  }
}

struct UpTile {             // one tile of comfy.utils.tiled_scale, output coordinates
  int b, oy, ox, oh, ow;    // image index, origin and size on the upscaled grid
  long long off, off2;      // float4 offsets of the tile's model output: pass 1 (input as is), pass 2 (input clamped)
};

// feather weight of output row/col `i` of a tile of length `len` (restated tiled_scale: ramps (t+1)/feather applied
// to both ends, successively, in fp32; a dimension not longer than the feather is left unmasked)
__device__ __forceinline__ float up_feather(float m, int i, int len, int feather) {
  if (feather >= len) return m;
  if (i < feather) m = __fmul_rn(m, (float)((double)(i + 1) / (double)feather));
  if (len - 1 - i < feather) m = __fmul_rn(m, (float)((double)(len - i) / (double)feather));
  return m;
}

// Blend of one pass: for every output pixel the tiles covering it, in the reference's accumulation order
// (itertools.product: y outer, x inner): out += reversal(ps) * mask; div += mask; result out / div.
__device__ __forceinline__ void up_blend_px(const float4* __restrict__ tiles_out, const UpTile* __restrict__ tl, int t0, int t1,
                                            int single, int second, int b, int y, int x, int feather, int kind, float* rgb) {
  if (single) {
    const UpTile t = tl[t0 + b];
    const float4 v = tiles_out[(second ? t.off2 : t.off) + (long long)y * t.ow + x];
    rgb[0] = up_reversal(v.x, kind); rgb[1] = up_reversal(v.y, kind); rgb[2] = up_reversal(v.z, kind);
    return;
  }
  float acc[3] = {0.f, 0.f, 0.f}, div = 0.f;
  for (int i = t0; i < t1; ++i) {
    const UpTile t = tl[i];
    if (t.b != b || y < t.oy || y >= t.oy + t.oh || x < t.ox || x >= t.ox + t.ow) continue;
    const int ly = y - t.oy, lx = x - t.ox;
    float m = 1.f;
    m = up_feather(m, ly, t.oh, feather);
    m = up_feather(m, lx, t.ow, feather);
    const float4 v = tiles_out[(second ? t.off2 : t.off) + (long long)ly * t.ow + lx];
    acc[0] = __fadd_rn(acc[0], __fmul_rn(up_reversal(v.x, kind), m));
    acc[1] = __fadd_rn(acc[1], __fmul_rn(up_reversal(v.y, kind), m));
    acc[2] = __fadd_rn(acc[2], __fmul_rn(up_reversal(v.z, kind), m));
    div = __fadd_rn(div, m);
  }
  rgb[0] = __fdiv_rn(acc[0], div); rgb[1] = __fdiv_rn(acc[1], div); rgb[2] = __fdiv_rn(acc[2], div);
}

__device__ __forceinline__ void up_rgb_to_ycbcr(const float* rgb, float* ycc) {
  // kornia.color.rgb_to_ycbcr, left-to-right fp32
  const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, rgb[0]), __fmul_rn(0.587f, rgb[1])), __fmul_rn(0.114f, rgb[2]));
  ycc[0] = y;
  ycc[1] = __fadd_rn(__fmul_rn(__fsub_rn(rgb[2], y), 0.564f), 0.5f);
  ycc[2] = __fadd_rn(__fmul_rn(__fsub_rn(rgb[0], y), 0.713f), 0.5f);
}

// Both passes blended -> (clamp(Y_unclamped, 0, 8), Cb_clamped, Cr_clamped) per pixel   (:189-213)
__global__ void up_blend_ycc_kernel(const float4* __restrict__ tiles_out,
                                    const UpTile* __restrict__ tl, int n_tiles, int single, int B, int OH, int OW,
                                    int feather, int kind, float* __restrict__ ycc /*[B][OH][OW][3]*/) {
  const long long total = (long long)B * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const int y = (int)((i / OW) % OH);
    const int b = (int)(i / ((long long)OW * OH));
    float ru[3], rc[3], yu[3], yc[3];
    up_blend_px(tiles_out, tl, 0, n_tiles, single, 0, b, y, x, feather, kind, ru);
    up_blend_px(tiles_out, tl, 0, n_tiles, single, 1, b, y, x, feather, kind, rc);
    up_rgb_to_ycbcr(ru, yu);
    up_rgb_to_ycbcr(rc, yc);
    ycc[i * 3 + 0] = fminf(fmaxf(yu[0], 0.f), 8.f);
    ycc[i * 3 + 1] = yc[1];
    ycc[i * 3 + 2] = yc[2];
  }
}

__device__ __forceinline__ void up_sort2(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
__device__ __forceinline__ float up_median9(float* v) {
  // 19-exchange median-of-9 network
  up_sort2(v[1], v[2]); up_sort2(v[4], v[5]); up_sort2(v[7], v[8]);
  up_sort2(v[0], v[1]); up_sort2(v[3], v[4]); up_sort2(v[6], v[7]);
  up_sort2(v[1], v[2]); up_sort2(v[4], v[5]); up_sort2(v[7], v[8]);
  up_sort2(v[0], v[3]); up_sort2(v[5], v[8]); up_sort2(v[4], v[7]);
  up_sort2(v[3], v[6]); up_sort2(v[1], v[4]); up_sort2(v[2], v[5]);
  up_sort2(v[4], v[7]); up_sort2(v[4], v[2]); up_sort2(v[6], v[4]);
  up_sort2(v[4], v[2]);
  return v[4];
}

// zero-padded 3x3 median of channel `c` of an interleaved [B][H][W][3] tensor (kornia.filters.median_blur)
__device__ __forceinline__ float up_median_at(const float* __restrict__ t, int b, int y, int x, int c, int H, int W) {
  float v[9];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xx = x + dx;
      v[(dy + 1) * 3 + dx + 1] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? t[(((long long)b * H + yy) * W + xx) * 3 + c] : 0.f;
    }
  return up_median9(v);
}

// median on Y, then the reference's un-clamped ycbcr_to_rgb (:20-48)
__global__ void up_recombine_kernel(const float* __restrict__ ycc, int B, int H, int W, float* __restrict__ out) {
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((long long)W * H));
    const float Y = up_median_at(ycc, b, y, x, 0, H, W);
    const float cb = __fsub_rn(ycc[i * 3 + 1], 0.5f), cr = __fsub_rn(ycc[i * 3 + 2], 0.5f);
    out[i * 3 + 0] = __fadd_rn(Y, __fmul_rn(1.403f, cr));
    out[i * 3 + 1] = __fsub_rn(__fsub_rn(Y, __fmul_rn(0.714f, cr)), __fmul_rn(0.344f, cb));
    out[i * 3 + 2] = __fadd_rn(Y, __fmul_rn(1.773f, cb));
  }
}

// per-channel zero-padded 3x3 median of an RGB image (small_blur output filter, :219-224)
__global__ void up_median_rgb_kernel(const float* __restrict__ in, int B, int H, int W, float* __restrict__ out) {
  const long long total = (long long)B * H * W * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const long long px = i / 3;
    const int x = (int)(px % W);
    const int y = (int)((px / W) % H);
    const int b = (int)(px / ((long long)W * H));
    out[i] = up_median_at(in, b, y, x, c, H, W);
  }
}

// local hot-spot fix (:229-258): mask = upscaled luma of the ORIGINAL image < 0.1; out = s (1 - mask) + clamp(s, -1, 1) mask
__global__ void up_local_fix_kernel(const float* __restrict__ img, int B, int H, int W, int scale, int method,
                                    float* __restrict__ s /*[B][H*scale][W*scale][3], in place*/) {
  const int OH = H * scale, OW = W * scale;
  const long long total = (long long)B * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const int y = (int)((i / OW) % OH);
    const int b = (int)(i / ((long long)OW * OH));
    auto luma = [&](int yy, int xx) {
      const float* p = img + (((long long)b * H + yy) * W + xx) * 3;
      return __fadd_rn(__fadd_rn(__fmul_rn(0.299f, p[0]), __fmul_rn(0.587f, p[1])), __fmul_rn(0.114f, p[2]));
    };
    float ys;
    if (method == HDRVAE_UPSCALE_BISLERP) {
      // comfy.utils.bislerp on the 1-channel luma (restated in oracle/upscaler_oracle.py): separable, width pass first.
      // Per axis: the two source pixels bilinear interpolation (align_corners = False) would blend, c1 = floor(src),
      // c2 = min(c1 + 1, last), ratio r = frac(src) with src = max((dst + 0.5) / scale - 0.5, 0) — exact in fp32 for the
      // 4x scale — then a SPHERICAL interpolation of the channel vectors; for scalars: same sign -> the first pixel,
      // opposite signs -> linear, a zero -> sine weights times the linearly interpolated magnitude.
      auto slerp1 = [](float b1, float b2, float r) {
        const float n1 = fabsf(b1), n2 = fabsf(b2);
        const float u1 = n1 == 0.f ? 0.f : b1 / n1, u2 = n2 == 0.f ? 0.f : b2 / n2;
        const float dot = u1 * u2;
        if (dot > 1.f - 1e-5f) return b1;
        if (dot < 1e-5f - 1.f) return __fadd_rn(__fmul_rn(b1, __fsub_rn(1.f, r)), __fmul_rn(b2, r));
        const float omega = acosf(dot), so = sinf(omega);
        float res = (sinf((1.f - r) * omega) / so) * u1 + (sinf(r * omega) / so) * u2;
        return res * (n1 * (1.f - r) + n2 * r);
      };
      const float fy = fmaxf((y + 0.5f) / scale - 0.5f, 0.f), fx = fmaxf((x + 0.5f) / scale - 0.5f, 0.f);
      const int y1 = min((int)fy, H - 1), x1 = min((int)fx, W - 1);
      const int y2 = min(y1 + 1, H - 1), x2 = min(x1 + 1, W - 1);
      const float ry = fy - floorf(fy), rx = fx - floorf(fx);
      const float t1 = slerp1(luma(y1, x1), luma(y1, x2), rx);
      const float t2 = slerp1(luma(y2, x1), luma(y2, x2), rx);
      ys = slerp1(t1, t2, ry);
    } else if (method == HDRVAE_UPSCALE_NEAREST_EXACT || method == HDRVAE_UPSCALE_AREA) {
      // nearest-exact: floor((dst + 0.5) / scale); "area" = adaptive average pooling, whose window for an integer
      // upscale factor is exactly that one source pixel
      ys = luma(min((int)floorf((y + 0.5f) / scale), H - 1), min((int)floorf((x + 0.5f) / scale), W - 1));
    } else if (method == HDRVAE_UPSCALE_BICUBIC) {
      // torch upsample_bicubic2d, align_corners = False: cubic convolution, A = -0.75, border indices clamped
      auto coef = [](float t, float* c) {
        const float A = -0.75f;
        auto c1 = [A](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };
        auto c2 = [A](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };
        c[0] = c2(t + 1.f); c[1] = c1(t); c[2] = c1(1.f - t); c[3] = c2(2.f - t);
      };
      const float fy = (y + 0.5f) / scale - 0.5f, fx = (x + 0.5f) / scale - 0.5f;
      const int iy = (int)floorf(fy), ix = (int)floorf(fx);
      float cy[4], cx[4];
      coef(fy - iy, cy); coef(fx - ix, cx);
      ys = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), H - 1);
        float rowv = 0.f;
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2) rowv += cx[b2] * luma(yy, min(max(ix - 1 + b2, 0), W - 1));
        ys += cy[a] * rowv;
      }
    } else {   // bilinear, align_corners = False (torch.nn.functional.interpolate)
      const float fy = fmaxf((y + 0.5f) / scale - 0.5f, 0.f), fx = fmaxf((x + 0.5f) / scale - 0.5f, 0.f);
      const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
      const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
      const float wy = fy - y0, wx = fx - x0;
      ys = (1.f - wy) * ((1.f - wx) * luma(y0, x0) + wx * luma(y0, x1)) + wy * ((1.f - wx) * luma(y1, x0) + wx * luma(y1, x1));
    }
    const float mask = ys < 0.1f ? 1.f : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = s[i * 3 + c];
      s[i * 3 + c] = __fadd_rn(__fmul_rn(v, __fsub_rn(1.f, mask)), __fmul_rn(fminf(fmaxf(v, -1.f), 1.f), mask));
    }
  }
}

static inline int up_grid(long long n) { return (int)std::min<long long>((n + 255) / 256, 148 * 32); }

}  // namespace hdrvae

using namespace hdrvae;

// ------------------------------------------------------------------------------------------------ model
struct UpRdb { PackedConv c[5]; };

struct hdrvae_upscaler {
  hdrvae_ctx* ctx = nullptr;      // device / SM count / conv implementation come from the decode context
  hdrvae_ctx store;               // owns this model's device allocations
  bool loaded = false;
  int nb = 0;
  PackedConv conv_first, conv_body, up1, up2, hr, last;
  std::vector<UpRdb> rdb;         // 3 per RRDB
};

namespace hdrvae {

struct UpPlan {                   // workspace of one forward over n tiles of h x w
  size_t img8, cat0, cat1, f_feat, f_a, f_b, f_c, u0, u1, u2, u3, total;
};
static UpPlan up_make_plan(int n, int h, int w) {
  UpPlan pl;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  const size_t P = (size_t)n * h * w;
  pl.img8 = take(P * 8 * 2);
  pl.cat0 = take(P * 192 * 2);
  pl.cat1 = take(P * 192 * 2);
  pl.f_feat = take(P * 64 * 4);
  pl.f_a = take(P * 64 * 4);
  pl.f_b = take(P * 64 * 4);
  pl.f_c = take(P * 64 * 4);
  pl.u0 = take(P * 64 * 2);
  pl.u1 = take(P * 4 * 64 * 2);
  pl.u2 = take(P * 16 * 64 * 2);
  pl.u3 = take(P * 16 * 64 * 2);
  pl.total = off;
  return pl;
}

// RRDBNet forward on n tiles (fp16 [n][h][w][8] in pl.img8) -> fp32 [n][4h][4w][4] at `out4`
static int up_forward(hdrvae_upscaler* up, int n, int h, int w, uint8_t* ws, const UpPlan& pl, float* out4, cudaStream_t s) {
  hdrvae_ctx* ctx = up->ctx;
  const int impl = ctx->conv_impl;
  uint16_t* cat[2] = {reinterpret_cast<uint16_t*>(ws + pl.cat0), reinterpret_cast<uint16_t*>(ws + pl.cat1)};
  float* feat = reinterpret_cast<float*>(ws + pl.f_feat);
  float* fb[3] = {reinterpret_cast<float*>(ws + pl.f_a), reinterpret_cast<float*>(ws + pl.f_b), reinterpret_cast<float*>(ws + pl.f_c)};
  int cur = 0;                                        // cat[cur] holds the fp16 copy of the current stream in channels 0..63
  {
    ConvIO io; io.x = ws + pl.img8; io.x_channels = 8; io.y = feat; io.y_dtype = DT_F32;
    io.y2 = cat[cur]; io.y2_channels = 192; io.y2_chan_off = 0;
    HDRVAE_TRY(run_conv(ctx, up->conv_first, io, n, h, w, impl, s));
  }
  // fp32 stream buffers: x0 (the RRDB input) lives in feat (first block) or fb[xi]; x1 -> fb[P], x2 -> fb[Q] and the
  // block output overwrites x1's buffer (dead by then), which makes it the next block's x0
  const float* x0 = feat;
  int xi = -1;
  for (int r = 0; r < up->nb; ++r) {
    const int P = (xi + 1) % 3, Q = (xi + 2) % 3;
    float* outs[3] = {fb[P], fb[Q], fb[P]};
    const float* xin = x0;
    for (int k = 0; k < 3; ++k) {
      const UpRdb& R = up->rdb[r * 3 + k];
      for (int j = 0; j < 4; ++j) {
        ConvIO io; io.x = cat[cur]; io.x_channels = 192; io.y = cat[cur]; io.y_dtype = DT_F16; io.y_channels = 192;
        io.y_chan_off = 64 + 32 * j; io.lrelu = 0.2f;
        HDRVAE_TRY(run_conv(ctx, R.c[j], io, n, h, w, impl, s));
      }
      ConvIO io; io.x = cat[cur]; io.x_channels = 192; io.y_dtype = DT_F32; io.y = outs[k];
      io.y2 = cat[cur ^ 1]; io.y2_channels = 192; io.y2_chan_off = 0;
      io.residual = xin; io.res_dtype = DT_F32;
      // k < 2: x_{k+1} = 0.2 conv5 + x_k (weights and bias packed x 0.2); end of the RRDB:
      // (0.2 conv5 + x2) 0.2 + x0 = 0.04 conv5 + 0.2 x2 + x0 (weights and bias packed x 0.04)
      if (k == 2) { io.res_scale = 0.2f; io.residual2 = x0; }
      HDRVAE_TRY(run_conv(ctx, R.c[4], io, n, h, w, impl, s));
      xin = outs[k];
      cur ^= 1;
    }
    x0 = outs[2];
    xi = P;
  }
  // trunk conv + long skip: t = conv_body(x) + feat; only its fp16 copy is needed (operand of conv_up1)
  {
    ConvIO io; io.x = cat[cur]; io.x_channels = 192; io.y = fb[(xi + 2) % 3]; io.y_dtype = DT_F32;
    io.residual = feat; io.res_dtype = DT_F32; io.y2 = ws + pl.u0;
    HDRVAE_TRY(run_conv(ctx, up->conv_body, io, n, h, w, impl, s));
  }
  {
    ConvIO io; io.x = ws + pl.u0; io.y = ws + pl.u1; io.y_dtype = DT_F16; io.lrelu = 0.2f;
    HDRVAE_TRY(run_conv(ctx, up->up1, io, n, h, w, impl, s));
  }
  {
    ConvIO io; io.x = ws + pl.u1; io.y = ws + pl.u2; io.y_dtype = DT_F16; io.lrelu = 0.2f;
    HDRVAE_TRY(run_conv(ctx, up->up2, io, n, 2 * h, 2 * w, impl, s));
  }
  {
    ConvIO io; io.x = ws + pl.u2; io.y = ws + pl.u3; io.y_dtype = DT_F16; io.lrelu = 0.2f;
    HDRVAE_TRY(run_conv(ctx, up->hr, io, n, 4 * h, 4 * w, impl, s));
  }
  {
    ConvIO io; io.x = ws + pl.u3; io.y = out4; io.y_dtype = DT_F32; io.y_channels = 4; io.n_store = 4;
    HDRVAE_TRY(run_conv(ctx, up->last, io, n, 4 * h, 4 * w, impl, s));
  }
  return 0;
}

}  // namespace hdrvae

// ------------------------------------------------------------------------------------------------ tiling plan
namespace hdrvae {

static const int kUpTile = 512, kUpOverlap = 64, kUpScale = 4;     // hdr_upscale_with_model.py:115-116, 4x ESRGAN
static const long long kUpMaxChunkPx = 8LL * 512 * 512;            // pixels of one forward batch (bounds the workspace: ~14 GB)

struct UpHostTile { int b, y0, x0, h, w; long long off, off2; };
struct UpGroup { int h, w; std::vector<int> tiles; };
// One forward batch: n tiles of a group, both passes at once ([n as-is inputs | n clamped inputs]: the two passes of
// hdr_upscale_with_model.py:181-186 share every launch, which halves the launch count and doubles the small grids)
struct UpChunk { int h, w, first, n; long long out_base; size_t origin0; };

static std::vector<int> up_positions(int size) {
  std::vector<int> p;
  if (size <= kUpTile) { p.push_back(0); return p; }
  for (int v = 0; v < size - kUpOverlap; v += kUpTile - kUpOverlap) p.push_back(v);
  return p;
}

struct UpTiling {
  bool single = false;
  std::vector<UpHostTile> tiles;      // reference accumulation order
  std::vector<UpGroup> groups;        // equal-shape tiles
  std::vector<UpChunk> chunks;        // forward batches
  std::vector<int> order;             // tile indices in chunk order
  long long out_px = 0;               // float4 elements of all tile outputs (both passes)
};

static UpTiling up_make_tiling(int B, int H, int W) {
  UpTiling t;
  t.single = H <= kUpTile && W <= kUpTile;
  const std::vector<int> py = up_positions(H), px = up_positions(W);
  for (int b = 0; b < B; ++b)
    for (int y : py)
      for (int x : px) {
        UpHostTile u;
        u.b = b;
        u.y0 = std::max(0, std::min(H - kUpOverlap, y)); u.h = std::min(kUpTile, H - u.y0);
        u.x0 = std::max(0, std::min(W - kUpOverlap, x)); u.w = std::min(kUpTile, W - u.x0);
        if (t.single) { u.y0 = u.x0 = 0; u.h = H; u.w = W; }
        u.off = u.off2 = 0;
        t.tiles.push_back(u);
      }
  for (size_t i = 0; i < t.tiles.size(); ++i) {
    const UpHostTile& u = t.tiles[i];
    size_t g = 0;
    for (; g < t.groups.size(); ++g)
      if (t.groups[g].h == u.h && t.groups[g].w == u.w) break;
    if (g == t.groups.size()) { UpGroup ng; ng.h = u.h; ng.w = u.w; t.groups.push_back(ng); }
    t.groups[g].tiles.push_back((int)i);
  }
  long long off = 0;
  for (UpGroup& g : t.groups) {
    const long long per = (long long)g.h * g.w, sz = 16 * per;
    const int cap = (int)std::max<long long>(1, std::min<long long>((long long)g.tiles.size(), kUpMaxChunkPx / (2 * per)));
    for (size_t i0 = 0; i0 < g.tiles.size(); i0 += cap) {
      UpChunk c;
      c.h = g.h; c.w = g.w; c.first = (int)t.order.size(); c.n = (int)std::min<size_t>(cap, g.tiles.size() - i0);
      c.out_base = off; c.origin0 = 2 * (size_t)t.order.size();
      for (int j = 0; j < c.n; ++j) {
        UpHostTile& u = t.tiles[g.tiles[i0 + j]];
        u.off = off + j * sz; u.off2 = off + (c.n + j) * sz;
        t.order.push_back(g.tiles[i0 + j]);
      }
      off += 2 * c.n * sz;
      t.chunks.push_back(c);
    }
  }
  t.out_px = off;
  return t;
}

struct UpWs { size_t descs, origins, out, ycc, tmp, fwd, total; };
static UpWs up_make_ws(const UpTiling& t, int B, int H, int W) {
  UpWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  w.descs = take(t.tiles.size() * sizeof(UpTile));
  w.origins = take(2 * t.tiles.size() * sizeof(int4));
  w.out = take((size_t)t.out_px * 16);
  w.ycc = take((size_t)B * 16 * H * W * 3 * 4);
  w.tmp = take((size_t)B * 16 * H * W * 3 * 4);
  size_t fwd = 0;
  for (const UpChunk& c : t.chunks) fwd = std::max(fwd, up_make_plan(2 * c.n, c.h, c.w).total);
  w.fwd = take(fwd);
  w.total = off;
  return w;
}

}  // namespace hdrvae

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int hdrvae_upscaler_create(hdrvae_ctx* ctx, hdrvae_upscaler** out) {
  HDRVAE_REQUIRE(ctx != nullptr && out != nullptr, "hdrvae_upscaler_create: null argument");
  hdrvae_upscaler* up = new hdrvae_upscaler();
  up->ctx = ctx;
  up->store.device = ctx->device; up->store.num_sms = ctx->num_sms;
  *out = up;
  return 0;
}

int hdrvae_upscaler_destroy(hdrvae_upscaler* up) {
  if (up == nullptr) return 0;
  cudaSetDevice(up->store.device);
  for (void* p : up->store.owned) cudaFree(p);
  delete up;
  return 0;
}

// Canonical (Real-ESRGAN) name of a state-dict key given in either layout; "" if it is not an RRDBNet tensor.
static std::string up_canonical(const std::string& k, int nb_hint) {
  if (k.compare(0, 6, "model.") != 0) return k;
  auto leaf = k.substr(k.rfind('.') + 1);
  const std::string head = k.substr(0, k.rfind('.'));
  if (head == "model.0") return "conv_first." + leaf;
  if (head == "model.3") return "conv_up1." + leaf;
  if (head == "model.6") return "conv_up2." + leaf;
  if (head == "model.8") return "conv_hr." + leaf;
  if (head == "model.10") return "conv_last." + leaf;
  int i = -1, r = -1, c = -1;
  if (sscanf(head.c_str(), "model.1.sub.%d.RDB%d.conv%d.0", &i, &r, &c) == 3)
    return "body." + std::to_string(i) + ".rdb" + std::to_string(r) + ".conv" + std::to_string(c) + "." + leaf;
  if (sscanf(head.c_str(), "model.1.sub.%d", &i) == 1 && head == "model.1.sub." + std::to_string(i)) {
    (void)nb_hint;
    return "conv_body." + leaf;
  }
  return "";
}

int hdrvae_upscaler_load_weights(hdrvae_upscaler* up, const hdrvae_weight_desc* descs, int n) {
  HDRVAE_REQUIRE(up != nullptr && descs != nullptr && n > 0, "hdrvae_upscaler_load_weights: null argument");
  HDRVAE_REQUIRE(!up->loaded, "hdrvae_upscaler_load_weights: model already loaded (create a new one)");
  hdrvae_ctx* st = &up->store;
  HDRVAE_CUDA_OK(cudaSetDevice(st->device));
  cudaStream_t s = nullptr;
  std::vector<void*> staging;
  int nb = 0;
  for (int i = 0; i < n; ++i) {
    const hdrvae_weight_desc& d = descs[i];
    HDRVAE_REQUIRE(d.name && d.data && d.ndim >= 1 && d.ndim <= 4, "hdrvae_upscaler_load_weights: bad descriptor %d", i);
    const std::string name = up_canonical(d.name, 0);
    if (name.empty()) continue;
    long long cnt = 1;
    std::vector<int64_t> shp;
    for (int k = 0; k < d.ndim; ++k) { cnt *= d.shape[k]; shp.push_back(d.shape[k]); }
    const size_t esz = d.dtype == HDRVAE_F32 ? 4 : 2;
    void* stage = nullptr;
    HDRVAE_CUDA_OK(cudaMalloc(&stage, cnt * esz));
    staging.push_back(stage);
    HDRVAE_CUDA_OK(cudaMemcpyAsync(stage, d.data, cnt * esz, cudaMemcpyDefault, s));
    float* f = nullptr;
    HDRVAE_TRY(dev_alloc(st, cnt * sizeof(float), (void**)&f));
    HDRVAE_TRY(launch_to_f32(stage, d.dtype, f, cnt, s));
    st->raw[name] = f;
    st->shapes[name] = shp;
    int bi = -1;
    if (sscanf(name.c_str(), "body.%d.", &bi) == 1) nb = std::max(nb, bi + 1);
  }
  HDRVAE_REQUIRE(nb >= 1, "upscaler state dict holds no RRDB blocks (body.N.rdbK.convJ / model.1.sub.N.RDBK.convJ.0)");
  up->nb = nb;
  auto conv = [&](const std::string& k, int cout, int cin, bool upsample, float scale, PackedConv* pc) -> int {
    auto it = st->shapes.find(k + ".weight");
    HDRVAE_REQUIRE(it != st->shapes.end() && st->raw.count(k + ".bias"), "upscaler state dict lacks %s.{weight,bias}", k.c_str());
    const auto& sh = it->second;
    HDRVAE_REQUIRE(sh.size() == 4 && sh[0] == cout && sh[1] == cin && sh[2] == 3 && sh[3] == 3,
                   "%s.weight has the wrong shape (want [%d,%d,3,3]: RRDBNet nf 64, gc 32)", k.c_str(), cout, cin);
    HDRVAE_TRY(pack_conv(st, st->raw[k + ".weight"], st->raw[k + ".bias"], cout, cin, 3, upsample, scale, DT_F16, pc, s));
    if (scale != 1.f) {
      up_scale_f32_kernel<<<1, 64, 0, s>>>(pc->bias, cout, scale);
      HDRVAE_CUDA_OK(cudaGetLastError());
    }
    return 0;
  };
  HDRVAE_TRY(conv("conv_first", 64, 3, false, 1.f, &up->conv_first));
  up->rdb.resize((size_t)nb * 3);
  for (int i = 0; i < nb; ++i)
    for (int r = 0; r < 3; ++r) {
      UpRdb& R = up->rdb[(size_t)i * 3 + r];
      const std::string base = "body." + std::to_string(i) + ".rdb" + std::to_string(r + 1) + ".conv";
      for (int j = 0; j < 4; ++j) HDRVAE_TRY(conv(base + std::to_string(j + 1), 32, 64 + 32 * j, false, 1.f, &R.c[j]));
      // residual scaling folded into conv5: 0.2 inside a dense block, 0.2 * 0.2 for the block that ends an RRDB
      HDRVAE_TRY(conv(base + "5", 64, 192, false, r == 2 ? 0.2f * 0.2f : 0.2f, &R.c[4]));
    }
  HDRVAE_TRY(conv("conv_body", 64, 64, false, 1.f, &up->conv_body));
  HDRVAE_TRY(conv("conv_up1", 64, 64, true, 1.f, &up->up1));
  HDRVAE_TRY(conv("conv_up2", 64, 64, true, 1.f, &up->up2));
  HDRVAE_TRY(conv("conv_hr", 64, 64, false, 1.f, &up->hr));
  HDRVAE_TRY(conv("conv_last", 3, 64, false, 1.f, &up->last));
  HDRVAE_CUDA_OK(cudaStreamSynchronize(s));
  for (void* p : staging) cudaFree(p);
  for (auto& kv : st->raw) cudaFree(kv.second);        // fp32 copies are no longer needed
  {
    std::vector<void*> keep;
    for (void* p : st->owned) {
      bool raw = false;
      for (auto& kv : st->raw) if (kv.second == p) { raw = true; break; }
      if (!raw) keep.push_back(p);
    }
    st->owned.swap(keep);
    st->raw.clear();
  }
  up->loaded = true;
  return 0;
}

int hdrvae_upscaler_blocks(hdrvae_upscaler* up) { return up ? up->nb : -1; }

int hdrvae_upscaler_forward_bytes(hdrvae_upscaler* up, int n, int h, int w, size_t* bytes) {
  HDRVAE_REQUIRE(up && bytes && n >= 1 && h >= 1 && w >= 1, "hdrvae_upscaler_forward_bytes: bad argument");
  const UpPlan pl = up_make_plan(n, h, w);
  *bytes = pl.total + (size_t)n * 16 * h * w * 16 + (size_t)n * sizeof(int4) + 4096;
  return 0;
}

int hdrvae_upscaler_forward(hdrvae_upscaler* up, const float* x_bhwc, int n, int h, int w, int reversal, float* y_bhwc,
                            void* workspace, size_t workspace_bytes, void* stream) {
  HDRVAE_REQUIRE(up && up->loaded && x_bhwc && y_bhwc && workspace, "hdrvae_upscaler_forward: null argument / no weights");
  HDRVAE_REQUIRE(reversal >= 0 && reversal <= 2, "hdrvae_upscaler_forward: reversal must be 0 (none), 1 (atanh) or 2 (logit)");
  size_t need = 0;
  HDRVAE_TRY(hdrvae_upscaler_forward_bytes(up, n, h, w, &need));
  HDRVAE_REQUIRE(workspace_bytes >= need, "hdrvae_upscaler_forward: workspace too small (%zu < %zu)", workspace_bytes, need);
  HDRVAE_CUDA_OK(cudaSetDevice(up->store.device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const UpPlan pl = up_make_plan(n, h, w);
  float* out4 = reinterpret_cast<float*>(ws + pl.total);
  int4* origins = reinterpret_cast<int4*>(ws + pl.total + (size_t)n * 16 * h * w * 16);
  std::vector<int4> host(n);
  for (int i = 0; i < n; ++i) host[i] = make_int4(i, 0, 0, 0);
  HDRVAE_CUDA_OK(cudaMemcpyAsync(origins, host.data(), n * sizeof(int4), cudaMemcpyHostToDevice, s));
  up_extract_tiles_kernel<<<up_grid((long long)n * h * w), 256, 0, s>>>(x_bhwc, h, w, origins, n, h, w, 0,
                                                                       reinterpret_cast<uint4*>(ws + pl.img8));
  HDRVAE_LAUNCHED();
  HDRVAE_TRY(up_forward(up, n, h, w, ws, pl, out4, s));
  const long long npx = (long long)n * 16 * h * w;
  up_finish_tiles_kernel<<<up_grid(npx), 256, 0, s>>>(reinterpret_cast<const float4*>(out4), npx, reversal, y_bhwc);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

int hdrvae_upscale_workspace_bytes(hdrvae_upscaler* up, int B, int H, int W, size_t* bytes) {
  HDRVAE_REQUIRE(up && bytes && B >= 1 && H >= 1 && W >= 1, "hdrvae_upscale_workspace_bytes: bad argument");
  const UpTiling t = up_make_tiling(B, H, W);
  *bytes = up_make_ws(t, B, H, W).total;
  return 0;
}

int hdrvae_upscale(hdrvae_upscaler* up, const float* image_bhwc, int B, int H, int W, int reversal, int small_blur,
                   int local_fix, int upscale_method, float* out_bhwc, void* workspace, size_t workspace_bytes,
                   void* stream) {
  HDRVAE_REQUIRE(up && up->loaded && image_bhwc && out_bhwc && workspace, "hdrvae_upscale: null argument / no weights");
  HDRVAE_REQUIRE(reversal == 1 || reversal == 2, "hdrvae_upscale: reversal must be 1 (atanh) or 2 (logit)");
  HDRVAE_REQUIRE(!local_fix || (upscale_method >= HDRVAE_UPSCALE_NEAREST_EXACT && upscale_method <= HDRVAE_UPSCALE_BISLERP),
                 "hdrvae_upscale: local_fix supports upscale_method nearest-exact, bilinear, area, bicubic and bislerp (got %d)", upscale_method);
  HDRVAE_CUDA_OK(cudaSetDevice(up->store.device));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const UpTiling t = up_make_tiling(B, H, W);
  const UpWs wsl = up_make_ws(t, B, H, W);
  HDRVAE_REQUIRE(workspace_bytes >= wsl.total, "hdrvae_upscale: workspace too small (%zu < %zu)", workspace_bytes, wsl.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int OH = H * kUpScale, OW = W * kUpScale;
  // tile descriptors (output coordinates) in the reference's accumulation order
  std::vector<UpTile> descs(t.tiles.size());
  for (size_t i = 0; i < t.tiles.size(); ++i) {
    const UpHostTile& u = t.tiles[i];
    descs[i].b = u.b; descs[i].oy = u.y0 * kUpScale; descs[i].ox = u.x0 * kUpScale;
    descs[i].oh = u.h * kUpScale; descs[i].ow = u.w * kUpScale; descs[i].off = u.off; descs[i].off2 = u.off2;
  }
  HDRVAE_CUDA_OK(cudaMemcpyAsync(ws + wsl.descs, descs.data(), descs.size() * sizeof(UpTile), cudaMemcpyHostToDevice, s));
  // small_blur's input filter (:176-178) is torchvision gaussian_blur(kernel 3, sigma 0.1): taps exp(-50) = 1.9e-22
  // around a centre tap that rounds to 1.0f, i.e. the identity on fp32 images; nothing to launch.
  int4* origins = reinterpret_cast<int4*>(ws + wsl.origins);
  std::vector<int4> host_or(2 * t.tiles.size());
  for (const UpChunk& c : t.chunks)
    for (int pass = 0; pass < 2; ++pass)
      for (int j = 0; j < c.n; ++j) {
        const UpHostTile& u = t.tiles[t.order[c.first + j]];
        host_or[c.origin0 + pass * c.n + j] = make_int4(u.b, u.y0, u.x0, pass);     // .w: clamp the input to [-1, 1]
      }
  HDRVAE_CUDA_OK(cudaMemcpyAsync(origins, host_or.data(), host_or.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
  float* out_all = reinterpret_cast<float*>(ws + wsl.out);
  for (const UpChunk& c : t.chunks) {
    const int n2 = 2 * c.n;
    const long long per = (long long)c.h * c.w;
    const UpPlan pl = up_make_plan(n2, c.h, c.w);
    uint8_t* fws = ws + wsl.fwd;
    up_extract_tiles_kernel<<<up_grid((long long)n2 * per), 256, 0, s>>>(image_bhwc, H, W, origins + c.origin0, n2, c.h, c.w, 0,
                                                                        reinterpret_cast<uint4*>(fws + pl.img8));
    HDRVAE_LAUNCHED();
    HDRVAE_TRY(up_forward(up, n2, c.h, c.w, fws, pl, out_all + c.out_base * 4, s));
  }
  const long long opx = (long long)B * OH * OW;
  float* ycc = reinterpret_cast<float*>(ws + wsl.ycc);
  float* tmp = reinterpret_cast<float*>(ws + wsl.tmp);
  up_blend_ycc_kernel<<<up_grid(opx), 256, 0, s>>>(reinterpret_cast<const float4*>(ws + wsl.out),
                                                    reinterpret_cast<const UpTile*>(ws + wsl.descs), (int)t.tiles.size(),
                                                    t.single ? 1 : 0, B, OH, OW, kUpOverlap * kUpScale, reversal, ycc);
  HDRVAE_LAUNCHED();
  up_recombine_kernel<<<up_grid(opx), 256, 0, s>>>(ycc, B, OH, OW, small_blur ? tmp : out_bhwc);
  HDRVAE_LAUNCHED();
  if (small_blur) {
    up_median_rgb_kernel<<<up_grid(opx * 3), 256, 0, s>>>(tmp, B, OH, OW, out_bhwc);
    HDRVAE_LAUNCHED();
  }
  if (local_fix) {
    up_local_fix_kernel<<<up_grid(opx), 256, 0, s>>>(image_bhwc, B, H, W, kUpScale, upscale_method, out_bhwc);
    HDRVAE_LAUNCHED();
  }
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
