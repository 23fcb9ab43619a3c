// Implicit-GEMM convolution / GEMM on tcgen05 tensor cores (sm_100a).
//
// One persistent, warp-specialised kernel serves every GEMM-shaped op of the Flux AE decoder
// (ComfyUI Decoder as driven by vae.decode, reference hdr_vae_decode.py:859,:1022):
//   * 3x3 convs       : 9 taps, each tap a TMA box load of the NHWC activation tile shifted by
//                       (dy,dx); out-of-bounds rows/cols are zero-filled by TMA = the conv padding;
//   * upsample convs  : nearest-2x folded into the load: 4 output phases, each a 2x2-tap conv on the
//                       SOURCE grid with pre-summed weights (launch per phase, strided output);
//   * 1x1 convs, QK^T, PV : 1 tap.
// Operands: fp16 or bf16 (kind::f16) for normalised activations x weights, or fp32 read as tf32
// (kind::tf32) where a conv consumes the raw fp32 residual stream.  Accumulation fp32 in TMEM.
//
// Roles (320 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = tcgen05.mma issuer,
// warps 2..9 = epilogue (TMEM -> registers -> scale/bias/residual -> global, + GroupNorm partial
// statistics of the output).  Two TMEM accumulator stages let the epilogue of tile i overlap the
// MMAs of tile i+1.
//
// Tile: 128 pixels (TH x TW patch) x BLOCK_N output channels, K step = one 128-byte swizzle row.
#include "common.cuh"
#include "ptx.cuh"

namespace hdrvae {

constexpr int kBlockM = 128;
constexpr int kRowBytes = 128;                    // K extent of one stage row (64 x 16-bit or 32 x tf32)
constexpr int kABytes = kBlockM * kRowBytes;      // 16 KB
constexpr int kEpilogueThreads = 256;             // 8 warps: 2 per TMEM lane quarter, each taking half of the columns
constexpr int kNumThreads = 64 + kEpilogueThreads;

template <int BLOCK_N>
struct TcConfig {
  static constexpr int kBBytes = BLOCK_N * kRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BLOCK_N;   // two accumulator stages (power of two: 256 / 512)
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * BLOCK_N * 4 /*bias*/ + 4 * 64 * 4 /*stats*/ +
                                    256 /*barriers*/ + 1024 /*align slack*/;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// Per GroupNorm group (CPG consecutive channels of this thread's 32-column chunk): sum and sum of squares
// over the warp's 32 pixels; lane 0 parks them in shared memory for the cross-warp sum.
template <int CPG>
__device__ __forceinline__ void emit_group_stats(const float (&f)[32], bool live, int lane, float* dst) {
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float s = 0.f, qq = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float t = live ? f[g * CPG + j] : 0.f;
      s += t;
      qq = fmaf(t, t, qq);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      qq += __shfl_xor_sync(0xffffffffu, qq, o);
    }
    if (lane == 0) { dst[g * 2 + 0] = s; dst[g * 2 + 1] = qq; }
  }
}

template <int BLOCK_N, bool kTf32>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  using Cfg = TcConfig<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kElemsPerRow = kTf32 ? 32 : 64;   // K elements per stage
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B needs 1024-byte aligned stage buffers.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  float* bias_s = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);      // [2][BLOCK_N]
  float* stat_s = bias_s + 2 * BLOCK_N;                                             // [4 warps][32 groups][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stat_s + 4 * 64);
  uint64_t* full_bar = bars;                      // [kStages]
  uint64_t* empty_bar = bars + kStages;           // [kStages]
  uint64_t* tmem_full_bar = bars + 2 * kStages;   // [2]
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], kEpilogueThreads / 32);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr_s);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = p.n_img * tiles_per_img * p.n_tiles_n;
  const int kb_per_tap = p.k_per_tap / kElemsPerRow;
  const int num_kb = p.ntaps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      const int mt = tile / p.n_tiles_n;
      const int img = mt / tiles_per_img;
      const int rem = mt - img * tiles_per_img;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = nt * BLOCK_N;
      for (int t = 0; t < p.ntaps; ++t) {
        const int xs = x0 + p.tap_dx[t], ys = y0 + p.tap_dy[t];
        for (int kb = 0; kb < kb_per_tap; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          ptx::tma_load_4d(smem_a + stage * kABytes, &tmA, &full_bar[stage], kb * kElemsPerRow, xs, ys, img);
          ptx::tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage],
                           (t * kb_per_tap + kb) * kElemsPerRow, n0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    const uint32_t idesc = kTf32 ? ptx::make_idesc(2u, kBlockM, BLOCK_N)
                                 : ptx::make_idesc(p.ab_dtype == DT_BF16 ? 1u : 0u, kBlockM, BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after_sync();
        const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_a + stage * kABytes));
        const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // one MMA consumes 32 bytes of K (16 x 16-bit or 8 x tf32) of the 128-byte swizzle row:
          // advance the start address by 32 B = +2 in 16-byte units
          if (kTf32) ptx::umma_tf32(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          else ptx::umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);            // frees the smem slot when these MMAs retire
        if (kb == num_kb - 1) ptx::umma_commit(&tmem_full_bar[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue (8 warps; 128 TMEM lanes x 2 column halves)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which half of the tile's columns this warp drains
    const int row = q * 32 + lane;          // accumulator row = pixel index inside the tile
    const int et = threadIdx.x - 64;        // 0..255
    const int cpg = p.n_cols >> 5;          // channels per GroupNorm group (stats only; n_cols % 128 == 0 there)
    constexpr int kChunksPerWarp = BLOCK_N / 64;
    const bool res_f32 = p.residual != nullptr && p.res_dtype == DT_F32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      const int mt = tile / p.n_tiles_n;
      const int img = mt / tiles_per_img;
      const int rem = mt - img * tiles_per_img;
      const int ty = rem / p.tiles_x;
      const int tx = rem - ty * p.tiles_x;
      const int n0 = nt * BLOCK_N;
      const int y = ty * p.TH + (row >> p.tw_log2);
      const int x = tx * p.TW + (row & (p.TW - 1));
      const bool valid = (y < p.H) && (x < p.W);
      const long long off = (long long)img * p.out_img_stride +
                            (long long)(y * p.sy + p.py) * p.out_row_stride +
                            (long long)(x * p.sx + p.px) * p.out_px_stride + n0;
      const int cbase = half * (BLOCK_N / 2);

      // fp32 residual of the first chunk: requested before anything else so it overlaps the MMA tail
      float4 rcur[8], rnext[8];
      const bool res_live = res_f32 && valid;
      if (res_live && (n0 + cbase) < p.n_cols) {
        const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + off + cbase);
#pragma unroll
        for (int j = 0; j < 8; ++j) rcur[j] = rp[j];
      }

      // stage the bias slice of this tile (double-buffered with the accumulator stage)
      float* bs = bias_s + acc * BLOCK_N;
      if (!p.bias_per_row) {
        for (int c = et; c < BLOCK_N; c += kEpilogueThreads) {
          const int col = n0 + c;
          bs[c] = (p.bias != nullptr && col < p.n_cols) ? __ldg(p.bias + col) : 0.f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // epilogue warps only
      const float row_bias = (p.bias_per_row && p.bias != nullptr && valid) ? __ldg(p.bias + x) : 0.f;
      const float scale = p.alpha * ((p.row_scale != nullptr && valid) ? __ldg(p.row_scale + x) : 1.f);

      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after_sync();

      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll
      for (int ci = 0; ci < kChunksPerWarp; ++ci) {
        const int c0 = cbase + ci * 32;
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_row + c0, v);
        // prefetch the next chunk's residual while the TMEM load is in flight
        if (ci + 1 < kChunksPerWarp && res_live && (n0 + c0 + 32) < p.n_cols) {
          const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + off + c0 + 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) rnext[j] = rp[j];
        }
        ptx::tmem_ld_wait(v);
        const bool cols_ok = (n0 + c0) < p.n_cols;       // n_cols is a multiple of 32 on every call site
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          f[j] = __uint_as_float(v[j]) * scale + (p.bias_per_row ? row_bias : bs[c0 + j]);
        if (valid && cols_ok) {
          if (res_f32) {
            // plain loads (above): the residual may alias the output (in-place add, same thread reads then writes)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[4 * j] += rcur[j].x; f[4 * j + 1] += rcur[j].y; f[4 * j + 2] += rcur[j].z; f[4 * j + 3] += rcur[j].w;
            }
          } else if (p.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) + off + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 r = rp[j];
              const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 t2;
                if (p.res_dtype == DT_BF16) t2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
                else t2 = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                f[j * 8 + e * 2 + 0] += t2.x;
                f[j * 8 + e * 2 + 1] += t2.y;
              }
            }
          }
          if (p.out2 != nullptr) {
            // second, scaled 16-bit copy of the output: the tensor-core operand of a conv that consumes this
            // (un-normalised) tensor directly; the power-of-two scale keeps fp16 far from overflow
            uint4* o2 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out2) + off + c0);
            const float s2 = p.out2_scale;
            if (p.out2_dtype == DT_BF16) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                o2[j] = make_uint4(pack_bf16x2(f[8 * j] * s2, f[8 * j + 1] * s2), pack_bf16x2(f[8 * j + 2] * s2, f[8 * j + 3] * s2),
                                   pack_bf16x2(f[8 * j + 4] * s2, f[8 * j + 5] * s2), pack_bf16x2(f[8 * j + 6] * s2, f[8 * j + 7] * s2));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                o2[j] = make_uint4(pack_f16x2(f[8 * j] * s2, f[8 * j + 1] * s2), pack_f16x2(f[8 * j + 2] * s2, f[8 * j + 3] * s2),
                                   pack_f16x2(f[8 * j + 4] * s2, f[8 * j + 5] * s2), pack_f16x2(f[8 * j + 6] * s2, f[8 * j + 7] * s2));
            }
          }
          if (p.out_dtype == DT_F32) {
            if (p.round_tf32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = round_tf32(f[j]);
            }
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + off + c0);
            if (p.out_dtype == DT_BF16) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                op[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                   pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                op[j] = make_uint4(pack_f16x2(f[8 * j], f[8 * j + 1]), pack_f16x2(f[8 * j + 2], f[8 * j + 3]),
                                   pack_f16x2(f[8 * j + 4], f[8 * j + 5]), pack_f16x2(f[8 * j + 6], f[8 * j + 7]));
            }
          }
        }
        if (p.stats != nullptr) {
          // GroupNorm partial statistics of what was just written (0 for masked pixels / columns)
          const bool live = valid && cols_ok;
          if (cpg == 4) emit_group_stats<4>(f, live, lane, stat_s + (q * 32 + c0 / 4) * 2);
          else if (cpg == 8) emit_group_stats<8>(f, live, lane, stat_s + (q * 32 + c0 / 8) * 2);
          else emit_group_stats<16>(f, live, lane, stat_s + (q * 32 + c0 / 16) * 2);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) rcur[j] = rnext[j];
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
      if (p.stats != nullptr) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const int groups_in_tile = BLOCK_N / cpg;        // 32 or 16
        if (et < groups_in_tile * 2) {
          const int gl = et >> 1, k = et & 1;
          const int gg = n0 / cpg + gl;                  // group index in the layer
          if (gg < 32) {
            const float t = ((stat_s[(0 * 32 + gl) * 2 + k] + stat_s[(1 * 32 + gl) * 2 + k]) +
                             (stat_s[(2 * 32 + gl) * 2 + k] + stat_s[(3 * 32 + gl) * 2 + k]));
            p.stats[(((long long)img * p.stats_chunks_per_img + p.stats_chunk0 + rem) * 32 + gg) * 2 + k] = t;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

static int make_maps(const GemmParams& p, int block_n, TensorMapPair* maps) {
  PFN_encodeTiled enc = get_encode_fn();
  HDRVAE_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const int eb = dt_bytes(p.ab_dtype);
  const int vec = 16 / eb;
  const int row_elems = kRowBytes / eb;
  const CUtensorMapDataType dt = p.ab_dtype == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                               : p.ab_dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  HDRVAE_REQUIRE((reinterpret_cast<uintptr_t>(p.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.b) & 15) == 0,
                 "gemm_tc: operand pointers must be 16-byte aligned");
  HDRVAE_REQUIRE(p.a_px_stride % vec == 0 && p.a_row_stride % vec == 0 && p.a_img_stride % vec == 0 && p.b_row_stride % vec == 0,
                 "gemm_tc: strides must be multiples of 16 bytes");
  {
    // A: {C, W, H, N}; the channel extent visible to TMA is k_per_tap (columns beyond are never addressed)
    cuuint64_t dims[4] = {(cuuint64_t)p.k_per_tap, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.n_img};
    cuuint64_t strides[3] = {(cuuint64_t)p.a_px_stride * eb, (cuuint64_t)p.a_row_stride * eb,
                             (cuuint64_t)p.a_img_stride * eb};
    cuuint32_t box[4] = {(cuuint32_t)row_elems, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&maps->a, dt, 4, const_cast<void*>(p.a), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HDRVAE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (W=%d H=%d N=%d K=%d)", (int)r, p.W, p.H,
                   p.n_img, p.k_per_tap);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.k_per_tap * p.ntaps, (cuuint64_t)(p.b_rows > 0 ? p.b_rows : p.n_cols)};
    cuuint64_t strides[1] = {(cuuint64_t)p.b_row_stride * eb};
    cuuint32_t box[2] = {(cuuint32_t)row_elems, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps->b, dt, 2, const_cast<void*>(p.b), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HDRVAE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d (K=%d cols=%d)", (int)r,
                   p.k_per_tap * p.ntaps, p.n_cols);
  }
  return 0;
}

// Chooses the pixel-patch shape (TW x TH = 128) that wastes the fewest MMA rows.
void choose_tile(int H, int W, GemmParams* p) {
  long long best = -1;
  for (int lg = 7; lg >= 0; --lg) {
    const int tw = 1 << lg, th = 128 >> lg;
    const long long cover = (long long)((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (best < 0 || cover < best) {
      best = cover;
      p->tw_log2 = lg; p->TW = tw; p->TH = th;
    }
  }
  p->tiles_x = (W + p->TW - 1) / p->TW;
  p->tiles_y = (H + p->TH - 1) / p->TH;
}

template <int BLOCK_N, bool kTf32>
static int launch_tc(const GemmParams& p_in, int num_sms, cudaStream_t stream) {
  GemmParams p = p_in;
  using Cfg = TcConfig<BLOCK_N>;
  p.n_tiles_n = (p.n_cols + BLOCK_N - 1) / BLOCK_N;
  TensorMapPair maps;
  HDRVAE_TRY(make_maps(p, BLOCK_N, &maps));
  static bool attr_set = false;
  if (!attr_set) {
    HDRVAE_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, kTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_set = true;
  }
  const long long num_tiles = (long long)p.n_img * p.tiles_x * p.tiles_y * p.n_tiles_n;
  const int grid = (int)(num_tiles < num_sms ? num_tiles : num_sms);
  gemm_tc_kernel<BLOCK_N, kTf32><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(maps.a, maps.b, p);
  HDRVAE_LAUNCHED();
  HDRVAE_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_gemm_tc(const GemmParams& p, int num_sms, cudaStream_t stream) {
  const int row_elems = kRowBytes / dt_bytes(p.ab_dtype);
  HDRVAE_REQUIRE(p.k_per_tap % row_elems == 0 && p.k_per_tap > 0, "gemm_tc: K per tap (%d) must be a multiple of %d",
                 p.k_per_tap, row_elems);
  HDRVAE_REQUIRE(p.n_cols % 32 == 0, "gemm_tc: output columns (%d) must be a multiple of 32", p.n_cols);
  HDRVAE_REQUIRE(p.TW * p.TH == kBlockM, "gemm_tc: tile must cover 128 pixels");
  HDRVAE_REQUIRE(p.stats == nullptr || (p.n_cols % 128 == 0 && p.n_cols <= 512),
                 "gemm_tc: GroupNorm statistics need 128/256/512 output channels");
  const bool tf32 = p.ab_dtype == DT_F32;
  if (p.n_cols <= 128) return tf32 ? launch_tc<128, true>(p, num_sms, stream) : launch_tc<128, false>(p, num_sms, stream);
  return tf32 ? launch_tc<256, true>(p, num_sms, stream) : launch_tc<256, false>(p, num_sms, stream);
}

}  // namespace hdrvae
