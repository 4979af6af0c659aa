// conv_streamk_tcgen05.cu -- a1 (+bias) for the 13x13 / 19x19 3x3 layers whose output is the float32 pre-BN tensor
// (the detection head, darknet.py:189-197, whose BN runs on batch statistics; and every such layer in training):
// the same implicit GEMM as conv_tcgen05.cu with 256 x 256 CTA tiles and a stream-K work split.
//
// Why.  conv_tc_kernel on these layers is bound by the operand bytes the L2 can deliver into the SMs -- measured
// 5 700 B/clk for the whole chip (l1tex__m_xbar2l1tex_read_bytes of L19: 2.41 GB in 215 us), not by the tensor pipe
// (46 % busy).  A 128 x 256 tile needs 48 KB per 64-deep K step = 85 flop per delivered byte; a 256 x 256 tile
// needs 64 KB for twice the flops = 128 flop/B, i.e. 1.5x fewer bytes for the same layer.  With K = 9 * 1024 the
// accumulator (2 x 128 x 256 fp32 = all 512 TMEM columns, single-buffered) is drained once per ~150 k clk, so the
// un-overlapped epilogue costs ~2 %.  But 256 x 256 tiles leave only 172 tiles for 148 SMs, so instead of whole
// tiles each CTA takes an equal, contiguous range of (tile, K-step) units ("stream-K"): a CTA's range covers the
// tail of one tile, possibly a whole tile, and the head of the next.  A segment that covers a tile's whole K range
// stores its result; partial segments add theirs into the zero-initialised output with red.global.add.v4.f32.
// Every CTA's range is at least one tile long, so a tile is shared by at most TWO CTAs and each output element
// receives at most two additions onto zero -- which is order-independent, so results stay bit-reproducible.
//
// Scope: ksize 1 or 3, stride 1, SAME; Cin % 64 == 0; Cout % 256 == 0; flags == Y2_CONV_OUT_F32 (no scale, no leaky,
// no pool: the raw conv + bias rows that y2_bn_stats / y2_affine_leaky_pool consume).  Anything else stays on
// conv_tc_kernel; y2_conv_fwd_bf16 chooses (env Y2_CONV_NO_STREAMK=1 disables this path).
#include "tc_common.cuh"

namespace y2 {

constexpr int SK_THREADS = 64 + 8 * 32;        // warps 0-7 epilogue (2 per TMEM lane quarter), 8 TMA, 9 MMA
constexpr int SK_STAGES = 3;                   // 64 KB each: A 2 x (128 px x 128 B), B 256 x 128 B
constexpr uint32_t SK_A_HALF = 128 * 128, SK_B_BYTES = 256 * 128, SK_STAGE = 2 * SK_A_HALF + SK_B_BYTES;
constexpr int SK_MAX_UNITS = 192;

struct SkArgs {
  const float* bias;
  float* y;
  long long M;
  int H, W, ldy, pad;
  int cin_p, cchunks, ksteps;       // K steps per tile = taps * cchunks
  int n_tiles, tiles;
  long long units;                  // tiles * ksteps
  uint32_t fd_nt_mul, fd_nt_shr, fd_w_mul, fd_w_shr, fd_h_mul, fd_h_shr, fd_ks_mul, fd_ks_shr;
};

struct SkUnit { int a_c0, kw, kh, b_k; };

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(SK_THREADS, 1)
conv_streamk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SkArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[SK_STAGES], empty_bar[SK_STAGES], tmem_full, tmem_empty;
  __shared__ uint32_t s_tmem_base;
  __shared__ SkUnit s_units[SK_MAX_UNITS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;

  // my contiguous range of (tile, K-step) units
  const long long u_begin = a.units * blockIdx.x / gridDim.x, u_end = a.units * (blockIdx.x + 1) / gridDim.x;

  if (warp == 8) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      for (int s = 0; s < SK_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(&tmem_full, 1);
      mbar_init(&tmem_empty, 8);
      fence_barrier_init();
    }
    for (int u = lane; u < a.ksteps; u += 32) {
      SkUnit d;
      const int tap = u / a.cchunks, cc = u - tap * a.cchunks;
      const int ks = 2 * a.pad + 1;
      d.a_c0 = cc * 64; d.kh = tap / ks; d.kw = tap - d.kh * ks; d.b_k = tap * a.cin_p + d.a_c0;
      s_units[u] = d;
    }
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long u = u_begin;
      while (u < u_end) {
        const uint32_t tile = (uint32_t)fdiv((uint32_t)u, a.fd_ks_mul, a.fd_ks_shr);       // units < 2^31 (host-checked)
        const int k0 = (int)(u - (long long)tile * a.ksteps);
        const int k1 = (int)min((long long)a.ksteps, k0 + (u_end - u));
        const uint32_t mt = fdiv(tile, a.fd_nt_mul, a.fd_nt_shr);
        const int nrow0 = (int)(tile - mt * (uint32_t)a.n_tiles) * 256;
        int w0[2], h0[2], n0[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t m = mt * 256u + (uint32_t)h * 128u;
          if ((long long)m >= a.M) m = 0;                    // half tile entirely past the end: load valid pixels, rows are masked
          const uint32_t row = fdiv(m, a.fd_w_mul, a.fd_w_shr);
          w0[h] = (int)(m - row * (uint32_t)a.W);
          const uint32_t img = fdiv(row, a.fd_h_mul, a.fd_h_shr);
          h0[h] = (int)(row - img * (uint32_t)a.H);
          n0[h] = (int)img;
        }
        for (int k = k0; k < k1; ++k) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_expect_tx(&full_bar[stage], SK_STAGE);
          const SkUnit d = s_units[k];
          const uint32_t sA = smem_base + stage * SK_STAGE;
          const uint32_t bar = smem_u32(&full_bar[stage]);
          tma_load_im2col_4d(sA, &tmA, bar, d.a_c0, w0[0] - a.pad, h0[0] - a.pad, n0[0], (uint16_t)d.kw, (uint16_t)d.kh);
          tma_load_im2col_4d(sA + SK_A_HALF, &tmA, bar, d.a_c0, w0[1] - a.pad, h0[1] - a.pad, n0[1], (uint16_t)d.kw,
                             (uint16_t)d.kh);
          tma_load_2d(sA + 2 * SK_A_HALF, &tmB, bar, d.b_k, nrow0);
          if (++stage == SK_STAGES) { stage = 0; phase ^= 1u; }
        }
        u += k1 - k0;
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(is_leader));
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc0 = make_smem_desc(smem_base, 16u, 8u * 128u, 2u);                    // SWIZZLE_128B, K-major
    const uint64_t bdesc0 = make_smem_desc(smem_base + 2 * SK_A_HALF, 16u, 8u * 128u, 2u);
    int stage = 0;
    uint32_t phase = 0, seg = 0;
    long long u = u_begin;
    while (u < u_end) {
      const uint32_t tile = (uint32_t)fdiv((uint32_t)u, a.fd_ks_mul, a.fd_ks_shr);
      const int k0 = (int)(u - (long long)tile * a.ksteps);
      const int k1 = (int)min((long long)a.ksteps, k0 + (u_end - u));
      mbar_wait(&tmem_empty, (seg & 1u) ^ 1u);                 // the previous segment's accumulators are drained
      tc_fence_after();
      uint32_t accum = 0;
      for (int k = k0; k < k1; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (is_leader) {
          const uint32_t soff = (uint32_t)(stage * SK_STAGE) >> 4;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
              umma_bf16(tmem_base + (uint32_t)(h * 256), adesc0 + soff + (uint32_t)(h * (SK_A_HALF >> 4)) + (uint32_t)(ks * 2),
                        bdesc0 + soff + (uint32_t)(ks * 2), idesc, accum);
            accum = 1;
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == SK_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (is_leader) umma_commit(&tmem_full);
      __syncwarp();
      u += k1 - k0;
      ++seg;
    }
  } else {
    // =========================== epilogue: +bias, store (whole K range) or red.add (partial) ===========================
    const int q = warp & 3, chalf = warp >> 2;                 // TMEM lane quarter, column half of each accumulator
    uint32_t seg = 0;
    long long u = u_begin;
    while (u < u_end) {
      const uint32_t tile = (uint32_t)fdiv((uint32_t)u, a.fd_ks_mul, a.fd_ks_shr);
      const int k0 = (int)(u - (long long)tile * a.ksteps);
      const int k1 = (int)min((long long)a.ksteps, k0 + (u_end - u));
      const bool whole = k0 == 0 && k1 == a.ksteps;
      const uint32_t mt = fdiv(tile, a.fd_nt_mul, a.fd_nt_shr);
      const int col0 = (int)(tile - mt * (uint32_t)a.n_tiles) * 256 + chalf * 128;
      mbar_wait(&tmem_full, seg & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const long long row = (long long)mt * 256 + h * 128 + q * 32 + lane;
        float* dst = a.y + (size_t)row * a.ldy + col0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256 + chalf * 128);
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          tmem_ld_wait();
          if (h == 1 && c == 96) {                             // my last chunk is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty);
          }
          if (row < a.M) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float f0 = __uint_as_float(v[i]), f1 = __uint_as_float(v[i + 1]), f2 = __uint_as_float(v[i + 2]),
                    f3 = __uint_as_float(v[i + 3]);
              if (k0 == 0 && a.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + c + i));
                f0 += b.x; f1 += b.y; f2 += b.z; f3 += b.w;
              }
              if (whole) *reinterpret_cast<float4*>(dst + c + i) = make_float4(f0, f1, f2, f3);
              else red_add_v4(dst + c + i, f0, f1, f2, f3);
            }
          }
          __syncwarp();
        }
      }
      u += k1 - k0;
      ++seg;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// returns Y2_OK and sets *handled = 1 when the layer was issued on this path
int conv_streamk_try(const y2_conv_params* p, cudaStream_t st, int* handled) {
  *handled = 0;
  if (getenv("Y2_CONV_NO_STREAMK")) return Y2_OK;
  if (p->flags != Y2_CONV_OUT_F32 || p->scale != nullptr) return Y2_OK;
  if (!(p->ksize == 1 || p->ksize == 3) || p->Cin % 64 != 0 || p->Cout % 256 != 0) return Y2_OK;
  if (p->H >= 64 && p->W >= 64) return Y2_OK;                   // large maps: the halo-patch mode of conv_tc_kernel wins
  const int ldy = p->ldy > 0 ? p->ldy : p->Cout;
  if (ldy % 4 != 0 || (reinterpret_cast<uintptr_t>(p->y) & 15) != 0) return Y2_OK;
  if (p->shift && (reinterpret_cast<uintptr_t>(p->shift) & 15) != 0) return Y2_OK;
  const int taps = p->ksize * p->ksize, cchunks = p->Cin / 64, ksteps = taps * cchunks;
  if (ksteps < 64 || ksteps > SK_MAX_UNITS) return Y2_OK;       // short K: the un-overlapped epilogue would show
  int rc = load_driver_entry_points();
  if (rc != Y2_OK) return rc;
  SkArgs a;
  memset(&a, 0, sizeof(a));
  a.bias = p->shift;
  a.y = reinterpret_cast<float*>(p->y);
  a.M = (long long)p->N * p->H * p->W;
  a.H = p->H; a.W = p->W; a.ldy = ldy; a.pad = p->ksize / 2;
  a.cin_p = p->Cin; a.cchunks = cchunks; a.ksteps = ksteps;
  a.n_tiles = p->Cout / 256;
  const long long m_tiles = (a.M + 255) / 256;
  a.tiles = (int)(m_tiles * a.n_tiles);
  a.units = (long long)a.tiles * ksteps;
  if ((a.tiles < g_num_sms / 2 && !getenv("Y2_CONV_FORCE_STREAMK")) || a.units >= (1ll << 31) || a.M + 256 >= (1ll << 31)) return Y2_OK;
  fastdiv_init((uint32_t)a.n_tiles, &a.fd_nt_mul, &a.fd_nt_shr);
  fastdiv_init((uint32_t)p->W, &a.fd_w_mul, &a.fd_w_shr);
  fastdiv_init((uint32_t)p->H, &a.fd_h_mul, &a.fd_h_shr);
  fastdiv_init((uint32_t)ksteps, &a.fd_ks_mul, &a.fd_ks_shr);
  // every CTA's range must be at least one tile long (<= 2 CTAs per tile -> order-independent additions)
  const int grid = a.tiles < g_num_sms ? a.tiles : g_num_sms;

  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    cuuint64_t strides[3] = {(cuuint64_t)p->Cin * 2, (cuuint64_t)p->W * p->Cin * 2, (cuuint64_t)p->H * p->W * p->Cin * 2};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    int lower[2] = {-a.pad, -a.pad};
    int upper[2] = {a.pad - (p->ksize - 1), a.pad - (p->ksize - 1)};
    CUresult r = g_encodeIm2col(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->x), dims, strides, lower, upper,
                                64, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS && g_driver_version <= 13010 && (size_t)a.M * p->Cin * 2 < 131072)
      reinterpret_cast<uint64_t*>(&tmA)[1] &= ~(1ull << 21);   // same small-tensor fix-up as conv_tcgen05.cu
    if (r != CUDA_SUCCESS) {
      set_error("conv_streamk: tensor map A encode failed (CUresult %d)", (int)r);
      return Y2_ERR_DRIVER;
    }
    const int Kp = taps * p->Cin;
    cuuint64_t bdims[2] = {(cuuint64_t)Kp, (cuuint64_t)p->Cout};
    cuuint64_t bstrides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t bbox[2] = {64, 256};
    cuuint32_t bestr[2] = {1, 1};
    r = g_encodeTiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p->w_packed), bdims, bstrides, bbox, bestr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_streamk: tensor map B encode failed (CUresult %d)", (int)r);
      return Y2_ERR_DRIVER;
    }
  }
  const size_t smem = (size_t)SK_STAGES * SK_STAGE + 1024;
  Y2_CUDA(cudaFuncSetAttribute(conv_streamk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // partial segments add into the output: zero the rows first (the columns [Cout, ldy) padding is left alone)
  if (ldy == p->Cout) {
    Y2_CUDA(cudaMemsetAsync(p->y, 0, (size_t)a.M * ldy * sizeof(float), st));
  } else {
    Y2_CUDA(cudaMemset2DAsync(p->y, (size_t)ldy * sizeof(float), 0, (size_t)p->Cout * sizeof(float), (size_t)a.M, st));
  }
  conv_streamk_kernel<<<grid, SK_THREADS, smem, st>>>(tmA, tmB, a);
  Y2_LAUNCHED();
  *handled = 1;
  return Y2_OK;
}

}  // namespace y2
