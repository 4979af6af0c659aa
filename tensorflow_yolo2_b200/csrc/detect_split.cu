// detect_split.cu -- a': region decode + score threshold + per-class NMS as TWO launches sized for the HBM roofline.
// (Absent from the reference -- SURVEY Appendix A.  IoU arithmetic: net_utils.py:222-260 via nms_common.cuh.  Same
// results as detect_fused.cu / y2_decode_region + y2_nms: boxes, dense thresholded scores, bit-identical keep lists.)
//
// Why two kernels.  detect_fused.cu gives each image to ONE CTA, so the image's load -> decode -> NMS chain is serial,
// 256 images on 148 SMs quantise to two rounds on most SMs, and every phase is latency-bound (ncu r1c: 26.7 us for
// 67 MB, IPC 0.3).  Here the streaming part -- 99 % of the bytes -- is spread over the whole chip with no per-image
// structure, and only the tiny candidate lists meet again per image:
//   detect_decode_kernel   one CTA per 32 CELLS (any image): ONE 16 000-byte bulk copy (cp.async.bulk, completion on an
//                          mbarrier) stages the 32 x 125 floats; 160 threads decode one (cell, anchor) each with the fast
//                          intrinsics; boxes go out as coalesced float4s, the 160 x 20 thresholded scores are staged in
//                          smem and leave as ONE 12 800-byte bulk store; every score above the threshold appends a
//                          64-bit key (~score_bits << 12 | box) to its (image, class) list in the workspace.
//   detect_nms_kernel      one CTA per image, one warp per class: count == 0 -> done; otherwise rank-sort the keys,
//                          fetch the candidates' boxes (L2 hits: just written), and run the same register-resident
//                          two-pass sweep as detect_fused.cu.  Counts are reset by their reader, so the workspace is
//                          all-zero between calls.  Images where a class overflows its list (only with a near-zero
//                          threshold) are redone by the CTA with the general algorithm of nms.cu on the dense scores.
#include "nms_body.cuh"

namespace y2 {

constexpr int DS_CELLS = 32;                 // cells per chunk (16 000 B in, 12 800 B of scores out)
constexpr int DS_CHUNKS = 2;                 // chunks per decode CTA
constexpr int DS_CAPK = 64;                  // candidates per (image, class) list
constexpr int DS_NMS_THREADS = 640;          // 20 warps: one per class; two CTAs per SM -> 256 images are one wave

// q = n / d for 0 <= n < 2^31 with host-computed (mul, shr); d == 1 encoded as mul == 0
__device__ __forceinline__ uint32_t fdiv32(uint32_t n, uint32_t mul, uint32_t shr) { return mul ? (__umulhi(n, mul) >> shr) : n; }
static inline void fastdiv32_init(uint32_t d, uint32_t* mul, uint32_t* shr) {
  if (d <= 1) { *mul = 0; *shr = 0; return; }
  uint32_t l = 0;
  while ((1u << l) < d) ++l;                       // ceil(log2 d)
  const uint32_t p = 31 + l;
  *mul = (uint32_t)(((1ull << p) + d - 1) / d);
  *shr = p - 32;
}

__device__ __forceinline__ float ds_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

template <int C, int A>
__global__ void __launch_bounds__(DS_CELLS* A) detect_decode_kernel(
    const float* __restrict__ net, const float* __restrict__ anchors, int S, float score_thresh, float* __restrict__ boxes,
    float* __restrict__ scores, unsigned long long* __restrict__ cand_keys, int* __restrict__ cand_cnt, long long total_cells,
    uint32_t fd_nc_mul, uint32_t fd_nc_shr, uint32_t fd_s_mul, uint32_t fd_s_shr) {
  constexpr int PER = 5 + C, CH = A * PER, NT = DS_CELLS * A;
  constexpr uint32_t IN_BYTES = DS_CELLS * CH * 4, OUT_BYTES = NT * C * 4;
  // DS_CHUNKS chunks per CTA: all bulk loads are issued up front, so chunk c + 1 lands while chunk c is decoded; the
  // thresholded scores of a chunk are staged IN PLACE (its input is dead once every thread holds its values in registers)
  __shared__ __align__(128) float s_buf[DS_CHUNKS][DS_CELLS * CH];
  __shared__ __align__(8) uint64_t s_bar[DS_CHUNKS];
  const int tid = threadIdx.x;
  const long long cta_cell0 = (long long)blockIdx.x * (DS_CELLS * DS_CHUNKS);
  if (tid == 0) {
#pragma unroll
    for (int c = 0; c < DS_CHUNKS; ++c) {
      const long long cell0 = cta_cell0 + c * DS_CELLS;
      if (total_cells - cell0 >= DS_CELLS) {                  // whole chunk: 16-byte granular (base alignment checked on the host)
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[c]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(IN_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_buf[c])),
                     "l"(net + cell0 * CH), "r"(IN_BYTES), "r"(bar)
                     : "memory");
      }
    }
  }
  __syncthreads();                                           // barriers initialised and armed before anyone polls them
  const float fs = (float)S;
  const int lc = tid / A, a = tid - lc * A;
  const float aw = anchors[2 * a + 0], ah = anchors[2 * a + 1];
#pragma unroll 1
  for (int c = 0; c < DS_CHUNKS; ++c) {
    const long long cell0 = cta_cell0 + c * DS_CELLS;
    if (cell0 >= total_cells) break;                         // block-uniform
    const int nc = (int)min((long long)DS_CELLS, total_cells - cell0);
    const bool bulk = nc == DS_CELLS;
    float* buf = s_buf[c];
    if (bulk) {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[c]);
      uint32_t ok = 0;
      long long t0 = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar) : "memory");
        if (!ok) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > 4000000000ll) { printf("y2 detect_decode: bulk load never completed (block %d)\n", blockIdx.x); __trap(); }
        }
      }
    } else {
      const float* src = net + cell0 * CH;
      for (int e = tid; e < nc * CH; e += NT) buf[e] = __ldcs(src + e);
      __syncthreads();
    }
    const bool active = tid < nc * A;
    float e[C];
    unsigned cm = 0;                                          // classes above the threshold (rare)
    uint32_t img = 0, cell = 0;
    if (active) {
      const uint32_t g = (uint32_t)(cell0 + lc);              // global cell (host: total_cells < 2^31)
      img = fdiv32(g, fd_nc_mul, fd_nc_shr);
      cell = g - img * (uint32_t)(S * S);
      const uint32_t i = fdiv32(cell, fd_s_mul, fd_s_shr), j = cell - i * (uint32_t)S;
      const float* q = buf + tid * PER;                       // stride 25 floats: conflict-free
      float mx = q[5];
#pragma unroll
      for (int k = 1; k < C; ++k) mx = fmaxf(mx, q[5 + k]);
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < C; ++k) { e[k] = __expf(q[5 + k] - mx); sum += e[k]; }
      const float w = __fdividef(ds_sigmoid(q[4]), sum);            // objectness / softmax denominator
      const float bx = __fdividef((float)j + ds_sigmoid(q[0]), fs);
      const float by = __fdividef((float)i + ds_sigmoid(q[1]), fs);
      const float bw = __fdividef(aw * __expf(q[2]), fs);
      const float bh = __fdividef(ah * __expf(q[3]), fs);
      reinterpret_cast<float4*>(boxes)[(size_t)cell0 * A + tid] = make_float4(bx, by, bw, bh);   // (img * nbox + cell * A + a)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float v = e[k] * w;
        cm |= v > score_thresh ? (1u << k) : 0u;
        e[k] = v > score_thresh ? v : 0.0f;
      }
    }
    __syncthreads();                                          // every thread holds its inputs in registers: the buffer is free
    if (active) {
      float* so = buf + tid * C;
      if ((C & 3) == 0) {
#pragma unroll
        for (int k = 0; k < C; k += 4) *reinterpret_cast<float4*>(so + k) = make_float4(e[k], e[k + 1], e[k + 2], e[k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < C; ++k) so[k] = e[k];
      }
      if (cm) {
        const unsigned b = cell * A + a;                       // box index within the image (< 4096)
        while (cm) {
          const int k = __ffs(cm) - 1;
          cm &= cm - 1;
          const int pos = atomicAdd(&cand_cnt[img * C + k], 1);
          if (pos < DS_CAPK)
            cand_keys[((size_t)img * C + k) * DS_CAPK + pos] = ((unsigned long long)(~__float_as_uint(so[k])) << 12) | b;
        }
      }
    }
    if (scores) {
      float* dst = scores + (size_t)cell0 * A * C;
      if (bulk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my generic-proxy smem writes -> visible to the bulk engine
        __syncthreads();
        if (tid == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                       "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(OUT_BYTES)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        __syncthreads();
        for (int e2 = tid; e2 < nc * A * C; e2 += NT) dst[e2] = buf[e2];
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must outlive the bulk stores' reads
}

template <int C>
__global__ void __launch_bounds__(DS_NMS_THREADS, 2) detect_nms_kernel(
    const float* __restrict__ boxes, const float* __restrict__ scores, unsigned long long* __restrict__ cand_keys,
    int* __restrict__ cand_cnt, int nbox, float score_thresh, float iou_thresh, int32_t* __restrict__ keep_idx,
    int32_t* __restrict__ keep_count, float* __restrict__ keep_score, int max_keep, size_t smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_cnt[C];
  __shared__ unsigned long long s_keys[C][DS_CAPK];
  __shared__ float4 s_c4[DS_NMS_THREADS / 32][32];            // per-warp corner table of the <= 32 candidate path
  __shared__ float s_ar[DS_NMS_THREADS / 32][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.x;
  if (tid < C) {
    int* p = cand_cnt + (size_t)img * C + tid;
    s_cnt[tid] = *p;
    *p = 0;                                                    // read once, reset: the workspace is all-zero between calls
  }
  __syncthreads();
  int32_t* kc = keep_count + (size_t)img * C;
  {
    bool overflow = false;
#pragma unroll
    for (int k = 0; k < C; ++k) overflow |= s_cnt[k] > DS_CAPK;
    if (overflow) {
      if (scores == nullptr) {
        for (int k = tid; k < C; k += DS_NMS_THREADS) kc[k] = -1;     // reported, not silently dropped
        return;
      }
      nms_image_fallback<DS_NMS_THREADS>(smem_raw, smem_bytes, boxes, scores, img, nbox, C, score_thresh, iou_thresh, keep_idx,
                                         keep_count, keep_score, max_keep);
      return;
    }
  }
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)img * nbox;
  for (int k = warp; k < C; k += DS_NMS_THREADS / 32) {
    const int m = s_cnt[k];
    if (m == 0) {
      if (lane == 0) kc[k] = 0;
      continue;
    }
    const unsigned long long* gk = cand_keys + ((size_t)img * C + k) * DS_CAPK;
    unsigned long long* keys = s_keys[k];
    const unsigned long long k0 = lane < m ? __ldcg(gk + lane) : ~0ull;
    const unsigned long long k1 = lane + 32 < m ? __ldcg(gk + lane + 32) : ~0ull;
    keys[lane] = k0;
    keys[lane + 32] = k1;
    __syncwarp();
    int r0 = 0, r1 = 0;
    for (int i = 0; i < m; ++i) {
      const unsigned long long o = keys[i];            // smem broadcast
      r0 += o < k0;
      r1 += o < k1;
    }
    __syncwarp();
    if (lane < m) keys[r0] = k0;                        // keys are unique (box index) -> ranks are a permutation
    if (lane + 32 < m) keys[r1] = k1;
    __syncwarp();
    int32_t* out = keep_idx + ((size_t)img * C + k) * max_keep;
    float* outs = keep_score ? keep_score + ((size_t)img * C + k) * max_keep : nullptr;
    int count = 0;
    if (m <= 32) {
      // one candidate per lane, everything in registers.  Pass 1: lane j collects the set of earlier candidates i that
      // would suppress it (IoU > thresh).  The IEEE division of get_iou only runs when some lane's intersection with
      // candidate i is non-empty (inter == 0 gives IoU 0 exactly, so skipping it cannot change a decision).
      const unsigned long long key = lane < m ? keys[lane] : 0ull;
      const int bi = (int)(key & 0xfffu);
      const Corner cj = to_corner(__ldcg(bx + (lane < m ? bi : 0)));
      // corners of all candidates in a per-warp smem table: the loop below reads candidate i by broadcast, has no
      // cross-lane exchange and no loop-carried dependency, so its iterations overlap (the shuffle / vote version cost
      // ~130 dependent clk per candidate and made this kernel latency-bound: 17 us for 60 k candidates)
      s_c4[warp][lane] = make_float4(cj.x1, cj.y1, cj.x2, cj.y2);
      s_ar[warp][lane] = cj.area;
      __syncwarp();
      unsigned supby = 0;
#pragma unroll 4
      for (int i = 0; i + 1 < m; ++i) {
        const float4 c4 = s_c4[warp][i];
        const float iw = fmaxf(0.0f, __fsub_rn(fminf(c4.z, cj.x2), fmaxf(c4.x, cj.x1)));
        const float ih = fmaxf(0.0f, __fsub_rn(fminf(c4.w, cj.y2), fmaxf(c4.y, cj.y1)));
        if (lane > i && lane < m && __fmul_rn(iw, ih) > 0.0f) {    // inter == 0 gives IoU 0 exactly: no division needed
          Corner ci;
          ci.x1 = c4.x; ci.y1 = c4.y; ci.x2 = c4.z; ci.y2 = c4.w; ci.area = s_ar[warp][i];
          if (iou_corner(ci, cj) > iou_thresh) supby |= 1u << i;
        }
      }
      __syncwarp();
      // Pass 2: greedy resolution in score order on bit masks
      unsigned alive = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
      for (int i = 0; i < m; ++i) {
        if (!((alive >> i) & 1u)) continue;            // warp-uniform
        alive &= ~__ballot_sync(0xffffffffu, (supby >> i) & 1u);
      }
      count = __popc(alive);
      if ((alive >> lane) & 1u) {
        const int slot = __popc(alive & ((1u << lane) - 1u));
        if (slot < max_keep) {
          out[slot] = bi;
          if (outs) outs[slot] = __uint_as_float(~(unsigned)(key >> 12));
        }
      }
    } else {
      // 33..64 candidates: two per lane (j = lane and lane + 32), still in registers; candidate i's corner and its
      // removed flag are broadcast from the owning lane
      const int b0 = (int)(keys[lane] & 0xfffu), b1 = lane + 32 < m ? (int)(keys[lane + 32] & 0xfffu) : 0;
      const Corner c0 = to_corner(__ldcg(bx + b0)), c1 = to_corner(__ldcg(bx + b1));
      bool rem0 = false, rem1 = false;
      for (int i = 0; i < m; ++i) {
        const int src = i & 31;
        const bool hi = i >= 32;                         // warp-uniform
        if (__shfl_sync(0xffffffffu, (int)(hi ? rem1 : rem0), src)) continue;
        const unsigned long long key = keys[i];
        if (lane == 0 && count < max_keep) {
          out[count] = (int)(key & 0xfffu);
          if (outs) outs[count] = __uint_as_float(~(unsigned)(key >> 12));
        }
        ++count;
        Corner ci;
        ci.x1 = __shfl_sync(0xffffffffu, hi ? c1.x1 : c0.x1, src);
        ci.y1 = __shfl_sync(0xffffffffu, hi ? c1.y1 : c0.y1, src);
        ci.x2 = __shfl_sync(0xffffffffu, hi ? c1.x2 : c0.x2, src);
        ci.y2 = __shfl_sync(0xffffffffu, hi ? c1.y2 : c0.y2, src);
        ci.area = __shfl_sync(0xffffffffu, hi ? c1.area : c0.area, src);
        if (lane > i && !rem0 && iou_corner(ci, c0) > iou_thresh) rem0 = true;                     // j = lane (< 32 < m)
        if (lane + 32 > i && lane + 32 < m && !rem1 && iou_corner(ci, c1) > iou_thresh) rem1 = true; // j = lane + 32
      }
    }
    if (lane == 0) kc[k] = count;
  }
}

static size_t ds_workspace_bytes(int N, int C) {
  const size_t cnt = (((size_t)N * C * sizeof(int)) + 255) & ~(size_t)255;
  return cnt + (size_t)N * C * DS_CAPK * sizeof(unsigned long long);
}

}  // namespace y2

using namespace y2;

extern "C" size_t y2_detect_workspace_bytes(int N, int C) { return N > 0 && C > 0 ? ds_workspace_bytes(N, C) : 0; }

extern "C" int y2_detect_split(const float* net, const float* anchors, int N, int S, int A, int C, float score_thresh,
                               float iou_thresh, float* boxes, float* scores, int32_t* keep_idx, int32_t* keep_count,
                               float* keep_score, int max_keep, void* workspace, size_t workspace_bytes, y2_stream_t stream) {
  Y2_ARG(net && anchors && boxes && keep_idx && keep_count && N > 0 && S > 0 && A > 0 && max_keep > 0);
  Y2_ARG((((uintptr_t)boxes) & 15) == 0 && score_thresh >= 0.0f);
  const int nbox = S * S * A;
  if (C != 20 || A != 5 || nbox > 4095 || (((uintptr_t)net) & 15) != 0 || (scores && (((uintptr_t)scores) & 15) != 0) ||
      (long long)N * S * S >= (1ll << 31) / 8) {
    set_error("y2_detect_split: C=%d A=%d / %d boxes / alignment unsupported (C == 20, A == 5, <= 4095 boxes, 16-byte aligned net and "
              "scores); use y2_detect_fused", C, A, nbox);
    return Y2_ERR_UNSUPPORTED;
  }
  if (!workspace || workspace_bytes < ds_workspace_bytes(N, C) || (((uintptr_t)workspace) & 255) != 0) {
    set_error("y2_detect_split: workspace too small or misaligned (%zu < %zu)", workspace_bytes, ds_workspace_bytes(N, C));
    return Y2_ERR_WORKSPACE;
  }
  int* cnt = reinterpret_cast<int*>(workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(workspace) +
                                                                   ((((size_t)N * C * sizeof(int)) + 255) & ~(size_t)255));
  cudaStream_t st = (cudaStream_t)stream;
  const long long total_cells = (long long)N * S * S;
  uint32_t nc_mul, nc_shr, s_mul, s_shr;
  fastdiv32_init((uint32_t)(S * S), &nc_mul, &nc_shr);
  fastdiv32_init((uint32_t)S, &s_mul, &s_shr);
  const int grid = (int)((total_cells + DS_CELLS * DS_CHUNKS - 1) / (DS_CELLS * DS_CHUNKS));
  detect_decode_kernel<20, 5><<<grid, DS_CELLS * 5, 0, st>>>(net, anchors, S, score_thresh, boxes, scores, keys, cnt, total_cells,
                                                              nc_mul, nc_shr, s_mul, s_shr);
  Y2_LAUNCHED();
  // dynamic smem only for the overflow fallback (bit-matrix NMS of nms.cu over all boxes of the image)
  int P = 1;
  while (P < nbox) P <<= 1;
  const size_t smem = (size_t)P * 8 + (size_t)nbox * sizeof(Corner) + 16 + 16 * 1024;   // small: three CTAs per SM
  if (smem > 32 * 1024)                                         // static (10 KB) + dynamic beyond 48 KB needs the opt-in
    Y2_CUDA(cudaFuncSetAttribute(detect_nms_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
  detect_nms_kernel<20><<<N, DS_NMS_THREADS, smem, st>>>(boxes, scores, keys, cnt, nbox, score_thresh, iou_thresh, keep_idx,
                                                          keep_count, keep_score, max_keep, smem);
  Y2_LAUNCHED();
  return Y2_OK;
}
