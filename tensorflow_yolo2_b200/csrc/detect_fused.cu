// detect_fused.cu -- a': region decode + score threshold + per-class NMS in ONE kernel, one CTA per image.
// (Absent from the reference -- SURVEY Appendix A; the reference decodes on the host with NumPy,
// yolo2_nets/net_utils.py:393-421, and has no NMS.  IoU arithmetic: net_utils.py:222-260 via nms_common.cuh.)
//
// Why fused: decode and NMS are HBM-bound and tiny (166 KB + 98 KB of algorithmic traffic per 13x13 image); as two
// kernels the [N, S*S*A, C] score tensor makes a round trip through memory and the NMS pays one CTA per
// (image, class).  At 256 images the whole job is ~15 us of HBM time, so what matters is the LATENCY of one image's
// chain and how few block-wide barriers it has:
//   1. the image (84.5 KB at 13x13; 128-cell chunks at 19x19) is staged with 128-bit streaming loads, all in flight
//      at once; one THREAD per (cell, anchor) reads its 5+C values from smem (stride 25 floats: conflict-free),
//      computes sigmoid / exp / softmax in registers with the fast intrinsics (ex2.approx / rcp.approx: ~5e-7
//      relative, the spec allows 1e-5), writes its box (smem + global) and its dense thresholded scores (global,
//      optional), and appends every score above the threshold to ITS CLASS's candidate list in smem as a 64-bit key
//      ~score_bits << 12 | box_index (smem atomics on 20 counters);
//   2. ONE barrier; then each warp owns whole classes: rank-sorts the class's keys (score desc, box index asc)
//      with warp-local reads, and runs the greedy sweep -- for <= 32 candidates entirely in registers (one candidate
//      per lane, corner broadcast by shuffle, suppression mask by ballot).
// Keep lists are bit-identical to the oracle's NMS on the kernel's own boxes / scores (unfused IEEE IoU ops).
// Images where some class has more than DF_CAPK candidates (only with a near-zero threshold) are redone in place by
// the same CTA with the general algorithm of nms.cu (nms_body.cuh) on the dense scores it has just written.
#include "nms_body.cuh"

namespace y2 {

constexpr int DF_THREADS = 448;          // 14 warps, <= 72 registers: two CTAs per SM overlap one image's load with
                                         // the other's compute
constexpr int DF_CAPK = 64;              // candidates per (image, class) held in smem
constexpr int DF_MAX_CHUNK_CELLS = 176;  // whole 13x13 image in one chunk; larger grids go in 128-cell chunks

__device__ __forceinline__ float df_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }

template <int C>
__global__ void __launch_bounds__(DF_THREADS, 2) detect_fused_kernel(
    const float* __restrict__ net, const float* __restrict__ anchors, int S, int A, float score_thresh, float iou_thresh,
    float* __restrict__ boxes, float* __restrict__ scores, int32_t* __restrict__ keep_idx,
    int32_t* __restrict__ keep_count, float* __restrict__ keep_score, int max_keep, int cells_per_chunk, size_t smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_cnt[C];
  __shared__ unsigned char s_removed[C * DF_CAPK];
  const int per = 5 + C, ch = A * per;
  const int ncell = S * S, nbox = ncell * A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.x;
  // smem: boxes float4[nbox] | keys u64[C][DF_CAPK] | chunk float[cells_per_chunk*ch + 4]
  float4* s_box = reinterpret_cast<float4*>(smem_raw);
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw + (size_t)nbox * 16);
  float* s_in = reinterpret_cast<float*>(s_keys + C * DF_CAPK);
  if (tid < C) s_cnt[tid] = 0;
  const float fs = (float)S;
  const float* src_img = net + (size_t)img * ncell * ch;
  for (int cell0 = 0; cell0 < ncell; cell0 += cells_per_chunk) {
    const int nc = min(cells_per_chunk, ncell - cell0);
    __syncthreads();                                   // previous chunk consumed (and s_cnt initialised)
    // stage [src, src + nfl) through the 16-byte aligned superset, keeping the misalignment in smem
    const float* src = src_img + (size_t)cell0 * ch;
    const int nfl = nc * ch;
    const int mis = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2);       // 0..3 floats
    {
      // whole 16-byte vectors [v0, nv) lie inside this chunk; the partial first / last vectors go float by float.
      // Loads are issued four deep per thread before the first dependent store: a plain load-store loop leaves ONE
      // DRAM round trip in flight per thread and turns the 12 iterations into 12 serial latencies (measured: 8 us).
      const float4* vsrc = reinterpret_cast<const float4*>(src - mis);
      const int nv = (mis + nfl) >> 2;
      const int v0 = mis ? 1 : 0;
      for (int base = v0 + tid; base < nv; base += 4 * DF_THREADS) {
        float4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (base + u * DF_THREADS < nv) r[u] = __ldcs(vsrc + base + u * DF_THREADS);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (base + u * DF_THREADS < nv) reinterpret_cast<float4*>(s_in)[base + u * DF_THREADS] = r[u];
      }
      if (mis && tid >= mis && tid < 4) s_in[tid] = __ldcs(src - mis + tid);
      for (int e = (nv << 2) + tid; e < mis + nfl; e += DF_THREADS) s_in[e] = __ldcs(src - mis + e);
    }
    __syncthreads();
    for (int t = tid; t < nc * A; t += DF_THREADS) {
      const int lc = t / A, a = t - lc * A;
      const int cell = cell0 + lc;
      const int i = cell / S, j = cell - i * S;
      const float* q = s_in + mis + (size_t)t * per;   // == (lc*A + a) * per
      float mx = q[5];
#pragma unroll
      for (int k = 1; k < C; ++k) mx = fmaxf(mx, q[5 + k]);
      float e[C];
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < C; ++k) { e[k] = __expf(q[5 + k] - mx); sum += e[k]; }
      const float w = __fdividef(df_sigmoid(q[4]), sum);              // objectness / softmax denominator
      const float bx = __fdividef((float)j + df_sigmoid(q[0]), fs);
      const float by = __fdividef((float)i + df_sigmoid(q[1]), fs);
      const float bw = __fdividef(anchors[2 * a + 0] * __expf(q[2]), fs);
      const float bh = __fdividef(anchors[2 * a + 1] * __expf(q[3]), fs);
      const int b = cell * A + a;
      const float4 box = make_float4(bx, by, bw, bh);
      s_box[b] = box;
      reinterpret_cast<float4*>(boxes)[(size_t)img * nbox + b] = box;
      unsigned cm = 0;                                  // classes above the threshold (rare: ~0.3 per box)
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float v = e[k] * w;
        cm |= v > score_thresh ? (1u << k) : 0u;
        e[k] = v > score_thresh ? v : 0.0f;
      }
      if (scores) {
        float* dst = scores + ((size_t)img * nbox + b) * C;
        if ((C & 3) == 0) {
#pragma unroll
          for (int k = 0; k < C; k += 4) __stcs(reinterpret_cast<float4*>(dst + k), make_float4(e[k], e[k + 1], e[k + 2], e[k + 3]));
        } else {
#pragma unroll
          for (int k = 0; k < C; ++k) dst[k] = e[k];
        }
      }
      if (__any_sync(__activemask(), cm != 0)) {
        // append (class, score, box) keys: the scores go back to the thread's own smem slot so that the loop over
        // the set bits can index them (a register array indexed at run time would live in local memory)
        float* qs = const_cast<float*>(q) + 5;
#pragma unroll
        for (int k = 0; k < C; ++k) qs[k] = e[k];
        while (cm) {
          const int k = __ffs(cm) - 1;
          cm &= cm - 1;
          const int pos = atomicAdd(&s_cnt[k], 1);
          if (pos < DF_CAPK) s_keys[k * DF_CAPK + pos] = ((unsigned long long)(~__float_as_uint(qs[k])) << 12) | (unsigned)b;
        }
      }
    }
  }
  __syncthreads();
  int32_t* kc = keep_count + (size_t)img * C;
  {
    bool overflow = false;
#pragma unroll
    for (int k = 0; k < C; ++k) overflow |= s_cnt[k] > DF_CAPK;
    if (overflow) {
      // rare (near-zero threshold): redo the image in place with the general bit-matrix algorithm of nms.cu on the
      // dense scores this CTA has just written (visible to the whole block after the barrier above)
      if (scores == nullptr) {
        for (int k = tid; k < C; k += DF_THREADS) kc[k] = -1;         // reported, not silently dropped
        return;
      }
      nms_image_fallback<DF_THREADS>(smem_raw, smem_bytes, boxes, scores, img, nbox, C, score_thresh, iou_thresh, keep_idx,
                                     keep_count, keep_score, max_keep);
      return;
    }
  }
  // ---- per class: rank sort + greedy sweep, one warp per class, no block-wide barrier from here on ----
  for (int k = warp; k < C; k += DF_THREADS / 32) {
    const int m = s_cnt[k];
    if (m == 0) {
      if (lane == 0) kc[k] = 0;
      continue;
    }
    unsigned long long* keys = s_keys + k * DF_CAPK;
    const unsigned long long k0 = lane < m ? keys[lane] : ~0ull;
    const unsigned long long k1 = lane + 32 < m ? keys[lane + 32] : ~0ull;
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
      const Corner cj = to_corner(s_box[lane < m ? bi : 0]);
      unsigned supby = 0;
      for (int i = 0; i + 1 < m; ++i) {
        Corner ci;
        ci.x1 = __shfl_sync(0xffffffffu, cj.x1, i);
        ci.y1 = __shfl_sync(0xffffffffu, cj.y1, i);
        ci.x2 = __shfl_sync(0xffffffffu, cj.x2, i);
        ci.y2 = __shfl_sync(0xffffffffu, cj.y2, i);
        const float iw = fmaxf(0.0f, __fsub_rn(fminf(ci.x2, cj.x2), fmaxf(ci.x1, cj.x1)));
        const float ih = fmaxf(0.0f, __fsub_rn(fminf(ci.y2, cj.y2), fmaxf(ci.y1, cj.y1)));
        const bool ov = lane > i && lane < m && __fmul_rn(iw, ih) > 0.0f;
        if (__any_sync(0xffffffffu, ov)) {
          ci.area = __shfl_sync(0xffffffffu, cj.area, i);
          if (ov && iou_corner(ci, cj) > iou_thresh) supby |= 1u << i;
        }
      }
      // Pass 2: greedy resolution in score order on bit masks
      unsigned alive = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
      for (int i = 0; i < m; ++i) {
        if (!((alive >> i) & 1u)) continue;            // warp-uniform
        alive &= ~__ballot_sync(0xffffffffu, (supby >> i) & 1u);
      }
      // kept candidates, in visiting order
      count = __popc(alive);
      if ((alive >> lane) & 1u) {
        const int slot = __popc(alive & ((1u << lane) - 1u));
        if (slot < max_keep) {
          out[slot] = bi;
          if (outs) outs[slot] = __uint_as_float(~(unsigned)(key >> 12));
        }
      }
    } else {
      unsigned char* removed = s_removed + k * DF_CAPK;
      removed[lane] = 0;
      removed[lane + 32] = 0;
      __syncwarp();
      for (int i = 0; i < m; ++i) {
        if (removed[i]) continue;                      // warp-uniform (smem broadcast)
        const unsigned long long key = keys[i];
        const int bi = (int)(key & 0xfffu);
        if (lane == 0 && count < max_keep) {
          out[count] = bi;
          if (outs) outs[count] = __uint_as_float(~(unsigned)(key >> 12));
        }
        ++count;
        const Corner ci = to_corner(s_box[bi]);
        for (int jn = i + 1 + lane; jn < m; jn += 32) {
          if (!removed[jn] && iou_corner(ci, to_corner(s_box[(int)(keys[jn] & 0xfffu)])) > iou_thresh) removed[jn] = 1;
        }
        __syncwarp();
      }
    }
    if (lane == 0) kc[k] = count;
  }
}

static size_t df_smem_bytes(int nbox, int cells_per_chunk, int ch, int C) {
  return (size_t)nbox * 16 + (size_t)C * DF_CAPK * 8 + ((size_t)cells_per_chunk * ch + 4) * 4;
}

}  // namespace y2

using namespace y2;

extern "C" int y2_detect_fused(const float* net, const float* anchors, int N, int S, int A, int C, float score_thresh,
                               float iou_thresh, float* boxes, float* scores, int32_t* keep_idx, int32_t* keep_count,
                               float* keep_score, int max_keep, y2_stream_t stream) {
  Y2_ARG(net && anchors && boxes && keep_idx && keep_count && N > 0 && S > 0 && A > 0 && max_keep > 0);
  Y2_ARG((((uintptr_t)boxes) & 15) == 0 && score_thresh >= 0.0f);
  const int nbox = S * S * A;
  if (C != 20 || nbox > 4095) {
    set_error("y2_detect_fused: C=%d / %d boxes unsupported (C == 20, <= 4095 boxes); use y2_decode_region + y2_nms", C, nbox);
    return Y2_ERR_UNSUPPORTED;
  }
  const int cells_per_chunk = S * S <= DF_MAX_CHUNK_CELLS ? S * S : 128;
  const size_t smem = df_smem_bytes(nbox, cells_per_chunk, A * (5 + C), C);
  Y2_ARG(smem <= 200 * 1024);
  cudaStream_t st = (cudaStream_t)stream;
  // (a per-DEVICE function attribute: set on every launch that needs it -- one thread may drive several GPUs)
  if (smem > 48 * 1024)
    Y2_CUDA(cudaFuncSetAttribute(detect_fused_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  detect_fused_kernel<20><<<N, DF_THREADS, smem, st>>>(net, anchors, S, A, score_thresh, iou_thresh, boxes, scores, keep_idx,
                                                       keep_count, keep_score, max_keep, cells_per_chunk, smem);
  Y2_LAUNCHED();
  return Y2_OK;
}
