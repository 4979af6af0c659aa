// detect_fused.cu -- a': region decode + score threshold + per-class NMS in ONE kernel, one CTA per image.
// (Absent from the reference -- SURVEY Appendix A; the reference decodes on the host with NumPy,
// yolo2_nets/net_utils.py:393-421, and has no NMS.  IoU arithmetic: net_utils.py:222-260 via nms_common.cuh.)
//
// Why fused: decode and NMS are HBM-bound and tiny (166 KB + 98 KB of algorithmic traffic per 13x13 image); as two
// kernels the [N, S*S*A, C] score tensor makes a round trip through memory and the NMS pays one CTA per
// (image, class).  Here an image's network output is streamed once through shared memory:
//   1. chunks of 64 cells are staged with coalesced 128-bit loads; one THREAD per (cell, anchor) then reads its
//      5+C values from smem (stride 5+C floats = odd -> bank-conflict-free), computes sigmoid / exp / softmax
//      serially in registers (no shuffles), writes its box to smem + global and its dense thresholded scores to
//      global (optional), and appends every (class, score, box) above the threshold to a candidate list in smem
//      as a 64-bit key  class << 44 | ~score_bits << 12 | box_index;
//   2. one bitonic sort of the keys orders candidates by (class asc, score desc, box index asc);
//   3. each warp takes whole class segments and runs the greedy sweep: the next surviving candidate is kept,
//      lanes test it against the rest of the segment and set "removed" flags.
// Keep lists are bit-identical to y2_decode_region + y2_nms (same float32 op order, same tie rule).
// Images with more than DF_CAP candidates (only with a near-zero threshold) are flagged keep_count = -1 and
// re-done by nms_kernel (bit-matrix algorithm) in a second, normally empty, launch.
#include "nms_common.cuh"

namespace y2 {

constexpr int DF_THREADS = 320;          // 64 cells x 5 anchors per chunk
constexpr int DF_CAP = 2048;             // candidate capacity per image

__device__ __forceinline__ float df_sigmoid(float v) { return 1.0f / (1.0f + expf(-v)); }

template <int C>
__global__ void __launch_bounds__(DF_THREADS) detect_fused_kernel(
    const float* __restrict__ net, const float* __restrict__ anchors, int S, int A, float score_thresh, float iou_thresh,
    float* __restrict__ boxes, float* __restrict__ scores, int32_t* __restrict__ keep_idx,
    int32_t* __restrict__ keep_count, float* __restrict__ keep_score, int max_keep, int cells_per_chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_count;
  __shared__ int s_seg[C + 1];
  const int per = 5 + C, ch = A * per;
  const int ncell = S * S, nbox = ncell * A;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.x;
  // smem: boxes float4[nbox] | keys u64[DF_CAP] | removed u8[DF_CAP] | chunk float[cells_per_chunk*ch]
  float4* s_box = reinterpret_cast<float4*>(smem_raw);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + (((size_t)nbox * 16 + 15) & ~(size_t)15));
  unsigned char* removed = reinterpret_cast<unsigned char*>(keys + DF_CAP);
  float* s_in = reinterpret_cast<float*>(removed + DF_CAP);
  if (tid == 0) s_count = 0;
  const float fs = (float)S;
  const float* src_img = net + (size_t)img * ncell * ch;
  for (int cell0 = 0; cell0 < ncell; cell0 += cells_per_chunk) {
    const int nc = min(cells_per_chunk, ncell - cell0);
    __syncthreads();                                   // previous chunk consumed (and s_count initialised)
    {
      const float* src = src_img + (size_t)cell0 * ch;
      const int nfl = nc * ch;
      if ((((uintptr_t)src) & 15) == 0) {
        const int nv = nfl >> 2;
        for (int v = tid; v < nv; v += DF_THREADS) reinterpret_cast<float4*>(s_in)[v] = __ldcs(reinterpret_cast<const float4*>(src) + v);
        for (int e = (nv << 2) + tid; e < nfl; e += DF_THREADS) s_in[e] = __ldcs(src + e);
      } else {
        for (int e = tid; e < nfl; e += DF_THREADS) s_in[e] = __ldcs(src + e);
      }
    }
    __syncthreads();
    for (int t = tid; t < nc * A; t += DF_THREADS) {
      const int lc = t / A, a = t - lc * A;
      const int cell = cell0 + lc;
      const int j = cell % S, i = cell / S;
      const float* q = s_in + (size_t)t * per;         // == (lc*A + a) * per
      // same expressions as decode_region_kernel (decode.cu)
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < C; ++k) mx = fmaxf(mx, q[5 + k]);
      float e[C];
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < C; ++k) { e[k] = expf(q[5 + k] - mx); sum += e[k]; }
      const float obj = df_sigmoid(q[4]);
      const float bx = ((float)j + df_sigmoid(q[0])) / fs;
      const float by = ((float)i + df_sigmoid(q[1])) / fs;
      const float bw = anchors[2 * a + 0] * expf(q[2]) / fs;
      const float bh = anchors[2 * a + 1] * expf(q[3]) / fs;
      const int b = cell * A + a;
      const float4 box = make_float4(bx, by, bw, bh);
      s_box[b] = box;
      reinterpret_cast<float4*>(boxes)[(size_t)img * nbox + b] = box;
      float sc[C];
#pragma unroll
      for (int k = 0; k < C; ++k) {
        const float v = obj * (e[k] / sum);
        sc[k] = v > score_thresh ? v : 0.0f;
        if (v > score_thresh) {
          const int pos = atomicAdd(&s_count, 1);
          if (pos < DF_CAP)
            keys[pos] = ((unsigned long long)k << 44) | ((unsigned long long)(~__float_as_uint(v)) << 12) | (unsigned)b;
        }
      }
      if (scores) {
        float* dst = scores + ((size_t)img * nbox + b) * C;
        if ((C & 3) == 0) {
#pragma unroll
          for (int k = 0; k < C; k += 4) __stcs(reinterpret_cast<float4*>(dst + k), make_float4(sc[k], sc[k + 1], sc[k + 2], sc[k + 3]));
        } else {
#pragma unroll
          for (int k = 0; k < C; ++k) dst[k] = sc[k];
        }
      }
    }
  }
  __syncthreads();
  const int n = s_count;
  int32_t* kc = keep_count + (size_t)img * C;
  if (n > DF_CAP) {                                     // rare: hand the image to the bit-matrix kernel
    for (int k = tid; k < C; k += DF_THREADS) kc[k] = -1;
    return;
  }
  if (n == 0) {
    for (int k = tid; k < C; k += DF_THREADS) kc[k] = 0;
    return;
  }
  // ---- sort by (class asc, score desc, index asc) ----
  int P2 = 1;
  while (P2 < n) P2 <<= 1;
  for (int i = n + tid; i < P2; i += DF_THREADS) keys[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (P2 >> 1); i += DF_THREADS) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const unsigned long long x = keys[lo], y = keys[hi];
        if ((x > y) == asc) { keys[lo] = y; keys[hi] = x; }
      }
      __syncthreads();
    }
  }
  // ---- class segments ----
  for (int k = tid; k <= C; k += DF_THREADS) s_seg[k] = n;
  for (int i = tid; i < n; i += DF_THREADS) removed[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += DF_THREADS) {
    const int k = (int)(keys[i] >> 44);
    if (i == 0 || (int)(keys[i - 1] >> 44) != k) s_seg[k] = i;
  }
  __syncthreads();
  // ---- greedy sweep, one warp per class ----
  for (int k = warp; k < C; k += DF_THREADS / 32) {
    const int beg = s_seg[k];
    int end = n;
    if (beg < n) {                                     // end = start of the next non-empty class
      for (int k2 = k + 1; k2 < C; ++k2)
        if (s_seg[k2] < n) { end = s_seg[k2]; break; }
    }
    int32_t* out = keep_idx + ((size_t)img * C + k) * max_keep;
    float* outs = keep_score ? keep_score + ((size_t)img * C + k) * max_keep : nullptr;
    int count = 0;
    if (beg < n) {
      for (int i = beg; i < end; ++i) {
        if (removed[i]) continue;                      // warp-uniform (smem broadcast)
        const unsigned long long key = keys[i];
        const int bi = (int)(key & 0xfffu);
        if (lane == 0 && count < max_keep) {
          out[count] = bi;
          if (outs) outs[count] = __uint_as_float(~(unsigned)(key >> 12));
        }
        ++count;
        const Corner ci = to_corner(s_box[bi]);
        for (int jn = i + 1 + lane; jn < end; jn += 32) {
          if (!removed[jn] && iou_corner(ci, to_corner(s_box[(int)(keys[jn] & 0xfffu)])) > iou_thresh) removed[jn] = 1;
        }
        __syncwarp();
      }
    }
    if (lane == 0) kc[k] = count;
  }
}

static size_t df_smem_bytes(int nbox, int cells_per_chunk, int ch) {
  return (((size_t)nbox * 16 + 15) & ~(size_t)15) + (size_t)DF_CAP * 8 + DF_CAP + (size_t)cells_per_chunk * ch * 4;
}

// defined in nms.cu
int launch_nms_flagged(const float* boxes, const float* scores, int N, int nbox, int C, float score_thresh, float iou_thresh,
                       int32_t* keep_idx, int32_t* keep_count, float* keep_score, int max_keep, cudaStream_t st);

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
  const int cells_per_chunk = DF_THREADS / A > 0 ? DF_THREADS / A : 1;
  const size_t smem = df_smem_bytes(nbox, cells_per_chunk, A * (5 + C));
  Y2_ARG(smem <= 200 * 1024);
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    Y2_CUDA(cudaFuncSetAttribute(detect_fused_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  detect_fused_kernel<20><<<N, DF_THREADS, smem, st>>>(net, anchors, S, A, score_thresh, iou_thresh, boxes, scores, keep_idx,
                                                       keep_count, keep_score, max_keep, cells_per_chunk);
  Y2_LAUNCHED();
  if (scores) return launch_nms_flagged(boxes, scores, N, nbox, C, score_thresh, iou_thresh, keep_idx, keep_count, keep_score, max_keep,
                                        st);
  return Y2_OK;
}
