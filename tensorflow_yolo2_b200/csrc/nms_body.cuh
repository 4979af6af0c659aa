// nms_body.cuh -- the per-(image, class) greedy NMS of nms.cu as a device function, so that detect_fused.cu can run
// it in place for an image whose candidate lists overflow its fast path.  Algorithm and parity notes: nms.cu.
#pragma once
#include "nms_common.cuh"

namespace y2 {

constexpr int NMS_MAX_MATRIX = 1024;   // 32 lanes x 32 bits

// smem layout (caller-provided, 16-byte aligned): keys[P] u64 | corners[n] (5 floats) | matrix[n * words] u32
// NT = threads of the calling CTA (all of them must call); smem_bytes bounds the bit-matrix path.
template <int NT>
__device__ __forceinline__ void nms_body_t(unsigned char* smem_raw, size_t smem_bytes, const float* __restrict__ boxes,
                                           const float* __restrict__ scores, int nbox, int C, float score_thresh,
                                           float iou_thresh, int32_t* __restrict__ keep_idx, int32_t* __restrict__ keep_count,
                                           float* __restrict__ keep_score, int max_keep, int P, int unit) {
  constexpr int NMS_THREADS = NT;
  __shared__ int s_n;
  __shared__ int s_count;
  __shared__ int s_next;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  const int tid = threadIdx.x;
  const int img = unit / C, k = unit % C;
  __syncthreads();                                   // a previous unit of this CTA may still be reading the shared state
  const float* sc = scores + (size_t)img * nbox * C + k;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (size_t)img * nbox;

  if (tid == 0) { s_n = 0; s_count = 0; }
  for (int i = tid; i < P; i += NMS_THREADS) keys[i] = ~0ull;
  __syncthreads();
  // 1. compaction (order irrelevant: sorted next)
  for (int b = tid; b < nbox; b += NMS_THREADS) {
    float s = sc[(size_t)b * C];
    if (s > score_thresh) {
      int pos = atomicAdd(&s_n, 1);
      keys[pos] = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned)b;
    }
  }
  __syncthreads();
  const int n = s_n;
  if (n == 0) {
    if (tid == 0) keep_count[unit] = 0;
    return;
  }
  // 2. bitonic sort over the smallest power of two >= n
  int P2 = 1;
  while (P2 < n) P2 <<= 1;
  for (int size = 2; size <= P2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (P2 >> 1); i += NMS_THREADS) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool asc = (lo & size) == 0;
        unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  Corner* corners = reinterpret_cast<Corner*>(smem_raw + (size_t)P * 8);
  for (int i = tid; i < n; i += NMS_THREADS) corners[i] = to_corner(bx[(unsigned)(keys[i] & 0xffffffffu)]);
  __syncthreads();
  int32_t* out = keep_idx + (size_t)unit * max_keep;
  float* outs = keep_score ? keep_score + (size_t)unit * max_keep : nullptr;

  const size_t mat_off = (size_t)P * 8 + (((size_t)n * sizeof(Corner) + 15) & ~(size_t)15);
  if (n <= NMS_MAX_MATRIX && mat_off + (size_t)n * ((n + 31) >> 5) * 4 <= smem_bytes) {
    // 3. suppression matrix, upper triangle only: row i, word w covers j = 32w..32w+31, j > i
    const int words = (n + 31) >> 5;
    unsigned* mat = reinterpret_cast<unsigned*>(smem_raw + (size_t)P * 8 + (((size_t)n * sizeof(Corner) + 15) & ~(size_t)15));
    for (int it = tid; it < n * words; it += NMS_THREADS) {
      int i = it / words, w = it - i * words;
      unsigned bits = 0u;
      if (32 * w + 31 > i) {
        Corner ci = corners[i];
        int j0 = 32 * w;
#pragma unroll 4
        for (int t = 0; t < 32; ++t) {
          int j = j0 + t;
          if (j > i && j < n && iou_corner(ci, corners[j]) > iou_thresh) bits |= 1u << t;
        }
      }
      mat[it] = bits;
    }
    __syncthreads();
    // 4. warp-ballot sweep
    if (tid < 32) {
      const int lane = tid;
      unsigned valid = 0u;           // bits of candidates that exist in my word
      if (lane < words) {
        int rem = n - 32 * lane;
        valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      }
      unsigned removed = 0u;
      int count = 0;
      while (true) {
        unsigned alive = valid & ~removed;
        unsigned ball = __ballot_sync(0xffffffffu, alive != 0u);
        if (ball == 0u) break;
        int wl = __ffs(ball) - 1;
        unsigned aw = __shfl_sync(0xffffffffu, alive, wl);
        int i = 32 * wl + (__ffs(aw) - 1);
        if (lane == 0 && count < max_keep) {
          out[count] = (int32_t)(keys[i] & 0xffffffffu);
          if (outs) outs[count] = __uint_as_float(~(unsigned)(keys[i] >> 32));
        }
        ++count;
        if (lane < words) removed |= mat[i * words + lane];
        if (lane == wl) removed |= 1u << (i & 31);   // visited
      }
      if (lane == 0) keep_count[unit] = count;
    }
  } else {
    // fallback: removed-bit array in smem, one block-wide pass per kept candidate
    unsigned* removed = reinterpret_cast<unsigned*>(smem_raw + (size_t)P * 8 + (((size_t)n * sizeof(Corner) + 15) & ~(size_t)15));
    const int words = (n + 31) >> 5;
    for (int w = tid; w < words; w += NMS_THREADS) removed[w] = 0u;
    __syncthreads();
    int cur = 0;
    while (true) {
      if (tid == 0) {
        int i = cur;
        while (i < n && ((removed[i >> 5] >> (i & 31)) & 1u)) ++i;
        s_next = i;
        if (i < n) {
          if (s_count < max_keep) {
            out[s_count] = (int32_t)(keys[i] & 0xffffffffu);
            if (outs) outs[s_count] = __uint_as_float(~(unsigned)(keys[i] >> 32));
          }
          ++s_count;
        }
      }
      __syncthreads();
      int i = s_next;
      if (i >= n) break;
      Corner ci = corners[i];
      for (int j = i + 1 + tid; j < n; j += NMS_THREADS)
        if (iou_corner(ci, corners[j]) > iou_thresh) atomicOr(&removed[j >> 5], 1u << (j & 31));
      cur = i + 1;
      __syncthreads();
    }
    if (tid == 0) keep_count[unit] = s_count;
  }
}

// shared with detect_fused.cu (in-kernel overflow path)
template <int NT>
__device__ void nms_image_fallback(unsigned char* smem, size_t smem_bytes, const float* boxes, const float* scores, int img,
                                   int nbox, int C, float score_thresh, float iou_thresh, int32_t* keep_idx,
                                   int32_t* keep_count, float* keep_score, int max_keep) {
  int P = 1;
  while (P < nbox) P <<= 1;
  for (int k = 0; k < C; ++k)
    nms_body_t<NT>(smem, smem_bytes, boxes, scores, nbox, C, score_thresh, iou_thresh, keep_idx, keep_count, keep_score, max_keep,
                   P, img * C + k);
}


}  // namespace y2
