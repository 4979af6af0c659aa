// nms.cu -- a': per-class greedy NMS (absent from the reference; semantics in SURVEY Appendix A and
// oracle/yolo2_oracle.py::nms_per_class).  One CTA per (image, class):
//   1. compact the candidates (score > score_thresh) of this class into shared memory,
//   2. bitonic-sort 64-bit keys (~score_bits << 32 | box_index): score descending, index ascending,
//   3. build the upper-triangular suppression bit matrix (IoU > thr) in shared memory,
//   4. one warp sweeps it: each lane owns one 32-bit word of the "removed" set, the next surviving
//      candidate is found with a ballot + ffs, its matrix row is OR-ed into the set.
// IoU uses the op order of yolo2_nets/net_utils.py:231-260 with explicitly rounded float32 ops
// (no FMA contraction), so keep lists are bit-identical to the NumPy oracle.
// Candidates beyond the bit-matrix capacity (n > NMS_MAX_MATRIX) fall back to an on-the-fly sweep.
#include "nms_common.cuh"

namespace y2 {

constexpr int NMS_THREADS = 256;
constexpr int NMS_MAX_MATRIX = 1024;   // 32 lanes x 32 bits

// smem layout (dynamic): keys[P] u64 | corners[n] (5 floats) | matrix[n * words] u32
__device__ __forceinline__ void nms_body(const float* __restrict__ boxes, const float* __restrict__ scores, int nbox, int C,
                                         float score_thresh, float iou_thresh, int32_t* __restrict__ keep_idx,
                                         int32_t* __restrict__ keep_count, float* __restrict__ keep_score, int max_keep,
                                         int P, int unit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
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

  if (n <= NMS_MAX_MATRIX) {
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

__global__ void __launch_bounds__(NMS_THREADS) nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                          int nbox, int C, float score_thresh, float iou_thresh,
                                                          int32_t* __restrict__ keep_idx, int32_t* __restrict__ keep_count,
                                                          int max_keep, int P) {
  nms_body(boxes, scores, nbox, C, score_thresh, iou_thresh, keep_idx, keep_count, nullptr, max_keep, P, (int)blockIdx.x);
}

// y2_detect_fused hand-over: persistent CTAs walk the images and redo only those flagged keep_count[img][0] < 0
// (more candidates than the fused kernel's list holds).  Normally nothing is flagged and every CTA exits at once.
__global__ void __launch_bounds__(NMS_THREADS) nms_flagged_kernel(const float* __restrict__ boxes,
                                                                  const float* __restrict__ scores, int N, int nbox, int C,
                                                                  float score_thresh, float iou_thresh,
                                                                  int32_t* __restrict__ keep_idx,
                                                                  int32_t* __restrict__ keep_count,
                                                                  float* __restrict__ keep_score, int max_keep, int P) {
  for (int img = blockIdx.x; img < N; img += gridDim.x) {
    if (keep_count[(size_t)img * C] >= 0) continue;          // uniform across the CTA; written by an earlier kernel
    __syncthreads();                                         // everyone has seen the flag before class 0 overwrites it
    for (int k = 0; k < C; ++k)
      nms_body(boxes, scores, nbox, C, score_thresh, iou_thresh, keep_idx, keep_count, keep_score, max_keep, P, img * C + k);
  }
}

static size_t nms_smem_bytes(int nbox, int* P_out) {
  int P = 1;
  while (P < nbox) P <<= 1;
  *P_out = P;
  size_t corners = ((size_t)nbox * sizeof(Corner) + 15) & ~(size_t)15;
  int nm = nbox < NMS_MAX_MATRIX ? nbox : NMS_MAX_MATRIX;
  size_t mat = (size_t)nm * ((nm + 31) / 32) * 4;
  size_t rem = (size_t)((nbox + 31) / 32) * 4;
  return (size_t)P * 8 + corners + (mat > rem ? mat : rem);
}

int launch_nms_flagged(const float* boxes, const float* scores, int N, int nbox, int C, float score_thresh, float iou_thresh,
                       int32_t* keep_idx, int32_t* keep_count, float* keep_score, int max_keep, cudaStream_t st) {
  int P;
  size_t smem = nms_smem_bytes(nbox, &P);
  if (smem > 200 * 1024) {
    set_error("y2_detect_fused: nbox=%d needs %zu B of shared memory in the overflow path", nbox, smem);
    return Y2_ERR_UNSUPPORTED;
  }
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    Y2_CUDA(cudaFuncSetAttribute(nms_flagged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = N < 148 ? N : 148;
  nms_flagged_kernel<<<grid, NMS_THREADS, smem, st>>>(boxes, scores, N, nbox, C, score_thresh, iou_thresh, keep_idx,
                                                      keep_count, keep_score, max_keep, P);
  Y2_LAUNCHED();
  return Y2_OK;
}

}  // namespace y2

using namespace y2;

extern "C" {

size_t y2_nms_workspace_bytes(int N, int nbox, int C) {
  (void)N; (void)nbox; (void)C;
  return 0;   // everything lives in shared memory
}

int y2_nms(const float* boxes, const float* scores, int N, int nbox, int C, float score_thresh, float iou_thresh,
           int32_t* keep_idx, int32_t* keep_count, int max_keep, void* workspace, size_t workspace_bytes,
           y2_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  Y2_ARG(boxes && scores && keep_idx && keep_count && N > 0 && nbox > 0 && C > 0 && max_keep > 0);
  Y2_ARG((((uintptr_t)boxes) & 15) == 0);
  Y2_ARG(score_thresh >= 0.0f);   // key ordering relies on positive scores
  int P;
  size_t smem = nms_smem_bytes(nbox, &P);
  if (smem > 200 * 1024) {
    set_error("y2_nms: nbox=%d needs %zu B of shared memory (> 200 KB)", nbox, smem);
    return Y2_ERR_UNSUPPORTED;
  }
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    Y2_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  nms_kernel<<<N * C, NMS_THREADS, smem, (cudaStream_t)stream>>>(boxes, scores, nbox, C, score_thresh, iou_thresh,
                                                                 keep_idx, keep_count, max_keep, P);
  Y2_LAUNCHED();
  return Y2_OK;
}

}  // extern "C"
