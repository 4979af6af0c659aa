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
#include "nms_body.cuh"

namespace y2 {

constexpr int NMS_THREADS = 256;
__global__ void __launch_bounds__(NMS_THREADS) nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                          int nbox, int C, float score_thresh, float iou_thresh,
                                                          int32_t* __restrict__ keep_idx, int32_t* __restrict__ keep_count,
                                                          int max_keep, int P) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  nms_body_t<NMS_THREADS>(nms_smem, (size_t)1 << 30, boxes, scores, nbox, C, score_thresh, iou_thresh, keep_idx, keep_count, nullptr,
                          max_keep, P, (int)blockIdx.x);
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
  // (a per-DEVICE function attribute: set on every launch that needs it -- one thread may drive several GPUs)
  if (smem > 48 * 1024)
    Y2_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nms_kernel<<<N * C, NMS_THREADS, smem, (cudaStream_t)stream>>>(boxes, scores, nbox, C, score_thresh, iou_thresh,
                                                                 keep_idx, keep_count, max_keep, P);
  Y2_LAUNCHED();
  return Y2_OK;
}

}  // extern "C"
