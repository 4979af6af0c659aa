// decode.cu -- box decode kernels.
//   y2_decode_ref_v1 : a8, the NumPy half of show_yolo_detection (yolo2_nets/net_utils.py:393-407,418)
//   y2_decode_region : a', YOLOv2 region-layer decode (absent from the reference)
// Both are one-warp-per-cell, HBM-bound: a cell's channels are staged with coalesced (128-bit where
// alignment allows) loads, reduced with warp shuffles, and written back coalesced.
#include "common.cuh"

namespace y2 {

// ---------------------------------------------------------------------------------------------
// REF_V1: channels [0:C] class scores (per cell), [C:C+B] confidences, [C+B:] B x (x,y,sqrt w,sqrt h)
// float32 op order = the NumPy expressions of net_utils.py:403-407 (separate add / div / mul).
// ---------------------------------------------------------------------------------------------
__global__ void decode_ref_v1_kernel(const float* __restrict__ net, int ncell_total, int S, int B, int C, float thresh,
                                     float* __restrict__ boxes, float* __restrict__ conf, uint8_t* __restrict__ keep,
                                     int32_t* __restrict__ cls) {
  const int lane = threadIdx.x & 31;
  const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cell >= ncell_total) return;
  const int ch = C + 5 * B;
  const float* p = net + (size_t)cell * ch;
  const int j = cell % S;             // column -> x offset   (config.py:40-42: off[i,j,b] = j)
  const int i = (cell / S) % S;       // row    -> y offset   (net_utils.py:404-405 transpose)
  // argmax over the class vector, first maximum wins (np.argmax)
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int k = lane; k < C; k += 32) {
    float v = p[k];
    if (v > best || (v == best && k < besti)) { best = v; besti = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if (lane == 0) cls[cell] = besti == 0x7fffffff ? 0 : besti;
  for (int b = lane; b < B; b += 32) {
    float cf = p[C + b];
    const float* bx = p + C + B + 4 * b;
    float fs = (float)S;
    float x = __fdiv_rn(__fadd_rn(bx[0], (float)j), fs);
    float y = __fdiv_rn(__fadd_rn(bx[1], (float)i), fs);
    float w = __fmul_rn(bx[2], bx[2]);
    float h = __fmul_rn(bx[3], bx[3]);
    size_t o = (size_t)cell * B + b;
    reinterpret_cast<float4*>(boxes)[o] = make_float4(x, y, w, h);
    conf[o] = cf;
    keep[o] = cf > thresh ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// REGION_V2: per anchor (tx,ty,tw,th,to,c_0..c_{C-1}).  CELLS_PER_BLOCK warps, one cell each.
// The block's cells are contiguous in memory: staged into smem with 128-bit loads when the
// block's byte range is 16-byte aligned (always true for A*(5+C)=125 and 8 cells per block),
// results staged in smem and written back with 128-bit stores.
// ---------------------------------------------------------------------------------------------
constexpr int RD_CELLS = 8;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

__global__ void __launch_bounds__(RD_CELLS * 32) decode_region_kernel(
    const float* __restrict__ net, const float* __restrict__ anchors, int ncell_total, int S, int A, int C,
    float thresh, float* __restrict__ boxes, float* __restrict__ scores) {
  extern __shared__ __align__(16) float smem[];
  const int per = 5 + C;
  const int ch = A * per;
  float* s_in = smem;                              // RD_CELLS * ch
  float* s_sc = s_in + ((RD_CELLS * ch + 3) & ~3); // RD_CELLS * A * C
  float* s_bx = s_sc + ((RD_CELLS * A * C + 3) & ~3);  // RD_CELLS * A * 4
  const int cell0 = blockIdx.x * RD_CELLS;
  const int ncell = min(RD_CELLS, ncell_total - cell0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- stage in ----
  {
    const float* src = net + (size_t)cell0 * ch;
    int nfl = ncell * ch;
    if ((((uintptr_t)src) & 15) == 0) {
      int nv = nfl >> 2;
      for (int v = tid; v < nv; v += blockDim.x)
        reinterpret_cast<float4*>(s_in)[v] = __ldg(reinterpret_cast<const float4*>(src) + v);
      for (int e = (nv << 2) + tid; e < nfl; e += blockDim.x) s_in[e] = __ldg(src + e);
    } else {
      for (int e = tid; e < nfl; e += blockDim.x) s_in[e] = __ldg(src + e);
    }
  }
  __syncthreads();

  if (warp < ncell) {
    const int cell = cell0 + warp;
    const int j = cell % S, i = (cell / S) % S;
    const float* p = s_in + warp * ch;
    const float fs = (float)S;
    for (int a = 0; a < A; ++a) {
      const float* q = p + a * per;
      // class softmax over lanes (C <= 32 per pass; loop for larger C)
      float mx = -INFINITY;
      for (int k = lane; k < C; k += 32) mx = fmaxf(mx, q[5 + k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.0f;
      for (int k = lane; k < C; k += 32) sum += expf(q[5 + k] - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float obj = sigmoidf_(q[4]);
      for (int k = lane; k < C; k += 32) {
        float sc = obj * (expf(q[5 + k] - mx) / sum);
        s_sc[(warp * A + a) * C + k] = sc > thresh ? sc : 0.0f;
      }
      if (lane == 0) {
        float bx = ((float)j + sigmoidf_(q[0])) / fs;
        float by = ((float)i + sigmoidf_(q[1])) / fs;
        float bw = anchors[2 * a + 0] * expf(q[2]) / fs;
        float bh = anchors[2 * a + 1] * expf(q[3]) / fs;
        float* o = s_bx + (warp * A + a) * 4;
        o[0] = bx; o[1] = by; o[2] = bw; o[3] = bh;
      }
    }
  }
  __syncthreads();

  // ---- stage out ----
  {
    float* dst = scores + (size_t)cell0 * A * C;
    int nfl = ncell * A * C;
    if ((((uintptr_t)dst) & 15) == 0 && (nfl & 3) == 0) {
      for (int v = tid; v < (nfl >> 2); v += blockDim.x)
        reinterpret_cast<float4*>(dst)[v] = reinterpret_cast<const float4*>(s_sc)[v];
    } else {
      for (int e = tid; e < nfl; e += blockDim.x) dst[e] = s_sc[e];
    }
    float4* dbx = reinterpret_cast<float4*>(boxes) + (size_t)cell0 * A;
    for (int v = tid; v < ncell * A; v += blockDim.x) dbx[v] = reinterpret_cast<const float4*>(s_bx)[v];
  }
}

}  // namespace y2

using namespace y2;

extern "C" {

int y2_decode_ref_v1(const float* net, int N, int S, int B, int C, float thresh, float* boxes, float* conf,
                     uint8_t* keep, int32_t* cls, y2_stream_t stream) {
  Y2_ARG(net && boxes && conf && keep && cls && N > 0 && S > 0 && B > 0 && C > 0);
  Y2_ARG((((uintptr_t)boxes) & 15) == 0);
  int ncell = N * S * S;
  decode_ref_v1_kernel<<<ceil_div(ncell, 4), 128, 0, (cudaStream_t)stream>>>(net, ncell, S, B, C, thresh, boxes, conf,
                                                                             keep, cls);
  Y2_LAUNCHED();
  return Y2_OK;
}

int y2_decode_region(const float* net, const float* anchors, int N, int S, int A, int C, float thresh, float* boxes,
                     float* scores, y2_stream_t stream) {
  Y2_ARG(net && anchors && boxes && scores && N > 0 && S > 0 && A > 0 && C > 0);
  Y2_ARG((((uintptr_t)boxes) & 15) == 0);
  int ncell = N * S * S;
  int ch = A * (5 + C);
  size_t smem = (size_t)(((RD_CELLS * ch + 3) & ~3) + ((RD_CELLS * A * C + 3) & ~3) + RD_CELLS * A * 4) * sizeof(float);
  Y2_ARG(smem <= 48 * 1024);
  decode_region_kernel<<<ceil_div(ncell, RD_CELLS), RD_CELLS * 32, smem, (cudaStream_t)stream>>>(
      net, anchors, ncell, S, A, C, thresh, boxes, scores);
  Y2_LAUNCHED();
  return Y2_OK;
}

}  // extern "C"
