// conv_simt.cu -- a1: exact float32 convolution (yolo2_nets/darknet.py:20-21 tf.nn.conv2d SAME,
// stride 1; :35 bias add).  Implicit GEMM on the FFMA pipe with fp32 accumulation:
//   M = N*H*W output pixels, Ncol = Cout, K = k*k*Cin (K index = tap*Cin + c, the HWIO row order).
// This is the spec's "fp32 path" (1e-5 parity) and the on-GPU cross-check of the tcgen05 kernel;
// the throughput path is conv_tcgen05.cu.
#include "common.cuh"

namespace y2 {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, int N,
                                                       int H, int W, int Cin, int Cout, int ksize) {
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const long long M = (long long)N * H * W;
  const int K = ksize * ksize * Cin;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int pad = ksize / 2;

  // A-load assignment: k_local fixed per thread, 4 pixels
  const int kl = t & 15;
  int pn[4], ph[4], pw[4];
  bool pv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + (t >> 4) + 16 * i;
    pv[i] = m < M;
    long long mm = pv[i] ? m : 0;
    pw[i] = (int)(mm % W);
    ph[i] = (int)((mm / W) % H);
    pn[i] = (int)(mm / ((long long)W * H));
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    int kk = k0 + kl;
    int tap = kk / Cin, c = kk - tap * Cin;
    int kh = tap / ksize, kw = tap - kh * ksize;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = 0.0f;
      int ih = ph[i] + kh - pad, iw = pw[i] + kw - pad;
      if (kk < K && pv[i] && ih >= 0 && ih < H && iw >= 0 && iw < W)
        v = x[(((size_t)pn[i] * H + ih) * W + iw) * Cin + c];
      As[kl][(t >> 4) + 16 * i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = t + 256 * i;
      int k = e >> 6, n = e & 63;
      float v = 0.0f;
      if (k0 + k < K && n0 + n < Cout) v = w[(size_t)(k0 + k) * Cout + n0 + n];
      Bs[k][n] = v;
    }
    __syncthreads();
    // two-level accumulation: 16-term partial sums, then one add into the running total (keeps the
    // fp32 rounding error growth ~sqrt(K/16) instead of ~sqrt(K) for K up to 9216)
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.0f;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
      float4 bq = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      b[0] = bq.x; b[1] = bq.y; b[2] = bq.z; b[3] = bq.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < Cout) y[(size_t)m * Cout + n] = acc[i][j] + (bias ? bias[n] : 0.0f);
    }
  }
}

}  // namespace y2

extern "C" int y2_conv_fwd_f32(const float* x, const float* w_hwio, const float* bias, float* y, int N, int H, int W,
                               int Cin, int Cout, int ksize, y2_stream_t stream) {
  Y2_ARG(x && w_hwio && y && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (ksize == 1 || ksize == 3));
  long long M = (long long)N * H * W;
  dim3 grid((unsigned)((M + y2::BM - 1) / y2::BM), (unsigned)((Cout + y2::BN - 1) / y2::BN));
  y2::conv_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w_hwio, bias, y, N, H, W, Cin, Cout, ksize);
  Y2_LAUNCHED();
  return Y2_OK;
}
