// loss.cu -- a6 + a7: the YOLO (v1-style) sum-squared loss of yolo2_nets/net_utils.py:263-372 and
// its gradient (what TF autodiff produces for that graph), fused in ONE kernel pass over the grid:
// one thread per cell, cells staged through shared memory with coalesced loads/stores, the
// gradient written in place of the staged predictions.  Loss terms: block partials + a fixed-order
// fold (deterministic).
//   net    [N,S,S,C+5B] : [0:C] class (per cell) | [C:C+B] confidence | B x (x, y, sqrt w, sqrt h)
//   labels [N,S,S,5+C]  : [0] responsible | [1:5] GT (cx,cy,w,h) pixels of the resized image | one-hot
// Tie rules of the backward pass follow TF (maximum: first arg wins on >=, minimum: on <=).
#include "common.cuh"

namespace y2 {

constexpr int LOSS_THREADS = 128;
constexpr int LOSS_MAXB = 8;

struct IouOut { float iou, gx, gy, gw, gh; };   // iou and d iou / d (cx, cy, w, h) of box 1

// net_utils.py:231-260 forward, then the reverse sweep.
__device__ __forceinline__ IouOut iou_fwd_bwd(float cx, float cy, float w, float h, float gcx, float gcy, float gw_,
                                              float gh_) {
  float x1a = cx - w / 2.0f, y1a = cy - h / 2.0f, x2a = cx + w / 2.0f, y2a = cy + h / 2.0f;
  float x1b = gcx - gw_ / 2.0f, y1b = gcy - gh_ / 2.0f, x2b = gcx + gw_ / 2.0f, y2b = gcy + gh_ / 2.0f;
  float lux = fmaxf(x1a, x1b), luy = fmaxf(y1a, y1b);
  float rdx = fminf(x2a, x2b), rdy = fminf(y2a, y2b);
  float dw = rdx - lux, dh = rdy - luy;
  float iw = fmaxf(0.0f, dw), ih = fmaxf(0.0f, dh);
  float inter = iw * ih;
  float sq1 = (x2a - x1a) * (y2a - y1a);
  float sq2 = (x2b - x1b) * (y2b - y1b);
  float uraw = sq1 + sq2 - inter;
  float uni = fmaxf(uraw, 1e-10f);
  float q = inter / uni;
  IouOut o;
  o.iou = fminf(fmaxf(q, 0.0f), 1.0f);
  // reverse, seeded with d iou = 1
  float g_q = (q >= 0.0f && q <= 1.0f) ? 1.0f : 0.0f;
  float g_inter = g_q / uni;
  float g_uni = -g_q * inter / (uni * uni);
  float g_uraw = (uraw >= 1e-10f) ? g_uni : 0.0f;
  float g_sq1 = g_uraw;
  g_inter -= g_uraw;
  float g_dw = (dw > 0.0f) ? g_inter * ih : 0.0f;
  float g_dh = (dh > 0.0f) ? g_inter * iw : 0.0f;
  float g_x1a = (x1a >= x1b) ? -g_dw : 0.0f;
  float g_y1a = (y1a >= y1b) ? -g_dh : 0.0f;
  float g_x2a = (x2a <= x2b) ? g_dw : 0.0f;
  float g_y2a = (y2a <= y2b) ? g_dh : 0.0f;
  float hh = y2a - y1a, ww = x2a - x1a;
  g_x2a += g_sq1 * hh;
  g_x1a -= g_sq1 * hh;
  g_y2a += g_sq1 * ww;
  g_y1a -= g_sq1 * ww;
  o.gx = g_x1a + g_x2a;
  o.gy = g_y1a + g_y2a;
  o.gw = (g_x2a - g_x1a) * 0.5f;
  o.gh = (g_y2a - g_y1a) * 0.5f;
  return o;
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_v1_kernel(
    const float* __restrict__ net, const float* __restrict__ labels, int ncell_total, int N, int S, int B, int C,
    float image_size, float lambda_coord, float lambda_noobj, float* __restrict__ ious_out,
    float* __restrict__ mask_out, float* __restrict__ dnet, double* __restrict__ partials) {
  extern __shared__ __align__(16) float smem[];
  const int ch = C + 5 * B, lab = 5 + C;
  float* s_net = smem;
  float* s_lab = smem + LOSS_THREADS * ch;
  __shared__ float s_red[4][LOSS_THREADS / 32];
  const int tid = threadIdx.x;
  const int cell0 = blockIdx.x * LOSS_THREADS;
  const int ncell = min(LOSS_THREADS, ncell_total - cell0);
  for (int e = tid; e < ncell * ch; e += LOSS_THREADS) s_net[e] = net[(size_t)cell0 * ch + e];
  for (int e = tid; e < ncell * lab; e += LOSS_THREADS) s_lab[e] = labels[(size_t)cell0 * lab + e];
  __syncthreads();

  float l_class = 0.0f, l_coord = 0.0f, l_obj = 0.0f, l_noobj = 0.0f;
  if (tid < ncell) {
    const int cell = cell0 + tid;
    const int j = cell % S, i = (cell / S) % S;
    float* p = s_net + tid * ch;
    const float* lb = s_lab + tid * lab;
    const float invN = 1.0f / (float)N;
    const float resp = lb[0];
    // class term (:290-297)
    for (int k = 0; k < C; ++k) {
      float d = resp * (p[k] - lb[5 + k]);
      l_class += d * d;
      p[k] = 2.0f * resp * d * invN;
    }
    // boxes (:302-334)
    const float fS = (float)S;
    const float gcx = lb[1] / image_size, gcy = lb[2] / image_size, gw = lb[3] / image_size, gh = lb[4] / image_size;
    float conf[LOSS_MAXB], bx[LOSS_MAXB], by[LOSS_MAXB], bw[LOSS_MAXB], bh[LOSS_MAXB];
    IouOut io[LOSS_MAXB];
    float mx = -INFINITY;
#pragma unroll
    for (int b = 0; b < LOSS_MAXB; ++b) {
      if (b < B) {
        conf[b] = p[C + b];
        bx[b] = p[C + B + 4 * b + 0];
        by[b] = p[C + B + 4 * b + 1];
        bw[b] = p[C + B + 4 * b + 2];
        bh[b] = p[C + B + 4 * b + 3];
        float px = (bx[b] + (float)j) / fS, py = (by[b] + (float)i) / fS;
        io[b] = iou_fwd_bwd(px, py, bw[b] * bw[b], bh[b] * bh[b], gcx, gcy, gw, gh);
        mx = fmaxf(mx, io[b].iou);
      }
    }
    const float grx = gcx * fS - (float)j, gry = gcy * fS - (float)i;
    const float grw = sqrtf(gw), grh = sqrtf(gh);
#pragma unroll
    for (int b = 0; b < LOSS_MAXB; ++b) {
      if (b < B) {
        float m = (io[b].iou >= mx ? 1.0f : 0.0f) * resp;     // :323-324 (ties mark every predictor)
        float nm = 1.0f - m;                                   // :325-326
        float dx = bx[b] - grx, dy = by[b] - gry, dw = bw[b] - grw, dh = bh[b] - grh;
        float m2 = m * m;
        l_coord += m2 * (dx * dx + dy * dy + dw * dw + dh * dh);
        float od = m * (conf[b] - io[b].iou);
        l_obj += od * od;
        float nd = nm * conf[b];
        l_noobj += nd * nd;
        // gradients
        float g_conf = (2.0f * m2 * (conf[b] - io[b].iou) + 2.0f * lambda_noobj * nm * nm * conf[b]) * invN;
        float g_iou = -2.0f * m2 * (conf[b] - io[b].iou) * invN;
        float cs = 2.0f * lambda_coord * invN * m2;
        p[C + b] = g_conf;
        p[C + B + 4 * b + 0] = cs * dx + g_iou * io[b].gx / fS;
        p[C + B + 4 * b + 1] = cs * dy + g_iou * io[b].gy / fS;
        p[C + B + 4 * b + 2] = cs * dw + g_iou * io[b].gw * 2.0f * bw[b];
        p[C + B + 4 * b + 3] = cs * dh + g_iou * io[b].gh * 2.0f * bh[b];
        if (ious_out) ious_out[(size_t)cell * B + b] = io[b].iou;
        if (mask_out) mask_out[(size_t)cell * B + b] = m;
      }
    }
  }
  // block reduction of the four terms
  float v[4] = {l_class, l_coord, l_obj, l_noobj};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if ((tid & 31) == 0) s_red[q][tid >> 5] = v[q];
  }
  __syncthreads();
  if (tid < 4) {
    double acc = 0.0;
    for (int wv = 0; wv < LOSS_THREADS / 32; ++wv) acc += (double)s_red[tid][wv];
    partials[(size_t)blockIdx.x * 4 + tid] = acc;
  }
  if (dnet)
    for (int e = tid; e < ncell * ch; e += LOSS_THREADS) dnet[(size_t)cell0 * ch + e] = s_net[e];
}

// a7 stand-alone: yolo2_nets/net_utils.py:222-260 get_iou on n box pairs (cx,cy,w,h), float32 with the
// NumPy op order (explicitly rounded ops, no FMA contraction) -> bit-identical to the oracle.
__global__ void iou_pairs_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, float* __restrict__ out,
                                 size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = b1[i], b = b2[i];
  float x1a = __fsub_rn(a.x, __fdiv_rn(a.z, 2.0f)), y1a = __fsub_rn(a.y, __fdiv_rn(a.w, 2.0f));
  float x2a = __fadd_rn(a.x, __fdiv_rn(a.z, 2.0f)), y2a = __fadd_rn(a.y, __fdiv_rn(a.w, 2.0f));
  float x1b = __fsub_rn(b.x, __fdiv_rn(b.z, 2.0f)), y1b = __fsub_rn(b.y, __fdiv_rn(b.w, 2.0f));
  float x2b = __fadd_rn(b.x, __fdiv_rn(b.z, 2.0f)), y2b = __fadd_rn(b.y, __fdiv_rn(b.w, 2.0f));
  float iw = fmaxf(0.0f, __fsub_rn(fminf(x2a, x2b), fmaxf(x1a, x1b)));
  float ih = fmaxf(0.0f, __fsub_rn(fminf(y2a, y2b), fmaxf(y1a, y1b)));
  float inter = __fmul_rn(iw, ih);
  float sq1 = __fmul_rn(__fsub_rn(x2a, x1a), __fsub_rn(y2a, y1a));
  float sq2 = __fmul_rn(__fsub_rn(x2b, x1b), __fsub_rn(y2b, y1b));
  float uni = fmaxf(__fsub_rn(__fadd_rn(sq1, sq2), inter), 1e-10f);
  out[i] = fminf(fmaxf(__fdiv_rn(inter, uni), 0.0f), 1.0f);
}

__global__ void loss_finalize_kernel(const double* __restrict__ partials, int nblocks, int N, float lambda_coord,
                                     float lambda_noobj, float* __restrict__ terms) {
  int q = threadIdx.x;
  __shared__ double s[4];
  if (q < 4) {
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += partials[(size_t)b * 4 + q];
    acc /= (double)N;
    if (q == 1) acc *= (double)lambda_coord;
    if (q == 3) acc *= (double)lambda_noobj;
    s[q] = acc;
    terms[q] = (float)acc;
  }
  __syncthreads();
  if (q == 0) terms[4] = (float)(s[0] + s[2] + s[3] + s[1]);   // :372 class + object + noobject + coord
}

// net_utils.py:337-342: the UNMASKED box deltas the reference logs as histograms (tf.summary.histogram('boxes_delta_x' ..
// 'boxes_delta_h'), :366-369): predict (x, y, sqrt w, sqrt h) minus the cell-relative ground truth (:330-334), for every
// cell and predictor.  deltas [N,S,S,B,4].
__global__ void loss_v1_box_deltas_kernel(const float* __restrict__ net, const float* __restrict__ labels, int total, int S, int B,
                                          int C, float image_size, float* __restrict__ deltas) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int b = t % B, cell = t / B;
  const int j = cell % S, i = (cell / S) % S;
  const float* p = net + (size_t)cell * (C + 5 * B) + C + B + 4 * b;
  const float* lb = labels + (size_t)cell * (5 + C);
  const float fS = (float)S;
  const float gcx = lb[1] / image_size, gcy = lb[2] / image_size, gw = lb[3] / image_size, gh = lb[4] / image_size;
  float4 d;
  d.x = p[0] - (gcx * fS - (float)j);
  d.y = p[1] - (gcy * fS - (float)i);
  d.z = p[2] - sqrtf(gw);
  d.w = p[3] - sqrtf(gh);
  reinterpret_cast<float4*>(deltas)[t] = d;
}

}  // namespace y2

using namespace y2;

// ----------------------------------------------------------------------------------------------------------------------
// ImageNet classifier loss (imagenet_train_darknet.py:50-54,60-61): logits = 7x7 average pool of the last layer
// (darknet.py:116-117), tf.nn.sparse_softmax_cross_entropy_with_logits, reduce_mean, accuracy = mean(argmax == label), and the
// gradient of the mean loss w.r.t. the PRE-pool map (softmax - onehot) / (N * HW) broadcast over the window -- one CTA per image.
// ----------------------------------------------------------------------------------------------------------------------
constexpr int XENT_THREADS = 256;

__device__ __forceinline__ float xent_block_reduce(float v, float* red, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();                                   // red[] may still be read by the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < XENT_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

__global__ void __launch_bounds__(XENT_THREADS) softmax_xent_kernel(const float* __restrict__ net, const int* __restrict__ labels,
                                                                    int N, int HW, int C, float* __restrict__ logits_out,
                                                                    float* __restrict__ losses, float* __restrict__ correct,
                                                                    float* __restrict__ dnet) {
  extern __shared__ float xs[];                      // C logits
  __shared__ float red[XENT_THREADS / 32];
  __shared__ int amax_s;
  const int n = blockIdx.x;
  const float* src = net + (size_t)n * HW * C;
  const float inv_hw = 1.0f / (float)HW;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += XENT_THREADS) {
    float acc = 0.0f;
    for (int p = 0; p < HW; ++p) acc += src[(size_t)p * C + c];
    acc *= inv_hw;
    xs[c] = acc;
    if (logits_out) logits_out[(size_t)n * C + c] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = xent_block_reduce(mx, red, true);
  if (threadIdx.x == 0) amax_s = C;
  __syncthreads();
  float se = 0.0f;
  for (int c = threadIdx.x; c < C; c += XENT_THREADS) {
    se += expf(xs[c] - mx);
    if (xs[c] == mx) atomicMin(&amax_s, c);          // tf.argmax: the first maximum
  }
  se = xent_block_reduce(se, red, false);
  const int lab = labels[n];
  const bool lab_ok = lab >= 0 && lab < C;           // TF raises on an out-of-range label; here it contributes NaN like TF-GPU
  if (threadIdx.x == 0) {
    losses[n] = lab_ok ? (logf(se) + mx - xs[lab]) : NAN;
    correct[n] = (amax_s == lab) ? 1.0f : 0.0f;
  }
  if (dnet) {
    const float gs = inv_hw / (float)N, inv_se = 1.0f / se;
    float* dst = dnet + (size_t)n * HW * C;
    for (int c = threadIdx.x; c < C; c += XENT_THREADS) {
      const float g = (expf(xs[c] - mx) * inv_se - (c == lab ? 1.0f : 0.0f)) * gs;
      for (int p = 0; p < HW; ++p) dst[(size_t)p * C + c] = g;
    }
  }
}

__global__ void softmax_xent_finalize_kernel(const float* __restrict__ losses, const float* __restrict__ correct, int N,
                                             float* __restrict__ terms) {
  double l = 0.0, a = 0.0;
  for (int i = threadIdx.x; i < N; i += 32) { l += losses[i]; a += correct[i]; }
  for (int o = 16; o > 0; o >>= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, o);
    a += __shfl_xor_sync(0xffffffffu, a, o);
  }
  if (threadIdx.x == 0) { terms[0] = (float)(l / N); terms[1] = (float)(a / N); }
}

extern "C" {

int y2_iou(const float* boxes1, const float* boxes2, float* iou, size_t n, y2_stream_t stream) {
  Y2_ARG(boxes1 && boxes2 && iou && n > 0);
  Y2_ARG((((uintptr_t)boxes1) & 15) == 0 && (((uintptr_t)boxes2) & 15) == 0);
  iou_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)boxes1, (const float4*)boxes2, iou, n);
  Y2_LAUNCHED();
  return Y2_OK;
}

size_t y2_loss_v1_workspace_bytes(int N, int S) {
  int ncell = N * S * S;
  return (size_t)ceil_div(ncell, LOSS_THREADS) * 4 * sizeof(double);
}

int y2_loss_v1_fwd_bwd(const float* net, const float* labels, int N, int S, int B, int C, float image_size,
                       float lambda_coord, float lambda_noobj, float* terms, float* ious, float* object_mask,
                       float* dnet, void* workspace, size_t workspace_bytes, y2_stream_t stream) {
  Y2_ARG(net && labels && terms && N > 0 && S > 0 && B > 0 && B <= LOSS_MAXB && C > 0 && image_size > 0.0f);
  if (!workspace || workspace_bytes < y2_loss_v1_workspace_bytes(N, S)) {
    set_error("y2_loss_v1_fwd_bwd: workspace too small");
    return Y2_ERR_WORKSPACE;
  }
  Y2_ARG((((uintptr_t)workspace) & 7) == 0);
  size_t smem = (size_t)LOSS_THREADS * (C + 5 * B + 5 + C) * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("y2_loss_v1_fwd_bwd: C=%d B=%d needs %zu B smem", C, B, smem);
    return Y2_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int ncell = N * S * S;
  int nblocks = ceil_div(ncell, LOSS_THREADS);
  loss_v1_kernel<<<nblocks, LOSS_THREADS, smem, st>>>(net, labels, ncell, N, S, B, C, image_size, lambda_coord,
                                                      lambda_noobj, ious, object_mask, dnet, (double*)workspace);
  Y2_LAUNCHED();
  loss_finalize_kernel<<<1, 32, 0, st>>>((const double*)workspace, nblocks, N, lambda_coord, lambda_noobj, terms);
  Y2_LAUNCHED();
  return Y2_OK;
}

int y2_loss_v1_box_deltas(const float* net, const float* labels, int N, int S, int B, int C, float image_size, float* deltas,
                          y2_stream_t stream) {
  Y2_ARG(net && labels && deltas && N > 0 && S > 0 && B > 0 && C > 0 && image_size > 0.0f);
  Y2_ARG((reinterpret_cast<uintptr_t>(deltas) & 15) == 0);
  const long long total = (long long)N * S * S * B;
  Y2_ARG(total < (1ll << 31));
  loss_v1_box_deltas_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(net, labels, (int)total, S, B, C,
                                                                                              image_size, deltas);
  Y2_LAUNCHED();
  return Y2_OK;
}

int y2_softmax_xent_fwd_bwd(const float* net, const int* labels, int N, int HW, int C, float* logits, float* losses, float* correct,
                            float* terms, float* dnet, y2_stream_t stream) {
  Y2_ARG(net && labels && losses && correct && terms && N > 0 && HW > 0 && C > 0);
  const size_t smem = (size_t)C * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("y2_softmax_xent_fwd_bwd: C=%d needs %zu B smem", C, smem);
    return Y2_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  softmax_xent_kernel<<<N, XENT_THREADS, smem, st>>>(net, labels, N, HW, C, logits, losses, correct, dnet);
  Y2_LAUNCHED();
  softmax_xent_finalize_kernel<<<1, 32, 0, st>>>(losses, correct, N, terms);
  Y2_LAUNCHED();
  return Y2_OK;
}
}  // extern "C"
