// region_loss.cu -- a': YOLOv2 region-layer loss forward + backward in ONE kernel.  Absent from the reference
// (its loss is the YOLOv1-style get_loss, loss.cu); definition = SURVEY.md Appendix A, restated by
// oracle/yolo2_oracle.py::region_loss_torch.  IoU arithmetic is the reference's get_iou (net_utils.py:222-260).
//
//   net      [N,S,S,A*(5+C)]  per anchor (tx, ty, tw, th, to, c_0..c_{C-1})
//   gt_boxes [N,G,4] normalised (cx,cy,w,h), gt_classes [N,G] int32, gt_counts [N] int32
//   per ground truth g: cell (i,j) = floor(gy*S), floor(gx*S); responsible anchor = arg-max IoU of (0,0,pw,ph) vs
//   (0,0,gw,gh), first maximum wins; the first ground truth claims a (cell, anchor) slot.
//   coord  lambda_coord*(2-gw*gh) * [(sig(tx)-tx*)^2 + (sig(ty)-ty*)^2 + (tw-tw*)^2 + (th-th*)^2]
//   obj    lambda_obj * (sig(to) - IoU(pred, gt))^2            (IoU is a constant target)
//   noobj  lambda_noobj * sig(to)^2  for unassigned predictions whose best IoU over the image's GT < ignore_thresh
//   class  lambda_class * sum_k (softmax(c)_k - onehot_k)^2
//   loss = mean over the batch of per-image sums (like net_utils.py:296);  dnet = d loss / d net.
//
// One warp per cell: the cell's A*(5+C) channels are staged in shared memory with coalesced loads, lanes scan the
// image's ground truths (shuffle max / min reductions), then lanes = classes for the softmax terms; the gradient
// overwrites the staged values and is written back coalesced.  Block partial sums in double, folded in a fixed
// order by a 1-block finalize launch (deterministic).
#include "common.cuh"

namespace y2 {

constexpr int RL_CELLS = 8;            // warps (cells) per block
constexpr int RL_MAXA = 16;

__device__ __forceinline__ float rl_sigmoid(float v) { return 1.0f / (1.0f + expf(-v)); }

// net_utils.py:231-260 (get_iou) on (cx,cy,w,h) boxes
__device__ __forceinline__ float rl_iou(float ax, float ay, float aw, float ah, float bx, float by, float bw, float bh) {
  const float x1a = ax - aw / 2.0f, y1a = ay - ah / 2.0f, x2a = ax + aw / 2.0f, y2a = ay + ah / 2.0f;
  const float x1b = bx - bw / 2.0f, y1b = by - bh / 2.0f, x2b = bx + bw / 2.0f, y2b = by + bh / 2.0f;
  const float iw = fmaxf(0.0f, fminf(x2a, x2b) - fmaxf(x1a, x1b));
  const float ih = fmaxf(0.0f, fminf(y2a, y2b) - fmaxf(y1a, y1b));
  const float inter = iw * ih;
  const float uni = fmaxf((x2a - x1a) * (y2a - y1a) + (x2b - x1b) * (y2b - y1b) - inter, 1e-10f);
  return fminf(fmaxf(inter / uni, 0.0f), 1.0f);
}

__global__ void __launch_bounds__(RL_CELLS * 32) region_loss_kernel(
    const float* __restrict__ net, const float* __restrict__ anchors, const float* __restrict__ gt_boxes,
    const int32_t* __restrict__ gt_classes, const int32_t* __restrict__ gt_counts, int ncell_total, int S, int A, int C,
    int G, float lambda_coord, float lambda_obj, float lambda_noobj, float lambda_class, float ignore_thresh, float invN,
    float* __restrict__ dnet, double* __restrict__ partials) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double s_part[RL_CELLS][4];
  const int per = 5 + C, ch = A * per;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cell0 = blockIdx.x * RL_CELLS;
  const int ncell = min(RL_CELLS, ncell_total - cell0);
  {
    const float* src = net + (size_t)cell0 * ch;
    const int nfl = ncell * ch;
    if ((((uintptr_t)src) & 15) == 0) {
      const int nv = nfl >> 2;
      for (int v = tid; v < nv; v += blockDim.x) reinterpret_cast<float4*>(smem)[v] = __ldg(reinterpret_cast<const float4*>(src) + v);
      for (int e = (nv << 2) + tid; e < nfl; e += blockDim.x) smem[e] = __ldg(src + e);
    } else {
      for (int e = tid; e < nfl; e += blockDim.x) smem[e] = __ldg(src + e);
    }
  }
  __syncthreads();
  double t_coord = 0.0, t_obj = 0.0, t_noobj = 0.0, t_cls = 0.0;      // lane 0 of each warp accumulates
  if (warp < ncell) {
    const int cell = cell0 + warp;
    const int j = cell % S, i = (cell / S) % S, n = cell / (S * S);
    float* p = smem + warp * ch;
    const float fs = (float)S;
    const int cnt = min(gt_counts[n], G);
    const float4* gtb = reinterpret_cast<const float4*>(gt_boxes) + (size_t)n * G;
    // ---- which ground truth owns each anchor slot of this cell (first one wins) ----
    int owner[RL_MAXA];
#pragma unroll
    for (int a = 0; a < RL_MAXA; ++a) owner[a] = 0x7fffffff;
    for (int g = lane; g < cnt; g += 32) {
      const float4 b = gtb[g];
      const int gj = min((int)(b.x * fs), S - 1), gi = min((int)(b.y * fs), S - 1);
      if (gi == i && gj == j) {
        int best_a = 0;
        float best = -1.0f;
        for (int a = 0; a < A; ++a) {
          const float pw = anchors[2 * a] / fs, ph = anchors[2 * a + 1] / fs;
          const float inter = fminf(pw, b.z) * fminf(ph, b.w);
          const float v = inter / (pw * ph + b.z * b.w - inter);
          if (v > best) { best = v; best_a = a; }
        }
#pragma unroll
        for (int a = 0; a < RL_MAXA; ++a)
          if (a == best_a) owner[a] = min(owner[a], g);
      }
    }
#pragma unroll
    for (int a = 0; a < RL_MAXA; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) owner[a] = min(owner[a], __shfl_xor_sync(0xffffffffu, owner[a], o));

#pragma unroll 1
    for (int a = 0; a < A; ++a) {
      float* q = p + a * per;
      const float tx = q[0], ty = q[1], tw = q[2], th = q[3], to = q[4];
      const float sx = rl_sigmoid(tx), sy = rl_sigmoid(ty), so = rl_sigmoid(to);
      const float pw = anchors[2 * a], ph = anchors[2 * a + 1];
      const float bx = ((float)j + sx) / fs, by = ((float)i + sy) / fs;
      const float bw = pw * expf(tw) / fs, bh = ph * expf(th) / fs;
      // best IoU of this prediction over the image's ground truths
      float best = 0.0f;
      for (int g = lane; g < cnt; g += 32) {
        const float4 b = gtb[g];
        best = fmaxf(best, rl_iou(bx, by, bw, bh, b.x, b.y, b.z, b.w));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
      int own = 0x7fffffff;
#pragma unroll
      for (int aa = 0; aa < RL_MAXA; ++aa)
        if (aa == a) own = owner[aa];
      // softmax over the classes (lanes), needed only when the slot is assigned
      float g_t[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      if (own != 0x7fffffff) {
        const float4 b = gtb[own];
        const int cls_id = gt_classes[(size_t)n * G + own];
        const float scale = lambda_coord * (2.0f - b.z * b.w);
        const float txt = b.x * fs - (float)j, tyt = b.y * fs - (float)i;
        const float twt = logf(b.z * fs / pw), tht = logf(b.w * fs / ph);
        const float iou = rl_iou(bx, by, bw, bh, b.x, b.y, b.z, b.w);
        const float dx = sx - txt, dy = sy - tyt, dw = tw - twt, dh = th - tht, dob = so - iou;
        if (lane == 0) {
          t_coord += (double)(scale * (dx * dx + dy * dy + dw * dw + dh * dh));
          t_obj += (double)(lambda_obj * dob * dob);
        }
        g_t[0] = scale * 2.0f * dx * sx * (1.0f - sx);
        g_t[1] = scale * 2.0f * dy * sy * (1.0f - sy);
        g_t[2] = scale * 2.0f * dw;
        g_t[3] = scale * 2.0f * dh;
        g_t[4] = lambda_obj * 2.0f * dob * so * (1.0f - so);
        // class term: L = lambda * sum_k (p_k - y_k)^2 ; dL/dc_m = 2*lambda*p_m*[(p_m - y_m) - sum_k (p_k - y_k) p_k]
        float mx = -INFINITY;
        for (int k = lane; k < C; k += 32) mx = fmaxf(mx, q[5 + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.0f;
        for (int k = lane; k < C; k += 32) sum += expf(q[5 + k] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        float sq = 0.0f, dot = 0.0f;
        for (int k = lane; k < C; k += 32) {
          const float pk = expf(q[5 + k] - mx) / sum;
          const float d = pk - (k == cls_id ? 1.0f : 0.0f);
          sq += d * d;
          dot += d * pk;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sq += __shfl_xor_sync(0xffffffffu, sq, o);
          dot += __shfl_xor_sync(0xffffffffu, dot, o);
        }
        if (lane == 0) t_cls += (double)(lambda_class * sq);
        __syncwarp();
        for (int k = lane; k < C; k += 32) {
          const float pk = expf(q[5 + k] - mx) / sum;
          const float d = pk - (k == cls_id ? 1.0f : 0.0f);
          q[5 + k] = 2.0f * lambda_class * pk * (d - dot) * invN;
        }
      } else {
        if (best < ignore_thresh) {
          if (lane == 0) t_noobj += (double)(lambda_noobj * so * so);
          g_t[4] = lambda_noobj * 2.0f * so * so * (1.0f - so);
        }
        for (int k = lane; k < C; k += 32) q[5 + k] = 0.0f;
      }
      __syncwarp();
      if (lane < 5) q[lane] = g_t[lane] * invN;
      __syncwarp();
    }
  }
  if (lane == 0) {
    s_part[warp][0] = t_coord; s_part[warp][1] = t_obj; s_part[warp][2] = t_noobj; s_part[warp][3] = t_cls;
  }
  __syncthreads();
  if (tid < 4) {
    double acc = 0.0;
    for (int w = 0; w < RL_CELLS; ++w) acc += s_part[w][tid];
    partials[(size_t)blockIdx.x * 4 + tid] = acc;
  }
  if (dnet) {
    float* dst = dnet + (size_t)cell0 * ch;
    const int nfl = ncell * ch;
    if ((((uintptr_t)dst) & 15) == 0 && (nfl & 3) == 0) {
      for (int v = tid; v < (nfl >> 2); v += blockDim.x) reinterpret_cast<float4*>(dst)[v] = reinterpret_cast<const float4*>(smem)[v];
    } else {
      for (int e = tid; e < nfl; e += blockDim.x) dst[e] = smem[e];
    }
  }
}

__global__ void region_loss_finalize_kernel(const double* __restrict__ partials, int nblocks, float invN,
                                            float* __restrict__ terms) {
  const int q = threadIdx.x;
  __shared__ double s[4];
  if (q < 4) {
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += partials[(size_t)b * 4 + q];
    acc *= (double)invN;
    s[q] = acc;
    terms[q] = (float)acc;
  }
  __syncthreads();
  if (q == 0) terms[4] = (float)(s[0] + s[1] + s[2] + s[3]);
}

}  // namespace y2

using namespace y2;

extern "C" {

size_t y2_region_loss_workspace_bytes(int N, int S) {
  return (size_t)ceil_div(N * S * S, RL_CELLS) * 4 * sizeof(double);
}

int y2_region_loss_fwd_bwd(const float* net, const float* anchors, const float* gt_boxes, const int32_t* gt_classes,
                           const int32_t* gt_counts, int N, int S, int A, int C, int G, float lambda_coord,
                           float lambda_obj, float lambda_noobj, float lambda_class, float ignore_thresh, float* terms,
                           float* dnet, void* workspace, size_t workspace_bytes, y2_stream_t stream) {
  Y2_ARG(net && anchors && gt_boxes && gt_classes && gt_counts && terms);
  Y2_ARG(N > 0 && S > 0 && A > 0 && A <= RL_MAXA && C > 0 && G > 0);
  Y2_ARG((((uintptr_t)gt_boxes) & 15) == 0);
  if (!workspace || workspace_bytes < y2_region_loss_workspace_bytes(N, S)) {
    set_error("y2_region_loss_fwd_bwd: workspace too small");
    return Y2_ERR_WORKSPACE;
  }
  Y2_ARG((((uintptr_t)workspace) & 7) == 0);
  const size_t smem = (size_t)RL_CELLS * A * (5 + C) * sizeof(float);
  Y2_ARG(smem <= 48 * 1024);
  cudaStream_t st = (cudaStream_t)stream;
  const int ncell = N * S * S;
  const int nblocks = ceil_div(ncell, RL_CELLS);
  region_loss_kernel<<<nblocks, RL_CELLS * 32, smem, st>>>(net, anchors, gt_boxes, gt_classes, gt_counts, ncell, S, A, C, G,
                                                           lambda_coord, lambda_obj, lambda_noobj, lambda_class,
                                                           ignore_thresh, 1.0f / (float)N, dnet, (double*)workspace);
  Y2_LAUNCHED();
  region_loss_finalize_kernel<<<1, 32, 0, st>>>((const double*)workspace, nblocks, 1.0f / (float)N, terms);
  Y2_LAUNCHED();
  return Y2_OK;
}

}  // extern "C"
