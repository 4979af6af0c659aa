// nms_common.cuh -- box arithmetic shared by nms.cu and detect_fused.cu.  IoU follows the op order of the reference's
// get_iou (yolo2_nets/net_utils.py:231-260) with explicitly rounded float32 ops (no FMA contraction), so that
// suppression decisions are bit-identical to the NumPy oracle.
#pragma once
#include "common.cuh"

namespace y2 {

struct Corner { float x1, y1, x2, y2, area; };

__device__ __forceinline__ Corner to_corner(float4 b) {
  Corner c;
  float hw = __fdiv_rn(b.z, 2.0f), hh = __fdiv_rn(b.w, 2.0f);
  c.x1 = __fsub_rn(b.x, hw);
  c.y1 = __fsub_rn(b.y, hh);
  c.x2 = __fadd_rn(b.x, hw);
  c.y2 = __fadd_rn(b.y, hh);
  c.area = __fmul_rn(__fsub_rn(c.x2, c.x1), __fsub_rn(c.y2, c.y1));
  return c;
}

__device__ __forceinline__ float iou_corner(const Corner& a, const Corner& b) {
  float lux = fmaxf(a.x1, b.x1), luy = fmaxf(a.y1, b.y1);
  float rdx = fminf(a.x2, b.x2), rdy = fminf(a.y2, b.y2);
  float iw = fmaxf(0.0f, __fsub_rn(rdx, lux)), ih = fmaxf(0.0f, __fsub_rn(rdy, luy));
  float inter = __fmul_rn(iw, ih);
  float uni = fmaxf(__fsub_rn(__fadd_rn(a.area, b.area), inter), 1e-10f);
  float q = __fdiv_rn(inter, uni);
  return fminf(fmaxf(q, 0.0f), 1.0f);
}

}  // namespace y2
