// boxmath.cuh — the two pairwise box-metric families of the reference, as device functions.
//
//   YOLO family   utils/tf_iou_utils.py:5-65  GetIOU   boxes x1,y1,x2,y2; no area clamp, plain '/',
//                 diou = iou - (u/c)^0.6, ciou = iou - (u/c + alpha*v), alpha = v/(1-iou+v+1e-8), c==0 -> iou
//   EffDet family efficientnet/utils/iou.py:26-100 get_iou  boxes y1,x1,y2,x2; clamped w/h, divide_no_nan,
//                 giou, diou = iou - |dc|^2/|diag|^2 via sqrt-then-square, ciou = diou - alpha*v
//
// Every arithmetic step is a single correctly rounded fp32 op (DM_* from detmath.h) in the reference's
// operation order, so results are bit-identical to the NumPy oracle (oracle/yolo.py, oracle/effdet.py).
#pragma once
#include "detmath.h"
#include "../../include/b200det.h"


// A box with the per-box terms of the metric hoisted out of the pair loop.
// c0..c3 are the four corner coordinates in the family's own order (xyxy for YOLO, yxyx for EffDet).
struct BoxT {
  float c0, c1, c2, c3;
  float area;  // YOLO: (c2-c0)*(c3-c1) unclamped; EffDet: max(0,w)*max(0,h)
  float at;    // YOLO ciou: atan(w/h); EffDet ciou: atan(divide_no_nan(w,h)); else unused
};

#define B200_CIOU_COEF 0.40528470277786255f /* fp32(4) / fp32(fp32(pi)^2), tf_iou_utils.py:55 */
#define B200_PI_F 3.1415927410125732f

__device__ __forceinline__ float bm_dnn(float x, float y) { return (y == 0.0f) ? 0.0f : DM_DIV(x, y); }

__device__ __forceinline__ BoxT bm_prep(float c0, float c1, float c2, float c3, int metric) {
  BoxT b;
  b.c0 = c0; b.c1 = c1; b.c2 = c2; b.c3 = c3;
  b.at = 0.0f;
  if (metric <= B200_METRIC_YOLO_CIOU) {
    float w = DM_SUB(c2, c0), h = DM_SUB(c3, c1);
    b.area = DM_MUL(w, h);
    if (metric == B200_METRIC_YOLO_CIOU) b.at = dm_atanf(DM_DIV(w, h));
  } else {
    float w = dm_max(0.0f, DM_SUB(c3, c1)), h = dm_max(0.0f, DM_SUB(c2, c0));
    b.area = DM_MUL(w, h);
    if (metric == B200_METRIC_EFF_CIOU) b.at = dm_atanf(bm_dnn(w, h));
  }
  return b;
}

// metric(b1, b2) with b1 the reference's first argument (the kept / "top" box in NMS, the
// prediction / anchor in the losses) and b2 the second.
__device__ __forceinline__ float bm_metric(const BoxT& a, const BoxT& b, int metric) {
  if (metric <= B200_METRIC_YOLO_CIOU) {
    // xyxy
    float iw = dm_max(DM_SUB(dm_min(a.c2, b.c2), dm_max(a.c0, b.c0)), 0.0f);
    float ih = dm_max(DM_SUB(dm_min(a.c3, b.c3), dm_max(a.c1, b.c1)), 0.0f);
    float inter = DM_MUL(iw, ih);
    float iou = DM_DIV(inter, DM_SUB(DM_ADD(a.area, b.area), inter));
    if (metric == B200_METRIC_YOLO_IOU) return iou;
    float uw = DM_SUB(dm_max(a.c2, b.c2), dm_min(a.c0, b.c0));
    float uh = DM_SUB(dm_max(a.c3, b.c3), dm_min(a.c1, b.c1));
    float c = DM_ADD(DM_MUL(uw, uw), DM_MUL(uh, uh));
    if (c == 0.0f) return iou;
    float dx = DM_SUB(DM_DIV(DM_ADD(a.c2, a.c0), 2.0f), DM_DIV(DM_ADD(b.c2, b.c0), 2.0f));
    float dy = DM_SUB(DM_DIV(DM_ADD(a.c3, a.c1), 2.0f), DM_DIV(DM_ADD(b.c3, b.c1), 2.0f));
    float u = DM_ADD(DM_MUL(dx, dx), DM_MUL(dy, dy));
    float d = DM_DIV(u, c);
    if (metric == B200_METRIC_YOLO_DIOU) return DM_SUB(iou, dm_powf(d, 0.6f));
    float da = DM_SUB(a.at, b.at);
    float v = DM_MUL(B200_CIOU_COEF, DM_MUL(da, da));
    float alpha = DM_DIV(v, DM_ADD(DM_ADD(DM_SUB(1.0f, iou), v), 1e-8f));
    return DM_SUB(iou, DM_ADD(d, DM_MUL(alpha, v)));
  }
  // yxyx: c0=ymin c1=xmin c2=ymax c3=xmax
  float iw = dm_max(0.0f, DM_SUB(dm_min(a.c3, b.c3), dm_max(a.c1, b.c1)));
  float ih = dm_max(0.0f, DM_SUB(dm_min(a.c2, b.c2), dm_max(a.c0, b.c0)));
  float inter = DM_MUL(iw, ih);
  float uni = DM_SUB(DM_ADD(a.area, b.area), inter);
  float iou = bm_dnn(inter, uni);
  if (metric == B200_METRIC_EFF_IOU) return iou;
  float eymin = dm_min(a.c0, b.c0), exmin = dm_min(a.c1, b.c1);
  float eymax = dm_max(a.c2, b.c2), exmax = dm_max(a.c3, b.c3);
  if (metric == B200_METRIC_EFF_GIOU) {
    float ew = dm_max(0.0f, DM_SUB(exmax, exmin));
    float eh = dm_max(0.0f, DM_SUB(eymax, eymin));
    float ea = DM_MUL(ew, eh);
    return DM_SUB(iou, bm_dnn(DM_SUB(ea, uni), ea));
  }
  float dy = DM_SUB(DM_DIV(DM_ADD(b.c0, b.c2), 2.0f), DM_DIV(DM_ADD(a.c0, a.c2), 2.0f));
  float dx = DM_SUB(DM_DIV(DM_ADD(b.c1, b.c3), 2.0f), DM_DIV(DM_ADD(a.c1, a.c3), 2.0f));
  float eu = DM_SQRT(DM_ADD(DM_MUL(dy, dy), DM_MUL(dx, dx)));
  float ey = DM_SUB(eymax, eymin), ex = DM_SUB(exmax, exmin);
  float dg = DM_SQRT(DM_ADD(DM_MUL(ey, ey), DM_MUL(ex, ex)));
  float diou = DM_SUB(iou, bm_dnn(DM_MUL(eu, eu), DM_MUL(dg, dg)));
  if (metric == B200_METRIC_EFF_DIOU) return diou;
  float q = DM_DIV(DM_SUB(a.at, b.at), B200_PI_F);
  float v = DM_MUL(4.0f, DM_MUL(q, q));
  float alpha = bm_dnn(v, DM_ADD(DM_SUB(1.0f, iou), v));
  return DM_SUB(diou, DM_MUL(alpha, v));
}

// Division-free reject for overlapping boxes.  Both overlap extents are positive here, so both boxes have positive width
// and height, the computed intersection I = iw*ih is <= both computed areas (fp32 rounding is monotone) and
// IoU = I / (s - I) with s = area_a + area_b > 2 I.  IoU < t  <=>  I (1 + t) < t s; testing against t = 0.999 thr leaves a
// 1e-3 relative margin over the few-ulp rounding of I, s, the division and of the metric itself (every metric of both
// families is its IoU minus non-negative terms, up to an ulp).  NMS on dense heads tests ~100 overlapping pairs for every
// one near the threshold: this keeps the two to four IEEE divisions of the full metric off all of them.
__device__ __forceinline__ bool bm_iou_clearly_below(float iw, float ih, float s, float thr) {
  return (iw * ih) * (1.0f + thr) < (0.999f * thr) * s;
}

// Every metric is <= its plain IoU, and IoU is exactly +0 when the boxes do not overlap and the
// union is positive and finite.  For a positive threshold such a pair can neither suppress
// (metric >= thr) nor be ignored-in-loss (metric >= thr); this is the cheap reject used before the
// full evaluation.  Anything degenerate (NaN/inf/zero union) returns false and takes the full path.
__device__ __forceinline__ bool bm_surely_below(const BoxT& a, const BoxT& b, int metric, float thr) {
  if (!(thr > 0.0f)) return false;
  float s = DM_ADD(a.area, b.area);
  if (!(s > 0.0f) || !(s < 3.0e38f)) return false;
  if (a.at != a.at || b.at != b.at) return false;  // NaN aspect term (0/0 box): full path decides
  if (metric <= B200_METRIC_YOLO_CIOU) {
    float iw = DM_SUB(dm_min(a.c2, b.c2), dm_max(a.c0, b.c0));
    float ih = DM_SUB(dm_min(a.c3, b.c3), dm_max(a.c1, b.c1));
    // both extents must be finite so the product max(iw,0)*max(ih,0) is exactly 0 (not 0*inf)
    if (!(dm_fabsf(iw) < 3.0e38f) || !(dm_fabsf(ih) < 3.0e38f)) return false;
    return (iw <= 0.0f) || (ih <= 0.0f) || bm_iou_clearly_below(iw, ih, s, thr);
  }
  float iw = DM_SUB(dm_min(a.c3, b.c3), dm_max(a.c1, b.c1));
  float ih = DM_SUB(dm_min(a.c2, b.c2), dm_max(a.c0, b.c0));
  if (!(dm_fabsf(iw) < 3.0e38f) || !(dm_fabsf(ih) < 3.0e38f)) return false;
  return (iw <= 0.0f) || (ih <= 0.0f) || bm_iou_clearly_below(iw, ih, s, thr);
}
