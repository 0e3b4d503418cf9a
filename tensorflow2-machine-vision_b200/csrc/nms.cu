// nms.cu — generic segmented NMS entry point (one CTA per segment); see nms.cuh for the algorithm.
#include "nms.cuh"
#include "../../include/b200det.h"

struct NmsBatchParams {
  const float* boxes;
  const float* scores;
  const int32_t* classes;
  const uint32_t* order_id;
  const int32_t* seg_offsets;
  NmsConfig cfg;
  int32_t* out_idx;
  int32_t* out_count;
};

template <int METRIC>
__global__ void __launch_bounds__(NMS_THREADS, 1) nms_batch_kernel(NmsBatchParams p) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  const int s = blockIdx.x;
  const int beg = p.seg_offsets[s], end = p.seg_offsets[s + 1];
  NmsSegment seg;
  seg.boxes = p.boxes + 4 * (size_t)beg;
  seg.scores = p.scores + beg;
  seg.classes = p.classes ? p.classes + beg : nullptr;
  seg.order_id = p.order_id ? p.order_id + beg : nullptr;
  seg.n = end - beg;
  int kept = nms_run_segment<METRIC>(seg, p.cfg, p.out_idx + (size_t)s * p.cfg.max_out, nms_smem);
  if (threadIdx.x == 0) p.out_count[s] = kept;
}

extern "C" int b200_nms(const float* boxes, const float* scores, const int32_t* classes, const uint32_t* order_id,
                        const int32_t* seg_offsets, int num_segments, int metric, int mode, float iou_thr,
                        int use_score_thr, float score_thr, int max_out, int32_t* out_idx, int32_t* out_count,
                        void* stream) {
  B200_REQUIRE(metric >= 0 && metric < B200_METRIC_COUNT, B200_ERR_BAD_ARG, "b200_nms: bad metric %d", metric);
  B200_REQUIRE(mode == B200_NMS_AGNOSTIC || mode == B200_NMS_BY_CLASS, B200_ERR_BAD_ARG, "b200_nms: bad mode %d", mode);
  B200_REQUIRE(num_segments >= 0, B200_ERR_BAD_ARG, "b200_nms: negative segment count");
  B200_REQUIRE(max_out >= 1 && max_out <= NMS_MAX_OUT_LIMIT, B200_ERR_UNSUPPORTED,
               "b200_nms: max_out %d outside [1, %d]", max_out, NMS_MAX_OUT_LIMIT);
  if (num_segments == 0) return B200_OK;
  B200_REQUIRE(boxes && scores && seg_offsets && out_idx && out_count, B200_ERR_BAD_ARG, "b200_nms: null pointer");
  B200_REQUIRE(mode != B200_NMS_BY_CLASS || classes, B200_ERR_BAD_ARG, "b200_nms: BY_CLASS needs classes");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, B200_ERR_BAD_ARG, "b200_nms: boxes must be 16-byte aligned");
  NmsBatchParams p;
  p.boxes = boxes; p.scores = scores; p.classes = classes; p.order_id = order_id; p.seg_offsets = seg_offsets;
  p.cfg.metric = metric; p.cfg.mode = mode; p.cfg.iou_thr = iou_thr; p.cfg.score_thr = score_thr;
  p.cfg.use_score_thr = use_score_thr; p.cfg.max_out = max_out;
  p.out_idx = out_idx; p.out_count = out_count;
  size_t smem = nms_smem_bytes(max_out);
#define NMS_LAUNCH(M)                                                                                          \
  B200_CUDA(cudaFuncSetAttribute(nms_batch_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
  nms_batch_kernel<M><<<num_segments, NMS_THREADS, smem, (cudaStream_t)stream>>>(p)
  NMS_DISPATCH_METRIC(metric, NMS_LAUNCH)
#undef NMS_LAUNCH
  B200_LAUNCH_CHECK();
  return B200_OK;
}
