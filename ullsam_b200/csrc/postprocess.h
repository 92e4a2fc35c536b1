// ullsam_b200 — mask post-processing (postprocess.cu): the step immediately downstream of the mask decoder
// (reference modeling/sam.py:133-162, train_joint_v2.py:1073-1079, app.py:635-644).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ullsam {

// F.interpolate(x[:, :, :crop_h, :crop_w], (out_h, out_w), mode="bilinear", align_corners=False) for n = B*C planes of
// in_h x in_w fp32, fused with the binarisation `> threshold`: writes fp32 logits (out_f32, may be null) and / or a
// uint8 0/1 mask (out_u8, may be null).  Returns nullptr or a static error string.
const char* launch_resize_masks(const float* x, int n, int in_h, int in_w, int crop_h, int crop_w, int out_h, int out_w,
                                float threshold, float* out_f32, uint8_t* out_u8, cudaStream_t stream);

// CalcIoU (train_joint_v2.py:666-696) on binarised predictions: per plane, intersection and union counts of
// (logit > threshold) against a 0/1 target (fp32); iou = (inter + smooth) / (union + smooth).
const char* launch_mask_iou(const float* logits, const float* target, int n, size_t hw, float threshold, float smooth, float* iou,
                            cudaStream_t stream);

}  // namespace ullsam
